set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_shards.py tests/test_gpu_population.py -m gpu -q -x > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log; tail -4 gpurun_out/r2k_pytest.log
python tests/run_cli_timing.py iterations=10000 profile=1 > gpurun_out/r2k_cli.json 2> gpurun_out/r2k_cli.err; cat gpurun_out/r2k_cli.json; tail -5 gpurun_out/r2k_cli.err
python -c "
import pstats; p=pstats.Stats('gpurun_out/cli_profile.pstats'); p.sort_stats('cumulative').print_stats(45)" > gpurun_out/r2k_cli_profile.txt 2>&1; head -80 gpurun_out/r2k_cli_profile.txt | cut -c1-150
