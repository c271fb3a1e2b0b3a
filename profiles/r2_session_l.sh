set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_shards.py tests/test_gpu_population.py -m gpu -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log; tail -4 gpurun_out/r2l_pytest.log
python tests/run_cli_timing.py iterations=10000 > gpurun_out/r2l_cli.json 2> gpurun_out/r2l_cli.err; cat gpurun_out/r2l_cli.json; tail -5 gpurun_out/r2l_cli.err
