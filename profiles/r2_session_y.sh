set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2y_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2y_pytest.log; tail -4 gpurun_out/r2y_pytest.log
for bh in 0 1 0 1; do
  NK_BIN_HITS=$bh python bench.py --case c1 --eta 5 --particles 2e7 --steps 30 --warmup 5 --no-cpu --sustained-steps 0 --e2e-calls 1 > gpurun_out/r2y_c1_$bh.json 2> gpurun_out/r2y_c1_$bh.err
  python -c "import json,sys; d=json.load(open('gpurun_out/r2y_c1_$bh.json')); r=d['roofline']; print('bin_hits=$bh c1', d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac'], r['kernel_share_of_step'])"
  NK_BIN_HITS=$bh python bench.py --particles 1e8 --steps 30 --warmup 5 --no-cpu --sustained-steps 0 --e2e-calls 1 > gpurun_out/r2y_film_$bh.json 2> gpurun_out/r2y_film_$bh.err
  python -c "import json,sys; d=json.load(open('gpurun_out/r2y_film_$bh.json')); r=d['roofline']; print('bin_hits=$bh film', d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac'], r['kernel_share_of_step'])"
done
