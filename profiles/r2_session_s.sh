set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2s_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2s_pytest.log; tail -4 gpurun_out/r2s_pytest.log
python bench.py --case c1 --eta 5 --particles 2e7 --steps 20 --warmup 5 --no-cpu --sustained-steps 0 --e2e-calls 1 > gpurun_out/r2s_c1.json 2> gpurun_out/r2s_c1.err
python -c "import json,sys; d=json.load(open('gpurun_out/r2s_c1.json')); r=d['roofline']; print('c1', d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac'], r['kernel_share_of_step'])"
python bench.py --particles 1e8 --steps 20 --warmup 5 --no-cpu --sustained-steps 0 --e2e-calls 1 > gpurun_out/r2s_film.json 2> gpurun_out/r2s_film.err
python -c "import json,sys; d=json.load(open('gpurun_out/r2s_film.json')); r=d['roofline']; print('film', d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac'], r['kernel_share_of_step'])"
