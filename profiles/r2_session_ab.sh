set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2ab_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2ab_pytest.log; tail -4 gpurun_out/r2ab_pytest.log
python bench.py --steps 30 --warmup 5 --no-cpu --sustained-steps 1000 --e2e-calls 1 > gpurun_out/r2ab_film.json 2> gpurun_out/r2ab_film.err
python -c "import json,sys; d=json.load(open('gpurun_out/r2ab_film.json')); r=d['roofline']; s=d['sustained']; print('film', d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac'], r['kernel_share_of_step'], 'sustained', s['value'], s['ms_per_step'])"
