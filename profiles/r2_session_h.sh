set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q -x > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log; tail -4 gpurun_out/r2h_pytest.log
python tests/run_intersection_bench.py sides=250,2500 rays=1e7 steps=20 > gpurun_out/r2h_intersect.jsonl 2> gpurun_out/r2h_intersect.err; cat gpurun_out/r2h_intersect.jsonl; tail -3 gpurun_out/r2h_intersect.err
for lib in libnk_b200.so libnk_b200_rare128.so; do
  NK_LIB=$PWD/nanokappa_b200/$lib python bench.py --case c1 --eta 5 --particles 2e7 --steps 20 --warmup 5 --no-cpu --sustained-steps 0 --e2e-calls 1 > gpurun_out/r2h_c1_$lib.json 2> gpurun_out/r2h_c1_$lib.err
  python -c "import json,sys; d=json.load(open('gpurun_out/r2h_c1_$lib.json')); r=d['roofline']; print('$lib c1', d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac'], r['kernel_share_of_step'])"
  NK_LIB=$PWD/nanokappa_b200/$lib python bench.py --particles 1e8 --steps 20 --warmup 5 --no-cpu --sustained-steps 0 --e2e-calls 1 > gpurun_out/r2h_film_$lib.json 2> gpurun_out/r2h_film_$lib.err
  python -c "import json,sys; d=json.load(open('gpurun_out/r2h_film_$lib.json')); r=d['roofline']; print('$lib film', d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac'], r['kernel_share_of_step'])"
done
