set -x
cd $GRAFT_REPO_ROOT
python bench.py --steps 20 --warmup 3 --no-cpu --e2e-calls 1 --sustained-steps 10000 > gpurun_out/r2_bench_1gpu_10000_steps.json 2> gpurun_out/r2long.err; tail -c 1500 gpurun_out/r2_bench_1gpu_10000_steps.json; tail -3 gpurun_out/r2long.err
