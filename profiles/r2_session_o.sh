set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "variants or sort_by_mode or slot" > gpurun_out/r2o_pytest.log 2>&1; tail -3 gpurun_out/r2o_pytest.log
python bench.py --particles 1.25e7 --slices 100 --material ge --steps 20 --warmup 3 --sustained-steps 2000 --e2e-calls 1 --no-cpu > gpurun_out/r2o_bench_1gpu_ge_s100.json 2> gpurun_out/r2o_ge.err; tail -c 1200 gpurun_out/r2o_bench_1gpu_ge_s100.json; tail -2 gpurun_out/r2o_ge.err
python bench.py --particles 1.25e6 --steps 50 --warmup 5 --sustained-steps 1000 --e2e-calls 2 --no-cpu > gpurun_out/r2o_bench_1gpu_1.25e6.json 2> gpurun_out/r2o_s.err; tail -c 700 gpurun_out/r2o_bench_1gpu_1.25e6.json
python bench.py --particles 1.25e8 --steps 20 --warmup 3 --sustained-steps 500 --e2e-calls 1 --no-cpu > gpurun_out/r2o_bench_1gpu_1.25e8.json 2> gpurun_out/r2o_b.err; tail -c 700 gpurun_out/r2o_bench_1gpu_1.25e8.json
