set -x
cd $GRAFT_REPO_ROOT
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29621 tests/run_multi_gpu_check.py > gpurun_out/r2n_check_${N}gpu.txt 2>&1; tail -5 gpurun_out/r2n_check_${N}gpu.txt
# configs[2]: Ge cell, 100 slices, 1e8 particles over 8 GPUs
$TR --master-port 29622 bench.py --gpus $N --particles 1.25e7 --slices 100 --material ge --steps 20 --warmup 3 --sustained-steps 500 --e2e-calls 1 > gpurun_out/r2n_bench_${N}gpu_ge_s100_1e8total.json 2> gpurun_out/r2n_ge.err; tail -c 600 gpurun_out/r2n_bench_${N}gpu_ge_s100_1e8total.json; tail -2 gpurun_out/r2n_ge.err
# configs[4] ends: 1e7 and 1e9 particles in total
$TR --master-port 29623 bench.py --gpus $N --particles 1.25e6 --steps 50 --warmup 5 --sustained-steps 1000 --e2e-calls 2 > gpurun_out/r2n_bench_${N}gpu_1e7total.json 2> gpurun_out/r2n_1e7.err; tail -c 600 gpurun_out/r2n_bench_${N}gpu_1e7total.json; tail -2 gpurun_out/r2n_1e7.err
$TR --master-port 29624 bench.py --gpus $N --particles 1.25e8 --steps 20 --warmup 3 --sustained-steps 500 --e2e-calls 1 > gpurun_out/r2n_bench_${N}gpu_1e9total.json 2> gpurun_out/r2n_1e9.err; tail -c 600 gpurun_out/r2n_bench_${N}gpu_1e9total.json; tail -2 gpurun_out/r2n_1e9.err
$TR --master-port 29625 tests/run_pcie_ceiling.py gb=2 > gpurun_out/r2n_pcie_${N}gpu.json 2> gpurun_out/r2n_pcie.err; cat gpurun_out/r2n_pcie_${N}gpu.json
$TR --master-port 29626 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2n_bench_${N}gpu_default.json 2> gpurun_out/r2n_default.err; tail -c 600 gpurun_out/r2n_bench_${N}gpu_default.json; tail -2 gpurun_out/r2n_default.err
