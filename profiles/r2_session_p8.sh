set -x
cd $GRAFT_REPO_ROOT
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29632 bench.py --gpus $N --particles 1.25e7 --slices 100 --material ge --steps 20 --warmup 3 --sustained-steps 1000 --e2e-calls 1 > gpurun_out/r2p_bench_${N}gpu_ge_s100_1e8total.json 2> gpurun_out/r2p_ge.err; tail -c 900 gpurun_out/r2p_bench_${N}gpu_ge_s100_1e8total.json; tail -2 gpurun_out/r2p_ge.err
python tests/run_multi_gpu_cli.py 8 > gpurun_out/r2p_cli_8gpu.txt 2>&1; tail -3 gpurun_out/r2p_cli_8gpu.txt
