set -x
cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/run_multi_gpu_check.py > gpurun_out/r2m_check_2gpu.txt 2>&1; tail -6 gpurun_out/r2m_check_2gpu.txt
python tests/run_multi_gpu_cli.py 2 > gpurun_out/r2m_cli_2gpu.txt 2>&1; tail -12 gpurun_out/r2m_cli_2gpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2m_bench_2gpu.json 2> gpurun_out/r2m_bench_2gpu.err; tail -c 2500 gpurun_out/r2m_bench_2gpu.json; tail -3 gpurun_out/r2m_bench_2gpu.err
