set -x
cd $GRAFT_REPO_ROOT
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29641 bench.py --gpus $N --steps 20 --warmup 3 --sustained-steps 500 --e2e-calls 2 > gpurun_out/r2s_${N}gpu_default.out 2> gpurun_out/r2s_${N}_a.err; grep "^{" gpurun_out/r2s_${N}gpu_default.out | tail -1 > gpurun_out/r2_final_bench_${N}gpu_1e8_per_gpu.json; tail -c 300 gpurun_out/r2_final_bench_${N}gpu_1e8_per_gpu.json
$TR --master-port 29642 bench.py --gpus $N --particles 1.25e6 --steps 50 --warmup 5 --sustained-steps 500 --e2e-calls 1 > gpurun_out/r2s_${N}gpu_small.out 2> gpurun_out/r2s_${N}_b.err; grep "^{" gpurun_out/r2s_${N}gpu_small.out | tail -1 > gpurun_out/r2_final_bench_${N}gpu_1.25e6_per_gpu.json
$TR --master-port 29643 bench.py --gpus $N --particles 1.25e8 --steps 20 --warmup 3 --sustained-steps 300 --e2e-calls 1 > gpurun_out/r2s_${N}gpu_big.out 2> gpurun_out/r2s_${N}_c.err; grep "^{" gpurun_out/r2s_${N}gpu_big.out | tail -1 > gpurun_out/r2_final_bench_${N}gpu_1.25e8_per_gpu.json
rm -f gpurun_out/*.out
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_final_bench_${N}gpu_*.json')):
    try:
        d=json.load(open(f)); s=d.get('sustained') or {}
        print(f.split('/')[-1], 'value %.3e'%d['value'], 'ms %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'e2e %.3e'%d['e2e']['value'], 'sustained %.3e'%s.get('value',0))
    except Exception as e: print(f, 'ERR', e)
PY
