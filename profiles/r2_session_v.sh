set -x
cd $GRAFT_REPO_ROOT
for lib in libnk_b200.so libnk_b200_pre2.so; do
  NK_STEP_TAB=0 NK_LIB=$PWD/nanokappa_b200/$lib python bench.py --particles 1e8 --steps 30 --warmup 5 --no-cpu --sustained-steps 0 --e2e-calls 1 > gpurun_out/r2v_direct_$lib.json 2> gpurun_out/r2v_direct_$lib.err
  python -c "import json,sys; d=json.load(open('gpurun_out/r2v_direct_$lib.json')); r=d['roofline']; print('$lib direct', d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac'])"
  NK_LIB=$PWD/nanokappa_b200/$lib python bench.py --case c1 --eta 5 --particles 2e7 --steps 30 --warmup 5 --no-cpu --sustained-steps 0 --e2e-calls 1 > gpurun_out/r2v_c1_$lib.json 2> gpurun_out/r2v_c1_$lib.err
  python -c "import json,sys; d=json.load(open('gpurun_out/r2v_c1_$lib.json')); r=d['roofline']; print('$lib c1', d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac'])"
  NK_LIB=$PWD/nanokappa_b200/$lib python bench.py --particles 1.25e7 --slices 100 --material ge --steps 30 --warmup 5 --no-cpu --sustained-steps 0 --e2e-calls 1 > gpurun_out/r2v_ge_$lib.json 2> gpurun_out/r2v_ge_$lib.err
  python -c "import json,sys; d=json.load(open('gpurun_out/r2v_ge_$lib.json')); r=d['roofline']; print('$lib ge s100', d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac'])"
done
NK_LIB=$PWD/nanokappa_b200/libnk_b200_pre2.so timeout 600 python -m pytest tests/test_gpu_parity_scale.py -m gpu -q -x -k "many_subvolumes" > gpurun_out/r2v_pytest.log 2>&1; tail -3 gpurun_out/r2v_pytest.log
