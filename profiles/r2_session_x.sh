set -x
cd $GRAFT_REPO_ROOT
for lib in libnk_b200.so libnk_b200_tg.so libnk_b200.so libnk_b200_tg.so; do
  NK_LIB=$PWD/nanokappa_b200/$lib python bench.py --case c1 --eta 5 --particles 2e7 --steps 30 --warmup 5 --no-cpu --sustained-steps 0 --e2e-calls 1 > gpurun_out/r2x_c1_$lib.json 2> gpurun_out/r2x_c1_$lib.err
  python -c "import json,sys; d=json.load(open('gpurun_out/r2x_c1_$lib.json')); r=d['roofline']; print('$lib c1', d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac'])"
done
NK_LIB=$PWD/nanokappa_b200/libnk_b200_tg.so timeout 600 python -m pytest tests -m gpu -q -x -k "c1 or parameters_test or c5 or c4 or c6 or variants" > gpurun_out/r2x_pytest.log 2>&1; tail -3 gpurun_out/r2x_pytest.log
