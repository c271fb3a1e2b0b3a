set -x
cd $GRAFT_REPO_ROOT
for lib in libnk_b200.so libnk_b200_sf.so libnk_b200.so libnk_b200_sf.so; do
  NK_LIB=$PWD/nanokappa_b200/$lib python bench.py --case c1 --eta 5 --particles 2e7 --steps 30 --warmup 5 --no-cpu --sustained-steps 0 --e2e-calls 1 > gpurun_out/r2zz_c1_$lib.json 2> gpurun_out/r2zz_c1_$lib.err
  python -c "import json,sys; d=json.load(open('gpurun_out/r2zz_c1_$lib.json')); r=d['roofline']; print('$lib c1', d['value'], d['ms_per_step'], r['avg_launch_ms'], r['kernel_share_of_step'])"
  NK_LIB=$PWD/nanokappa_b200/$lib python bench.py --particles 1e8 --steps 30 --warmup 5 --no-cpu --sustained-steps 0 --e2e-calls 1 > gpurun_out/r2zz_film_$lib.json 2> gpurun_out/r2zz_film_$lib.err
  python -c "import json,sys; d=json.load(open('gpurun_out/r2zz_film_$lib.json')); r=d['roofline']; print('$lib film', d['value'], d['ms_per_step'], r['avg_launch_ms'], r['kernel_share_of_step'])"
done
