set -x
cd $GRAFT_REPO_ROOT
for lib in libnk_b200.so libnk_b200_tab2.so libnk_b200.so libnk_b200_tab2.so; do
  NK_LIB=$PWD/nanokappa_b200/$lib python bench.py --particles 1e8 --steps 40 --warmup 5 --no-cpu --sustained-steps 300 --e2e-calls 1 > gpurun_out/r2t_film_$lib.json 2> gpurun_out/r2t_film_$lib.err
  python -c "import json,sys; d=json.load(open('gpurun_out/r2t_film_$lib.json')); r=d['roofline']; s=d['sustained']; print('$lib film', d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac'], 'sustained', s['value'], s['kstep_avg_ms'])"
done
NK_LIB=$PWD/nanokappa_b200/libnk_b200_tab2.so timeout 600 python -m pytest tests -m gpu -q -x -k "variants or film or fullsize or readme or tau_slab" > gpurun_out/r2t_pytest.log 2>&1; tail -3 gpurun_out/r2t_pytest.log
