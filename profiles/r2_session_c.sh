set -x
cd $GRAFT_REPO_ROOT
python tests/run_intersection_bench.py sides=250,2500 rays=1e7 steps=20 > gpurun_out/r2c_intersect.jsonl 2> gpurun_out/r2c_intersect.err; tail -3 gpurun_out/r2c_intersect.jsonl; tail -3 gpurun_out/r2c_intersect.err
python bench.py --case c1 --eta 5 --particles 2e7 --steps 20 --warmup 5 --no-cpu --sustained-steps 0 --e2e-calls 1 > gpurun_out/r2c_bench_c1.json 2> gpurun_out/r2c_bench_c1.err; tail -c 1500 gpurun_out/r2c_bench_c1.json
NK_STEP_TAB=0 python bench.py --particles 1e8 --steps 20 --warmup 3 --no-cpu --sustained-steps 0 --e2e-calls 1 > gpurun_out/r2c_bench_direct.json 2> gpurun_out/r2c_bench_direct.err; tail -c 700 gpurun_out/r2c_bench_direct.json
# ncu captures (each command has exited 0 above or is checked by the wrapper)
ncu --set full --clock-control none --import-source on -k regex:'k_rare|k_step_tab' -s 12 -c 2 -o gpurun_out/r2c_film python bench.py --particles 1e8 --steps 3 --warmup 3 --no-cpu --e2e-calls 1 --sustained-steps 0 > gpurun_out/r2c_ncu_film.log 2>&1; tail -2 gpurun_out/r2c_ncu_film.log
ncu --set full --clock-control none --import-source on -k regex:'k_rare|k_step' -s 12 -c 2 -o gpurun_out/r2c_c1 python bench.py --case c1 --eta 5 --particles 2e7 --steps 3 --warmup 3 --no-cpu --e2e-calls 1 --sustained-steps 0 > gpurun_out/r2c_ncu_c1.log 2>&1; tail -2 gpurun_out/r2c_ncu_c1.log
ncu --set full --clock-control none --import-source on -k regex:k_init_collisions -s 1 -c 1 -o gpurun_out/r2c_isect_f1000 python tests/run_intersection_bench.py sides=250 rays=4e6 reps=1 > gpurun_out/r2c_ncu_i1.log 2>&1; tail -2 gpurun_out/r2c_ncu_i1.log
ncu --set full --clock-control none --import-source on -k regex:k_init_collisions -s 1 -c 1 -o gpurun_out/r2c_isect_f10000 python tests/run_intersection_bench.py sides=2500 rays=1e6 reps=1 > gpurun_out/r2c_ncu_i2.log 2>&1; tail -2 gpurun_out/r2c_ncu_i2.log
ls -la gpurun_out/
