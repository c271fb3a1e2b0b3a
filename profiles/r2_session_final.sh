set -x
cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/r2z_bench_default.json 2> gpurun_out/r2z_bench_default.err; tail -c 400 gpurun_out/r2z_bench_default.json; tail -3 gpurun_out/r2z_bench_default.err
python bench.py --case c1 --eta 5 --particles 2e7 --steps 20 --warmup 5 --no-cpu --sustained-steps 500 --e2e-calls 1 > gpurun_out/r2z_bench_c1_diagnostic.json 2> gpurun_out/r2z_c1.err; tail -c 300 gpurun_out/r2z_bench_c1_diagnostic.json
NK_STEP_TAB=0 python bench.py --steps 20 --warmup 3 --no-cpu --sustained-steps 0 --e2e-calls 1 > gpurun_out/r2z_bench_direct_kernel.json 2> gpurun_out/r2z_direct.err
python bench.py --steps 5 --warmup 3 --no-cpu --e2e-calls 1 --sustained-steps 0 > gpurun_out/r2z_plain_for_ncu.json 2> gpurun_out/r2z_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2z_launches_1e8.csv python bench.py --steps 5 --warmup 3 --no-cpu --e2e-calls 1 --sustained-steps 0 > gpurun_out/r2z_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_rare|k_step_tab|k_mode_tables' -s 12 -c 3 -o gpurun_out/r2z_film python bench.py --particles 1e8 --steps 3 --warmup 3 --no-cpu --e2e-calls 1 --sustained-steps 0 > gpurun_out/r2z_ncu_film.log 2>&1; tail -2 gpurun_out/r2z_ncu_film.log
ncu --set full --clock-control none --import-source on -k regex:'k_rare|k_step' -s 8 -c 2 -o gpurun_out/r2z_c1 python bench.py --case c1 --eta 5 --particles 2e7 --steps 3 --warmup 3 --no-cpu --e2e-calls 1 --sustained-steps 0 > gpurun_out/r2z_ncu_c1.log 2>&1; tail -2 gpurun_out/r2z_ncu_c1.log
ncu --set full --clock-control none --import-source on -k regex:k_init_collisions -s 1 -c 1 -o gpurun_out/r2z_isect_f1000 python tests/run_intersection_bench.py sides=250 rays=4e6 reps=1 > gpurun_out/r2z_ncu_i1.log 2>&1; tail -2 gpurun_out/r2z_ncu_i1.log
ncu --set full --clock-control none --import-source on -k regex:k_init_collisions -s 1 -c 1 -o gpurun_out/r2z_isect_f10000 python tests/run_intersection_bench.py sides=2500 rays=1e6 reps=1 > gpurun_out/r2z_ncu_i2.log 2>&1; tail -2 gpurun_out/r2z_ncu_i2.log
ncu --set full --clock-control none -k regex:'k_sort_permute|k_sort_rank|k_sort_hist' -s 3 -c 3 -o gpurun_out/r2z_sort python tests/run_order_decay.py 1e8 pools= > gpurun_out/r2z_ncu_sort.log 2>&1; tail -2 gpurun_out/r2z_ncu_sort.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2z_bench_reference_arm.json 2> gpurun_out/r2z_ref.err; tail -c 900 gpurun_out/r2z_bench_reference_arm.json; tail -3 gpurun_out/r2z_ref.err
ls -la gpurun_out | tail -20
