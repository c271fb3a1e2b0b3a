set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python bench.py > $O/r2_bench_default.json 2> $O/r2z_bench_default.err; tail -c 300 $O/r2_bench_default.json; tail -3 $O/r2z_bench_default.err
python bench.py --case c1 --eta 5 --particles 2e7 --steps 20 --warmup 5 --no-cpu --sustained-steps 500 --e2e-calls 1 > $O/r2_bench_c1_diagnostic.json 2> $O/r2z_c1.err
NK_STEP_TAB=0 python bench.py --steps 20 --warmup 3 --no-cpu --sustained-steps 0 --e2e-calls 1 > $O/r2_bench_direct_kernel.json 2> $O/r2z_direct.err
python bench.py --steps 5 --warmup 3 --no-cpu --e2e-calls 1 --sustained-steps 0 > $O/r2z_plain_for_ncu.json 2> $O/r2z_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r2_launches_1e8.csv python bench.py --steps 5 --warmup 3 --no-cpu --e2e-calls 1 --sustained-steps 0 > $O/r2z_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_rare|k_step_tab|k_mode_tables' -s 12 -c 3 -o /tmp/r2_film python bench.py --particles 1e8 --steps 3 --warmup 3 --no-cpu --e2e-calls 1 --sustained-steps 0 > $O/r2z_ncu_film.log 2>&1
python profiles/ncu_summary.py /tmp/r2_film.ncu-rep > $O/r2_film_kstep_tab_krare_kmodetables_ncu_summary.txt 2>&1
python profiles/ncu_hotspots.py /tmp/r2_film.ncu-rep k_rare 30 > $O/r2_film_krare_hotspots.txt 2>&1
python profiles/ncu_hotspots.py /tmp/r2_film.ncu-rep k_step_tab 25 > $O/r2_film_kstep_tab_hotspots.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_rare|k_step' -s 8 -c 2 -o /tmp/r2_c1 python bench.py --case c1 --eta 5 --particles 2e7 --steps 3 --warmup 3 --no-cpu --e2e-calls 1 --sustained-steps 0 > $O/r2z_ncu_c1.log 2>&1
python profiles/ncu_summary.py /tmp/r2_c1.ncu-rep > $O/r2_c1_kstep_krare_ncu_summary.txt 2>&1
python profiles/ncu_hotspots.py /tmp/r2_c1.ncu-rep k_rare 30 > $O/r2_c1_krare_hotspots.txt 2>&1
python profiles/ncu_hotspots.py /tmp/r2_c1.ncu-rep 'k_step<' 30 > $O/r2_c1_kstep_hotspots.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_init_collisions -s 1 -c 1 -o /tmp/r2_isect_f1000 python tests/run_intersection_bench.py sides=250 rays=4e6 reps=1 > $O/r2z_ncu_i1.log 2>&1
python profiles/ncu_summary.py /tmp/r2_isect_f1000.ncu-rep > $O/r2_isect_f1000_ncu_summary.txt 2>&1
python profiles/ncu_hotspots.py /tmp/r2_isect_f1000.ncu-rep k_init_collisions 20 > $O/r2_isect_f1000_hotspots.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_init_collisions -s 1 -c 1 -o /tmp/r2_isect_f10000 python tests/run_intersection_bench.py sides=2500 rays=1e6 reps=1 > $O/r2z_ncu_i2.log 2>&1
python profiles/ncu_summary.py /tmp/r2_isect_f10000.ncu-rep > $O/r2_isect_f10000_ncu_summary.txt 2>&1
ncu --set full --clock-control none -k regex:'k_sort_permute|k_sort_rank|k_sort_hist' -s 3 -c 3 -o /tmp/r2_sort python tests/run_order_decay.py 1e8 pools= > $O/r2z_ncu_sort.log 2>&1
python profiles/ncu_summary.py /tmp/r2_sort.ncu-rep > $O/r2_sort_kernels_ncu_summary.txt 2>&1
cp /tmp/r2_film.ncu-rep $O/r2_film.ncu-rep
du -sh $O; ls -la $O | tail -30
