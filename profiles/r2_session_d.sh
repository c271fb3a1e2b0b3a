set -x
cd $GRAFT_REPO_ROOT
python tests/run_intersection_bench.py sides=250,2500 rays=1e7 steps=20 > gpurun_out/r2d_intersect.jsonl 2> gpurun_out/r2d_intersect.err; tail -3 gpurun_out/r2d_intersect.jsonl; tail -3 gpurun_out/r2d_intersect.err
ncu --set full --clock-control none --import-source on -k regex:'k_rare|k_step_tab' -s 8 -c 2 -o gpurun_out/r2d_film python bench.py --particles 1e8 --steps 3 --warmup 3 --no-cpu --e2e-calls 1 --sustained-steps 0 > gpurun_out/r2d_ncu_film.log 2>&1; tail -2 gpurun_out/r2d_ncu_film.log
ncu --set full --clock-control none --import-source on -k regex:'k_rare|k_step' -s 8 -c 2 -o gpurun_out/r2d_c1 python bench.py --case c1 --eta 5 --particles 2e7 --steps 3 --warmup 3 --no-cpu --e2e-calls 1 --sustained-steps 0 > gpurun_out/r2d_ncu_c1.log 2>&1; tail -2 gpurun_out/r2d_ncu_c1.log
