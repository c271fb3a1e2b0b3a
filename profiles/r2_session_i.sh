set -x
cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:'k_rare' -s 4 -c 1 -o gpurun_out/r2i_c1_rare python bench.py --case c1 --eta 5 --particles 2e7 --steps 3 --warmup 3 --no-cpu --e2e-calls 1 --sustained-steps 0 > gpurun_out/r2i_ncu_c1.log 2>&1; tail -2 gpurun_out/r2i_ncu_c1.log
ncu --set full --clock-control none --import-source on -k regex:k_init_collisions -s 1 -c 1 -o gpurun_out/r2i_isect_f1000 python tests/run_intersection_bench.py sides=250 rays=4e6 reps=1 > gpurun_out/r2i_ncu_i1.log 2>&1; tail -2 gpurun_out/r2i_ncu_i1.log
