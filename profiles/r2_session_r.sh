set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests -m gpu -q -x -k "stl or c9 or contains or arbitrary or find_boundary or init_collisions" > gpurun_out/r2r_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2r_pytest.log; tail -4 gpurun_out/r2r_pytest.log
timeout 300 python tests/run_intersection_bench.py sides=250,2500 rays=1e7 steps=20 > gpurun_out/r2r_intersect.jsonl 2> gpurun_out/r2r_intersect.err; cat gpurun_out/r2r_intersect.jsonl; tail -3 gpurun_out/r2r_intersect.err
