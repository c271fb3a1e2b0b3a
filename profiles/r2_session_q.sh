set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_population.py -m gpu -q -x -k "contains or arbitrary" > gpurun_out/r2q_pytest.log 2>&1; tail -25 gpurun_out/r2q_pytest.log
