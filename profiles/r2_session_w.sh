set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2w_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2w_pytest.log; tail -5 gpurun_out/r2w_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w_smoke.log 2>&1; tail -2 gpurun_out/r2w_smoke.log
