set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q -x > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log; tail -6 gpurun_out/r2j_pytest.log
python tests/run_cli_timing.py iterations=10000 > gpurun_out/r2j_cli.json 2> gpurun_out/r2j_cli.err; cat gpurun_out/r2j_cli.json; tail -5 gpurun_out/r2j_cli.err
python tests/run_pcie_ceiling.py gb=4 > gpurun_out/r2j_pcie_1gpu.json 2> gpurun_out/r2j_pcie.err; cat gpurun_out/r2j_pcie_1gpu.json
