#!/bin/bash
# round 2, last validation: GPU suite + the default bench line with the PCIe ceiling measured inside the run
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/ac_pytest.txt
cat gpurun_out/ac_pytest.txt
timeout 900 python bench.py > gpurun_out/ac_bench_default.json 2> gpurun_out/ac_bench_default.err
echo "bench rc $?"
tail -c 3000 gpurun_out/ac_bench_default.json
tail -5 gpurun_out/ac_bench_default.err
