#!/bin/bash
# re-entry check: full gpu test suite, default bench, DSMEM + line-grouped probe microbenches
mkdir -p gpurun_out
T0=$SECONDS
timeout 120 ./ntsm_b200/bin/microbench 24 > gpurun_out/microbench_dsmem.txt 2>&1; cat gpurun_out/microbench_dsmem.txt
echo "microbench done ($((SECONDS-T0)) s)"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? ($((SECONDS-T0)) s)"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r6_bench.json 2> gpurun_out/r6_bench.log; echo "bench rc=$? ($((SECONDS-T0)) s)"; tail -3 gpurun_out/r6_bench.log
cat gpurun_out/r6_bench.json
echo "total $((SECONDS-T0)) s"
