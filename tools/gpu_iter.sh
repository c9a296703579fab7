#!/bin/bash
# quick iteration: parity tests for both kernel variants + microbench(count) + short bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest(v2) rc=$?"; tail -3 gpurun_out/pytest_gpu.log
NTSM_KERNEL=0 python -m pytest tests -m gpu -x -q -k "oracle or properties" > gpurun_out/pytest_gpu_v1.log 2>&1; echo "pytest(v1) rc=$?"; tail -2 gpurun_out/pytest_gpu_v1.log
ntsm_b200/bin/microbench 4 2>&1 | tee gpurun_out/microbench_count.log
python bench.py --steps 5 --warmup 3 --gbases ${1:-20} --no-cpu 2> gpurun_out/bench_iter.log | tee gpurun_out/bench_iter.json
tail -2 gpurun_out/bench_iter.log
