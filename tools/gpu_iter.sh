#!/bin/bash
# quick iteration: parity tests for both kernel variants + microbench(count) + short bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest(default variant) rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for V in 0 1; do NTSM_KERNEL=$V python -m pytest tests -m gpu -x -q -k "oracle or properties" > gpurun_out/pytest_gpu_v$V.log 2>&1; echo "pytest(variant $V) rc=$?"; tail -2 gpurun_out/pytest_gpu_v$V.log; done
ntsm_b200/bin/microbench 4 2>&1 | tee gpurun_out/microbench_count.log
python bench.py --steps 5 --warmup 3 --gbases ${1:-20} --no-cpu 2> gpurun_out/bench_iter.log | tee gpurun_out/bench_iter.json
tail -2 gpurun_out/bench_iter.log
