#!/bin/bash
# first contact with the GPU: smoke, GPU parity tests, nvidia-smi facts
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu_info.txt 2>&1
nproc >> gpurun_out/gpu_info.txt; free -g >> gpurun_out/gpu_info.txt; lscpu | head -20 >> gpurun_out/gpu_info.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
ntsm_b200/bin/microbench 7 > gpurun_out/microbench.log 2>&1; tail -60 gpurun_out/microbench.log
