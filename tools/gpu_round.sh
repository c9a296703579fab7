#!/bin/bash
# one GPU-box visit: host facts, parity of every kernel variant, variant sweep, full bench
mkdir -p gpurun_out
{ nproc; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA|Flags" | cut -c1-300; free -g; df -h /tmp /dev/shm | cat; nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current --format=csv; } > gpurun_out/host.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest default rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for V in "NTSM_KERNEL=2" "NTSM_GATE_M=13" "NTSM_GATE_M=12" "NTSM_KERNEL=0"; do
  env $V python -m pytest tests -m gpu -x -q -k "oracle or properties or fixture" > gpurun_out/pytest_gpu_$V.log 2>&1; echo "pytest $V rc=$?"; tail -1 gpurun_out/pytest_gpu_$V.log
done
for V in "NTSM_KERNEL=2" "NTSM_GATE_M=14" "NTSM_GATE_M=13" "NTSM_GATE_M=12"; do
  env $V python bench.py --steps 5 --warmup 3 --gbases ${1:-20} --kernel-only 2>/dev/null | tee -a gpurun_out/sweep.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$V', round(d['value'],1), 'Gbases/s', d['check'])"
done
python bench.py 2> gpurun_out/bench_err.log | tee gpurun_out/bench_full.json | cut -c1-3000
tail -8 gpurun_out/bench_err.log
