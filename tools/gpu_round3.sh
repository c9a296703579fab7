#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/sweep.jsonl
for V in "NTSM_KERNEL=3"; do
  env $V python -m pytest tests -m gpu -x -q -k "oracle or properties or fixture" > gpurun_out/pytest_gpu_$V.log 2>&1; echo "pytest $V rc=$?"; tail -1 gpurun_out/pytest_gpu_$V.log
done
for V in "NTSM_KERNEL=2" "NTSM_KERNEL=3" "NTSM_TAIL_POOL=0" "NTSM_GATE_THREADS=768" "NTSM_GATE_M=13"; do
  env $V python bench.py --steps 5 --warmup 3 --gbases ${1:-20} --kernel-only 2>/dev/null | tee -a gpurun_out/sweep.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$V', round(d['value'],1), 'Gbases/s', d['check'])"
done
