#!/bin/bash
# parity of the default kernel, variant timings, the five-config CLI parity run
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest default rc=$?"; tail -3 gpurun_out/pytest_gpu.log
rm -f gpurun_out/sweep.jsonl
for V in "NTSM_KERNEL=2" "NTSM_KERNEL=3"; do
  env $V python bench.py --steps 5 --warmup 3 --gbases ${1:-20} --kernel-only 2>/dev/null | tee -a gpurun_out/sweep.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$V', round(d['value'],1), 'Gbases/s', d['check'])"
done
python tools/config_parity.py --check --scale 1 2>&1 | tail -12
