#!/bin/bash
# seed kernel: parity (all gpu tests with NTSM_KERNEL=4), shape sweep vs gate2, DSMEM microbench
mkdir -p gpurun_out
rm -f gpurun_out/sweep4.jsonl
T0=$SECONDS
NTSM_KERNEL=4 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_k4.log 2>&1; echo "pytest k4 rc=$? ($((SECONDS-T0)) s)"; tail -3 gpurun_out/pytest_gpu_k4.log
for V in "NTSM_KERNEL=3" "NTSM_KERNEL=4 NTSM_SEED_CFG=0" "NTSM_KERNEL=4 NTSM_SEED_CFG=1" "NTSM_KERNEL=4 NTSM_SEED_CFG=2" "NTSM_KERNEL=4 NTSM_SEED_CFG=3"; do
  env $V timeout 300 python bench.py --steps 5 --warmup 3 --gbases ${1:-20} --kernel-only 2>/dev/null | tee -a gpurun_out/sweep4.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$V', round(d['value'],1), 'Gbases/s', d['check'])"
done
timeout 120 ./ntsm_b200/bin/microbench 8 > gpurun_out/microbench_dsmem.txt 2>&1; cat gpurun_out/microbench_dsmem.txt
echo "total $((SECONDS-T0)) s"
