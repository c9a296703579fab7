#!/bin/bash
# paired-seed kernel: parity (full gpu suite, default variant 5), shape sweep vs the strided-seed kernel, ncu
mkdir -p gpurun_out
rm -f gpurun_out/sweep5.jsonl
T0=$SECONDS
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? ($((SECONDS-T0)) s)"; tail -3 gpurun_out/pytest_gpu.log
for V in "NTSM_KERNEL=4 NTSM_SEED_CFG=1" "NTSM_KERNEL=5 NTSM_SEED_CFG=0" "NTSM_KERNEL=5 NTSM_SEED_CFG=1" "NTSM_KERNEL=5 NTSM_SEED_CFG=2" "NTSM_KERNEL=5 NTSM_SEED_CFG=3"; do
  env $V timeout 300 python bench.py --steps 5 --warmup 3 --gbases ${1:-20} --kernel-only 2>/dev/null | tee -a gpurun_out/sweep5.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$V', round(d['value'],1), 'Gbases/s', d['check'])"
done
echo "sweep done ($((SECONDS-T0)) s)"
bash tools/gpu_ncu.sh ${2:-r01v8}
echo "total $((SECONDS-T0)) s"
