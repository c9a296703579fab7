#!/bin/bash
# folded pair table: parity (full gpu suite), fold x level-2-size sweep on the bench workload, ncu of the default
mkdir -p gpurun_out
rm -f gpurun_out/sweep_fold.jsonl
T0=$SECONDS
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? ($((SECONDS-T0)) s)"; tail -3 gpurun_out/pytest_gpu.log
for V in "NTSM_PAIR_FOLD=0" "NTSM_PAIR_FOLD=1" "NTSM_PAIR_FOLD=2" "NTSM_PAIR_FOLD=3" "NTSM_PAIR_FOLD=1 NTSM_FILTER_BITS=26" "NTSM_PAIR_FOLD=1 NTSM_FILTER_BITS=25" "NTSM_PAIR_FOLD=2 NTSM_FILTER_BITS=26" "NTSM_PAIR_FOLD=1 NTSM_SEED_CFG=0" "NTSM_PAIR_FOLD=1 NTSM_SEED_CFG=2"; do
  env $V timeout 300 python bench.py --steps 5 --warmup 3 --gbases ${1:-20} --kernel-only 2>/dev/null | tee -a gpurun_out/sweep_fold.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$V', round(d['value'],1), 'Gbases/s', d['check'])"
done
echo "sweep done ($((SECONDS-T0)) s)"
bash tools/gpu_ncu.sh ${2:-r01v9}
echo "total $((SECONDS-T0)) s"
