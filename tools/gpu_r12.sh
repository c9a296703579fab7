#!/bin/bash
# cfg5 regime (10^6 synthetic sites, ~26 M k-mers, tables past L2): kernel-resident throughput of every k=19 variant
mkdir -p gpurun_out
rm -f gpurun_out/sweep_cfg5.jsonl
T0=$SECONDS
for V in "NTSM_KERNEL=5" "NTSM_KERNEL=5 NTSM_PAIR_FOLD=0" "NTSM_KERNEL=4" "NTSM_KERNEL=3" "NTSM_KERNEL=1" "NTSM_KERNEL=0"; do
  env $V timeout 400 python bench.py --steps 3 --warmup 3 --gbases ${1:-10} --kernel-only --synthetic-sites 1000000 2> gpurun_out/cfg5_last.log | tee -a gpurun_out/sweep_cfg5.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$V', d['kernel'], round(d['value'],1), 'Gbases/s', d['n_kmers'], 'k-mers filter 2^%d' % d['filter_bits'], d['check'])" || tail -3 gpurun_out/cfg5_last.log
  echo "  ($((SECONDS-T0)) s)"
done
echo "total $((SECONDS-T0)) s"
