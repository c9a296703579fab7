#!/bin/bash
# L2 fetch granularity experiment: cfg5 (DRAM-bound random probes) and cfg2 (L2-resident tables)
mkdir -p gpurun_out
: > gpurun_out/r02b_l2fetch.jsonl
for g in 0 32 64 128; do
  if [ $g = 0 ]; then unset NTSM_L2_FETCH; else export NTSM_L2_FETCH=$g; fi
  python bench.py --steps 3 --warmup 3 --gbases 10 --kernel-only --synthetic-sites 1000000 >> gpurun_out/r02b_l2fetch.jsonl 2>> gpurun_out/r02b.log
done
for g in 0 32 128; do
  if [ $g = 0 ]; then unset NTSM_L2_FETCH; else export NTSM_L2_FETCH=$g; fi
  python bench.py --steps 3 --warmup 3 --gbases 20 --kernel-only >> gpurun_out/r02b_l2fetch.jsonl 2>> gpurun_out/r02b.log
done
grep "L2 fetch" gpurun_out/r02b.log | sort | uniq -c
cut -c1-200 gpurun_out/r02b_l2fetch.jsonl
