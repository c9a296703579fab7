#!/bin/bash
# compute-sanitizer over the kernels that are new in round 2 (device packer, fused peer combine, wide seeds,
# run-time-k pair kernel, merge): memcheck, then racecheck (the warp-pooled tail shares candidates through shared
# memory) and initcheck on a smaller subset -> gpurun_out/<tag>_memcheck.txt
TAG=${1:-r02}
mkdir -p gpurun_out
OUT=gpurun_out/${TAG}_memcheck.txt
: > $OUT
SEL="device_packed_reads_vs_oracle or host_register or group_finalize or (dense_hits and 2) or (other_k_vs_oracle and (17 or 21 or 31)) or add_counts_abi or insert_count_vs_oracle_k19"
echo "=== memcheck: pytest -k \"$SEL\"" >> $OUT
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --target-processes all \
  python -m pytest tests/test_gpu_parity.py tests/test_merge.py -q -m gpu -x -k "$SEL" >> $OUT 2>&1
echo "memcheck rc=$?" >> $OUT
for tool in racecheck initcheck; do
  echo "=== $tool: pytest -k \"insert_count_vs_oracle_k19 or host_register or (group_finalize and 2)\"" >> $OUT
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --target-processes all \
    python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "insert_count_vs_oracle_k19 or host_register or (group_finalize and 2)" >> $OUT 2>&1
  echo "$tool rc=$?" >> $OUT
done
grep -a "rc=\|ERROR SUMMARY\|passed\|failed" $OUT
