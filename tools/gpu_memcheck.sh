#!/bin/bash
# compute-sanitizer memcheck over the kernels that are new in round 2 (device packer, fused peer combine, wide seeds,
# run-time-k pair kernel, merge) -> gpurun_out/<tag>_memcheck.txt
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --target-processes all \
  python -m pytest tests/test_gpu_parity.py tests/test_merge.py -q -m gpu -x \
  -k "device_packed_reads_vs_oracle or host_register or group_finalize or (dense_hits and 2) or (other_k_vs_oracle and (17 or 21 or 31)) or add_counts_abi or insert_count_vs_oracle_k19" \
  > gpurun_out/${TAG}_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/${TAG}_memcheck.txt
grep -c "Invalid\|ERROR SUMMARY" gpurun_out/${TAG}_memcheck.txt
tail -6 gpurun_out/${TAG}_memcheck.txt
