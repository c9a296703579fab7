#!/bin/bash
# the multi-sample matrix path on the GPU box: its parity tests, then memcheck + racecheck over a subset
# -> gpurun_out/<tag>_multi_pytest.log, <tag>_multi_sanitizer.txt        usage: bash tools/gpu_multi_path.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi.py -q -m gpu > gpurun_out/${TAG}_multi_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_multi_pytest.log
tail -25 gpurun_out/${TAG}_multi_pytest.log
OUT=gpurun_out/${TAG}_multi_sanitizer.txt
: > $OUT
SEL="(abi_matches_reference and (overlap_dupes or missing_precision or window_past or panel300 or repeated_site)) or (fuzzed and 2-25)"
for tool in memcheck racecheck; do
  echo "=== $tool: pytest tests/test_multi.py -k \"$SEL\"" >> $OUT
  timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 --target-processes all \
    python -m pytest tests/test_multi.py -q -m gpu -x -k "$SEL" >> $OUT 2>&1
  echo "$tool rc=$?" >> $OUT
done
grep -a "rc=\|ERROR SUMMARY\|passed\|failed" $OUT
