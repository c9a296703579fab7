#!/bin/bash
# round 2, session c: full GPU test suite on the restructured library, then sweeps and a first new-style bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_configs.py > gpurun_out/r02c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
tail -5 gpurun_out/r02c_pytest.log
: > gpurun_out/r02c_sweep.jsonl
for o in "l2_persist=1" "l2_persist=0"; do
  python bench.py --steps 3 --warmup 3 --gbases 20 --kernel-only --opt $o >> gpurun_out/r02c_sweep.jsonl 2>> gpurun_out/r02c_sweep.log
  python bench.py --steps 3 --warmup 3 --gbases 10 --kernel-only --synthetic-sites 1000000 --opt $o >> gpurun_out/r02c_sweep.jsonl 2>> gpurun_out/r02c_sweep.log
done
for k in 17 21 25 31; do
  python bench.py --steps 3 --warmup 3 --gbases 20 --kernel-only --k $k >> gpurun_out/r02c_sweep.jsonl 2>> gpurun_out/r02c_sweep.log
done
cut -c1-330 gpurun_out/r02c_sweep.jsonl
timeout 900 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.log
echo "bench rc=$?"
tail -3 gpurun_out/r02c_bench.log
cat gpurun_out/r02c_bench.json
