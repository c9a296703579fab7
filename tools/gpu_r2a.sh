#!/bin/bash
# round 2, first GPU session: config parity tests under pytest, cfg5 kernel ncu capture, and a live default bench
mkdir -p gpurun_out
python -m pytest tests/test_configs.py -q -m gpu -x > gpurun_out/r02a_pytest_configs.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest_configs.log
ncu --set full --clock-control none --import-source on -k regex:count_kernel -s 3 -c 1 -o gpurun_out/r02a_cfg5_count -f \
    python bench.py --steps 2 --warmup 3 --gbases 4 --kernel-only --synthetic-sites 1000000 > gpurun_out/r02a_cfg5_ncu.log 2>&1
python bench.py --steps 3 --warmup 3 --gbases 10 --kernel-only --synthetic-sites 1000000 > gpurun_out/r02a_cfg5_bench.json 2> gpurun_out/r02a_cfg5_bench.log
tail -3 gpurun_out/r02a_pytest_configs.log
cat gpurun_out/r02a_cfg5_bench.json
