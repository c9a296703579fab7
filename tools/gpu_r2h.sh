#!/bin/bash
# wide paired-seed kernel (large panels): parity tests, cfg5 through the CLI, cfg5 kernel-resident rate + ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_configs.py -q -m gpu -x -k "variant or dense or large_panel or cfg5" > gpurun_out/r02i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log
tail -4 gpurun_out/r02i_pytest.log
: > gpurun_out/r02i_cfg5.jsonl
for kn in 2 1; do
  python bench.py --steps 3 --warmup 3 --gbases 10 --kernel-only --synthetic-sites 1000000 --opt kernel=$kn >> gpurun_out/r02i_cfg5.jsonl 2>> gpurun_out/r02h.log
done
python bench.py --steps 3 --warmup 3 --gbases 20 --kernel-only --opt kernel=2 >> gpurun_out/r02i_cfg5.jsonl 2>> gpurun_out/r02h.log
cut -c1-330 gpurun_out/r02i_cfg5.jsonl
ncu --set full --clock-control none --import-source on -k regex:count_kernel -s 3 -c 1 -o gpurun_out/r02i_cfg5_wide -f \
    python bench.py --steps 2 --warmup 3 --gbases 4 --kernel-only --synthetic-sites 1000000 > gpurun_out/r02i_ncu.log 2>&1
ls -la gpurun_out | grep r02h
