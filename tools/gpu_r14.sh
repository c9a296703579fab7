#!/bin/bash
# (1) compute-sanitizer memcheck over the kernel-variant, fold, dense-seed and -m tests; (2) cfg3-like kernel-resident number
mkdir -p gpurun_out
T0=$SECONDS
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variant or fold or dense_hits or m_cap_exact or stops_after or insert_count_vs_oracle_k19 or other_k" > gpurun_out/r14_memcheck.log 2>&1; echo "memcheck rc=$? ($((SECONDS-T0)) s)"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r14_memcheck.log | tail -8
python bench.py --kernel-only --steps 5 --warmup 3 --gbases 20 --read-len 20000 --err 0.075 2>/dev/null | tee gpurun_out/r14_cfg3_kernel.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg3-like 20 kb reads, 7.5 % substitutions:', d['kernel'], round(d['value'],1), 'Gbases/s', d['check'])"
python bench.py --kernel-only --steps 5 --warmup 3 --gbases 20 --genome-mb 30 2>/dev/null | tee gpurun_out/r14_cfg1_kernel.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg1-like 30 Mb genome (dense hits):', d['kernel'], round(d['value'],1), 'Gbases/s', d['check'])"
echo "total $((SECONDS-T0)) s"
