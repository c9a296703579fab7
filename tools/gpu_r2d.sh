#!/bin/bash
# round 2, session d: tests (incl. drop-in + merge), ncu of the count kernel with/without the L2 window, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_configs.py > gpurun_out/r02d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
tail -4 gpurun_out/r02d_pytest.log
for p in 1 0; do
ncu --set full --clock-control none --import-source on -k regex:count_kernel -s 3 -c 1 -o gpurun_out/r02d_count_persist$p -f \
    python bench.py --steps 2 --warmup 3 --gbases 6 --kernel-only --opt l2_persist=$p > gpurun_out/r02d_ncu_persist$p.log 2>&1
done
timeout 900 python bench.py > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.log
echo "bench rc=$?"
grep "e2e ascii\|FASTQ files" gpurun_out/r02d_bench.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02d_bench.json'))
for k in ('value','e2e','e2e_gz','e2e_ascii','e2e_ascii_host_pack_only','e2e_ascii_device_pack_only','e2e_packed','cpu_baseline'):
    v=d[k]; print(k, v if not isinstance(v,dict) else (v['value'], v.get('ms_per_step')))
print(d['check'], d['parity_vs_reference_on_cpu_sample'])
PY
