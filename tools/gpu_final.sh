#!/bin/bash
# round-end evidence in one gpurun call: GPU suite, smoke, default bench + reference arm, ncu launch list + full capture
TAG=${1:-r02k}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.log; echo "reference rc=$?"
bash tools/gpu_ncu.sh ${TAG} > /dev/null 2>&1
python - <<PY
import json
d = json.load(open('gpurun_out/${TAG}_bench.json'))
for k in ('value', 'e2e', 'e2e_gz', 'e2e_ascii', 'e2e_ascii_hybrid', 'e2e_ascii_host_pack_only', 'e2e_ascii_device_pack_only', 'e2e_packed', 'cpu_baseline'):
    v = d.get(k); print(k, v if not isinstance(v, dict) else (round(v['value'], 3), v.get('ms_per_step')))
print(d['roofline']['frac'], d['other_regimes_kernel_resident'])
print(d['check'], d['parity_vs_reference_on_cpu_sample'])
r = json.load(open('gpurun_out/${TAG}_bench_reference.json')); print('reference', r['value'], r['ms_per_step'])
PY
