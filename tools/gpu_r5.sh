#!/bin/bash
# full gpu test suite, short bench (all legs), five-config CLI parity
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? ($((SECONDS-T0)) s)"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --gbases ${1:-30} > gpurun_out/r5_bench.json 2> gpurun_out/r5_bench.log; echo "bench rc=$?"; tail -2 gpurun_out/r5_bench.log
python -c "
import json; d=json.load(open('gpurun_out/r5_bench.json'))
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'packed', round(d['e2e_packed']['value'],1), 'files', d['e2e_files'] and round(d['e2e_files']['value'],1), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'],4), d['check'], d['parity_vs_reference_on_cpu_sample'])"
timeout 900 python tools/config_parity.py --check --scale 1 2>&1 | tail -14
echo "total $((SECONDS-T0)) s"
