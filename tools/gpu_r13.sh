#!/bin/bash
# exact -m stop: gpu suite (fixtures now byte-compared for -m too), cfg4m / cfg5 through the CLI
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? ($((SECONDS-T0)) s)"; tail -15 gpurun_out/pytest_gpu.log
timeout 900 python tools/config_parity.py --check --configs ${1:-cfg4m,cfg5} --out gpurun_out/r13_config_parity.json 2> gpurun_out/r13_config_parity.log; echo "config parity rc=$? ($((SECONDS-T0)) s)"
python -c "
import json; d=json.load(open('gpurun_out/r13_config_parity.json'))
print('all_ok', d['all_ok'])
for k,v in d['configs'].items(): print(k, v['verdict'], 'ours %.2fs' % v['tool_seconds'], 'bases', v.get('bases'), 'hits', v.get('hits'), 'ref', v['reference'].get('bases'), v['reference'].get('hits'))
"
echo "total $((SECONDS-T0)) s"
