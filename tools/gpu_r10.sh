#!/bin/bash
# own inflate + BGZF-parallel byte source: gpu suite (incl. the damaged-gzip fixtures through the CLI), five-config parity, cfg4 with zlib for comparison
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? ($((SECONDS-T0)) s)"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python tools/config_parity.py --check --out gpurun_out/${1:-r10}_config_parity.json 2> gpurun_out/${1:-r10}_config_parity.log; echo "config parity rc=$? ($((SECONDS-T0)) s)"
python -c "
import json; d=json.load(open('gpurun_out/${1:-r10}_config_parity.json'))
print('all_ok', d['all_ok'])
for k,v in d['configs'].items(): print(k, v['verdict'], 'ours %.2fs' % v['tool_seconds'], 'ref %.2fs' % v['reference']['tool_seconds'], 'Gbases/s', round(v.get('gbases_per_s',0),3))
"
NTSM_INFLATE=zlib timeout 600 python tools/config_parity.py --check --configs cfg4 --out gpurun_out/${1:-r10}_config_parity_zlib.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/${1:-r10}_config_parity_zlib.json'))
for k,v in d['configs'].items(): print('zlib-inflate', k, v['verdict'], 'ours %.2fs' % v['tool_seconds'])
"
echo "total $((SECONDS-T0)) s"
