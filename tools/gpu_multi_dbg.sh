#!/bin/bash
# CLI across the GPUs of one process: 1-GPU vs N-GPU counts files, with a diff summary and phase timing
mkdir -p gpurun_out
python - <<'PY'
import sys, os, subprocess
sys.path.insert(0, '.')
import torch
from ntsm_b200 import synth
import bench
wc, wl = synth.panel_windows(bench.PANEL)
g = synth.Genome(200_000_000, wc, wl, 3, 'cuda')
codes = synth.sample_reads(g, 2_000_000, 150, 0.01, 9).cpu()
os.makedirs('/tmp/cli', exist_ok=True)
paths = bench.write_fastq_files(codes, 4, '/tmp/cli')
outs = {}
env = dict(os.environ, NTSM_TIMING='1')
for gp in ('1', str(torch.cuda.device_count()), '1', str(torch.cuda.device_count())):
    p = subprocess.run(['ntsm_b200/bin/ntsmCount', '--gpus', gp, '--batch-bases', '4000000', '-t', '4', '-s', bench.PANEL] + paths, capture_output=True, env=env)
    err = p.stderr.decode()
    print('gpus', gp, 'rc', p.returncode, 'stdout bytes', len(p.stdout))
    print('   ' + '\n   '.join(l for l in err.strip().splitlines() if 'timing' in l or 'Total' in l or 'Time' in l or 'NCCL' in l))
    outs.setdefault(gp, []).append(p.stdout)
a, b = outs['1'][0], outs[str(torch.cuda.device_count())][0]
print('1-GPU runs identical:', outs['1'][0] == outs['1'][1], ' N-GPU runs identical:', b == outs[str(torch.cuda.device_count())][1])
print('1 vs N identical:', a == b)
if a != b:
    la, lb = a.decode(errors='replace').splitlines(), b.decode(errors='replace').splitlines()
    print('lines', len(la), len(lb), 'head A', la[:3], 'head B', lb[:3])
    nd = 0
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            nd += 1
            if nd <= 8: print('  line', i, repr(x), '|', repr(y))
    print('differing lines', nd)
PY
