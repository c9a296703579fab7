#!/bin/bash
python - <<'PY'
import sys, os, subprocess
sys.path.insert(0, '.')
import torch
from ntsm_b200 import synth
import bench
wc, wl = synth.panel_windows(bench.PANEL)
g = synth.Genome(200_000_000, wc, wl, 3, 'cuda')
codes = synth.sample_reads(g, 400_000, 150, 0.01, 9).cpu()
os.makedirs('/tmp/cli', exist_ok=True)
paths = bench.write_fastq_files(codes, 4, '/tmp/cli')
ref = subprocess.run(['oracle/_ref/ntsmCount', '-t', '4', '-s', bench.PANEL] + paths, capture_output=True).stdout
def diff(a, b):
    la, lb = a.decode().splitlines(), b.decode().splitlines()
    d = [(x, y) for x, y in zip(la, lb) if x != y]
    return len(la), len(lb), len(d), d[:4]
for args in (['--gpus', '1'], ['--gpus', '2'], ['--gpus', '2', '-t', '1'], ['--gpus', '2', '--batch-bases', '4000000'], ['--gpus', '1', '--batch-bases', '4000000'], ['--gpus', '2', '--batch-bases', '4000000', '-t', '1']):
    base = ['-t', '4'] if '-t' not in args else []
    for rep in range(2):
        p = subprocess.run(['ntsm_b200/bin/ntsmCount'] + args + base + ['-s', bench.PANEL] + paths, capture_output=True)
        print(args, 'rc', p.returncode, 'same as reference:', p.stdout == ref, diff(p.stdout, ref) if p.stdout != ref else '')
        print('   ', [l for l in p.stderr.decode().splitlines() if l.startswith('Total')][:3])
PY
