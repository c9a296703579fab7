#!/bin/bash
# multi-GPU check: NCCL parity test + torchrun bench at N ranks.  usage: tools/gpu_multi.sh N [gbases]
N=${1:-2}; GB=${2:-20}
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_multi_rank.py -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --gbases $GB --no-cpu > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.log; echo "bench N=$N rc=$?"
tail -4 gpurun_out/bench_n$N.log; cat gpurun_out/bench_n$N.json
# the CLI across GPUs in one process: same counts file for 1 and N GPUs
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
outs = []
for gp in ('1', str(torch.cuda.device_count())):
    p = subprocess.run(['ntsm_b200/bin/ntsmCount', '--gpus', gp, '--batch-bases', '4000000', '-t', '4', '-s', bench.PANEL] + paths, capture_output=True)
    print('gpus', gp, 'rc', p.returncode, p.stderr.decode().strip().splitlines()[-1])
    outs.append(p.stdout)
print('CLI 1-GPU vs N-GPU counts files identical:', outs[0] == outs[1], len(outs[0]))
PY
