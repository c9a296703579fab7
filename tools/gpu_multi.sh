#!/bin/bash
# multi-GPU session: NCCL parity tests, torchrun bench at N ranks (weak + strong scaling, N-GPU = 1-GPU identity),
# the reference arm, and the CLI's NCCL-free --gpus N against --gpus 1.   usage: tools/gpu_multi.sh N [tag]
N=${1:-2}; TAG=${2:-r02}
mkdir -p gpurun_out
nvidia-smi -L | head -8
python -m pytest tests/test_multi_rank.py tests/test_gpu_parity.py -m gpu -x -q -k "nccl or all_gpus or group_finalize" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 \
    > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.log; echo "bench N=$N rc=$?"
grep "e2e ascii\|FASTQ files" gpurun_out/${TAG}_bench_n$N.log | head -4
python - <<PY
import json
d = json.load(open('gpurun_out/${TAG}_bench_n$N.json'))
for k in ('value', 'e2e', 'e2e_ascii', 'e2e_ascii_host_pack_only', 'e2e_ascii_device_pack_only', 'e2e_packed', 'strong_scaling', 'cpu_baseline'):
    v = d.get(k); print(k, v if not isinstance(v, dict) else (round(v['value'], 3), v.get('ms_per_step')))
print(d['check'])
PY
# the CLI across the GPUs of one process (no communicator): same counts file for 1 and N GPUs, and the wall time of each
python - <<'PY'
import sys, os, subprocess, time, shutil
sys.path.insert(0, '.')
import torch
import bench
d = '/dev/shm/ntsm_cli'
shutil.rmtree(d, ignore_errors=True)
paths = bench.make_fastq_files('cuda', 16_000_000, 3, 300, 8, d)
outs = []
for gp in ('1', str(torch.cuda.device_count())):
    t = time.time()
    p = subprocess.run(['ntsm_b200/bin/ntsmCount', '--gpus', gp, '-t', '16', '-s', bench.PANEL] + paths, capture_output=True, env=dict(os.environ, NTSM_TIMING='1'))
    print('gpus', gp, 'rc', p.returncode, 'wall %.2f s' % (time.time() - t), [l for l in p.stderr.decode().splitlines() if 'timing' in l or 'Time:' in l])
    outs.append(p.stdout)
print('CLI 1-GPU vs N-GPU counts files identical:', outs[0] == outs[1], len(outs[0]))
shutil.rmtree(d, ignore_errors=True)
PY
