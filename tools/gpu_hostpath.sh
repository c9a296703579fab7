#!/bin/bash
# host-only experiment on the GPU box: does the mapped plain-file source scale across PROCESSES where it does not across threads?
mkdir -p gpurun_out
g++ -O3 -std=c++17 -pthread -I ntsm_b200/csrc tools/hostpath_bench.cpp ntsm_b200/csrc/{fastx,gzsource,inflate,pargz,pack}.cpp -lz -o /tmp/hostpath_bench || exit 1
python - <<'PY'
import numpy as np
n = 3_000_000
rng = np.random.default_rng(1)
codes = rng.integers(0, 4, (n, 150), dtype=np.uint8)
rec = np.empty((n, 315), np.uint8)
rec[:, 0] = ord('@'); rec[:, 1] = ord('r')
idx = np.arange(n)
for d in range(8): rec[:, 2 + d] = (idx // 10 ** (7 - d)) % 10 + 48
rec[:, 10] = 10; rec[:, 11:161] = np.frombuffer(b'ACGT', np.uint8)[codes]; rec[:, 161] = 10; rec[:, 162] = ord('+'); rec[:, 163] = 10
rec[:, 164:314] = ord('I'); rec[:, 314] = 10
rec.tofile('/dev/shm/hp_0.fq')
PY
for i in $(seq 1 15); do cp /dev/shm/hp_0.fq /dev/shm/hp_$i.fq; done
files=$(for i in $(seq 0 15); do echo /dev/shm/hp_$i.fq; done)
{
for rep in 1 2 3; do
  echo "16 threads, gzread:  $(NTSM_INFLATE=zlib /tmp/hostpath_bench pack $files)"
  echo "16 threads, mapped:  $(/tmp/hostpath_bench pack $files)"
  echo "16 processes, mapped: $(/tmp/hostpath_bench procs $files)   (7.2 Gbases)"
  echo "16 processes, gzread: $(NTSM_INFLATE=zlib /tmp/hostpath_bench procs $files)   (7.2 Gbases)"
done
} > gpurun_out/r02r_hostpath_procs.txt 2>&1
cat gpurun_out/r02r_hostpath_procs.txt
rm -f /dev/shm/hp_*.fq
