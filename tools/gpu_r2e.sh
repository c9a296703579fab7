#!/bin/bash
# host-side experiment on the GPU box: plain FASTQ through gzread vs the mapped source (with / without
# MADV_POPULATE_READ), 1..16 parser threads; and the k-mer bitmap size sweep for the count kernel
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
{
for nt in 1 4 8 16; do
  files=$(for i in $(seq 0 $((nt-1))); do echo /dev/shm/hp_$i.fq; done)
  for rep in 1 2; do
  echo "threads $nt gzread:        $(NTSM_INFLATE=zlib /tmp/hostpath_bench pack $files)"
  echo "threads $nt mapped:        $(/tmp/hostpath_bench pack $files)"
  echo "threads $nt mapped nopop:  $(NTSM_MAP_POPULATE=0 /tmp/hostpath_bench pack $files)"
  done
done
} > gpurun_out/r02e_hostpath.txt 2>&1
cat gpurun_out/r02e_hostpath.txt
rm -f /dev/shm/hp_*.fq
: > gpurun_out/r02e_filterbits.jsonl
for fb in 24 25 26 27 28; do
  python bench.py --steps 3 --warmup 3 --gbases 20 --kernel-only --opt filter_bits=$fb --opt l2_persist=0 >> gpurun_out/r02e_filterbits.jsonl 2>> gpurun_out/r02e.log
done
cut -c1-120 gpurun_out/r02e_filterbits.jsonl
