#!/bin/bash
# mapped plain-file source at 16 parser threads: which way of setting up the mapping scales on the GPU box's VM
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
cat /sys/kernel/mm/transparent_hugepage/shmem_enabled /sys/kernel/mm/transparent_hugepage/enabled 2>&1
files=$(for i in $(seq 0 15); do echo /dev/shm/hp_$i.fq; done)
{
for rep in 1 2 3; do
  echo "gzread:            $(NTSM_INFLATE=zlib /tmp/hostpath_bench pack $files)"
  echo "mapped default:    $(/tmp/hostpath_bench pack $files)"
  echo "mapped shared:     $(NTSM_MAP_VARIANT=shared /tmp/hostpath_bench pack $files)"
  echo "mapped shared nopop: $(NTSM_MAP_VARIANT=shared NTSM_MAP_POPULATE=0 /tmp/hostpath_bench pack $files)"
  echo "mapped MAP_POPULATE: $(NTSM_MAP_VARIANT=populate NTSM_MAP_POPULATE=0 /tmp/hostpath_bench pack $files)"
  echo "mapped noseq:      $(NTSM_MAP_VARIANT=noseq /tmp/hostpath_bench pack $files)"
  echo "mapped noseq nopop: $(NTSM_MAP_VARIANT=noseq NTSM_MAP_POPULATE=0 /tmp/hostpath_bench pack $files)"
  echo "mapped stretch 256: $(NTSM_MAP_STRETCH=256 /tmp/hostpath_bench pack $files)"
  echo "mapped stretch 2:  $(NTSM_MAP_STRETCH=2 /tmp/hostpath_bench pack $files)"
  echo "parse only mapped: $(/tmp/hostpath_bench parse $files)"
  echo "parse only gzread: $(NTSM_INFLATE=zlib /tmp/hostpath_bench parse $files)"
done
} > gpurun_out/r02f_hostpath.txt 2>&1
cat gpurun_out/r02f_hostpath.txt | sed 's/pack: 16 file(s).thread(s), 48000000 reads, 7.200 Gbases in//'
rm -f /dev/shm/hp_*.fq
