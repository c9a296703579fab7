#!/bin/bash
# where does cfg4 (8 x .fq.gz, -t 8) spend its time: own inflate (mmap / read / no madvise) vs zlib, alternating on one box; BGZF helper scaling
mkdir -p gpurun_out
T0=$SECONDS
python tools/config_parity.py --check --configs cfg4 --tmp /tmp/cp --out gpurun_out/r11_cfg4.json 2>&1 | tail -2
F=$(ls /tmp/cp/cfg4/*.fq.gz | tr '\n' ' ')
P=/tmp/cp/human_sites_n10.fa
nproc; grep -m1 "model name" /proc/cpuinfo; df -h /tmp | tail -1
run() { # label env...
  local label=$1; shift
  for i in 1 2 3; do
    env NTSM_TIMING=1 "$@" ./ntsm_b200/bin/ntsmCount -t ${T:-8} -s $P $F 2> gpurun_out/r11_err.txt | sha256sum | cut -c1-12 | tr '\n' ' '
    echo "$label t=${T:-8} $(grep -E 'Time:' gpurun_out/r11_err.txt | head -1) | $(grep -iE 'phase|timing' gpurun_out/r11_err.txt | tr '\n' ';' | cut -c1-300)"
  done
}
run fast-mmap NTSM_X=1
run zlib NTSM_INFLATE=zlib
run fast-read NTSM_GZ_INPUT=read
run fast-mmap-noadvise NTSM_GZ_INPUT=mmap_plain
run fast-mmap NTSM_X=1
run zlib NTSM_INFLATE=zlib
T=16 run fast-mmap NTSM_X=1
echo "cfg4 runs done ($((SECONDS-T0)) s)"
# BGZF: two lanes re-blocked with python, -t 2 (no helpers) vs -t 16 (7 helpers per file)
python - <<'PY'
import gzip, struct, zlib, sys
for lane in (0, 1):
    data = gzip.open('/tmp/cp/cfg4/lane%d_R1.fq.gz' % lane).read()
    out = bytearray()
    for i in list(range(0, len(data), 0xff00)) + [len(data)]:
        c = data[i:i + 0xff00]
        co = zlib.compressobj(6, zlib.DEFLATED, -15); body = co.compress(c) + co.flush()
        out += b"\x1f\x8b\x08\x04\0\0\0\0\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, 25 + len(body)) + body + struct.pack("<II", zlib.crc32(c), len(c))
    open('/tmp/cp/bgzf_lane%d.fq.gz' % lane, 'wb').write(out)
PY
F="/tmp/cp/bgzf_lane0.fq.gz /tmp/cp/bgzf_lane1.fq.gz"
T=2 run bgzf-t2 NTSM_X=1
T=16 run bgzf-t16 NTSM_X=1
T=2 run bgzf-zlib-t2 NTSM_INFLATE=zlib
echo "total $((SECONDS-T0)) s"
