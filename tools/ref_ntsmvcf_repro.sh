#!/bin/bash
# Why SURVEY 8(f) rank 4 (ntsmVCF / MultiCount::printNormMatrix) has nothing to pin against: the UNMODIFIED reference
# tool, compiled from /root/reference like oracle/_ref/ntsmCount, segfaults on a minimal well-formed input.
# VCFConvert's constructor builds its MultiCount before any sample ID is known (src/VCFConvert.hpp:42), so
# m_matCounts is sized for ZERO samples (src/MultiCount.hpp:266) and the first insertCount writes out of bounds.
# Build container only (needs /root/reference).   usage: bash tools/ref_ntsmvcf_repro.sh
set -e
REF=${REF:-/root/reference}
D=$(mktemp -d)
cd $D
printf '#define PACKAGE_NAME "ntsm"\n#define GIT_REVISION "663f9a5"\n' > config.h
g++ -O1 -std=c++11 -fopenmp -I. -I$REF -I$REF/src -I$REF/vendor $REF/src/ntSeqMatchVCF.cpp $REF/src/Options.cpp -o ntsmVCF -lz -pthread
python3 - <<'PY'
import random
random.seed(1)
g = [random.choice("ACGT") for _ in range(2000)]
sites = [(500, "rs1"), (1200, "rs2")]
for pos, _ in sites:
    g[pos - 1] = "A"
g = "".join(g)
open("ref.fa", "w").write(">chr1\n" + g + "\n")
with open("sites.fa", "w") as fh:
    for pos, name in sites:
        w = g[pos - 16: pos + 15]
        v = w[:15] + "C" + w[16:]
        fh.write(">%s ref\n%s\n>%s var\n%s\n" % (name, "N".join(w[j:j + 19] for j in range(13)), name, "N".join(v[j:j + 19] for j in range(13))))
with open("s.vcf", "w") as fh:
    fh.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\tS2\tS3\n")
    fh.write("chr1\t500\trs1\tA\tC\t.\tPASS\t.\tGT\t0|0\t0|1\t1|1\n")
    fh.write("chr1\t1200\trs2\tA\tC\t.\tPASS\t.\tGT\t0|1\t1|1\t0|0\n")
PY
set +e
./ntsmVCF -s sites.fa -r ref.fa -p out s.vcf
echo "reference ntsmVCF exit code: $?   (139 = SIGSEGV)"
