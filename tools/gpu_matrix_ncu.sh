#!/bin/bash
# ncu evidence for the matrix path's kernels on the bench_matrix workload: launch list of one whole job + one full capture of
# the fill kernel (a batch in the middle of the job) and of the two norm-matrix kernels.   usage: tools/gpu_matrix_ncu.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:multi_|vcf_|occ_" --csv --log-file gpurun_out/${TAG}_matrix_launches.csv \
    python tools/bench_matrix.py --profile > gpurun_out/${TAG}_matrix_ncu_run.log 2>&1
grep -c . gpurun_out/${TAG}_matrix_launches.csv
ncu --set full --clock-control none --import-source on -k "regex:multi_fill_kernel|multi_norm" -s 45 -c 4 -o gpurun_out/${TAG}_matrix -f \
    python tools/bench_matrix.py --profile > gpurun_out/${TAG}_matrix_ncu_full.log 2>&1
ls -la gpurun_out/ | grep ${TAG}_matrix
