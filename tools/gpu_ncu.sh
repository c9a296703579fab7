#!/bin/bash
# ncu evidence for the count kernel on the bench workload (3.1 Gb genome, 6 Gbases): launch list of the
# timed step + one full capture.  usage: tools/gpu_ncu.sh <tag> [extra bench.py args]
TAG=${1:-r02}; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:count_kernel|site_reduce|set_u64|pack_ascii" --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --gbases 6 --kernel-only "$@" > gpurun_out/${TAG}_ncu_bench.log 2>&1
grep -c . gpurun_out/${TAG}_launches.csv
ncu --set full --clock-control none --import-source on -k regex:count_kernel -s 3 -c 1 -o gpurun_out/${TAG}_count -f \
    python bench.py --steps 2 --warmup 3 --gbases 6 --kernel-only "$@" > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/ | grep ${TAG}
