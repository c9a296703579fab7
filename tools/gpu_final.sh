#!/bin/bash
# round-end evidence in one call: gpu suite, the default bench line + reference arm, ncu launch list of the bench's timed job, one full capture
TAG=${1:-r01v15}
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? ($((SECONDS-T0)) s)"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log; echo "bench rc=$? ($((SECONDS-T0)) s)"; tail -4 gpurun_out/${TAG}_bench.log
cat gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.log; echo "ref rc=$? ($((SECONDS-T0)) s)"
cat gpurun_out/${TAG}_ref.json
# launch list: the same job the bench times (100 Gbases resident, zero counts -> count kernel -> site reduce), kernels only
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:count_kernel|site_reduce|set_u64" --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --kernel-only > gpurun_out/${TAG}_ncu_bench.log 2>&1
grep -c . gpurun_out/${TAG}_launches.csv; echo "launch list done ($((SECONDS-T0)) s)"
ncu --set full --clock-control none --import-source on -k regex:count_kernel -s 3 -c 1 -o gpurun_out/${TAG}_count -f \
    python bench.py --steps 2 --warmup 3 --gbases 6 --kernel-only > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/ | grep ${TAG}
echo "total $((SECONDS-T0)) s"
