#!/bin/bash
# bench + profiles on one B200.  usage: tools/gpu_bench.sh <tag> [gbases]
TAG=${1:-r01}; GB=${2:-100}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${TAG}_clocks.csv &
SMI=$!
python bench.py --steps 5 --warmup 3 --gbases $GB > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log; echo "bench rc=$?"
kill $SMI
tail -3 gpurun_out/${TAG}_bench.log; cat gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.log; cat gpurun_out/${TAG}_bench_ref.json
# launch list of the same command at a small size (ncu serialises and replays; shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --gbases 4 --genome-mb 100 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
grep -c . gpurun_out/${TAG}_launches.csv
# full capture of the count kernel
ncu --set full --clock-control none --import-source on -k regex:count_kernel -s 3 -c 2 -o gpurun_out/${TAG}_count -f \
    python bench.py --steps 2 --warmup 3 --gbases 4 --genome-mb 100 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/
