#!/bin/bash
# full default bench line (our arm), then the reference arm, as the driver runs them
mkdir -p gpurun_out
T0=$SECONDS
timeout 1200 python bench.py > gpurun_out/${1:-r8}_bench.json 2> gpurun_out/${1:-r8}_bench.log; echo "bench rc=$? ($((SECONDS-T0)) s)"; tail -12 gpurun_out/${1:-r8}_bench.log
cat gpurun_out/${1:-r8}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${1:-r8}_ref.json 2> gpurun_out/${1:-r8}_ref.log; echo "ref rc=$? ($((SECONDS-T0)) s)"
cat gpurun_out/${1:-r8}_ref.json
echo "total $((SECONDS-T0)) s"
