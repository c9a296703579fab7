#!/bin/bash
# compare kernel variants on the real bench workload + ncu of the default
mkdir -p gpurun_out
for V in ${VARIANTS:-1 2}; do
  NTSM_KERNEL=$V python bench.py --steps 5 --warmup 3 --gbases ${1:-20} --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('variant $V value %.1f Gbases/s kernel_ms %.3f frac %.4f e2e %.1f TK %d hits %d' % (d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['check']['TK'], d['check']['hits']))"
done
bash tools/gpu_ncu.sh ${2:-cmp} > /dev/null 2>&1
