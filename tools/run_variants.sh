#!/bin/bash
# on the GPU box: compile-time variants of the bench line
for v in "-DPB_OS_MINB=6" "-DPB_OS_MINB=10" "-DPB_OS_WARPS=4 -DPB_OS_MINB=4" "-DPB_OS_WARPS=1 -DPB_OS_MINB=16"; do
  PB200_NVCC_EXTRA="$v" python -m patolette_b200.build --force > /dev/null 2>&1
  python bench.py --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/var.json 2>gpurun_out/var.err
  python - "$v" <<'PY'
import json,sys
d=json.load(open('gpurun_out/var.json')); k=d['roofline']['kernels']
print(sys.argv[1], round(d['ms_per_step'],2), {n:k[n]['ms'] for n in k if 'summary' in n})
PY
done
