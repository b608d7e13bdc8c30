#!/bin/bash
# on the GPU box: parity tests, then the bench line (tag = $1)
tag=${1:-sX}; shift
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
python bench.py --steps 5 --warmup 3 "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${tag}_bench.json')); k=d['roofline']['kernels']
    print(round(d['value'],1),'Mpx/s', round(d['ms_per_step'],2),'ms  e2e', round(d['e2e']['value'],1), {n:k[n]['ms'] for n in k})
    print(d['stage_ms']); print(d['ordered_sums'])
except Exception as e: print('ERR', e)
PY
