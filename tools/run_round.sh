#!/bin/bash
# on the GPU box: parity tests, bench (both arms), full-size configs, ncu launch list
tag=${1:-r01}
[ -n "$SKIP_TESTS" ] || { python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log; }
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -2 gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
python tools/time_configs.py --profile FULL > gpurun_out/${tag}_full_configs.jsonl 2> gpurun_out/${tag}_full_configs.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 900 --csv --log-file gpurun_out/${tag}_ncu_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1
gzip -9 -f gpurun_out/${tag}_ncu_launches.csv
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json')); k=d['roofline']['kernels']
print(round(d['value'],1),'Mpx/s', round(d['ms_per_step'],2),'ms  e2e', round(d['e2e']['value'],1), 'cpu', d.get('cpu_baseline',{}).get('value'))
print({n:k[n]['ms'] for n in k}); print(d['roofline']['stages'])
for l in open('gpurun_out/${tag}_full_configs.jsonl'): print(l[:1200])
PY
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
