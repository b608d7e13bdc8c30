#!/bin/bash
# on the GPU box (round 2): parity tests, bench line, ncu launch list, ncu --set full captures of the top kernels
tag=${1:-r02z}
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -2 gpurun_out/${tag}_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
# launch list of the bench command itself (per-launch times are cold-cache and serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_ncu_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${tag}_ncu_launches.log 2>&1
gzip -9 -f gpurun_out/${tag}_ncu_launches.csv
# the dominant streaming kernel at the bench's own size (first launch: the level below GQ, all 268 M pixels)
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_scatter2" -c 1 -f -o gpurun_out/${tag}_scatter2_c4 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${tag}_scatter2_c4.log 2>&1
# the other kernels of the path at 8192^2
timeout 600 ncu --set full --clock-control none --import-source on \
    -k "regex:k_color|k_ord_blocksum_raw|k_ord_fast|k_dots_minmax|k_buckets_hist|k_scatter2|k_permute_tile|k_riemersma_spec4|k_ord_resolve" \
    -c 16 -f -o gpurun_out/${tag}_kernels_8192 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --side 8192 > gpurun_out/${tag}_kernels_8192.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:k_riemersma_spec4|k_permute_tile|k_unpermute_tile" \
    -c 3 -f -o gpurun_out/${tag}_dither_8192 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --side 8192 > gpurun_out/${tag}_dither_8192.log 2>&1
ls -la gpurun_out/${tag}_*
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json')); k=d['roofline']['kernels']
print(round(d['value'],1),'Mpx/s', round(d['ms_per_step'],2),'ms  e2e', round(d['e2e']['value'],1), 'cpu', d.get('cpu_baseline',{}).get('value'))
print(d['roofline']['kernel'], d['roofline']['frac'], {n:k[n]['ms'] for n in k}); print(d['stage_ms'])
PY
# rows N2 / N3 (later in the round): eigen + saliency tests are part of `pytest -m gpu`; stage timings and the scans' timeline
python tools/time_saliency.py --ref-side 2048 > gpurun_out/${tag}_saliency_time.jsonl 2> gpurun_out/${tag}_saliency_time.err
python tools/mbd_timeline.py --side 4096 > gpurun_out/${tag}_mbd_timeline.txt 2>&1
python tools/time_weighted_dither.py > gpurun_out/${tag}_weighted_dither.jsonl 2>&1
