python -m pytest tests -m gpu -x -q > gpurun_out/s12_tests.log 2>&1; tail -3 gpurun_out/s12_tests.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s12_bench.json 2> gpurun_out/s12_bench.err
for v in 6 8 10; do PB200_NVCC_EXTRA="-DPB_OS_MINB=$v" python -m patolette_b200.build --force > /dev/null 2>&1; python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s12_bench_minb$v.json 2> gpurun_out/s12_bench_minb$v.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s12_bench*.json')):
    try:
        d=json.load(open(f)); k=d['roofline']['kernels']
        print(f, round(d['ms_per_step'],2), {n:k[n]['ms'] for n in k if 'ord' in n or 'chains' in n})
    except Exception as e: print(f, 'ERR', e)
PY
