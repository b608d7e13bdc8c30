#!/bin/bash
# on the GPU box: compile-time variants "file.cu:-DFLAG=.. -DFLAG2=.." (only that object is rebuilt), full bench line each
cd patolette_b200
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --fmad=false -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden,-O2 --expt-relaxed-constexpr -DPATOLETTE_B200_BUILD"
for spec in "$@"; do
  f=${spec%%:*}; v=${spec#*:}
  nvcc $FLAGS $v -c csrc/$f -o build/$f.o 2>/dev/null || exit 1
  nvcc -shared -o libpatolette_b200.so build/*.o -lcudart_static -ldl -lpthread -lrt 2>/dev/null || exit 1
  ( cd .. && python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-extras $BENCH_ARGS > gpurun_out/var.json 2>gpurun_out/var.err
    python - "$spec" <<'PY'
import json,sys
d=json.load(open('gpurun_out/var.json')); k=d['roofline']['kernels']
print(sys.argv[1], round(d['ms_per_step'],2), d['stage_ms']['lq'], {n:k[n]['ms'] for n in list(k)[:9]})
PY
  )
  nvcc $FLAGS -c csrc/$f -o build/$f.o 2>/dev/null   # back to the default build of that file
done
nvcc -shared -o libpatolette_b200.so build/*.o -lcudart_static -ldl -lpthread -lrt 2>/dev/null
