#!/bin/bash
# on the GPU box: compile-time variants of pb_certify.cu (only that object is rebuilt), bench at --side 8192
cd patolette_b200
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --fmad=false -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden,-O2 --expt-relaxed-constexpr -DPATOLETTE_B200_BUILD"
for v in "$@"; do
  nvcc $FLAGS $v -c csrc/pb_certify.cu -o build/pb_certify.cu.o || exit 1
  nvcc -shared -o libpatolette_b200.so build/*.o -lcudart_static -ldl -lpthread -lrt || exit 1
  ( cd .. && python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-extras --side 8192 > gpurun_out/var.json 2>gpurun_out/var.err
    python - "$v" <<'PY'
import json,sys
d=json.load(open('gpurun_out/var.json')); k=d['roofline']['kernels']
print(sys.argv[1], round(d['ms_per_step'],2), {n:k[n] for n in k if 'hist' in n}, d.get('split_routes'))
PY
  )
done
