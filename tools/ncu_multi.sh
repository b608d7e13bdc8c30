#!/bin/bash
# Runs ON the GPU box: one `ncu --set full` capture of the first <count> launches matching <regex>, report kept as
# gpurun_out/<tag>.ncu-rep (read back with `ncu -i ... --page raw --csv`).   tools/ncu_multi.sh <tag> <regex> <count> [bench args]
set -u
tag=$1; regex=$2; count=$3; shift 3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:${regex}" -c "$count" -f -o gpurun_out/${tag} \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras "$@" > gpurun_out/${tag}.log 2>&1
ls -la gpurun_out/${tag}.ncu-rep
