#!/bin/bash
# Runs ON the GPU box (under gpurun): one `ncu --set full` capture of selected kernels, exported to compact CSVs
# (gpurun only copies back <= 64 MiB, a full report with source is larger).
#   tools/ncu_capture.sh <tag> <kernel-regex> <skip> <count> [bench args...]
set -u
tag=$1; regex=$2; skip=$3; count=$4; shift 4
out=gpurun_out
mkdir -p $out
rep=/tmp/${tag}.ncu-rep
ncu --set full --clock-control none --import-source on -k "regex:${regex}" -s "$skip" -c "$count" -f -o /tmp/${tag} \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > $out/${tag}.log 2>&1
ncu -i $rep --page raw --csv 2>/dev/null | gzip -9 > $out/${tag}_raw.csv.gz
ncu -i $rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > $out/${tag}_source_sass.csv.gz
ncu -i $rep --page source --csv --print-source cuda 2>/dev/null | gzip -9 > $out/${tag}_source_cuda.csv.gz
ncu -i $rep --page details --csv 2>/dev/null | gzip -9 > $out/${tag}_details.csv.gz
ls -la $rep $out/${tag}_*
