#!/usr/bin/env python
"""Build libpatolette_b200.so in-tree with nvcc for sm_100a.

    python -m patolette_b200.build [--force]

No torch involved: the library is plain CUDA C++ behind a C ABI (include/patolette_b200.h).
Flags that matter for parity: --fmad=false (every FMA in the sources is explicit; the
reference is a generic x86-64 build without contraction) and default IEEE div/sqrt.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libpatolette_b200.so")
CU = ["pb_color.cu", "pb_chain.cu", "pb_ordered.cu", "pb_parallel.cu", "pb_certify.cu", "pb_nngrid.cu", "pb_kmeans.cu", "pb_dither.cu", "pb_saliency.cu", "pb_eigen.cu", "pb_pipeline.cu"]
CPP = ["pb_lapack.cpp", "pb_prof.cpp", "pb_pool.cpp", "pb_xfer.cpp", "pb_hostpool.cpp", "pb_nccl.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false",
              "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2,-ffp-contract=off", "--expt-relaxed-constexpr", "-DPATOLETTE_B200_BUILD"]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return [os.path.join(CSRC, f) for f in CU + CPP]


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "patolette_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("patolette_b200 build failed")
    return r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    cc = nvcc()
    objs, jobs = [], []
    for src in sources():
        obj = os.path.join(BUILD, os.path.basename(src) + ".o")
        objs.append(obj)
        extra = (["-Xptxas", "-v"] if verbose else []) + os.environ.get("PB200_NVCC_EXTRA", "").split()
        jobs.append([cc] + NVCC_FLAGS + extra + ["-c", src, "-o", obj])
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        logs = list(ex.map(_run, jobs))
    if verbose:
        sys.stderr.write("".join(logs))
    # C ABI symbols are exported through extern "C" + default visibility in pb_pipeline.cu
    _run([cc, "-shared", "-o", LIB] + objs + ["-lcudart_static", "-ldl", "-lpthread", "-lrt"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
