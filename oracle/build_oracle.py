#!/usr/bin/env python
"""Compile the plain-C restatement oracle/patolette_oracle.c into
oracle/_build/libpatolette_oracle.so (TEST INFRASTRUCTURE ONLY).

Links the scipy-bundled OpenBLAS for the one LAPACK entry point the reference
itself calls (dsyev_, math/eigen.c:50) and for the dgemv self-test.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "patolette_oracle.c")
OUT = os.path.join(HERE, "_build")


def lib_path() -> str:
    return os.path.join(OUT, "libpatolette_oracle.so")


def stale() -> bool:
    so = lib_path()
    return (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(SRC)


def scipy_openblas() -> str:
    import scipy
    libs = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs",
                                  "libscipy_openblas-*.so"))
    if not libs:
        raise RuntimeError("scipy-bundled OpenBLAS not found")
    return os.path.realpath(libs[0])


def build(force: bool = False) -> str:
    so = lib_path()
    if not force and not stale():
        return so
    os.makedirs(OUT, exist_ok=True)
    blas = scipy_openblas()
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-fopenmp", "-fvisibility=hidden",
           "-Wall", "-Wno-unused-function", SRC, "-o", so, blas,
           f"-Wl,-rpath,{os.path.dirname(blas)}", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("oracle compile failed")
    if r.stderr.strip():
        sys.stderr.write(r.stderr)
    return so


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
