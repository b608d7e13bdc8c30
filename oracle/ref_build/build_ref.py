#!/usr/bin/env python
"""Compile the UNMODIFIED reference (patolette C + the vendored faiss KMeans slice)
in place from /root/reference into oracle/_ref/libpatolette_ref.so.

Test infrastructure only: the result is the strongest oracle we have (the
reference's own code), used by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing in patolette_b200/ links or
loads it.  No reference source is copied: gcc/g++ read the files where they
lie; only object files and the .so land in oracle/_ref/ (git-ignored).

Substitutions:
  * BLAS/LAPACK  -> the OpenBLAS bundled with scipy (symbols prefixed scipy_).
  * FLANN        -> exact brute-force shim (shim/flann_shim.c); FLANN is not
                    installed here and is un-vendored/unpinned upstream.
The reference's own build system (CMake + scikit-build) is NOT run.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref")
REF = os.environ.get("PATOLETTE_REFERENCE", "/root/reference")

FAISS_SOURCES = [
    "faiss/Clustering.cpp", "faiss/Index.cpp", "faiss/IndexFlat.cpp",
    "faiss/IndexFlatCodes.cpp", "faiss/VectorTransform.cpp",
    "faiss/utils/distances.cpp", "faiss/utils/distances_simd.cpp",
    "faiss/utils/random.cpp", "faiss/utils/utils.cpp", "faiss/utils/Heap.cpp",
    "faiss/utils/sorting.cpp", "faiss/utils/extra_distances.cpp",
    "faiss/utils/partitioning.cpp",
    "faiss/utils/distances_fused/distances_fused.cpp",
    "faiss/utils/distances_fused/simdlib_based.cpp",
    "faiss/utils/distances_fused/avx512.cpp",
    "faiss/impl/AuxIndexStructures.cpp", "faiss/impl/FaissException.cpp",
    "faiss/impl/IDSelector.cpp", "faiss/impl/CodePacker.cpp",
    "faiss/impl/kmeans1d.cpp", "faiss/impl/ProductQuantizer.cpp",
    "c_api/Clustering_c.cpp", "c_api/error_impl.cpp",
]
BLAS_RENAMES = ["sgemm_", "dgemm_", "ssyrk_", "sgesvd_", "dgesvd_", "sgeqrf_",
                "sorgqr_", "dsyev_", "sgelsd_", "sgetrf_", "sgetri_"]


def scipy_openblas() -> str:
    import scipy
    libs = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs",
                                  "libscipy_openblas-*.so"))
    if not libs:
        raise RuntimeError("scipy-bundled OpenBLAS not found")
    return os.path.realpath(libs[0])


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("compile failed")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "lib", "src"))


def lib_path() -> str:
    return os.path.join(OUT, "libpatolette_ref.so")


def build(force: bool = False) -> str:
    so = lib_path()
    if os.path.exists(so) and not force:
        return so
    if not available():
        raise RuntimeError(f"reference tree not present at {REF}")
    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)
    shim = os.path.join(HERE, "shim")
    jobs, objs = [], []
    c_srcs = sorted(glob.glob(os.path.join(REF, "lib", "src", "**", "*.c"), recursive=True))
    for src in c_srcs + [os.path.join(shim, "flann_shim.c")]:
        obj = os.path.join(OUT, "obj", "c_" + os.path.relpath(src, "/").replace("/", "_") + ".o")
        objs.append(obj)
        jobs.append(["gcc", "-O2", "-fPIC", "-fopenmp", "-w", f"-I{shim}", f"-I{REF}/lib/include",
                     f"-I{REF}/lib", "-Ddsyev_=scipy_dsyev_", "-c", src, "-o", obj])
    for rel in FAISS_SOURCES:
        src = os.path.join(REF, "lib", "faiss", rel)
        obj = os.path.join(OUT, "obj", "f_" + rel.replace("/", "_") + ".o")
        objs.append(obj)
        jobs.append(["g++", "-std=c++17", "-O2", "-fPIC", "-fopenmp", "-w", f"-I{REF}/lib/faiss",
                     "-DFINTEGER=int"] + [f"-D{s}=scipy_{s}" for s in BLAS_RENAMES]
                    + ["-c", src, "-o", obj])
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(run, jobs))
    blas = scipy_openblas()
    run(["g++", "-shared", "-fopenmp", "-o", so] + objs
        + [blas, f"-Wl,-rpath,{os.path.dirname(blas)}", "-lm"])
    return so


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
