#!/usr/bin/env python
"""Freeze outputs of the REFERENCE'S OWN CODE (oracle/_ref) at BASELINE.json's own sizes.

    python tests/golden/make_golden_big.py [case ...]        (no argument: every case not yet frozen)

Writes / extends tests/golden/golden_big.json: per case the sha256 of the size_t map bytes and of the
F-order palette bytes plus the palette as hex floats - no maps (they are 0.1-2 GB).  These are the pins
the `-m gpu` parity tests of tests/test_gpu_big.py compare the CUDA path with; the CPU side needs
seconds (C2) to tens of minutes and ~45 GB of RAM (C4), which is why they are frozen here instead of
being recomputed on the GPU box.  Results do not depend on OMP_NUM_THREADS (faiss' reductions are per
centroid, the exact-NN stand-in is per pixel); OpenBLAS is pinned to one thread (oracle/reflib.py).
"""
import os
os.environ["OPENBLAS_NUM_THREADS"] = "1"
os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
import hashlib
import json
import platform
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import scipy  # noqa: E402
from oracle.reflib import RefLib  # noqa: E402
from synth import BIG_CASES, make_case  # noqa: E402

OUT = os.path.join(HERE, "golden_big.json")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = RefLib()
    out = {"meta": {}, "cases": {}}
    if os.path.exists(OUT):
        with open(OUT) as f:
            out = json.load(f)
    out["meta"] = {
        "generator": "tests/golden/make_golden_big.py", "source": "oracle/_ref (reference compiled in place)",
        "numpy": np.__version__, "scipy_openblas": scipy.__version__, "glibc": platform.libc_ver()[1],
        "machine": platform.machine(), "nn_backend": "exact brute-force FLANN shim (lowest index on ties)",
    }
    names = sys.argv[1:] or [n for n in BIG_CASES if n not in out["cases"]]
    for name in names:
        spec = BIG_CASES[name]
        colors, weights, kw = make_case(spec)
        t0 = time.time()
        code, pal, pmap = ref.quantize(spec["w"], spec["h"], colors, spec["K"], weights=weights, **kw)
        dt = time.time() - t0
        del colors
        entry = {"spec": spec, "exit_code": code, "cpu_seconds": round(dt, 1), "omp_threads": int(os.environ["OMP_NUM_THREADS"]),
                 "palette_sha256": sha(pal.ravel(order="F")),
                 "palette_hex": [[float(v).hex() for v in row] for row in pal]}
        if pmap is not None:
            entry["map_sha256"] = sha(pmap)
            entry["map_distinct"] = int(len(np.unique(pmap)))
            entry["map_head"] = [int(v) for v in pmap[:16]]
        out["cases"][name] = entry
        print(name, code, f"{dt:.1f}s", entry.get("map_sha256", "-")[:12], entry["palette_sha256"][:12], flush=True)
        with open(OUT, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
