#!/usr/bin/env python
"""Freeze outputs of the REFERENCE'S OWN CODE (oracle/_ref, built in place from
/root/reference by oracle/ref_build/build_ref.py) for the cases in tests/synth.py.

    OMP_NUM_THREADS=1 python tests/golden/make_golden.py

Writes tests/golden/golden.json (per case: sha256 of the size_t map bytes and of the
F-order palette bytes, the palette itself as hex floats, and for small images the
map).  The reference ships no tests or fixtures of its own, so
these vectors are the pin.  Library provenance is recorded in the "meta" block.
"""
import os
os.environ["OPENBLAS_NUM_THREADS"] = "1"  # see oracle/reflib.py
os.environ.setdefault("OMP_NUM_THREADS", "1")
import hashlib
import json
import os
import platform
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import scipy  # noqa: E402
from oracle.reflib import RefLib  # noqa: E402
from synth import GOLDEN_CASES, make_case  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = RefLib()
    out = {"meta": {
        "generator": "tests/golden/make_golden.py", "source": "oracle/_ref (reference compiled in place)",
        "numpy": np.__version__, "scipy_openblas": scipy.__version__, "glibc": platform.libc_ver()[1],
        "machine": platform.machine(), "nn_backend": "exact brute-force FLANN shim (lowest index on ties)",
    }, "cases": {}}
    for name, spec in GOLDEN_CASES.items():
        colors, weights, kw = make_case(spec)
        code, pal, pmap = ref.quantize(spec["w"], spec["h"], colors, spec["K"], weights=weights, **kw)
        entry = {"spec": spec, "exit_code": code,
                 "palette_sha256": sha(pal.ravel(order="F")),
                 "palette_hex": [[float(v).hex() for v in row] for row in pal]}
        if pmap is not None:
            entry["map_sha256"] = sha(pmap)
            entry["map_distinct"] = int(len(np.unique(pmap)))
            if pmap.size <= 4096:
                entry["map"] = [int(v) for v in pmap]
        out["cases"][name] = entry
        print(name, code, entry.get("map_sha256", "-")[:12], entry["palette_sha256"][:12])
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
