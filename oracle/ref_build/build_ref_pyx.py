#!/usr/bin/env python
"""Compile the reference's OWN Python wrapper (src/patolette/patolette.pyx: `quantize`, and the saliency code
`raster_scan[_inv]`, `mbd`, `get_weights`, :54-313) in place from /root/reference into
oracle/_ref/pyx/patolette.<abi>.so, linked against oracle/_ref/libpatolette_ref.so (build_ref.py).

Test infrastructure only (row N3's pin): the module is imported by tests/ and by tests/golden/make_golden_saliency.py
to check oracle/saliency_port.py and to produce golden weights.  No reference source is copied: Cython reads the
.pyx where it lies; the generated C file and the .so land in oracle/_ref/pyx/ (git-ignored).  The reference's
build system (scikit-build + CMake) is NOT run.

Substitution: scikit-image is not installed -> shim/skimage (our restatement of `rgb2lab`, see its docstring).
Everything else the wrapper calls is the real thing: numpy, scipy's `cdist`, the reference's C library.
"""
from __future__ import annotations

import importlib
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref", "pyx")
REF = os.environ.get("PATOLETTE_REFERENCE", "/root/reference")
PYX = os.path.join(REF, "src", "patolette", "patolette.pyx")
SHIM = os.path.join(HERE, "shim")


def available() -> bool:
    return os.path.isfile(PYX)


def so_path() -> str:
    return os.path.join(OUT, "patolette" + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force: bool = False) -> str:
    so = so_path()
    if os.path.exists(so) and not force:
        return so
    if not available():
        raise RuntimeError(f"reference wrapper not present at {PYX}")
    from oracle.ref_build import build_ref
    ref_so = build_ref.build()
    import numpy as np
    os.makedirs(OUT, exist_ok=True)
    c_file = os.path.join(OUT, "patolette.c")
    r = subprocess.run([sys.executable, "-m", "cython", "-3", PYX, "-o", c_file], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("cython failed")
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-w", f"-I{sysconfig.get_paths()['include']}", f"-I{np.get_include()}",
           f"-I{REF}/lib/include", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION", c_file, "-o", so, ref_so,
           f"-Wl,-rpath,{os.path.dirname(ref_so)}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("compile failed")
    return so


def load():
    """The compiled reference wrapper as a module (prebuilt file: works without /root/reference)."""
    so = so_path()
    if not os.path.exists(so):
        build()
    for p in (SHIM, OUT):
        if p not in sys.path:
            sys.path.insert(0, p)
    return importlib.import_module("patolette")


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    print(build(force="--force" in sys.argv))
    m = load()
    print(sorted(n for n in dir(m) if not n.startswith("_"))[:12])
