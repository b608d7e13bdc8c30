"""ctypes access to the two CPU checkers (TEST INFRASTRUCTURE ONLY).

  * ``RefLib``    - oracle/_ref/libpatolette_ref.so: the reference's own C/C++
                    compiled in place by oracle/ref_build/build_ref.py.
  * ``OracleLib`` - oracle/_build/libpatolette_oracle.so: our plain-C restatement
                    of the same algorithm (oracle/patolette_oracle.c).

Both export the reference's public C ABI (lib/include/patolette.h:22-35), so one
``quantize()`` helper drives either.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the
product package (patolette_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# The reference's projections go through cblas_dgemv (quantize/sort.c:43).  A multi-threaded
# OpenBLAS splits the rows between threads and each range gets its own differently-rounded
# scalar tail, so results would depend on the thread count.  Pin the BLAS to one thread
# (must happen before the library is first loaded); faiss' own OpenMP threading is unaffected.
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")

ColorSpace_sRGB, ColorSpace_CIELuv, ColorSpace_ICtCp = 0, 1, 2


class QuantizationOptions(C.Structure):
    """lib/include/patolette.h:13-20 (x86-64 layout: offsets 0,1,4,8,16,24; sizeof 32)."""
    _fields_ = [
        ("dither", C.c_bool),
        ("palette_only", C.c_bool),
        ("color_space", C.c_int),
        ("kmeans_niter", C.c_int),
        ("kmeans_max_samples", C.c_size_t),
        ("verbose", C.c_bool),
    ]


def _bind_public_abi(lib):
    lib.patolette.restype = None
    lib.patolette.argtypes = [
        C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
        C.POINTER(QuantizationOptions), C.c_void_p, C.c_void_p, C.POINTER(C.c_int),
    ]
    lib.get_patolette_exit_code_info_message.restype = C.c_char_p
    lib.get_patolette_exit_code_info_message.argtypes = [C.c_int]
    lib.patolette_create_default_options.restype = C.POINTER(QuantizationOptions)
    lib.patolette_create_default_options.argtypes = []


def quantize_with(lib, width, height, colors, palette_size, dither=True, palette_only=False,
                  color_space=ColorSpace_ICtCp, kmeans_niter=32, kmeans_max_samples=512 ** 2,
                  weights=None, verbose=False):
    """Mirror of src/patolette/patolette.pyx:332-466 minus the saliency step, plus an
    explicit ``weights`` (the C ABI has it, the Python surface derives it from tile_size).
    Returns (exit_code, palette[K,3] F-order f64, palette_map[N] uintp | None)."""
    n = width * height
    data = np.asfortranarray(colors, dtype=np.float64)
    assert data.shape == (n, 3)
    palette = np.zeros((palette_size, 3), dtype=np.float64, order="F")
    pmap = None if palette_only else np.zeros(n, dtype=np.uintp)
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    opts = QuantizationOptions(bool(dither), bool(palette_only), int(color_space),
                               int(kmeans_niter), int(kmeans_max_samples), bool(verbose))
    code = C.c_int(0)
    # each library binds its own (layout-identical) options class: cast to whatever it declared
    popts = C.cast(C.pointer(opts), lib.patolette.argtypes[5])
    lib.patolette(width, height, data.ctypes.data if n else None,
                  None if w is None else w.ctypes.data, palette_size, popts,
                  palette.ctypes.data if palette_size else None,
                  None if pmap is None else pmap.ctypes.data, C.byref(code))
    return code.value, palette, pmap


class RefLib:
    """The reference's own code (oracle/_ref)."""

    def __init__(self, build_if_missing=True):
        from oracle.ref_build import build_ref
        path = build_ref.lib_path()
        if not os.path.exists(path):
            if not (build_if_missing and build_ref.available()):
                raise FileNotFoundError(path)
            build_ref.build()
        self.path = path
        self.lib = C.CDLL(path)
        _bind_public_abi(self.lib)

    def quantize(self, *a, **k):
        return quantize_with(self.lib, *a, **k)


class OracleLib:
    """Our plain-C restatement (oracle/patolette_oracle.c)."""

    def __init__(self, build_if_missing=True):
        from oracle import build_oracle
        path = build_oracle.lib_path()
        if not os.path.exists(path) or build_oracle.stale():
            if not build_if_missing:
                raise FileNotFoundError(path)
            build_oracle.build()
        self.path = path
        self.lib = C.CDLL(path)
        _bind_public_abi(self.lib)

    def quantize(self, *a, **k):
        return quantize_with(self.lib, *a, **k)
