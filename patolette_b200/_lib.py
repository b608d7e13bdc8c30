"""ctypes binding of libpatolette_b200.so (include/patolette_b200.h)."""
from __future__ import annotations

import ctypes as C
import glob
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpatolette_b200.so")


class QuantizationOptions(C.Structure):
    """patolette__QuantizationOptions (reference lib/include/patolette.h:13-20)."""
    _fields_ = [
        ("dither", C.c_bool),
        ("palette_only", C.c_bool),
        ("color_space", C.c_int),
        ("kmeans_niter", C.c_int),
        ("kmeans_max_samples", C.c_size_t),
        ("verbose", C.c_bool),
    ]


# every symbol include/patolette_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "patolette": (None, [C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                         C.POINTER(QuantizationOptions), C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "get_patolette_exit_code_info_message": (C.c_char_p, [C.c_int]),
    "patolette_create_default_options": (C.POINTER(QuantizationOptions), []),
    "patolette_b200_set_device": (C.c_int, [C.c_int]),
    "patolette_b200_device_count": (C.c_int, []),
    "patolette_b200_set_lapack": (None, [C.c_char_p]),
    "patolette_b200_lapack_source": (C.c_char_p, []),
    "patolette_b200_color_transform": (C.c_int, [C.c_int, C.c_void_p, C.c_size_t]),
    "patolette_b200_pow": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_size_t]),
    "patolette_b200_quantize_clusters": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                                   C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "patolette_b200_nearest": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "patolette_b200_kmeans": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "patolette_b200_dither": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "patolette_b200_last_timings": (C.c_int, [C.c_void_p]),
    "patolette_b200_interleaved": (None, [C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                          C.POINTER(QuantizationOptions), C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "patolette_b200_device": (None, [C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                     C.POINTER(QuantizationOptions), C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "patolette_b200_set_stream": (C.c_int, [C.c_void_p, C.c_int]),
    "patolette_b200_release_cache": (C.c_size_t, []),
    "patolette_b200_ordered_counts": (C.c_int, [C.c_void_p, C.c_int]),
    "patolette_b200_split_counts": (C.c_int, [C.c_void_p, C.c_int]),
    "patolette_b200_ordered_chain_debug": (C.c_int, [C.c_void_p, C.c_int]),
    "patolette_b200_set_option": (C.c_int, [C.c_char_p, C.c_longlong]),
    "patolette_b200_set_sharding": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "patolette_b200_gq_cuts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "patolette_b200_profile_enable": (C.c_int, [C.c_int]),
    "patolette_b200_profile_json": (C.c_size_t, [C.c_char_p, C.c_size_t]),
    "patolette_b200_profile_timeline": (C.c_size_t, [C.c_char_p, C.c_size_t]),
    "patolette_b200_fp64_peak": (C.c_double, []),
    "patolette_b200_u8": (None, [C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(QuantizationOptions),
                                 C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "patolette_b200_quantize": (None, [C.c_size_t, C.c_size_t, C.c_void_p, C.c_int, C.c_double, C.c_size_t,
                                       C.POINTER(QuantizationOptions), C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                       C.POINTER(C.c_int)]),
    "patolette_b200_saliency_weights": (C.c_int, [C.c_size_t, C.c_size_t, C.c_void_p, C.c_double, C.c_void_p, C.c_int]),
    "patolette_b200_saliency_mbd": (C.c_int, [C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]),
    "patolette_b200_last_saliency_ms": (C.c_double, []),
    "patolette_b200_eigen3": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "patolette_b200_comm_unique_id": (C.c_int, [C.c_char_p]),
    "patolette_b200_comm_init": (C.c_int, [C.c_int, C.c_int, C.c_char_p]),
    "patolette_b200_comm_destroy": (None, []),
    "patolette_b200_comm_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "patolette_b200_shard_range": (C.c_int, [C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "patolette_b200_sharded": (None, [C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(QuantizationOptions),
                                      C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
}

_lib = None


def _find_lapack() -> str | None:
    env = os.environ.get("PATOLETTE_B200_LAPACK")
    if env:
        return env
    try:
        import scipy
        libs = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas-*.so"))
        if libs:
            return os.path.realpath(libs[0])
    except Exception:
        pass
    return None


def load():
    """Load the CUDA library.  There is NO CPU fallback: a missing .so is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing - build it with `python -m patolette_b200.build` (needs nvcc). "
            "patolette_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header drift apart
        fn.restype = res
        fn.argtypes = args
    lapack = _find_lapack()
    if lapack:
        lib.patolette_b200_set_lapack(lapack.encode())
    _lib = lib
    return lib
