"""GPU suite (-m gpu), part 2: parity at BASELINE.json's own sizes and option classes.

  * frozen hashes of the REFERENCE'S OWN OUTPUT (tests/golden/golden_big.json, made on the CPU by
    tests/golden/make_golden_big.py from oracle/_ref): C2 at 4096^2, C3 at 8192^2, C4 at 4096^2 and at
    16384^2, a C5-class weighted K=1024 run with full-image KMeans and dither, an image-like weighted CIELuv
    run, K=8192 - the CUDA path must give the same palette bits and the same map;
  * the 7 wider configurations of tests/synth.py::WIDER against the oracle restatement, live;
  * the reference build itself (oracle/_ref, shipped prebuilt to the GPU box) run LIVE on this box's CPU
    against the CUDA path - the dgemv / sgemm / dsyev behaviour the goldens pin is host-CPU dependent, this
    is the check that "identical to the reference" holds on the box the product runs on;
  * palette sizes beyond the shared-memory palette paths, the KMeans size limit, the LAPACK fail-closed rule.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import ROOT, sha
from synth import BIG_CASES, WIDER, make_case

pytestmark = pytest.mark.gpu


def cuda_quantize(lib, w, h, colors, K, weights=None, **kw):
    from oracle.reflib import quantize_with
    return quantize_with(lib, w, h, colors, K, weights=weights, **kw)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


@pytest.fixture(scope="module")
def golden_big():
    with open(os.path.join(ROOT, "tests", "golden", "golden_big.json")) as f:
        return json.load(f)["cases"]


def _palette_from_hex(entry):
    return np.array([[float.fromhex(v) for v in row] for row in entry["palette_hex"]])


BIG_ON_GPU = [n for n in BIG_CASES if n != "k70000_512"]  # K above PB_KMEANS_MAX_K with KMeans: see test_kmeans_size_limit


@pytest.mark.parametrize("name", BIG_ON_GPU)
def test_big_case_matches_frozen_reference(cuda_lib, golden_big, name):
    if name not in golden_big:
        pytest.skip(f"{name} has no frozen reference output (tests/golden/make_golden_big.py {name})")
    entry = golden_big[name]
    spec = BIG_CASES[name]
    colors, weights, kw = make_case(spec)
    code, pal, pmap = cuda_quantize(cuda_lib, spec["w"], spec["h"], colors, spec["K"], weights=weights, **kw)
    del colors
    assert code == entry["exit_code"] == 0
    want = _palette_from_hex(entry)
    diff = bits(pal) != bits(want)
    assert not diff.any(), f"{int(diff.sum())} palette values differ from the reference (max abs {np.abs(pal - want).max()})"
    assert sha(pal.ravel(order="F")) == entry["palette_sha256"]
    assert [int(v) for v in pmap[:16]] == entry["map_head"]
    assert int(len(np.unique(pmap))) == entry["map_distinct"]
    assert sha(pmap) == entry["map_sha256"], "palette_map differs from the reference's"


@pytest.mark.parametrize("name", sorted(WIDER))
def test_wider_case_matches_oracle(cuda_lib, oracle, name):
    spec = WIDER[name]
    colors, weights, kw = make_case(spec)
    a = oracle.quantize(spec["w"], spec["h"], colors, spec["K"], weights=weights, **kw)
    b = cuda_quantize(cuda_lib, spec["w"], spec["h"], colors, spec["K"], weights=weights, **kw)
    assert a[0] == b[0] == 0
    assert np.array_equal(bits(a[1]), bits(b[1])), "palette bits differ"
    if a[2] is None:
        assert b[2] is None
    else:
        assert np.array_equal(a[2], b[2]), f"{int((a[2] != b[2]).sum())} map entries differ"


LIVE = {
    "512_k16_ictcp": dict(w=512, h=512, K=16, seed=0, color_space=2, dither=False, kmeans_niter=0),
    "300x200_k64_weighted_kmeans_dither": dict(w=300, h=200, K=64, seed=5, color_space=2, dither=True, kmeans_niter=5, weighted=True),
    "1024_k256_ictcp_kmeans_dither": dict(w=1024, h=1024, K=256, seed=51, color_space=2, dither=True, kmeans_niter=4),
    "640x480_k128_luv_imagelike": dict(w=640, h=480, K=128, seed=52, color_space=1, dither=False, kmeans_niter=0, image_like=True),
    "777x333_k200_srgb_dither": dict(w=777, h=333, K=200, seed=53, color_space=0, dither=True, kmeans_niter=2),
}


@pytest.mark.parametrize("name", sorted(LIVE))
def test_reference_build_live_on_this_box(cuda_lib, reflib, name):
    """oracle/_ref is the reference's own C/C++ (compiled in place in the authoring container, shipped as a
    built .so).  Here it runs on THIS box's CPU - whatever dgemv / sgemm kernels this CPU makes OpenBLAS
    pick - and the CUDA output must equal it bit for bit."""
    spec = LIVE[name]
    colors, weights, kw = make_case(spec)
    a = reflib.quantize(spec["w"], spec["h"], colors, spec["K"], weights=weights, **kw)
    b = cuda_quantize(cuda_lib, spec["w"], spec["h"], colors, spec["K"], weights=weights, **kw)
    assert a[0] == b[0] == 0
    nd = int((bits(a[1]) != bits(b[1])).sum())
    assert nd == 0, f"{nd} palette values differ from the reference running on this box ({_cpu()})"
    assert np.array_equal(a[2], b[2]), f"{int((a[2] != b[2]).sum())} map entries differ from the reference on this box ({_cpu()})"


def _cpu():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown cpu"


def test_eigen_solver_is_the_builtin_restatement_and_host_lapack_agrees(cuda_lib):
    """Default: pb_dsyev3.h (row N2).  "host_lapack" = 1 calls the real dsyev_ instead: same palette, same map."""
    src = cuda_lib.patolette_b200_lapack_source().decode()
    assert src.startswith("builtin-dsyev3"), src
    colors, weights, kw = make_case(dict(w=256, h=192, K=64, seed=5, color_space=2, dither=True, kmeans_niter=3, weighted=True))
    a = cuda_quantize(cuda_lib, 256, 192, colors, 64, weights=weights, **kw)
    assert cuda_lib.patolette_b200_set_option(b"host_lapack", 1) == 0
    try:
        src = cuda_lib.patolette_b200_lapack_source().decode()
        assert src != "builtin-jacobi" and "dsyev" in src, src
        b = cuda_quantize(cuda_lib, 256, 192, colors, 64, weights=weights, **kw)
    finally:
        cuda_lib.patolette_b200_set_option(b"host_lapack", 0)
    assert a[0] == 0 and b[0] == 0
    assert np.array_equal(a[1].view(np.uint64), b[1].view(np.uint64)) and np.array_equal(a[2], b[2])


def test_missing_lapack_fails_closed(cuda_lib):
    """"host_lapack" mode without a dsyev_ -> exit code -1, not a silently different palette; allow_jacobi opts in."""
    from patolette_b200 import _lib
    colors, weights, kw = make_case(dict(w=64, h=64, K=8, seed=3, color_space=2, dither=False, kmeans_niter=0))
    real = _lib._find_lapack()
    saved = os.environ.pop("PATOLETTE_B200_LAPACK", None)
    try:
        cuda_lib.patolette_b200_set_option(b"host_lapack", 1)
        cuda_lib.patolette_b200_set_lapack(b"/nonexistent/liblapack.so")
        if cuda_lib.patolette_b200_lapack_source().decode() != "builtin-jacobi":
            pytest.skip("a system LAPACK is installed on this box: the fall-through search found it")
        code, pal, pmap = cuda_quantize(cuda_lib, 64, 64, colors, 8, **kw)
        assert code == -1
        assert cuda_lib.patolette_b200_set_option(b"allow_jacobi", 1) == 0
        code, pal, pmap = cuda_quantize(cuda_lib, 64, 64, colors, 8, **kw)
        assert code == 0 and len(np.unique(pmap)) == 8
    finally:
        cuda_lib.patolette_b200_set_option(b"allow_jacobi", 0)
        cuda_lib.patolette_b200_set_option(b"host_lapack", 0)
        cuda_lib.patolette_b200_set_lapack(real.encode() if real else None)
        if saved is not None:
            os.environ["PATOLETTE_B200_LAPACK"] = saved
    code, pal, pmap = cuda_quantize(cuda_lib, 64, 64, colors, 8, **kw)
    assert code == 0


def test_kmeans_size_limit(cuda_lib, golden_big):
    """palette_size above PB_KMEANS_MAX_K (50000) with kmeans_niter > 0: a dedicated exit code, not a silently
    un-refined palette.  The same K without KMeans runs (global-memory palette paths)."""
    spec = BIG_CASES["k70000_512"]
    colors, weights, kw = make_case(spec)
    code, pal, pmap = cuda_quantize(cuda_lib, spec["w"], spec["h"], colors, spec["K"], **kw)
    assert code == -6
    assert b"not supported" in cuda_lib.get_patolette_exit_code_info_message(-6)


def test_k70000_without_kmeans_matches_oracle(cuda_lib, oracle):
    """K = 70000 > 65535 on a 300 x 300 image, no KMeans: tree, palette and exact map for a palette that fits
    neither shared memory nor 16 bits."""
    w = h = 300
    K = 70000
    colors, weights, kw = make_case(dict(w=w, h=h, K=K, seed=8, color_space=2, dither=False, kmeans_niter=0))
    a = oracle.quantize(w, h, colors, K, **kw)
    b = cuda_quantize(cuda_lib, w, h, colors, K, **kw)
    assert a[0] == b[0] == 0
    assert np.array_equal(bits(a[1]), bits(b[1])), "palette bits differ"
    assert np.array_equal(a[2], b[2])


def test_k20000_kmeans_dither_matches_oracle(cuda_lib, oracle):
    """K = 20000: above the shared-memory limits of the dither (48 B x K), the brute-force map (24 B x K) and the
    KMeans assignment (16 B x K) - the global-memory palette routes - against the oracle."""
    w, h, K = 400, 300, 20000
    colors, weights, kw = make_case(dict(w=w, h=h, K=K, seed=9, color_space=2, dither=True, kmeans_niter=2))
    a = oracle.quantize(w, h, colors, K, **kw)
    b = cuda_quantize(cuda_lib, w, h, colors, K, **kw)
    assert a[0] == b[0] == 0
    assert np.array_equal(bits(a[1]), bits(b[1])), "palette bits differ"
    assert np.array_equal(a[2], b[2]), f"{int((a[2] != b[2]).sum())} map entries differ"


@pytest.mark.parametrize("spec", [
    dict(w=512, h=512, K=16, color_space=2, dither=False, kmeans_niter=0),
    dict(w=640, h=400, K=256, color_space=2, dither=True, kmeans_niter=4),
    dict(w=333, h=222, K=300, color_space=1, dither=True, kmeans_niter=0),          # 16-bit map
    dict(w=1600, h=1200, K=64, color_space=0, dither=False, kmeans_niter=2, weighted=True),
])
def test_u8_ingest_equals_f64_abi(cuda_lib, spec):
    """N1: uint8 RGB in, / 255 on the device, narrow map out == the f64 ABI fed rgb / 255.0 (README.md:155-158)."""
    import patolette_b200 as pb
    from synth import saliency_like_weights
    w, h, K = spec["w"], spec["h"], spec["K"]
    rng = np.random.default_rng(w + h)
    yy, xx = np.mgrid[0:h, 0:w]
    rgb = np.stack([(xx * 255 // w), (yy * 255 // h), ((xx + yy) * 255 // (w + h))], -1).reshape(-1, 3)
    rgb = np.clip(rgb + rng.integers(-40, 41, rgb.shape), 0, 255).astype(np.uint8)
    weights = saliency_like_weights(w, h, 3) if spec.get("weighted") else None
    kw = dict(dither=spec["dither"], color_space=spec["color_space"], kmeans_niter=spec["kmeans_niter"])
    colors = rgb.astype(np.float64)
    colors /= 255
    ok, pal, pmap, msg = pb.quantize(w, h, colors, K, tile_size=0, weights=weights, **kw)
    ok8, pal8, map8, msg8 = pb.quantize_u8(w, h, rgb, K, weights=weights, **kw)
    assert ok and ok8, (msg, msg8)
    assert map8.dtype == (np.uint8 if K <= 256 else np.uint16)
    assert np.array_equal(bits(pal), bits(pal8)), "palette bits differ"
    assert np.array_equal(pmap, map8.astype(np.uintp))
    # a map type that cannot hold the indices is refused
    from patolette_b200 import _lib
    code = C.c_int(0)
    opts = _lib.QuantizationOptions(False, False, 2, 0, 512 ** 2, False)
    palbuf = np.zeros((300, 3), order="F")
    small = np.zeros(w * h, dtype=np.uint8)
    cuda_lib.patolette_b200_u8(w, h, rgb.ctypes.data, None, 300, C.byref(opts), palbuf.ctypes.data, small.ctypes.data, 1, 0, C.byref(code))
    assert code.value == -3
