"""CPU suite: the oracle restatement against the frozen reference outputs and, when the
reference itself is available, against oracle/_ref bit for bit."""
import ctypes as C

import numpy as np
import pytest

from conftest import sha
from synth import GOLDEN_CASES, WIDER, make_case


def _check_against_golden(entry, code, pal, pmap):
    assert code == entry["exit_code"]
    want = np.array([[float.fromhex(v) for v in row] for row in entry["palette_hex"]])
    np.testing.assert_array_equal(np.isnan(pal), np.isnan(want))
    assert sha(pal.ravel(order="F")) == entry["palette_sha256"], "palette bits differ from the reference"
    if "map_sha256" in entry:
        assert sha(pmap) == entry["map_sha256"], "palette_map differs from the reference"
    else:
        assert pmap is None


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_oracle_matches_golden(oracle, golden, name):
    spec = GOLDEN_CASES[name]
    colors, weights, kw = make_case(spec)
    code, pal, pmap = oracle.quantize(spec["w"], spec["h"], colors, spec["K"], weights=weights, **kw)
    _check_against_golden(golden["cases"][name], code, pal, pmap)


@pytest.mark.parametrize("name", ["512_k16_ictcp", "300x200_k64_weighted", "256_k256_imagelike_luv", "4x4_k8_kmeans_seq"])
def test_oracle_matches_reference_build(oracle, reflib, name):
    spec = GOLDEN_CASES[name]
    colors, weights, kw = make_case(spec)
    a = reflib.quantize(spec["w"], spec["h"], colors, spec["K"], weights=weights, **kw)
    b = oracle.quantize(spec["w"], spec["h"], colors, spec["K"], weights=weights, **kw)
    assert a[0] == b[0]
    assert np.array_equal(a[1].view(np.uint64), b[1].view(np.uint64))
    assert np.array_equal(a[2], b[2])


@pytest.mark.parametrize("name", sorted(WIDER))
def test_oracle_matches_reference_build_wider(oracle, reflib, name):
    spec = WIDER[name]
    colors, weights, kw = make_case(spec)
    a = reflib.quantize(spec["w"], spec["h"], colors, spec["K"], weights=weights, **kw)
    b = oracle.quantize(spec["w"], spec["h"], colors, spec["K"], weights=weights, **kw)
    assert a[0] == b[0]
    assert np.array_equal(np.asarray(a[1]).view(np.uint64), np.asarray(b[1]).view(np.uint64)), "palette bits differ"
    if a[2] is None:
        assert b[2] is None
    else:
        assert np.array_equal(a[2], b[2]), f"{int((a[2] != b[2]).sum())} map entries differ"


def test_reference_build_matches_golden(reflib, golden):
    """The goldens were produced by oracle/_ref in the authoring container; on another host CPU
    OpenBLAS may pick other dgemv/sgemm kernels, so this documents (xfail) rather than gates."""
    spec = GOLDEN_CASES["512_k16_ictcp"]
    colors, weights, kw = make_case(spec)
    code, pal, pmap = reflib.quantize(spec["w"], spec["h"], colors, spec["K"], weights=weights, **kw)
    if sha(pmap) != golden["cases"]["512_k16_ictcp"]["map_sha256"]:
        pytest.xfail("host BLAS on this box evaluates dgemv/sgemm differently from the golden host")


def test_exit_codes(oracle):
    z = np.zeros((0, 3))
    assert oracle.quantize(0, 5, z, 4)[0] == -2
    assert oracle.quantize(2, 2, np.zeros((4, 3)), 0)[0] == -3
    msg = oracle.lib.get_patolette_exit_code_info_message
    assert msg(0) == b"Quantization successful."
    assert msg(-3) == b"Palette size should be greater than 0."


def test_dgemv_formula_matches_live_blas(oracle):
    """sort.c:43 delegates to cblas_dgemv; the oracle (and the CUDA path) hard-code the per-row
    formula fma(a0,x0,a1*x1)+a2*x2.  Re-derive it against the BLAS on this box, including row
    counts that exercise the kernel's tail handling."""
    f = oracle.lib.orc_selftest_dgemv
    f.restype = C.c_long
    f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    rng = np.random.default_rng(5)
    bad = 0
    for n in (1, 2, 3, 5, 7, 64, 1001, 65537, 300001):
        a = np.asfortranarray(rng.random((n, 3)) * 2 - 0.5)
        x = rng.standard_normal(3)
        bad += f(a.ctypes.data, n, x.ctypes.data)
    if bad:
        pytest.xfail(f"{bad} rows: this host's BLAS dgemv kernel rounds differently from the pinned formula")


def test_stage_pca_matches_numpy_left_to_right(oracle):
    """strictly sequential accumulation reproduces the mean bit for bit."""
    n = 20001
    rng = np.random.default_rng(2)
    c = np.asfortranarray(rng.random((n, 3)))
    w = 1 + rng.random(n) * 7
    mean = np.zeros(3); vcov = np.zeros(9); axis = np.zeros(3)
    f = oracle.lib.orc_pca
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    assert f(c.ctypes.data, n, w.ctypes.data, None, n, mean.ctypes.data, vcov.ctypes.data, axis.ctypes.data) == 0
    wsum = np.add.accumulate(w)[-1]
    want = np.array([np.add.accumulate(c[:, j] * w)[-1] * (1 / wsum) for j in range(3)])
    assert np.array_equal(mean.view(np.uint64), want.view(np.uint64))
    assert abs(np.linalg.norm(axis) - 1) < 1e-12
