"""Row N2: the 3 x 3 symmetric eigen solve (lib/src/math/eigen.c:83-140).  patolette_b200/csrc/pb_dsyev3.h restates
LAPACK's dsyev('V', 'L', 3) operation for operation; here it is compared with the real dsyev_ (scipy's OpenBLAS - the
LAPACK the oracle's reference build links), eigenvalues and eigenvectors bit for bit:
  * CPU: the header compiled for the host, 3 million matrices of five adversarial families (tests/native/test_dsyev3.cpp),
    and the C-ABI stage entry's host instantiation;
  * GPU: the device instantiation (k_eigen3) on a million matrices against the host one and against dsyev_."""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def scipy_openblas():
    import scipy
    libs = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas-*.so"))
    return os.path.realpath(libs[0]) if libs else None


def covariances(n: int, seed: int) -> np.ndarray:
    """n symmetric 3 x 3 matrices (column-major, lower triangle significant): sample covariances of few colours
    (8-bit, luminance-dominated, rank deficient), generic symmetric ones, structured zeros and ties, extreme norms."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, 9))
    k = n // 4
    pts = rng.integers(0, 256, (k, 12, 3)) / 255.0                      # 8-bit colours
    lum = rng.random((k, 12, 1)) * np.array([0.57, 0.59, 0.57]) + 0.01 * rng.random((k, 12, 3))
    for j, p in enumerate((pts, lum)):
        c = p - p.mean(axis=1, keepdims=True)
        cov = np.einsum("nij,nik->njk", c, c) / 12.0
        out[j * k:(j + 1) * k] = cov.reshape(k, 9)
    g = rng.standard_normal((k, 3, 3)) * np.exp(3 * rng.standard_normal((k, 1, 1)))
    out[2 * k:3 * k] = (g + g.transpose(0, 2, 1)).reshape(k, 9)
    rest = n - 3 * k
    vals = np.array([0.0, 1.0, -1.0, 0.5, 0.25, 1e-160, 1e150, 3.0])
    s = vals[rng.integers(0, len(vals), (rest, 3, 3))]
    out[3 * k:] = np.tril(s).reshape(rest, 9) + np.tril(s, -1).transpose(0, 2, 1).reshape(rest, 9)
    return out


def real_dsyev(mats: np.ndarray):
    lib = C.CDLL(scipy_openblas())
    f = lib.scipy_dsyev_
    f.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.c_void_p,
                  C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_size_t, C.c_size_t]
    n3, lwork, info = C.c_int(3), C.c_int(128), C.c_int(0)
    work = np.empty(128)
    z = mats.copy()
    w = np.empty((len(mats), 3))
    for i in range(len(mats)):
        f(b"V", b"L", C.byref(n3), z[i].ctypes.data, C.byref(n3), w[i].ctypes.data, work.ctypes.data, C.byref(lwork),
          C.byref(info), 1, 1)
    return w, z


def ours(mats: np.ndarray, on_device: int):
    from patolette_b200 import _lib
    lib = _lib.load()
    a = np.ascontiguousarray(mats)
    w = np.empty((len(a), 3))
    z = np.empty((len(a), 9))
    info = np.empty(len(a), dtype=np.int32)
    assert lib.patolette_b200_eigen3(a.ctypes.data, len(a), w.ctypes.data, z.ctypes.data, info.ctypes.data, on_device) == 0
    return w, z, info


@pytest.mark.skipif(scipy_openblas() is None, reason="scipy's OpenBLAS not found")
def test_header_matches_real_dsyev_on_millions(tmp_path):
    exe = str(tmp_path / "test_dsyev3")
    blas = scipy_openblas()
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-I", os.path.join(ROOT, "patolette_b200", "csrc"),
                    os.path.join(ROOT, "tests", "native", "test_dsyev3.cpp"), "-o", exe, blas,
                    f"-Wl,-rpath,{os.path.dirname(blas)}"], check=True)
    r = subprocess.run([exe, "3"], capture_output=True, text=True, env={**os.environ, "OPENBLAS_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr
    assert r.stdout.strip().endswith("OK 3000000")


@pytest.mark.skipif(scipy_openblas() is None, reason="scipy's OpenBLAS not found")
def test_stage_entry_host_instantiation_matches_real_dsyev():
    mats = covariances(40000, 1)
    w0, z0 = real_dsyev(mats)
    w1, z1, info = ours(mats, 0)
    assert not info.any()
    assert np.array_equal(w0.view(np.uint64), w1.view(np.uint64))
    assert np.array_equal(z0.view(np.uint64), z1.view(np.uint64))


@pytest.mark.skipif(scipy_openblas() is None, reason="scipy's OpenBLAS not found")
def test_special_matrices_match_real_dsyev():
    """Zeros, identity, repeated eigenvalues, a single off-diagonal entry, subnormal / near-overflow entries, NaN and
    Inf: same bits as dsyev_ (and no endless loop: dsteqr's iteration cap is part of the restatement)."""
    def sym(a11, a21, a31, a22, a32, a33):
        return [a11, a21, a31, a21, a22, a32, a31, a32, a33]
    nan, inf = float("nan"), float("inf")
    mats = np.array([sym(0, 0, 0, 0, 0, 0), sym(1, 0, 0, 1, 0, 1), sym(2, 0, 0, 2, 0, 1), sym(1, 1, 1, 1, 1, 1),
                     sym(1e-320, 0, 0, 1e-320, 0, 0), sym(1e308, 1e308, 0, 1e308, 0, 1e308), sym(0, 1, 0, 0, 0, 0),
                     sym(0, 0, 1, 0, 0, 0), sym(0, 0, 0, 0, 1, 0), sym(5, 0, 0, -3, 0, 7), sym(1, 1e-200, 0, 1, 0, 1),
                     sym(3, 0, 4, 0, 0, -3), sym(nan, 0, 0, 1, 0, 1), sym(1, inf, 0, 1, 0, 1), sym(1, 0, 0, nan, 0, 1)],
                    dtype=np.float64)
    w0, z0 = real_dsyev(mats)
    w1, z1, info = ours(mats, 0)
    assert np.array_equal(w0.view(np.uint64), w1.view(np.uint64))
    assert np.array_equal(z0.view(np.uint64), z1.view(np.uint64))


@pytest.mark.gpu
def test_device_instantiation_matches_host_and_real_dsyev():
    mats = covariances(1 << 20, 2)
    w1, z1, i1 = ours(mats, 0)
    w2, z2, i2 = ours(mats, 1)
    assert np.array_equal(i1, i2)
    assert np.array_equal(w1.view(np.uint64), w2.view(np.uint64)), "device eigenvalues differ from the host instantiation"
    assert np.array_equal(z1.view(np.uint64), z2.view(np.uint64)), "device eigenvectors differ from the host instantiation"
    if scipy_openblas() is not None:
        w0, z0 = real_dsyev(mats[:50000])
        assert np.array_equal(w0.view(np.uint64), w2[:50000].view(np.uint64))
        assert np.array_equal(z0.view(np.uint64), z2[:50000].view(np.uint64))
