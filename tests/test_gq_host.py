"""CPU suite: the host side of GQ - Wu's dynamic programme over the 512 buckets (global.c:189-298) - runs its
n loop on a few host threads.  Every thread count must return the same cuts (each n keeps the reference's own
scan over t), on tables that drive the programme through all of its iterations."""
import ctypes as C
import time

import numpy as np
import pytest

from synth import image_like_colors, uniform_colors


def bucket_table(colors):
    """per-bucket sums the way the GPU hands them to the host: 512 x 10 doubles + 513 class starts"""
    c = np.asarray(colors, dtype=np.float64)
    d = c - c.mean(0)
    w, v = np.linalg.eigh(d.T @ d)
    dots = c @ v[:, -1]
    b = np.minimum((512 * ((dots - dots.min()) / (dots.max() - dots.min()))).astype(np.int64), 511)
    order = np.argsort(b, kind="stable")
    cs = np.searchsorted(b[order], np.arange(513)).astype(np.uint32)
    hs = np.zeros((512, 10))
    terms = np.stack([c[:, 0], c[:, 1], c[:, 2], (c ** 2).sum(1), c[:, 0] * c[:, 0], c[:, 0] * c[:, 1], c[:, 1] * c[:, 1],
                      c[:, 0] * c[:, 2], c[:, 1] * c[:, 2], c[:, 2] * c[:, 2]], 1)
    np.add.at(hs, b, terms)
    return np.ascontiguousarray(hs), cs


@pytest.mark.parametrize("kind,K", [("image_like", 256), ("image_like", 7), ("uniform", 256), ("two_blobs", 64), ("gradient", 1024),
                                     ("gradient", 5), ("gradient", 12), ("gradient", 600)])
def test_gq_cuts_do_not_depend_on_the_thread_count(kind, K):
    from patolette_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(len(kind) + K)
    if kind == "image_like":
        colors = image_like_colors(400, 300, 5)
    elif kind == "uniform":
        colors = uniform_colors(400, 300, 6)
    elif kind == "two_blobs":
        colors = np.concatenate([0.2 + 0.05 * rng.standard_normal((60_000, 3)), 0.8 + 0.02 * rng.standard_normal((40_000, 3))])
    else:
        t = np.linspace(0, 1, 150_000)[:, None]
        colors = t * np.array([[1.0, 0.6, 0.2]]) + 0.01 * rng.random((150_000, 3))
    hs, cs = bucket_table(colors)
    res, ms = {}, {}
    try:
        for threads in (1, 2, 3, 8):
            assert lib.patolette_b200_set_option(b"gq_threads", threads) == 0
            q = np.zeros(16, dtype=np.uintp)
            t0 = time.perf_counter()
            cells = lib.patolette_b200_gq_cuts(hs.ctypes.data, cs.ctypes.data, K, q.ctypes.data)
            ms[threads] = (time.perf_counter() - t0) * 1e3
            res[threads] = (cells, q.tolist())
    finally:
        lib.patolette_b200_set_option(b"gq_threads", 0)
    assert res[1][0] >= 1 and res[1][1][0] == 0 and res[1][1][res[1][0]] == 512
    assert all(np.diff(res[1][1][:res[1][0] + 1]) > 0), "cuts must ascend"
    for threads in (2, 3, 8):
        assert res[threads] == res[1], f"{threads} threads: {res[threads]} != {res[1]}"
    try:  # the reference's full (max(K, 512) + 1)^2 table against the 513 x 13 one used by default
        assert lib.patolette_b200_set_option(b"gq_full_table", 1) == 0
        assert lib.patolette_b200_set_option(b"gq_threads", 1) == 0
        q = np.zeros(16, dtype=np.uintp)
        cells = lib.patolette_b200_gq_cuts(hs.ctypes.data, cs.ctypes.data, K, q.ctypes.data)
        assert (cells, q.tolist()) == res[1]
    finally:
        lib.patolette_b200_set_option(b"gq_full_table", 0)
        lib.patolette_b200_set_option(b"gq_threads", 0)
    print(f"{kind} K={K}: {res[1][0]} cells, ms by threads {ms}")


def test_host_pool_survives_many_alternating_calls():
    """The parked worker threads are reused across calls and thread counts (no lost wake-ups, no stale jobs)."""
    from patolette_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(3)
    t = np.linspace(0, 1, 60_000)[:, None]
    hs, cs = bucket_table(t * np.array([[0.9, 0.5, 0.3]]) + 0.02 * rng.random((60_000, 3)))
    want = None
    try:
        for it in range(300):
            assert lib.patolette_b200_set_option(b"gq_threads", 1 + (it * 7) % 9) == 0
            q = np.zeros(16, dtype=np.uintp)
            cells = lib.patolette_b200_gq_cuts(hs.ctypes.data, cs.ctypes.data, 256, q.ctypes.data)
            got = (cells, q.tolist())
            want = want or got
            assert got == want, it
    finally:
        lib.patolette_b200_set_option(b"gq_threads", 0)
