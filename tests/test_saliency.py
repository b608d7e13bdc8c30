"""Row N3: the saliency weights of the reference's Python wrapper (src/patolette/patolette.pyx:54-313).

CPU part: oracle/saliency_port.py (numpy restatement) against tests/golden/golden_saliency.npz - outputs of the
reference's OWN compiled wrapper (tests/golden/make_golden_saliency.py) - and, where the compiled wrapper is present
(oracle/_ref/pyx, built in place from /root/reference), against it live.
GPU part: pb_saliency.cu through the C ABI against the same goldens, the port and the live wrapper: the distance map
bit for bit, the weights to WEIGHT_RTOL."""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from synth import SALIENCY_CASES, saliency_case_colors, scene_colors, image_like_colors  # noqa: E402

# The weights go through pow / cbrt / exp / sqrt and sums whose order differs between numpy (pairwise, BLAS) and a GPU
# (per-CTA partials): relative agreement, not bit parity.  Observed: <= 1e-13; the bound leaves two orders of margin.
WEIGHT_RTOL = 1e-11
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_saliency.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def ref_wrapper():
    """The reference's own compiled wrapper, or None where it was never built (it cannot be built without /root/reference)."""
    from oracle.ref_build import build_ref_pyx
    if not os.path.exists(build_ref_pyx.so_path()) and not build_ref_pyx.available():
        return None
    try:
        return build_ref_pyx.load()
    except Exception:  # noqa: BLE001 - e.g. scipy missing on some box: the goldens still pin the port
        return None


# ---------------------------------------------------------------------------------------------------------- CPU

@pytest.mark.parametrize("name", list(SALIENCY_CASES))
def test_port_matches_reference_goldens(gold, name):
    from oracle import saliency_port as sp
    spec = SALIENCY_CASES[name]
    img = saliency_case_colors(spec).reshape(spec["h"], spec["w"], 3)
    d = sp.mbd(np.mean(img, axis=2).astype(np.float32), 3)
    assert np.array_equal(d.view(np.uint32), gold[name + "/mbd"].view(np.uint32)), "distance map differs from the reference's"
    w = sp.get_weights(img, spec["tile"], sal=d)
    np.testing.assert_allclose(w, gold[name + "/weights"], rtol=1e-12, atol=0)


def test_port_matches_live_reference_wrapper():
    ref = ref_wrapper()
    if ref is None:
        pytest.skip("compiled reference wrapper not present (no /root/reference on this box)")
    from oracle import saliency_port as sp
    for (w, h, seed, tile) in [(53, 41, 31, 12.0), (24, 90, 32, 64.0)]:
        img = scene_colors(w, h, seed).reshape(h, w, 3)
        m = np.mean(img, axis=2).astype(np.float32)
        assert np.array_equal(sp.mbd(m.copy(), 3), np.asarray(ref.mbd(m.copy(), 3)))
        np.testing.assert_allclose(sp.get_weights(img, tile), np.asarray(ref.get_weights(img, tile)), rtol=1e-12, atol=0)


def _skewed_tile_scan(img, L, U, D, inverse):
    """The schedule of k_mbd_pass (pb_saliency.cu) in plain Python: 32-row groups, skewed 32 x 32 tiles (tile k, step j,
    lane t -> column 32 k + j - t), left neighbour in a per-lane register, upper neighbour = the lane above's previous
    step, lane 0's upper neighbours from the row above the group.  Groups run one after the other here (any order
    that respects the hand-off gives the same result)."""
    f32 = np.float32
    rows, cols = img.shape
    R, Cn = (rows - 3, cols - 3) if inverse else (rows - 2, cols - 2)
    if R <= 0 or Cn <= 0:
        return
    row_of = (lambda i: rows - 2 - i) if inverse else (lambda i: 1 + i)
    col_of = (lambda c: cols - 2 - c) if inverse else (lambda c: 1 + c)
    for g in range((R + 31) // 32):
        uleft, lleft, myU, myL = [f32(0)] * 32, [f32(0)] * 32, [f32(0)] * 32, [f32(0)] * 32
        for t in range(32):
            if g * 32 + t < R:
                y0 = cols - 1 if inverse else 0
                uleft[t], lleft[t] = U[row_of(g * 32 + t), y0], L[row_of(g * 32 + t), y0]
        uprow = row_of(g * 32) + 1 if inverse else row_of(g * 32) - 1
        for k in range((Cn + 62) // 32):
            tile = {}
            for r in range(32):  # tile in
                for lane in range(32):
                    ir, c = g * 32 + r, 32 * k + lane - r
                    if ir < R and 0 <= c < Cn:
                        p = (row_of(ir), col_of(c))
                        tile[r, lane] = [img[p], D[p], U[p], L[p]]
            up = {lane: (U[uprow, col_of(32 * k + lane)], L[uprow, col_of(32 * k + lane)]) for lane in range(32) if 32 * k + lane < Cn}
            for j in range(32):
                shU, shL = [f32(0)] + myU[:-1], [f32(0)] + myL[:-1]
                for lane in range(32):
                    c = 32 * k + j - lane
                    if not (g * 32 + lane < R and 0 <= c < Cn):
                        continue
                    upU, upL = up[j] if lane == 0 else (shU[lane], shL[lane])
                    ix, d, cu, cl = tile[lane, j]
                    u1, l1, u2, l2 = max(upU, ix), min(upL, ix), max(uleft[lane], ix), min(lleft[lane], ix)
                    b1, b2 = f32(u1 - l1), f32(u2 - l2)
                    keep, use1 = d <= b1 and d <= b2, b1 < d and b1 <= b2
                    nD, nU, nL = (d, cu, cl) if keep else ((b1, u1, l1) if use1 else (b2, u2, l2))
                    tile[lane, j] = [ix, nD, nU, nL]
                    uleft[lane], lleft[lane], myU[lane], myL[lane] = nU, nL, nU, nL
            for (r, lane), (ix, d, u, l) in tile.items():  # tile out
                p = (row_of(g * 32 + r), col_of(32 * k + lane - r))
                D[p], U[p], L[p] = d, u, l


@pytest.mark.parametrize("shape", [(70, 45), (36, 5), (33, 67)])
def test_skewed_tile_schedule_equals_the_raster_scans(shape):
    """The wavefront schedule the CUDA kernel uses visits every pixel after its two neighbours: same distance map as the
    port's literal raster scans (row counts around a 32-row group, column counts around a tile)."""
    from oracle import saliency_port as sp
    rng = np.random.default_rng(shape[0])
    img = np.cumsum(rng.normal(size=shape), axis=1).astype(np.float32)
    L, U = img.copy(), img.copy()
    D = np.full(img.shape, np.inf, dtype=np.float32)
    D[0, :] = 0; D[-1, :] = 0; D[:, 0] = 0; D[:, -1] = 0
    for it in range(3):
        _skewed_tile_scan(img, L, U, D, it % 2 == 0)
    assert np.array_equal(D, sp.mbd(img, 3))


def test_tile_size_validation_without_gpu():
    import patolette_b200 as pb
    assert pb.quantize(2, 2, np.zeros((4, 3)), 2, tile_size=-1)[3] == pb.bad_tile_size
    with pytest.raises(ValueError):
        pb.saliency_weights(4, 4, np.zeros((16, 3)), tile_size=0)
    msg = pb._lib.load().get_patolette_exit_code_info_message(-7)
    assert msg.startswith(b"Saliency weights")


# ---------------------------------------------------------------------------------------------------------- GPU

@pytest.mark.gpu
@pytest.mark.parametrize("name", list(SALIENCY_CASES))
def test_gpu_saliency_matches_reference_goldens(gold, name):
    import patolette_b200 as pb
    spec = SALIENCY_CASES[name]
    colors = saliency_case_colors(spec)
    d = pb.saliency_mbd(spec["w"], spec["h"], colors)
    assert np.array_equal(d.view(np.uint32), gold[name + "/mbd"].view(np.uint32)), "distance map differs from the reference's"
    w = pb.saliency_weights(spec["w"], spec["h"], colors, spec["tile"])
    np.testing.assert_allclose(w, gold[name + "/weights"], rtol=WEIGHT_RTOL, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1000, 700), (33, 2100), (2100, 37), (64, 64), (4, 4), (5, 4), (129, 97)])
def test_gpu_mbd_bit_exact_multi_warp(shape):
    """Row counts around the 32-row warp groups and column counts around the 32-column skewed tiles, against the
    reference's compiled scans where present (fast), else the numpy port on a crop-sized image."""
    import patolette_b200 as pb
    w, h = shape
    ref = ref_wrapper()
    if ref is None and w * h > 20000:
        w, h = max(4, w // 8), max(4, h // 8)
    colors = scene_colors(w, h, 40 + w % 7) if min(w, h) > 8 else image_like_colors(w, h, 41)
    m = np.mean(colors.reshape(h, w, 3), axis=2).astype(np.float32)
    if ref is not None:
        want = np.asarray(ref.mbd(m.copy(), 3))
    else:
        from oracle import saliency_port as sp
        want = sp.mbd(m.copy(), 3)
    got = pb.saliency_mbd(w, h, colors)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_gpu_weights_match_live_reference_wrapper_at_size():
    ref = ref_wrapper()
    if ref is None:
        pytest.skip("compiled reference wrapper not present")
    import patolette_b200 as pb
    w, h, tile = 1024, 768, 512.0
    colors = scene_colors(w, h, 50)
    want = np.asarray(ref.get_weights(colors.reshape(h, w, 3), tile))
    got = pb.saliency_weights(w, h, colors, tile)
    np.testing.assert_allclose(got, want, rtol=WEIGHT_RTOL, atol=0)


@pytest.mark.gpu
def test_gpu_quantize_with_tile_size_is_the_weighted_path(gold):
    """quantize(tile_size = T) == quantize(weights = saliency_weights(T)) bit for bit (f64, interleaved and uint8
    ingest), and the CPU oracle given the same weights returns the same palette and map."""
    import patolette_b200 as pb
    from oracle.reflib import OracleLib
    spec = SALIENCY_CASES["scene_160x120_t512"]
    w, h, tile = spec["w"], spec["h"], 24.0
    colors = saliency_case_colors(spec)
    wts = pb.saliency_weights(w, h, colors, tile)
    kw = dict(dither=True, color_space=pb.ColorSpace_ICtCp, kmeans_niter=4)
    ok, pal, pmap, msg = pb.quantize(w, h, colors, 32, tile_size=tile, **kw)
    assert ok, msg
    assert pb.last_timings()["saliency"] > 0
    ok2, pal2, pmap2, _ = pb.quantize(w, h, colors, 32, tile_size=0, weights=wts, **kw)
    assert ok2 and np.array_equal(pmap, pmap2) and np.array_equal(pal.view(np.uint64), pal2.view(np.uint64))
    ok3, pal3, pmap3, _ = pb.quantize(w, h, np.asfortranarray(colors), 32, tile_size=tile, **kw)  # planar ingest
    assert ok3 and np.array_equal(pmap, pmap3) and np.array_equal(pal.view(np.uint64), pal3.view(np.uint64))
    u8 = np.round(colors * 255).astype(np.uint8)
    ok4, pal4, pmap4, _ = pb.quantize_u8(w, h, u8, 32, tile_size=tile, **kw)
    assert ok4 and np.array_equal(pmap, pmap4.astype(np.uintp)) and np.array_equal(pal.view(np.uint64), pal4.view(np.uint64))
    code, opal, omap = OracleLib().quantize(w, h, colors, 32, dither=True, color_space=2, kmeans_niter=4, weights=wts)
    assert code == 0 and np.array_equal(pmap, omap) and np.array_equal(pal.view(np.uint64), opal.view(np.uint64))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["scene_96x64_t32", "scene_160x120_t512"])
def test_gpu_quantize_with_tile_size_against_reference_wrapper_output(gold, name):
    """End to end against the reference's own quantize(..., tile_size): the weights agree to ~1e-13, so palettes
    agree far below a colour step and the maps almost everywhere (not bit for bit: see WEIGHT_RTOL)."""
    import patolette_b200 as pb
    spec = SALIENCY_CASES[name]
    kw = {"scene_96x64_t32": dict(K=16, dither=False, kmeans_niter=0),
          "scene_160x120_t512": dict(K=32, dither=True, kmeans_niter=4)}[name]
    colors = saliency_case_colors(spec)
    ok, pal, pmap, msg = pb.quantize(spec["w"], spec["h"], colors, kw["K"], dither=kw["dither"], color_space=pb.ColorSpace_ICtCp,
                                     tile_size=spec["tile"], kmeans_niter=kw["kmeans_niter"])
    assert ok, msg
    np.testing.assert_allclose(pal, gold[name + "/palette"], rtol=0, atol=1e-6)
    assert np.mean(pmap == gold[name + "/map"].astype(np.uintp)) > 0.995


@pytest.mark.gpu
def test_gpu_saliency_rejects_what_the_reference_raises_on():
    import patolette_b200 as pb
    with pytest.raises(ValueError):  # a side <= 3: mbd() returns None and get_weights raises (patolette.pyx:154-155)
        pb.saliency_weights(3, 50, np.random.default_rng(0).random((150, 3)), 8.0)
    with pytest.raises(ValueError):  # 0.1 * sqrt(9 * 9) < 1: empty border strips
        pb.quantize(9, 9, np.random.default_rng(0).random((81, 3)), 4, tile_size=8.0)
    with pytest.raises(ValueError):  # 12 x 4000: the strips (21 wide) do not fit 12 rows, the reference's reshape raises
        pb.saliency_weights(4000, 12, np.random.default_rng(0).random((48000, 3)), 8.0)
    ok, pal, pmap, msg = pb.quantize(16, 16, np.full((256, 3), 0.5), 4, tile_size=8.0)  # singular strip covariance
    assert not ok and pal is None
