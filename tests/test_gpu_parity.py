"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs, against the frozen reference outputs, and - at sizes the oracle cannot
reach in seconds - through size-independent properties.

Bars (BASELINE.json north_star): palette_map bit-exact, palette floats bit-exact (tighter than
the stated 1 ULP), pow/colour transforms bit-exact, dithered maps bit-exact (tighter than the
stated +-1 LSB).  Nothing here reads /root/reference."""
import ctypes as C

import numpy as np
import pytest

from conftest import sha
from synth import GOLDEN_CASES, make_case, uniform_colors, image_like_colors, saliency_like_weights

pytestmark = pytest.mark.gpu


def cuda_quantize(lib, w, h, colors, K, weights=None, **kw):
    from oracle.reflib import quantize_with
    return quantize_with(lib, w, h, colors, K, weights=weights, **kw)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_same_floats(a, b, what):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    nan = np.isnan(a) & np.isnan(b)  # NaN payloads are not comparable across x86 and the GPU
    same = (bits(a) == bits(b)) | nan
    assert same.all(), f"{what}: {int((~same).sum())} of {same.size} values differ, max abs diff " \
                       f"{np.nanmax(np.abs(a - b)[~same])}"


# ---------------------------------------------------------------------------------- pow / colour
@pytest.fixture(scope="module")
def libm_pow(tmp_path_factory):
    """out[i] = pow(x[i], y) through the host libm (numpy's own pow may use SVML, not glibc)."""
    import os, subprocess
    from conftest import ROOT
    so = tmp_path_factory.mktemp("libm") / "libm_helper.so"
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", os.path.join(ROOT, "tests", "native", "libm_helper.c"), "-o", str(so), "-lm"], check=True)
    lib = C.CDLL(str(so))
    lib.libm_pow.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_size_t]
    lib.libm_pow.restype = None

    def run(x, y, out):
        lib.libm_pow(x.ctypes.data, y, out.ctypes.data, x.size)
    return run


@pytest.mark.parametrize("y", [2.4, 1 / 2.4, 0.1593017578125, 78.84375, 1 / 0.1593017578125, 1 / 78.84375, 1 / 3, 3.0])
def test_device_pow_is_bit_exact_with_host_libm(cuda_lib, libm_pow, y):
    rng = np.random.default_rng(int(y * 1000) % 97)
    n = 1 << 20
    x = np.concatenate([rng.random(n // 4), rng.random(n // 4) * 1e4, rng.random(n // 4) * 1e-4,
                        np.exp((rng.random(n // 8) - 0.5) * 1400), -rng.random(n // 16),
                        np.array([0.0, 1.0, np.inf, 5e-324, 1e-310, 1e308])])
    out = np.empty_like(x)
    assert cuda_lib.patolette_b200_pow(x.ctypes.data, y, out.ctypes.data, x.size) == 0
    want = np.empty_like(x)
    libm_pow(x, y, want)
    assert_same_floats(out, want, f"pow(x, {y})")


@pytest.mark.parametrize("which", range(6))
def test_colour_transforms_bit_exact(cuda_lib, oracle, which):
    n = 200_003
    rng = np.random.default_rng(which)
    src = rng.random((n, 3))
    if which == 2:  # ICtCp input: take it from real sRGB colours
        tmp = np.asfortranarray(src); oracle.lib.orc_color_transform(0, tmp.ctypes.data_as(C.c_void_p), C.c_size_t(n)); src = tmp
    if which == 3:
        tmp = np.asfortranarray(src); oracle.lib.orc_color_transform(1, tmp.ctypes.data_as(C.c_void_p), C.c_size_t(n)); src = tmp
    if which == 5:
        src = src * 1.2 - 0.1  # out-of-gamut linear values exercise the clamps
    a = np.asfortranarray(src).copy(order="F"); b = a.copy(order="F")
    oracle.lib.orc_color_transform(which, a.ctypes.data_as(C.c_void_p), C.c_size_t(n))
    assert cuda_lib.patolette_b200_color_transform(which, b.ctypes.data, n) == 0
    assert_same_floats(b, a, f"colour transform {which}")


# ---------------------------------------------------------------------------------- stages
@pytest.mark.parametrize("n,K,weighted,image_like", [(100_003, 32, False, False), (65_536, 64, True, False),
                                                      (50_000, 16, False, True), (3, 8, False, False),
                                                      (200_000, 256, True, True)])
def test_cluster_tree_matches_oracle(cuda_lib, oracle, n, K, weighted, image_like):
    """GQ + LQ on colours already in the quantisation space: identical partition and centres."""
    w = h = None
    side = int(np.ceil(np.sqrt(n)))
    base = image_like_colors(side, side, 3)[:n] if image_like else uniform_colors(side, side, 3)[:n]
    planar = np.asfortranarray(base)
    wts = saliency_like_weights(side, side, 3)[:n].copy() if weighted else None
    res = {}
    for name, lib, fn in (("oracle", oracle.lib, "orc_quantize_clusters"), ("cuda", cuda_lib, "patolette_b200_quantize_clusters")):
        labels = np.zeros(n, dtype=np.uint32); centers = np.zeros((K, 3)); cnt = C.c_size_t(0); gq = C.c_size_t(0)
        f = getattr(lib, fn)
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        rc = f(planar.ctypes.data, n, None if wts is None else wts.ctypes.data, K, labels.ctypes.data, centers.ctypes.data, C.byref(cnt), C.byref(gq))
        assert rc == 0
        res[name] = (labels, centers[:cnt.value].copy(), cnt.value, gq.value)
    assert res["cuda"][2:] == res["oracle"][2:], "cluster counts differ"
    assert np.array_equal(res["cuda"][0], res["oracle"][0]), "cluster membership differs"
    assert_same_floats(res["cuda"][1], res["oracle"][1], "cluster centres")


@pytest.mark.parametrize("kind", ["dyadic", "narrow", "signed", "spiky", "sorted"])
def test_ordered_sums_adversarial(cuda_lib, oracle, kind):
    """Inputs chosen to stress the binade-speculative ordered sums (pb_ordered.cu): exact ties
    (dyadic values), catastrophic cancellation (narrow range), sign changes of the running sum,
    isolated huge terms, and monotone data.  Partition and centres must still be bit-identical."""
    n, K = 150_001, 24
    rng = np.random.default_rng(hash(kind) % 1000)
    if kind == "dyadic":
        c = rng.integers(0, 257, (n, 3)) / 256.0
    elif kind == "narrow":
        c = 0.5 + (rng.random((n, 3)) - 0.5) * 1e-9
    elif kind == "signed":
        c = rng.standard_normal((n, 3)) * np.array([1.0, 1e-3, 50.0])
    elif kind == "spiky":
        c = rng.random((n, 3)) * 1e-3
        c[rng.choice(n, 40, replace=False)] += rng.random((40, 3)) * 1e4
    else:
        c = np.sort(rng.random((n, 3)), axis=0)
    planar = np.asfortranarray(c)
    wts = (1 + 1000 * rng.random(n) ** 4) if kind in ("signed", "spiky") else None
    res = {}
    for name, lib, fn in (("oracle", oracle.lib, "orc_quantize_clusters"), ("cuda", cuda_lib, "patolette_b200_quantize_clusters")):
        labels = np.zeros(n, dtype=np.uint32); centers = np.zeros((K, 3)); cnt = C.c_size_t(0); gq = C.c_size_t(0)
        f = getattr(lib, fn)
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        assert f(planar.ctypes.data, n, None if wts is None else wts.ctypes.data, K, labels.ctypes.data, centers.ctypes.data, C.byref(cnt), C.byref(gq)) == 0
        res[name] = (labels, centers[:cnt.value].copy(), cnt.value, gq.value)
    assert res["cuda"][2:] == res["oracle"][2:]
    assert np.array_equal(res["cuda"][0], res["oracle"][0]), "cluster membership differs"
    assert_same_floats(res["cuda"][1], res["oracle"][1], "cluster centres")


def test_nearest_map_matches_oracle(cuda_lib, oracle):
    n, K = 300_001, 256
    rng = np.random.default_rng(8)
    planar = np.asfortranarray(rng.random((n, 3)))
    pal = rng.random((K, 3)); pal[17] = pal[3]  # duplicate entry: lowest index must win
    a = np.zeros(n, dtype=np.uintp); b = np.zeros(n, dtype=np.uintp)
    oracle.lib.orc_fill_palette_map_nearest(planar.ctypes.data_as(C.c_void_p), C.c_size_t(n), pal.ctypes.data_as(C.c_void_p), C.c_size_t(K), a.ctypes.data_as(C.c_void_p))
    assert cuda_lib.patolette_b200_nearest(planar.ctypes.data, n, pal.ctypes.data, K, b.ctypes.data) == 0
    assert np.array_equal(a, b)
    assert 17 not in b


@pytest.mark.parametrize("kind", ["uniform", "narrow", "lattice_ties", "clustered_far", "degenerate_axis", "outliers_nan"])
def test_nearest_candidate_lists_equal_brute_force(cuda_lib, oracle, kind):
    """The candidate-list 1-NN (pb_nngrid.cu) against the CUDA brute force and the oracle on inputs chosen to
    break a careless pruning rule: a pixel range tiny next to the coordinates, exact ties on a lattice
    (lowest index must win), palette entries far outside the pixels' box, a constant channel, non-finite values."""
    n, K = 400_003, 256
    rng = np.random.default_rng(len(kind))
    if kind == "uniform":
        px = rng.random((n, 3)); pal = rng.random((K, 3))
    elif kind == "narrow":
        px = 0.5 + (rng.random((n, 3)) - 0.5) * 1e-9; pal = 0.5 + (rng.random((K, 3)) - 0.5) * 1e-9
    elif kind == "lattice_ties":
        px = rng.integers(0, 64, (n, 3)) / 64.0          # every pixel sits on the half-way planes of the palette lattice
        pal = (rng.integers(0, 8, (K, 3)) * 8 + 4) / 64.0  # many duplicates and equidistant entries
    elif kind == "clustered_far":
        px = rng.random((n, 3)) * 0.1 + 0.45
        pal = np.concatenate([rng.random((K // 2, 3)) * 50 - 25, rng.random((K - K // 2, 3)) * 0.1 + 0.45])
    elif kind == "degenerate_axis":
        px = rng.random((n, 3)); px[:, 1] = 0.25; pal = rng.random((K, 3))
    else:
        px = rng.random((n, 3)); pal = rng.random((K, 3))
        px[::50_000] = 1e12; px[7] = np.nan; px[8, 2] = -3.0; pal[5] = 1e9
    planar = np.asfortranarray(px)
    got = np.zeros(n, dtype=np.uintp); brute = np.zeros(n, dtype=np.uintp)
    assert cuda_lib.patolette_b200_nearest(planar.ctypes.data, n, pal.ctypes.data, K, got.ctypes.data) == 0
    try:
        assert cuda_lib.patolette_b200_set_option(b"nn_grid", 0) == 0
        assert cuda_lib.patolette_b200_nearest(planar.ctypes.data, n, pal.ctypes.data, K, brute.ctypes.data) == 0
    finally:
        cuda_lib.patolette_b200_set_option(b"nn_grid", 1)
    assert np.array_equal(got, brute), f"{int((got != brute).sum())} pixels differ from the brute force"
    if kind != "outliers_nan":  # NaN ordering is the brute force's own convention; the oracle is compared on finite data
        want = np.zeros(n, dtype=np.uintp)
        oracle.lib.orc_fill_palette_map_nearest(planar.ctypes.data_as(C.c_void_p), C.c_size_t(n), pal.ctypes.data_as(C.c_void_p),
                                                C.c_size_t(K), want.ctypes.data_as(C.c_void_p))
        assert np.array_equal(got, want)


@pytest.mark.parametrize("n,K,weighted,mppc", [(40_000, 64, False, 10_000), (300_000, 32, True, 1024), (16, 8, False, 8192),
                                                 (8, 8, False, 100), (5, 8, False, 100)])
def test_kmeans_matches_oracle(cuda_lib, oracle, n, K, weighted, mppc):
    rng = np.random.default_rng(n)
    x = rng.random((n, 3)).astype(np.float32)
    w = (1 + 50 * rng.random(n)).astype(np.float32) if weighted else None
    c0 = x[rng.choice(n, K, replace=n < K)].copy() + np.float32(0.01)
    ca, cb = c0.copy(), c0.copy()
    f = oracle.lib.orc_kmeans
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    ra = f(x.ctypes.data, n, K, ca.ctypes.data, None if w is None else w.ctypes.data, 5, mppc)
    rb = cuda_lib.patolette_b200_kmeans(x.ctypes.data, n, K, cb.ctypes.data, None if w is None else w.ctypes.data, 5, mppc)
    assert (ra == 0) == (rb == 0)
    assert np.array_equal(ca.view(np.uint32), cb.view(np.uint32)), f"max diff {np.abs(ca - cb).max()}"


@pytest.mark.parametrize("W,H,K", [(64, 64, 16), (100, 37, 40), (1, 7, 4), (130, 257, 256), (1, 1, 3), (700, 500, 8), (512, 384, 2),
                                   (1500, 1100, 64), (2048, 2048, 300)])
def test_dither_matches_oracle(cuda_lib, oracle, W, H, K):
    n = W * H
    rng = np.random.default_rng(W * 1000 + H)
    planar = np.asfortranarray(rng.random((n, 3)))
    if K == 2:  # flat image + two far colours: speculative segments lock on slowly, repairs must kick in
        planar = np.asfortranarray(np.tile([[0.37, 0.52, 0.61]], (n, 1)) + 1e-3 * rng.random((n, 3)))
    pal = rng.random((K, 3))
    a = np.full(n, 7, dtype=np.uintp); b = np.full(n, 7, dtype=np.uintp)
    oracle.lib.orc_dither_riemersma(planar.ctypes.data_as(C.c_void_p), C.c_size_t(W), C.c_size_t(H), pal.ctypes.data_as(C.c_void_p), C.c_size_t(K), a.ctypes.data_as(C.c_void_p))
    assert cuda_lib.patolette_b200_dither(planar.ctypes.data, W, H, pal.ctypes.data, K, b.ctypes.data) == 0
    assert np.array_equal(a, b), f"{int((a != b).sum())} of {n} indices differ, first at {int(np.argmax(a != b))}"


@pytest.mark.parametrize("kind", ["in_gamut", "palette_off_to_one_side", "narrow"])
def test_dither_candidate_lists_equal_brute_force(cuda_lib, oracle, kind):
    """The dither's per-step exact search through the per-cell candidate lists against its own brute force and
    the oracle: a palette that covers the pixels, one that sits off to one side (errors accumulate and the
    queries leave the grid: fallback path), and a pixel range tiny next to the coordinates."""
    W, H, K = 320, 200, 128
    n = W * H
    rng = np.random.default_rng(len(kind) + 40)
    if kind == "in_gamut":
        px = rng.random((n, 3)); pal = rng.random((K, 3))
    elif kind == "palette_off_to_one_side":
        px = rng.random((n, 3)); pal = rng.random((K, 3)) * 0.3
    else:
        px = 0.4 + 1e-7 * rng.random((n, 3)); pal = 0.4 + 1e-7 * rng.random((K, 3))
    planar = np.asfortranarray(px)
    got = np.full(n, 7, dtype=np.uintp); brute = np.full(n, 7, dtype=np.uintp); want = np.full(n, 7, dtype=np.uintp)
    assert cuda_lib.patolette_b200_dither(planar.ctypes.data, W, H, pal.ctypes.data, K, got.ctypes.data) == 0
    try:
        assert cuda_lib.patolette_b200_set_option(b"dither_grid", 0) == 0
        assert cuda_lib.patolette_b200_dither(planar.ctypes.data, W, H, pal.ctypes.data, K, brute.ctypes.data) == 0
    finally:
        cuda_lib.patolette_b200_set_option(b"dither_grid", 1)
    oracle.lib.orc_dither_riemersma(planar.ctypes.data_as(C.c_void_p), C.c_size_t(W), C.c_size_t(H), pal.ctypes.data_as(C.c_void_p),
                                    C.c_size_t(K), want.ctypes.data_as(C.c_void_p))
    assert np.array_equal(got, brute), f"{int((got != brute).sum())} indices differ from the brute-force search"
    assert np.array_equal(got, want), f"{int((got != want).sum())} indices differ from the oracle"
    # the other speculation kernel (one warp per chain instead of four lanes per chain): same map
    warp = np.full(n, 7, dtype=np.uintp)
    try:
        assert cuda_lib.patolette_b200_set_option(b"dither_subwarp", 0) == 0
        assert cuda_lib.patolette_b200_dither(planar.ctypes.data, W, H, pal.ctypes.data, K, warp.ctypes.data) == 0
    finally:
        cuda_lib.patolette_b200_set_option(b"dither_subwarp", 1)
    assert np.array_equal(warp, want), f"{int((warp != want).sum())} indices of the warp-per-chain kernel differ from the oracle"
    # per-pixel permutation kernels and the n / 2048 segment length instead of the tile kernels / one-wave sizing
    for knob in (b"dither_tiles", b"dither_one_wave"):
        other = np.full(n, 7, dtype=np.uintp)
        try:
            assert cuda_lib.patolette_b200_set_option(knob, 0) == 0
            assert cuda_lib.patolette_b200_dither(planar.ctypes.data, W, H, pal.ctypes.data, K, other.ctypes.data) == 0
        finally:
            cuda_lib.patolette_b200_set_option(knob, 1)
        assert np.array_equal(other, want), f"{knob}: {int((other != want).sum())} indices differ from the oracle"


# ---------------------------------------------------------------------------------- end to end
@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_pipeline_matches_golden_and_oracle(cuda_lib, oracle, golden, name):
    spec = GOLDEN_CASES[name]
    colors, weights, kw = make_case(spec)
    code, pal, pmap = cuda_quantize(cuda_lib, spec["w"], spec["h"], colors, spec["K"], weights=weights, **kw)
    ocode, opal, omap = oracle.quantize(spec["w"], spec["h"], colors, spec["K"], weights=weights, **kw)
    assert code == ocode == golden["cases"][name]["exit_code"]
    assert_same_floats(pal, opal, "palette vs oracle")
    if omap is None:
        assert pmap is None
    else:
        assert np.array_equal(pmap, omap), f"{int((pmap != omap).sum())} map entries differ from the oracle"
        assert sha(pmap) == golden["cases"][name]["map_sha256"], "map differs from the frozen reference output"
    if not np.isnan(opal).any():
        assert sha(pal.ravel(order="F")) == golden["cases"][name]["palette_sha256"]


def test_python_quantize_surface_on_gpu(cuda_lib, oracle):
    import patolette_b200 as pb
    colors = uniform_colors(96, 80, 21)
    ok, pal, pmap, msg = pb.quantize(96, 80, colors, 24, dither=False, color_space=pb.ColorSpace_ICtCp, tile_size=0, kmeans_niter=0)
    assert ok and msg == "Quantization successful."
    assert pal.shape == (24, 3) and pal.flags.f_contiguous and pmap.dtype == np.uintp and pmap.shape == (96 * 80,)
    code, opal, omap = oracle.quantize(96, 80, colors, 24, dither=False, color_space=2, kmeans_niter=0)
    assert np.array_equal(pmap, omap)
    assert_same_floats(pal, opal, "palette")
    t = pb.last_timings()
    assert t["launches"] > 0 and t["total"] > 0
    ok, pal2, pmap2, _ = pb.quantize(96, 80, colors, 24, dither=False, tile_size=0, kmeans_niter=0, palette_only=True)
    assert ok and pmap2 is None
    ocode, opal2, _ = oracle.quantize(96, 80, colors, 24, dither=False, kmeans_niter=0, palette_only=True)
    assert_same_floats(pal2, opal2, "palette_only palette (quantisation space, bug B2)")


def test_unused_palette_rows_are_minus_one(cuda_lib):
    colors = uniform_colors(5, 3, 1)
    code, pal, pmap = cuda_quantize(cuda_lib, 5, 3, colors, 64, dither=False, color_space=2, kmeans_niter=0)
    assert code == 0 and (pal[15:] == -1).all() and (pal[:15] != -1).any() and pmap.max() < 15


def _cluster_tree(lib, fn, planar, n, K, wts=None):
    labels = np.zeros(n, dtype=np.uint32); centers = np.zeros((K, 3)); cnt = C.c_size_t(0); gq = C.c_size_t(0)
    f = getattr(lib, fn)
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    assert f(planar.ctypes.data, n, None if wts is None else wts.ctypes.data, K, labels.ctypes.data, centers.ctypes.data,
             C.byref(cnt), C.byref(gq)) == 0
    return labels, centers[:cnt.value].copy(), cnt.value, gq.value


@pytest.mark.parametrize("route", ["default", "split_exact", "split_redo_all", "prefix_5_slabs", "prefix_one_cta", "fused", "no_raw_moments", "gq_chain_warp", "ord_gather", "scatter_warp", "classic_summary", "no_term_dump", "one_slot", "no_overlap", "overlap"])
def test_cluster_tree_1p5M_every_route_matches_oracle(cuda_lib, oracle, route):
    """1.5 M pixels, K=48 (clusters long enough for the group records and the two-level resolve, hovering
    off-diagonal sums, two-parity records): the ordered-sum machinery has several routes to the same bits -
    replays from the term dump or from the planes (dump exhausted), half-batches on two streams or one.
    Every route must give the oracle's partition and centres bit for bit."""
    n, K = 1_500_000, 48
    side = int(np.ceil(np.sqrt(n)))
    planar = np.asfortranarray(image_like_colors(side, side, 17)[:n] * 0.5 + 0.5 * uniform_colors(side, side, 18)[:n])
    want = _cluster_tree(oracle.lib, "orc_quantize_clusters", planar, n, K)
    opts = {"default": [], "split_exact": [(b"split_certify", 0)], "split_redo_all": [(b"split_certify", 2)], "prefix_5_slabs": [(b"prefix_slabs", 5)], "prefix_one_cta": [(b"prefix_slabs", 0)], "fused": [(b"fused_pass", 1)], "no_raw_moments": [(b"raw_moments", 0)], "gq_chain_warp": [(b"gq_chain_cta", 0)], "ord_gather": [(b"sorted_payload", 0)], "scatter_warp": [(b"scatter_cta", 0)], "classic_summary": [(b"fast_summary", 0)], "no_term_dump": [(b"dump_cap", 0)], "one_slot": [(b"dump_cap", 1)],
            "no_overlap": [(b"overlap", 0)], "overlap": [(b"overlap", 1)]}[route]
    try:
        for k, v in opts:
            assert cuda_lib.patolette_b200_set_option(k, v) == 0
        got = _cluster_tree(cuda_lib, "patolette_b200_quantize_clusters", planar, n, K)
    finally:
        cuda_lib.patolette_b200_set_option(b"dump_cap", -1)
        cuda_lib.patolette_b200_set_option(b"overlap", -1)
        cuda_lib.patolette_b200_set_option(b"split_certify", 1)
        cuda_lib.patolette_b200_set_option(b"prefix_slabs", 1)
        cuda_lib.patolette_b200_set_option(b"fast_summary", 1)
        cuda_lib.patolette_b200_set_option(b"fused_pass", 0)
        cuda_lib.patolette_b200_set_option(b"raw_moments", 1)
        cuda_lib.patolette_b200_set_option(b"gq_chain_cta", 1)
        cuda_lib.patolette_b200_set_option(b"sorted_payload", 1)
        cuda_lib.patolette_b200_set_option(b"scatter_cta", 1)
    assert got[2:] == want[2:], "cluster counts differ"
    assert np.array_equal(got[0], want[0]), "cluster membership differs"
    assert_same_floats(got[1], want[1], "cluster centres")


def _split_counts(lib, reset=False):
    c = (C.c_ulonglong * 4)()
    assert lib.patolette_b200_split_counts(c, 1 if reset else 0) == 0
    return {"certified": c[0], "refused": c[1], "redone": c[2]}


@pytest.mark.parametrize("kind", ["two_blobs", "mirror", "few_colours", "constant", "two_pixels_per_bucket", "weighted_risky",
                                  "weighted_integer", "weighted_small", "huge_offset"])
def test_split_certificate_adversarial(cuda_lib, oracle, kind):
    """The certified route (pb_certify.cu) finds a split's optimal bucket from unordered per-bucket sums and proves
    that the reference's first maximum is the same; whatever it cannot prove is re-evaluated exactly.  Inputs chosen
    against the proof: plateaus of empty buckets between two blobs (the FIRST maximum must win), mirrored data (near
    ties of the objective), a handful of distinct colours (rows of equal buckets: the warp reductions), a constant
    image (round-robin buckets), weights whose fraction makes size_t += double round up, integer weights, weights
    below 1, and colours far from the origin (the error bound scales with max |c|).  Partition and centres must be
    the oracle's bit for bit, whichever route each cluster took."""
    n, K = 120_000, 40
    rng = np.random.default_rng(len(kind) * 7 + 1)
    wts = None
    if kind == "two_blobs":
        c = np.where(rng.random((n, 1)) < 0.4, 0.1, 0.8) + rng.random((n, 3)) * 0.02
    elif kind == "mirror":
        half = rng.random((n // 2, 3)) * 0.5
        c = np.concatenate([half, 1.0 - half])
    elif kind == "few_colours":
        pal = rng.integers(0, 256, (6, 3)) / 255.0
        c = pal[rng.integers(0, 6, n) * (rng.random(n) < 0.97)]
        c[::1000] += rng.random((len(c[::1000]), 3)) * 1e-3
    elif kind == "constant":
        c = np.full((n, 3), 0.25)
    elif kind == "two_pixels_per_bucket":
        n, K = 1024, 16
        c = np.repeat(np.linspace(0.0, 1.0, 512), 2)[:, None] * np.array([1.0, 0.5, 0.25]) + rng.random((1024, 3)) * 1e-6
    elif kind == "weighted_risky":
        c = rng.random((n, 3))
        wts = np.floor(1 + 50 * rng.random(n)) + np.where(rng.random(n) < 0.3, 1.0 - 2.0 ** -rng.integers(20, 53, n), 0.5)
    elif kind == "weighted_integer":
        c = rng.random((n, 3))
        wts = np.floor(1 + 20 * rng.random(n) ** 3)
    elif kind == "weighted_small":
        c = rng.random((n, 3))
        wts = rng.random(n) * 1.5  # some below 1: floor(w) = 0, sizes may stay 0 (local.c:157-163 branches)
    else:
        c = 1e6 + rng.random((n, 3))
    planar = np.asfortranarray(c)
    want = _cluster_tree(oracle.lib, "orc_quantize_clusters", planar, n, K, wts)
    _split_counts(cuda_lib, reset=True)
    got = _cluster_tree(cuda_lib, "patolette_b200_quantize_clusters", planar, n, K, wts)
    counts = _split_counts(cuda_lib)
    assert got[2:] == want[2:], f"cluster counts differ ({counts})"
    assert np.array_equal(got[0], want[0]), f"cluster membership differs ({counts})"
    assert_same_floats(got[1], want[1], "cluster centres")
    assert counts["redone"] == counts["refused"], counts


def test_split_certificate_is_the_common_route(cuda_lib):
    """On the bench's kind of data (uniform noise) nearly every cluster is certified; with every certificate refused
    ("split_certify" = 2) every cluster is evaluated twice and the result is the same."""
    W = H = 1024; K = 64
    colors = uniform_colors(W, H, 5)
    _split_counts(cuda_lib, reset=True)
    code, pal, pmap = cuda_quantize(cuda_lib, W, H, colors, K, dither=False, color_space=2, kmeans_niter=0)
    c1 = _split_counts(cuda_lib, reset=True)
    assert code == 0 and c1["certified"] >= 100 and c1["refused"] <= c1["certified"] // 10, c1
    try:
        assert cuda_lib.patolette_b200_set_option(b"split_certify", 2) == 0
        code2, pal2, pmap2 = cuda_quantize(cuda_lib, W, H, colors, K, dither=False, color_space=2, kmeans_niter=0)
        c2 = _split_counts(cuda_lib, reset=True)
    finally:
        cuda_lib.patolette_b200_set_option(b"split_certify", 1)
    assert code2 == 0 and c2["certified"] == 0 and c2["redone"] == c2["refused"] > 100, c2
    assert np.array_equal(pmap, pmap2) and np.array_equal(bits(pal), bits(pal2))


# ---------------------------------------------------------------------------------- large-size properties
def test_config2_scale_properties(cuda_lib):
    """2048 x 2048, K=256, ICtCp, dither off (BASELINE config 2 at a quarter of the pixels): size-independent
    invariants the domain offers.  (1) the map is idempotent under re-quantisation of the palette image in the NN
    space: every pixel is assigned to its exact nearest palette entry (checked with an f64 brute force on a sample);
    (2) determinism: two runs give identical bits; (3) all K entries are used on uniform noise."""
    W = H = 2048; K = 256
    colors = uniform_colors(W, H, 2)
    code, pal, pmap = cuda_quantize(cuda_lib, W, H, colors, K, dither=False, color_space=2, kmeans_niter=0)
    assert code == 0
    code2, pal2, pmap2 = cuda_quantize(cuda_lib, W, H, colors, K, dither=False, color_space=2, kmeans_niter=0)
    assert np.array_equal(pmap, pmap2) and np.array_equal(bits(pal), bits(pal2))
    assert len(np.unique(pmap)) == K and pmap.max() == K - 1
    assert ((pal >= 0) & (pal <= 1)).all()
    # reconstruct in ICtCp on the device and verify nearest-ness on a 20k sample
    idx = np.random.default_rng(0).choice(W * H, 20_000, replace=False)
    sample = np.asfortranarray(colors[idx]); assert cuda_lib.patolette_b200_color_transform(0, sample.ctypes.data, idx.size) == 0
    p = np.asfortranarray(pal.copy()); assert cuda_lib.patolette_b200_color_transform(0, p.ctypes.data, K) == 0
    d = ((sample[:, None, :] - p[None, :, :]) ** 2).sum(-1)
    best = d.min(1)
    chosen = d[np.arange(idx.size), pmap[idx]]
    # the returned palette went ICtCp -> sRGB -> ICtCp, so allow the round-trip's rounding in the comparison
    assert (chosen <= best * (1 + 1e-9) + 1e-18).all()


def test_dither_kmeans_1024_matches_oracle(cuda_lib, oracle):
    """1024 x 1024, K=256, CIELuv, KMeans 3 iterations on the default subsample, Riemersma dither:
    every stage of BASELINE config 3 at a size the CPU oracle finishes in seconds - bit-exact."""
    W = H = 1024; K = 256
    colors = image_like_colors(W, H, 31)
    kw = dict(dither=True, color_space=1, kmeans_niter=3)
    code, pal, pmap = cuda_quantize(cuda_lib, W, H, colors, K, **kw)
    ocode, opal, omap = oracle.quantize(W, H, colors, K, **kw)
    assert code == ocode == 0
    assert_same_floats(pal, opal, "palette")
    assert np.array_equal(pmap, omap), f"{int((pmap != omap).sum())} dithered indices differ"


def test_config3_scale_8192_properties(cuda_lib):
    """8192 x 8192 (BASELINE config 3/4 pixel count class), K=256, ICtCp, dither off: the pipeline runs at
    67 M pixels (u32 positions, packed scratch tables, speculative batches), is deterministic, uses all K
    entries, and agrees with itself when the image is fed as weights == 1 (the weighted code path must give
    the same partition as the unweighted one: w*c == c, sum of ones == n)."""
    W = H = 8192; K = 256
    colors = uniform_colors(W, H, 3)
    code, pal, pmap = cuda_quantize(cuda_lib, W, H, colors, K, dither=False, color_space=2, kmeans_niter=0)
    assert code == 0 and len(np.unique(pmap)) == K
    ones = np.ones(W * H)
    code2, pal2, pmap2 = cuda_quantize(cuda_lib, W, H, colors, K, weights=ones, dither=False, color_space=2, kmeans_niter=0)
    assert code2 == 0
    assert np.array_equal(pmap, pmap2)
    assert np.array_equal(bits(pal), bits(pal2))
