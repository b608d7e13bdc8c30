"""Seeded synthetic inputs shared by the golden generator, the tests and bench.py."""
from __future__ import annotations

import numpy as np


def uniform_colors(width: int, height: int, seed: int) -> np.ndarray:
    """N x 3 uniform [0,1) f64 ("sRGB by definition")."""
    return np.random.default_rng(seed).random((width * height, 3))


def uniform_colors_slice(width: int, height: int, seed: int, first: int, count: int) -> np.ndarray:
    """Rows [first, first + count) of uniform_colors(width, height, seed) without generating the rest: every f64
    draw consumes one 64-bit output of PCG64, and the bit generator can jump."""
    rng = np.random.default_rng(seed)
    rng.bit_generator.advance(3 * first)
    return rng.random((count, 3))


def image_like_colors(width: int, height: int, seed: int) -> np.ndarray:
    """Smooth gradients + noise, quantised to 8 bits (/255): many duplicate colours and exact ties,
    which uniform noise never produces."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:height, 0:width]
    base = np.stack([xx / max(width, 1), yy / max(height, 1), (xx + yy) / max(width + height, 1)], -1).reshape(-1, 3)
    c = base * 0.8 + 0.2 * rng.random((width * height, 3))
    return np.round(c * 255) / 255


def scene_colors(width: int, height: int, seed: int) -> np.ndarray:
    """A photograph-like test scene for the saliency weights: soft background gradient, two textured objects off the
    border, 8-bit values (/255).  The minimum-barrier distances, the border statistics and the centre prior all have
    something to see (uniform noise has no structure, a pure gradient no objects)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float64)
    u, v = xx / max(width - 1, 1), yy / max(height - 1, 1)
    img = np.stack([0.25 + 0.35 * u, 0.30 + 0.25 * v, 0.55 - 0.25 * u * v], -1)
    for cx, cy, rx, ry, col in ((0.38, 0.45, 0.17, 0.24, (0.85, 0.30, 0.15)), (0.70, 0.62, 0.12, 0.10, (0.10, 0.65, 0.35))):
        inside = ((u - cx) / rx) ** 2 + ((v - cy) / ry) ** 2 < 1.0
        tex = 0.08 * np.sin(40 * u + 25 * v)[..., None]
        img = np.where(inside[..., None], np.asarray(col)[None, None, :] + tex, img)
    img = np.clip(img + 0.04 * rng.standard_normal(img.shape), 0.0, 1.0)
    return (np.round(img * 255) / 255).reshape(-1, 3)


def saliency_like_weights(width: int, height: int, seed: int) -> np.ndarray:
    """w = 1 + 1024 * s^2 in [1, 1025] (mirrors 1 + sal^2 * N / tile^2 of patolette.pyx:313)."""
    rng = np.random.default_rng(seed + 1000)
    yy, xx = np.mgrid[0:height, 0:width]
    s = 0.25 * (1 + np.sin(6 * np.pi * xx / width)) * (1 + np.cos(4 * np.pi * yy / height)) * rng.random((height, width))
    return (1 + 1024 * s ** 2).reshape(-1)


# name -> kwargs.  Small enough for the CPU checkers to finish in seconds.
GOLDEN_CASES = {
    "c1_512_k16_srgb": dict(w=512, h=512, K=16, seed=0, color_space=0, dither=False, kmeans_niter=0),
    "512_k16_ictcp": dict(w=512, h=512, K=16, seed=0, color_space=2, dither=False, kmeans_niter=0),
    "512_k16_cieluv": dict(w=512, h=512, K=16, seed=0, color_space=1, dither=False, kmeans_niter=0),
    "512_k16_ictcp_dither": dict(w=512, h=512, K=16, seed=0, color_space=2, dither=True, kmeans_niter=0),
    "512_k16_ictcp_kmeans10": dict(w=512, h=512, K=16, seed=0, color_space=2, dither=False, kmeans_niter=10),
    "64_k256_ictcp": dict(w=64, h=64, K=256, seed=0, color_space=2, dither=False, kmeans_niter=0),
    "1x1_k4_dither": dict(w=1, h=1, K=4, seed=0, color_space=2, dither=True, kmeans_niter=0),
    "5x3_k64_kmeans": dict(w=5, h=3, K=64, seed=3, color_space=2, dither=True, kmeans_niter=3),
    "4x4_k8_kmeans_seq": dict(w=4, h=4, K=8, seed=4, color_space=0, dither=False, kmeans_niter=3),
    "300x200_k64_weighted": dict(w=300, h=200, K=64, seed=5, color_space=2, dither=True, kmeans_niter=5, weighted=True),
    "300x200_k64_luv_weighted": dict(w=300, h=200, K=64, seed=6, color_space=1, dither=False, kmeans_niter=0, weighted=True),
    "256_k64_imagelike": dict(w=256, h=256, K=64, seed=7, color_space=2, dither=False, kmeans_niter=0, image_like=True),
    "256_k256_imagelike_luv": dict(w=256, h=256, K=256, seed=8, color_space=1, dither=True, kmeans_niter=4,
                                   image_like=True, weighted=True),
    "400x300_k32_palette_only": dict(w=400, h=300, K=32, seed=9, color_space=0, dither=True, kmeans_niter=4,
                                     image_like=True, palette_only=True),
    "333x77_k100_srgb_dither": dict(w=333, h=77, K=100, seed=10, color_space=0, dither=True, kmeans_niter=0),
    "256_k200_fullkmeans": dict(w=256, h=256, K=200, seed=11, color_space=2, dither=False, kmeans_niter=6,
                                kmeans_max_samples=256 * 256),
    "constant_image": dict(w=32, h=32, K=8, seed=12, color_space=2, dither=True, kmeans_niter=0, constant=True),
    "two_colors": dict(w=40, h=25, K=16, seed=13, color_space=2, dither=False, kmeans_niter=2, two_colors=True),
}


# Configurations beyond the goldens, small: restatement vs the reference's own code live on the CPU
# (tests/test_oracle.py) and CUDA vs the oracle on the GPU (tests/test_gpu_big.py).
WIDER = {
    "k1024_weighted_kmeans_dither": dict(w=320, h=240, K=1024, seed=41, color_space=2, dither=True, kmeans_niter=4, weighted=True),
    "gradient_many_gq_cells": dict(w=400, h=300, K=200, seed=42, color_space=1, dither=False, kmeans_niter=0, image_like=True),
    "k_above_pixel_count": dict(w=9, h=7, K=100, seed=43, color_space=2, dither=True, kmeans_niter=2),
    "srgb_kmeans_full_image": dict(w=200, h=150, K=37, seed=44, color_space=0, dither=True, kmeans_niter=6,
                                   kmeans_max_samples=200 * 150),
    "luv_weighted_palette_only": dict(w=256, h=128, K=64, seed=45, color_space=1, dither=False, kmeans_niter=3, weighted=True,
                                      palette_only=True),
    "single_row": dict(w=4096, h=1, K=16, seed=46, color_space=2, dither=True, kmeans_niter=0),
    "k2": dict(w=300, h=300, K=2, seed=47, color_space=2, dither=True, kmeans_niter=1),
}

# BASELINE.json's own configurations (uniform sRGB, seed = config index) and their class at a
# size the CPU reference finishes: frozen as hashes by tests/golden/make_golden_big.py (golden_big.json).
BIG_CASES = {
    "c2_4096_k256_ictcp": dict(w=4096, h=4096, K=256, seed=1, color_space=2, dither=False, kmeans_niter=0),
    "c3_8192_k256_cieluv_dither": dict(w=8192, h=8192, K=256, seed=2, color_space=1, dither=True, kmeans_niter=0),
    "c4_16384_k256_ictcp_kmeans10_dither": dict(w=16384, h=16384, K=256, seed=3, color_space=2, dither=True, kmeans_niter=10),
    "c4_4096_k256_ictcp_kmeans10_dither": dict(w=4096, h=4096, K=256, seed=3, color_space=2, dither=True, kmeans_niter=10),
    "c5_2048_k1024_weighted_fullkmeans_dither": dict(w=2048, h=2048, K=1024, seed=4, color_space=2, dither=True, kmeans_niter=10,
                                                     kmeans_max_samples=2048 * 2048, weighted=True),
    "imagelike_2048_k256_luv_weighted_dither": dict(w=2048, h=2048, K=256, seed=5, color_space=1, dither=True, kmeans_niter=4,
                                                    image_like=True, weighted=True),
    "k8192_1024": dict(w=1024, h=1024, K=8192, seed=6, color_space=2, dither=True, kmeans_niter=2),
    "k70000_512": dict(w=512, h=512, K=70000, seed=7, color_space=2, dither=False, kmeans_niter=2),
}


def make_case(spec: dict):
    """-> (colors, weights | None, quantize kwargs)"""
    w, h = spec["w"], spec["h"]
    if spec.get("constant"):
        colors = np.tile(np.array([[0.25, 0.5, 0.75]]), (w * h, 1))
    elif spec.get("two_colors"):
        rng = np.random.default_rng(spec["seed"])
        pick = rng.integers(0, 2, w * h)
        colors = np.where(pick[:, None] == 0, np.array([[0.1, 0.2, 0.3]]), np.array([[0.9, 0.6, 0.2]]))
    elif spec.get("image_like"):
        colors = image_like_colors(w, h, spec["seed"])
    else:
        colors = uniform_colors(w, h, spec["seed"])
    weights = saliency_like_weights(w, h, spec["seed"]) if spec.get("weighted") else None
    kw = dict(dither=spec["dither"], palette_only=spec.get("palette_only", False), color_space=spec["color_space"],
              kmeans_niter=spec["kmeans_niter"], kmeans_max_samples=spec.get("kmeans_max_samples", 512 ** 2))
    return np.ascontiguousarray(colors, dtype=np.float64), weights, kw


# Saliency (row N3) cases: name -> kwargs.  Frozen from the reference's own wrapper by tests/golden/make_golden_saliency.py.
SALIENCY_CASES = {
    "scene_96x64_t32": dict(w=96, h=64, seed=21, tile=32.0, kind="scene"),
    "scene_37x29_t8": dict(w=37, h=29, seed=22, tile=8.0, kind="scene"),
    "imagelike_70x130_t16": dict(w=70, h=130, seed=23, tile=16.0, kind="image_like"),
    "scene_160x120_t512": dict(w=160, h=120, seed=24, tile=512.0, kind="scene"),
    "uniform_48x40_t10": dict(w=48, h=40, seed=25, tile=10.0, kind="uniform"),
}


def saliency_case_colors(spec) -> np.ndarray:
    f = {"scene": scene_colors, "image_like": image_like_colors, "uniform": uniform_colors}[spec["kind"]]
    return f(spec["w"], spec["h"], spec["seed"])
