"""TEST INFRASTRUCTURE ONLY - numpy restatement of the reference's saliency weights (row N3).

Follows src/patolette/patolette.pyx:54-313 (`raster_scan`, `raster_scan_inv`, `mbd`, `get_weights`) statement by
statement.  **Parity unpinned**: the reference calls `skimage.color.rgb2lab`, and scikit-image is not installed in this
image (nor vendored under /root/reference), so the reference's own weights cannot be produced here.  `rgb2lab` below
restates scikit-image's published algorithm (skimage/color/colorconv.py: `rgb2xyz` - sRGB companding with the 0.04045 /
12.92 / 2.4 constants and the `xyz_from_rgb` matrix - and `xyz2lab` with the D65 / 2-degree white point (0.95047, 1,
1.08883), the 0.008856 threshold, 7.787 x + 16/116 below it and cbrt above).  The product (pb_saliency.cu) is compared
with this file within a floating-point tolerance, the MBD distance map bit for bit (it only uses float32 min / max /
subtract, which have one result).

Only tests/ may import this module."""
from __future__ import annotations

from math import exp, floor, sqrt

import numpy as np

XYZ_FROM_RGB = np.array([[0.412453, 0.357580, 0.180423],
                         [0.212671, 0.715160, 0.072169],
                         [0.019334, 0.119193, 0.950227]])
WHITE_D65_2 = np.array([0.95047, 1.0, 1.08883])


def rgb2lab(rgb: np.ndarray) -> np.ndarray:
    """skimage.color.rgb2lab(rgb) for float input in [0, 1] (illuminant D65, observer 2)."""
    arr = np.array(rgb, dtype=np.float64, copy=True)
    mask = arr > 0.04045
    arr[mask] = np.power((arr[mask] + 0.055) / 1.055, 2.4)
    arr[~mask] /= 12.92
    xyz = arr @ XYZ_FROM_RGB.T
    arr = xyz / WHITE_D65_2
    mask = arr > 0.008856
    arr[mask] = np.cbrt(arr[mask])
    arr[~mask] = 7.787 * arr[~mask] + 16.0 / 116.0
    x, y, z = arr[..., 0], arr[..., 1], arr[..., 2]
    return np.stack([116.0 * y - 16.0, 500.0 * (x - y), 200.0 * (y - z)], axis=-1)


def _scan(img, L, U, D, inverse: bool):
    """patolette.pyx:54-151: one raster scan (x, y ascending from 1) or inverse scan (descending, stopping ABOVE 1)."""
    rows, cols = img.shape
    if inverse:
        xs, ys, dx = range(rows - 2, 1, -1), range(cols - 2, 1, -1), 1
    else:
        xs, ys, dx = range(1, rows - 1), range(1, cols - 1), -1
    f32 = np.float32
    for x in xs:
        for y in ys:
            ix, d = img[x, y], D[x, y]
            u1, l1 = U[x + dx, y], L[x + dx, y]
            u2, l2 = U[x, y + dx], L[x, y + dx]
            b1 = f32(max(u1, ix) - min(l1, ix))
            b2 = f32(max(u2, ix) - min(l2, ix))
            if d <= b1 and d <= b2:
                continue
            if b1 < d and b1 <= b2:
                D[x, y], U[x, y], L[x, y] = b1, max(u1, ix), min(l1, ix)
            else:
                D[x, y], U[x, y], L[x, y] = b2, max(u2, ix), min(l2, ix)


def mbd(img: np.ndarray, iters: int = 3):
    """patolette.pyx:153-201 (pure Python loops: use small images)."""
    if img.shape[0] <= 3 or img.shape[1] <= 3:
        return None
    img = img.astype(np.float32)
    L, U = img.copy(), img.copy()
    D = np.full(img.shape, np.inf, dtype=np.float32)
    D[0, :] = 0; D[-1, :] = 0; D[:, 0] = 0; D[:, -1] = 0
    for it in range(iters):
        _scan(img, L, U, D, inverse=(it % 2 == 0))
    return D


def mahalanobis_to(points: np.ndarray, mean: np.ndarray, vi: np.ndarray) -> np.ndarray:
    """scipy.spatial.distance.cdist(points, mean[None], 'mahalanobis', VI=vi)[:, 0]."""
    d = points - mean[None, :]
    return np.sqrt(np.einsum("ij,jk,ik->i", d, vi, d))


def get_weights(img: np.ndarray, tile_size: float, sal: np.ndarray | None = None):
    """patolette.pyx:203-313.  img: H x W x 3 float64 sRGB in [0, 1].  `sal` lets a test supply the MBD map."""
    img_mean = np.mean(img, axis=2).astype(np.float32)
    if sal is None:
        sal = mbd(img_mean, 3)
    rows, cols = img.shape[0], img.shape[1]
    img_size = sqrt(rows * cols)
    bt = int(floor(0.1 * img_size))
    lab = rgb2lab(img)
    strips = [lab[0:bt, :, :], lab[rows - bt - 1:-1, :, :], lab[:, 0:bt, :], lab[:, cols - bt - 1:-1, :]]  # "left", "right", "top", "bottom"
    unrolled = lab.reshape(rows * cols, 3)
    us = []
    for px in strips:
        mean = np.mean(px, axis=(0, 1))
        vi = np.linalg.inv(np.cov(px.reshape(-1, 3).T))
        u = mahalanobis_to(unrolled, mean, vi).reshape(rows, cols)
        us.append(u / float(np.float32(np.max(u))))  # cdef float max_u_*
    u_max = np.maximum(np.maximum(np.maximum(us[0], us[1]), us[2]), us[3])
    u_final = (us[0] + us[1] + us[2] + us[3]) - u_max
    u_max_final = float(np.float32(np.max(u_final)))
    sal_max = float(np.float32(np.max(sal)))
    s = sal / np.float32(sal_max) + u_final / u_max_final  # float32 array / python float stays float32
    s = s / np.max(s)
    xv, yv = np.meshgrid(np.arange(cols), np.arange(rows))
    w2, h2 = rows / 2.0, cols / 2.0
    C = 1 - np.sqrt(np.power(xv - h2, 2) + np.power(yv - w2, 2)) / sqrt(w2 ** 2 + h2 ** 2)
    s = s * C
    s = s / np.max(s)
    s = np.vectorize(lambda v: 1.0 / (1.0 + exp(-10.0 * (v - 0.5))))(s)
    return 1 + np.reshape(s, -1) ** 2 * (rows * cols) / tile_size ** 2
