"""See skimage/__init__.py (stand-in, test infrastructure only)."""
import numpy as np

_XYZ_FROM_RGB = np.array([[0.412453, 0.357580, 0.180423],
                          [0.212671, 0.715160, 0.072169],
                          [0.019334, 0.119193, 0.950227]])
_WHITE_D65_2 = np.array([0.95047, 1.0, 1.08883])


def rgb2lab(rgb):
    """float sRGB in [0, 1], channels last -> CIELab (D65, 2 degree observer)."""
    arr = np.array(rgb, dtype=np.float64, copy=True)
    mask = arr > 0.04045
    arr[mask] = np.power((arr[mask] + 0.055) / 1.055, 2.4)
    arr[~mask] /= 12.92
    arr = (arr @ _XYZ_FROM_RGB.T) / _WHITE_D65_2
    mask = arr > 0.008856
    arr[mask] = np.cbrt(arr[mask])
    arr[~mask] = 7.787 * arr[~mask] + 16.0 / 116.0
    x, y, z = arr[..., 0], arr[..., 1], arr[..., 2]
    return np.stack([116.0 * y - 16.0, 500.0 * (x - y), 200.0 * (y - z)], axis=-1)
