"""TEST INFRASTRUCTURE ONLY - stand-in for scikit-image, which the reference's Python wrapper imports
(src/patolette/patolette.pyx:4-5) and which is not installed in this image (no network, not in the wheelhouse).

Only the two functions the wrapper uses exist: `skimage.color.rgb2lab` and `skimage.util.img_as_float`, restated
from scikit-image's published algorithm (skimage/color/colorconv.py `rgb2xyz` + `xyz2lab`, illuminant D65,
observer 2).  This is OUR code, not scikit-image's: the reference's saliency weights computed through it are pinned
to the reference for everything except `rgb2lab` itself.  Used by oracle/ref_build/build_ref_pyx.py."""
