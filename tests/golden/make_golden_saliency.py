#!/usr/bin/env python
"""Freeze outputs of the REFERENCE'S OWN WRAPPER (src/patolette/patolette.pyx compiled in place by
oracle/ref_build/build_ref_pyx.py, scikit-image replaced by the stand-in of oracle/ref_build/shim/skimage) for the
saliency cases of tests/synth.py (row N3).

    OMP_NUM_THREADS=1 python tests/golden/make_golden_saliency.py

Writes tests/golden/golden_saliency.npz: per case `<name>/mbd` (float32 [h, w], the reference's `mbd()`),
`<name>/weights` (float64 [h * w], the reference's `get_weights()`), and for the end-to-end cases `<name>/palette`
(float64 [K, 3]) and `<name>/map` (uint16) of the reference's `quantize(..., tile_size)`.
"""
import os
os.environ["OPENBLAS_NUM_THREADS"] = "1"
os.environ.setdefault("OMP_NUM_THREADS", "1")
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.ref_build import build_ref_pyx  # noqa: E402
from synth import SALIENCY_CASES, saliency_case_colors  # noqa: E402

E2E = {"scene_96x64_t32": dict(K=16, dither=False, kmeans_niter=0),
       "scene_160x120_t512": dict(K=32, dither=True, kmeans_niter=4)}


def main():
    ref = build_ref_pyx.load()
    out = {}
    for name, spec in SALIENCY_CASES.items():
        w, h = spec["w"], spec["h"]
        colors = saliency_case_colors(spec)
        img = np.reshape(colors, (h, w, 3))
        out[name + "/mbd"] = np.asarray(ref.mbd(np.mean(img, axis=2).astype(np.float32), 3))
        out[name + "/weights"] = np.asarray(ref.get_weights(img, spec["tile"]), dtype=np.float64)
        print(name, float(out[name + "/mbd"].max()), float(out[name + "/weights"].min()), float(out[name + "/weights"].max()))
        if name in E2E:
            kw = E2E[name]
            ok, pal, pmap, msg = ref.quantize(w, h, colors, kw["K"], dither=kw["dither"], palette_only=False, color_space=2,
                                              tile_size=spec["tile"], kmeans_niter=kw["kmeans_niter"])
            assert ok, msg
            out[name + "/palette"] = np.asarray(pal, dtype=np.float64)
            out[name + "/map"] = np.asarray(pmap).astype(np.uint16)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_saliency.npz"), **out)


if __name__ == "__main__":
    main()
