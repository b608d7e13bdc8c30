#!/usr/bin/env python
"""Why is the dither stage slower on a saliency-weighted run?  Per-kernel profile of quantize_u8 at one size with
(a) no weights, (b) tile_size = 512 (device saliency weights), (c) synthetic saliency-like weights.
    python tools/time_weighted_dither.py [--side 8192]   -> one JSON line per arm"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=8192)
    ap.add_argument("--K", type=int, default=256)
    args = ap.parse_args()
    import patolette_b200 as pb
    from patolette_b200 import _lib
    from synth import saliency_like_weights
    lib = _lib.load()
    side, K = args.side, args.K
    n = side * side
    rgb8 = np.random.default_rng(3).integers(0, 256, (n, 3), dtype=np.uint8)
    arms = {"unweighted": dict(), "tile_size=512": dict(tile_size=512),
            "synthetic weights": dict(weights=saliency_like_weights(side, side, 3))}
    kw = dict(dither=True, color_space=pb.ColorSpace_ICtCp, kmeans_niter=10)
    for name, extra in arms.items():
        pb.quantize_u8(side, side, rgb8, K, **kw, **extra)
        lib.patolette_b200_profile_enable(1)
        ok, pal, pmap, msg = pb.quantize_u8(side, side, rgb8, K, **kw, **extra)
        buf = C.create_string_buffer(1 << 18)
        lib.patolette_b200_profile_json(buf, len(buf))
        lib.patolette_b200_profile_enable(0)
        prof = json.loads(buf.value.decode())
        t = pb.last_timings()
        d = {k: round(v["ms"], 3) for k, v in prof.items() if any(s in k for s in ("riemersma", "permute", "hilbert", "nngrid", "cand"))}
        pal_used = pal[pal[:, 0] >= 0]
        dup = len(pal_used) - len(np.unique(np.round(pal_used, 12), axis=0))
        print(json.dumps({"arm": name, "ok": bool(ok), "stage_ms": {k: round(v, 2) for k, v in t.items()}, "dither_kernels": d,
                          "palette_rows": int(len(pal_used)), "duplicate_palette_rows": int(dup),
                          "palette_min": [round(float(x), 4) for x in pal_used.min(axis=0)],
                          "palette_max": [round(float(x), 4) for x in pal_used.max(axis=0)]}), flush=True)


if __name__ == "__main__":
    main()
