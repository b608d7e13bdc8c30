#!/usr/bin/env python
"""Per-kernel times of the saliency stage (row N3) on the GPU box, and - where the compiled reference wrapper
travelled along (oracle/_ref/pyx) - the reference's own get_weights() on the box's CPU at a bounded size.

    python tools/time_saliency.py [--sides 4096 16384] [--ref-side 1024]      -> one JSON line per size"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sides", type=int, nargs="*", default=[4096, 16384])
    ap.add_argument("--ref-side", type=int, default=1024, help="0: skip the reference's CPU timing")
    args = ap.parse_args()
    import patolette_b200 as pb
    from patolette_b200 import _lib
    from synth import scene_colors
    lib = _lib.load()
    pb.saliency_weights(256, 256, scene_colors(256, 256, 1), 512.0)  # warm-up: context, module load
    for side in args.sides:
        n = side * side
        colors = scene_colors(side, side, 7) if side <= 4096 else np.random.default_rng(7).random((n, 3))
        planar = np.asfortranarray(colors)
        del colors
        out = np.empty(n)
        lib.patolette_b200_saliency_weights(side, side, planar.ctypes.data, 512.0, out.ctypes.data, 0)
        lib.patolette_b200_profile_enable(1)
        t0 = time.perf_counter()
        rc = lib.patolette_b200_saliency_weights(side, side, planar.ctypes.data, 512.0, out.ctypes.data, 0)
        wall = time.perf_counter() - t0
        buf = C.create_string_buffer(1 << 16)
        lib.patolette_b200_profile_json(buf, len(buf))
        lib.patolette_b200_profile_enable(0)
        prof = json.loads(buf.value.decode())
        print(json.dumps({"side": side, "rc": rc, "saliency_ms": round(lib.patolette_b200_last_saliency_ms(), 3),
                          "wall_ms_with_copies": round(wall * 1e3, 1),
                          "Mpixels/s (device stage)": round(n / lib.patolette_b200_last_saliency_ms() / 1e3, 1),
                          "kernels": {k: {"ms": round(v["ms"], 3), "launches": v["launches"]} for k, v in prof.items()},
                          "weights": [float(out.min()), float(out.max())]}), flush=True)
        del planar, out
    if args.ref_side <= 0:
        return
    try:
        from oracle.ref_build import build_ref_pyx
        ref = build_ref_pyx.load()
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"reference": "unavailable", "why": str(e)[:200]}))
        return
    side = args.ref_side
    img = scene_colors(side, side, 7).reshape(side, side, 3)
    t0 = time.perf_counter()
    wr = np.asarray(ref.get_weights(img, 512.0))
    dt = time.perf_counter() - t0
    wg = pb.saliency_weights(side, side, img.reshape(-1, 3), 512.0)
    print(json.dumps({"reference_get_weights": {"side": side, "seconds": round(dt, 3), "Mpixels/s": round(side * side / dt / 1e6, 3),
                                                "what": "the reference's compiled wrapper (patolette.pyx get_weights, numpy + scipy + skimage stand-in) on this box's CPU"},
                      "max_rel_diff_vs_gpu": float(np.max(np.abs(wr - wg) / wr))}), flush=True)


if __name__ == "__main__":
    main()
