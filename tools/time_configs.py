#!/usr/bin/env python
"""Stage timings of the BASELINE configs (or scaled versions) through the public quantize() API.
Not a bench contract file - a survey tool whose output is committed under profiles/."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import patolette_b200 as pb  # noqa: E402
from synth import saliency_like_weights, uniform_colors  # noqa: E402

CONFIGS = [
    ("C1 512^2 K=16 sRGB", dict(side=512, K=16, cs=0, dither=False, km=0)),
    ("C2 4096^2 K=256 ICtCp", dict(side=4096, K=256, cs=2, dither=False, km=0)),
    ("C3/4 2048^2 K=256 CIELuv dither", dict(side=2048, K=256, cs=1, dither=True, km=0)),
    ("C3 8192^2 K=256 CIELuv no-dither", dict(side=8192, K=256, cs=1, dither=False, km=0)),
    ("C4 4096^2 K=256 ICtCp kmeans10 subsample", dict(side=4096, K=256, cs=2, dither=False, km=10)),
    ("C4 4096^2 K=256 ICtCp kmeans10 full-N", dict(side=4096, K=256, cs=2, dither=False, km=10, full=True)),
    ("C5 4096^2 K=1024 ICtCp weighted kmeans10 full-N", dict(side=4096, K=1024, cs=2, dither=False, km=10, full=True, w=True)),
    ("C4 16384^2 K=256 ICtCp no-dither", dict(side=16384, K=256, cs=2, dither=False, km=0)),
    ("FULL C3 8192^2 K=256 CIELuv dither", dict(side=8192, K=256, cs=1, dither=True, km=0, reps=2)),
    ("FULL C4 16384^2 K=256 ICtCp kmeans10 dither", dict(side=16384, K=256, cs=2, dither=True, km=10, reps=2)),
]
PROFILE = "--profile" in sys.argv
if PROFILE:
    sys.argv.remove("--profile")
import ctypes as C  # noqa: E402
from patolette_b200 import _lib  # noqa: E402


only = sys.argv[1:]
out = []
for name, c in CONFIGS:
    if only and not any(o in name for o in only):
        continue
    side = c["side"]
    colors = uniform_colors(side, side, 1)
    w = saliency_like_weights(side, side, 1) if c.get("w") else None
    kw = dict(dither=c["dither"], color_space=c["cs"], tile_size=0, kmeans_niter=c["km"],
              kmeans_max_samples=side * side if c.get("full") else 512 ** 2, weights=w)
    best = None
    for rep in range(c.get("reps", 2)):
        t0 = time.perf_counter()
        ok, pal, pmap, msg = pb.quantize(side, side, colors, c["K"], **kw)
        dt = time.perf_counter() - t0
        assert ok, msg
        if best is None or dt < best[0]:
            best = (dt, pb.last_timings())
    rec = {"config": name, "wall_s": round(best[0], 4), "Mpx_per_s": round(side * side / best[0] / 1e6, 2),
           "stage_ms": {k: round(v, 2) for k, v in best[1].items()}, "distinct": int(len(np.unique(pmap)))}
    if PROFILE:  # one more call with the per-kernel CUDA-event profiler on
        lib = _lib.load()
        lib.patolette_b200_profile_enable(1)
        ok, pal, pmap, msg = pb.quantize(side, side, colors, c["K"], **kw)
        buf = C.create_string_buffer(1 << 18)
        lib.patolette_b200_profile_json(buf, len(buf))
        lib.patolette_b200_profile_enable(0)
        prof = json.loads(buf.value.decode())
        rec["kernels_ms"] = {k: [round(v["ms"], 2), v["launches"]] for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:16]}
        top = max(prof.items(), key=lambda kv: kv[1]["ms"])
        rec["top_each_ms"] = {top[0]: [round(x, 2) for x in top[1].get("each", [])]}
    print(json.dumps(rec), flush=True)
    out.append(rec)
    del colors, pmap
