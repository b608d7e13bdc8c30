#!/usr/bin/env python
"""GPU debug: where the resolving warps of the ordered sums spend their cycles, per chain, for a
single-cluster run (K=1: GQ root passes only) and for a K=256 tree."""
import ctypes as C
import json
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from patolette_b200 import _lib
from synth import uniform_colors

lib = _lib.load()
side = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
colors = uniform_colors(side, side, 1)
planar = np.asfortranarray(colors)
n = side * side
assert lib.patolette_b200_color_transform(0, planar.ctypes.data, n) == 0
for K in (1, 256):
    for rep in range(2):
        cnt = (C.c_ulonglong * 16)(); dbg = (C.c_ulonglong * 35)()
        lib.patolette_b200_ordered_counts(cnt, 1); lib.patolette_b200_ordered_chain_debug(dbg, 1)
        count = C.c_size_t(0); gq = C.c_size_t(0)
        centers = np.zeros((max(K, 16), 3))
        assert lib.patolette_b200_quantize_clusters(planar.ctypes.data, n, None, K, None, centers.ctypes.data, C.byref(count), C.byref(gq)) == 0
        lib.patolette_b200_ordered_counts(cnt, 0); lib.patolette_b200_ordered_chain_debug(dbg, 0)
    d = np.array(list(dbg), dtype=np.float64).reshape(7, 5)
    print(json.dumps({"K": K, "clusters": count.value, "counts": [int(x) for x in cnt],
                      "per_chain_Mcycles_scan_walk_replay": (d[:, :3] / 1e6).round(2).tolist(),
                      "per_chain_replays": d[:, 3].astype(int).tolist(), "per_chain_records_walked_singly": d[:, 4].astype(int).tolist()}))
