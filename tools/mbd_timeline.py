#!/usr/bin/env python
"""Debug: where does a tile's time go in k_mbd_pass?  PB_MBD_DEBUG=<file> makes the library dump per-tile time stamps of
the first scan's first four CTAs; this runs one saliency_mbd at --side and prints per-role medians (microseconds).
    python tools/mbd_timeline.py [--side 4096]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=4096)
    args = ap.parse_args()
    path = "/tmp/mbd_dbg.bin"
    import patolette_b200 as pb
    side = args.side
    colors = np.random.default_rng(1).random((side * side, 3))
    from patolette_b200 import _lib
    pb.saliency_mbd(side, side, colors)
    pb.saliency_mbd(side, side, colors)
    print(f"stage (k_sal_prepare + three scans): {_lib.load().patolette_b200_last_saliency_ms():.2f} ms")
    os.environ["PB_MBD_DEBUG"] = path
    pb.saliency_mbd(side, side, colors)
    del os.environ["PB_MBD_DEBUG"]
    report(path)


def report(path):
    raw = np.fromfile(path, dtype=np.uint64)
    nt = len(raw) // 32
    t = raw.reshape(4, nt, 8).astype(np.float64)
    valid = t[0, :, 0] > 0
    nv = int(valid.sum())
    t = t[:, :nv] / 1e3  # us
    t0 = t[0, 0, 2]
    names = ["compute start", "compute end", "loader start", "copies issued", "edge seen", "tile landed", "storer start", "storer end"]
    for g in range(4):
        d = t[g]
        print(f"group {g}: first compute start {d[0, 0] - t0:9.1f} us, last compute end {d[-1, 1] - t0:9.1f} us, tiles {nv}")
        print("   per-tile period (compute start k+1 - k): median %.2f us, p90 %.2f" % (np.median(np.diff(d[:, 0])), np.percentile(np.diff(d[:, 0]), 90)))
        print("   compute %.2f | loader: issue %.2f, edge wait %.2f, landing %.2f, total %.2f | storer %.2f | landed->compute start %.2f | compute end->storer start %.2f"
              % (np.median(d[:, 1] - d[:, 0]), np.median(d[:, 3] - d[:, 2]), np.median(d[:, 4] - d[:, 3]), np.median(d[:, 5] - d[:, 4]),
                 np.median(d[:, 5] - d[:, 2]), np.median(d[:, 7] - d[:, 6]), np.median(d[:, 0] - d[:, 5]), np.median(d[:, 6] - d[:, 1])))
        if g > 0:
            lag = d[:, 0] - t[g - 1][:, 0]
            print("   lag behind group %d at the same tile: median %.1f us" % (g - 1, np.median(lag)))
            k = np.arange(nv - 2)
            print("   edge seen(k) - predecessor compute end(k+1): median %.2f us" % np.median(d[k, 4] - t[g - 1][k + 1, 1]))


if __name__ == "__main__":
    main()
