#!/usr/bin/env python
"""One device-resident step of the bench workload under the library's event profiler in TIMELINE mode (the split loop
keeps its two streams): writes gpurun_out/<tag>_timeline.txt ("kernel stream start_ms end_ms") and prints where the
step's time goes - per stream busy time, time when both / one / no stream has a kernel in flight, the largest gaps.
   python tools/timeline.py <tag> [side]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from patolette_b200 import _lib
from patolette_b200._lib import QuantizationOptions
from synth import uniform_colors

tag = sys.argv[1]; side = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
lib = _lib.load()
n = side * side; K = 256
cols = torch.from_numpy(np.asfortranarray(uniform_colors(side, side, 3)).T.copy()).cuda()  # 3 x n planes
dmap = torch.empty(n, dtype=torch.int64, device="cuda")
pal = np.zeros((K, 3), order="F")
opt = QuantizationOptions(dither=True, palette_only=False, color_space=2, kmeans_niter=10, kmeans_max_samples=512 * 512, verbose=False)
code = C.c_int(0)
def step():
    lib.patolette_b200_device(side, side, C.c_void_p(cols.data_ptr()), None, K, C.byref(opt), pal.ctypes.data, C.c_void_p(dmap.data_ptr()), C.byref(code))
    assert code.value == 0
for _ in range(2): step()
lib.patolette_b200_set_option(b"prof_timeline", 1)
lib.patolette_b200_profile_enable(1)
step(); torch.cuda.synchronize()
buf = C.create_string_buffer(1 << 22)
lib.patolette_b200_profile_timeline(buf, len(buf))
lib.patolette_b200_profile_enable(0)
lib.patolette_b200_set_option(b"prof_timeline", 0)
text = buf.value.decode()
open(os.path.join(ROOT, "gpurun_out", f"{tag}_timeline.txt"), "w").write(text)
ev = [(l.split()[0], int(l.split()[1]), float(l.split()[2]), float(l.split()[3])) for l in text.splitlines()]
end = max(e[3] for e in ev)
print("launches", len(ev), "span ms", round(end, 2))
for sid in sorted({e[1] for e in ev}):
    print(" stream", sid, "busy ms", round(sum(e[3] - e[2] for e in ev if e[1] == sid), 2), "launches", sum(1 for e in ev if e[1] == sid))
# coverage: sweep
pts = sorted([(e[2], 1) for e in ev] + [(e[3], -1) for e in ev])
cov = {0: 0.0, 1: 0.0, 2: 0.0}; depth = 0; last = 0.0
for t, d in pts:
    cov[min(depth, 2)] += t - last; last = t; depth += d
print(" time with 0 / 1 / 2+ kernels in flight:", {k: round(v, 2) for k, v in cov.items()})
by = {}
for nm, sid, a, b in ev: by[nm] = by.get(nm, 0.0) + (b - a)
print(" per kernel (ms, incl. waiting for the other stream):", {k: round(v, 1) for k, v in sorted(by.items(), key=lambda kv: -kv[1])[:14]})
