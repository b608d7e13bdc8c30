#!/usr/bin/env python
"""Offline (CPU) analysis of the ordered-sum block summaries on the chains a real quantisation produces.

Builds the cluster tree of a synthetic image with the CPU oracle, forms the mean / centred-moment term
sequences of every cluster exactly as the kernels do, and runs the host build of pb_span.h over them
(tools/ubench/span_stats.cpp): how many blocks are accepted, parity-dependent, unusable - and why.
Test infrastructure only (uses oracle/)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.reflib import OracleLib  # noqa: E402
from synth import uniform_colors  # noqa: E402

NAMES = ["blocks", "accepted", "wrong", "unus_range", "unus_big", "unus_up", "unus_tie", "sensitive", "plainified",
         "unit_changes", "interval_fail", "first_block", "replayed_elems",
         "groups", "groups_plain_one_unit", "groups_usable_one_unit", "groups_usable", "runs", "run_records"]


def build():
    so = "/tmp/libspan_stats.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-DPB_SPAN_REASONS", "-I",
                    os.path.join(ROOT, "patolette_b200", "csrc"), os.path.join(ROOT, "tools", "ubench", "span_stats.cpp"),
                    "-o", so], check=True)
    lib = C.CDLL(so)
    lib.span_stats.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_void_p]
    return lib


def chains_of(c, w):
    """Term sequences of one cluster (rows in ascending pixel order)."""
    ww = np.ones(len(c)) if w is None else w
    t_mean = [ww] + [c[:, j] * ww for j in range(3)]
    wsum = np.add.accumulate(ww)[-1]
    mean = np.array([np.add.accumulate(t)[-1] for t in t_mean[1:]]) * (1.0 / wsum)
    d = c - mean
    wd = d * ww[:, None]
    t_cov = [wd[:, j] * d[:, k] for j, k in ((0, 0), (1, 0), (1, 1), (2, 0), (2, 1), (2, 2))]
    t_dist = ((d[:, 0] ** 2 + d[:, 1] ** 2) + d[:, 2] ** 2) * ww
    return {"mean": t_mean, "diag": [t_cov[0], t_cov[2], t_cov[5], t_dist], "offdiag": [t_cov[1], t_cov[3], t_cov[4]]}


def main():
    side = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    lib = build()
    orc = OracleLib().lib
    colors = uniform_colors(side, side, 1)
    planar = np.asfortranarray(colors).copy(order="F")
    orc.orc_color_transform(0, planar.ctypes.data_as(C.c_void_p), C.c_size_t(side * side))  # sRGB -> ICtCp
    n = side * side
    f = orc.orc_quantize_clusters
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    for OB, PER in ((512, 16),):
        tot = {g: np.zeros(len(NAMES), dtype=np.int64) for g in ("mean", "diag", "offdiag")}
        elems = {g: 0 for g in tot}
        for K in (1, 4, 16, 64, 256):
            if K == 1:
                labels = np.zeros(n, dtype=np.uint32)
                cnt = 1
            else:
                labels = np.zeros(n, dtype=np.uint32); centers = np.zeros((K, 3)); cc = C.c_size_t(0); gq = C.c_size_t(0)
                assert f(planar.ctypes.data, n, None, K, labels.ctypes.data, centers.ctypes.data, C.byref(cc), C.byref(gq)) == 0
                cnt = cc.value
            order = np.argsort(labels, kind="stable")
            bounds = np.searchsorted(labels[order], np.arange(cnt + 1))
            for j in range(cnt):
                idx = order[bounds[j]:bounds[j + 1]]
                if len(idx) < 4096:
                    continue
                ch = chains_of(planar[idx], None)
                for g, seqs in ch.items():
                    for t in seqs:
                        t = np.ascontiguousarray(t)
                        out = np.zeros(len(NAMES), dtype=np.int64)
                        lib.span_stats(t.ctypes.data, len(t), OB, PER, out.ctypes.data)
                        tot[g] += out
                        elems[g] += len(t)
        print(f"--- OB={OB} PER={PER} ({side}x{side}, clusters of K=1,4,16,64,256)")
        for g, o in tot.items():
            b = max(o[0], 1)
            print(f"  {g:8s} blocks {o[0]:8d} accepted {100*o[1]/b:6.2f}% wrong {o[2]} | unusable: range {100*o[3]/b:5.2f}% (first blocks {100*o[11]/b:4.2f}%) "
                  f"big {100*o[4]/b:5.2f}% up {100*o[5]/b:5.2f}% tie {100*o[6]/b:5.2f}% | two-parity {100*o[7]/b:5.2f}% plainified {100*o[8]/b:5.2f}% "
                  f"unit changes {100*o[9]/b:5.2f}% interval fails {100*o[10]/b:5.2f}% | replayed elements {100*o[12]/max(elems[g],1):5.2f}%")
            gr = max(o[13], 1)
            print(f"           groups {o[13]:7d}: all plain, one unit {100*o[14]/gr:5.1f}% | all usable, one unit {100*o[15]/gr:5.1f}% | "
                  f"all usable {100*o[16]/gr:5.1f}% | level-1 runs per group {o[17]/gr:4.2f}, records per run {o[18]/max(o[17],1):5.1f}")


if __name__ == "__main__":
    main()
