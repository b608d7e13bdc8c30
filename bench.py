#!/usr/bin/env python
"""bench.py - headline metric of the patolette pixel-array hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): Mpixels/s of one full quantisation at K=256.  A "step" is one pass of
the whole hot path (colour transform -> GQ -> LQ split loop -> palette -> nearest map) over one
synthetic image.  Workload at N=1 = BASELINE config[1]: 4096 x 4096 uniform-random sRGB f64,
K=256, ICtCp, dither off, kmeans off.

  value  : device-resident throughput - inputs already in HBM (patolette_b200_device), timed with
           CUDA events on the stream the library launches on, max over ranks.
  e2e    : same metric through the reference-facing C ABI patolette() with HOST buffers (pinned),
           host->device and device->host copies inside the timed region.
  roofline: the dominant kernel of the step (by accumulated CUDA-event time from the library's
           per-kernel profiler, measured in a separate untimed profiling step on one stream), algorithmic
           bytes / duration against MEASURED_PEAKS.json; `roofline.stages` gives the same for the covariance
           stage as a whole (all k_ord_* kernels), the assignment kernel and the projection/sort/partition
           group; `traffic` is the ncu DRAM byte count per launch of that kernel (profiles/r01_ncu_summary.json).
  cpu_baseline: the reference's own code (oracle/_ref) - or the oracle port when the prebuilt
           .so is absent - timed on this box's host cores on a bounded sample (rank 0, N=1).

N > 1 (torchrun, one rank per GPU): the LQ tree does not shard yet (DESIGN.md section multi-GPU) -
each rank quantises its own image ("replicas", weak scaling); no data-path collective.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOAD = dict(name="4096x4096 uniform sRGB f64, K=256, ICtCp, dither off, kmeans off (BASELINE config[1])",
                w=4096, h=4096, K=256, seed=1, color_space=2, dither=False, kmeans_niter=0)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--side", type=int, default=0, help="override the image side (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard", action="store_true",
                    help="N > 1: ONE image, ordered sums chain-sharded over the ranks (strong scaling) instead of replicas")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "250", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def reduce_max_ms(ms: float, dist, device="cuda") -> float:
    """Step time of the job = the slowest rank's (contract: max over ranks, measured on the device)."""
    if dist is None:
        return ms
    import torch
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(n_pixels: int, world: int, ms_per_step: float) -> float:
    """Whole-job Mpixels/s: every rank processes its own n_pixels image per step (replicas, weak scaling)."""
    return world * n_pixels / (ms_per_step * 1e-3) / 1e6


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def time_reference(side: int, steps: int, warmup: int):
    """The reference's CPU implementation (oracle/_ref, else the oracle port) on a side x side sample."""
    cores = host_cores()
    os.environ["OMP_NUM_THREADS"] = str(cores)      # faiss + the exact-NN stand-in scale; patolette's C does not
    os.environ["OPENBLAS_NUM_THREADS"] = "1"        # see oracle/reflib.py
    from oracle.reflib import OracleLib, RefLib
    from synth import uniform_colors
    try:
        lib, kind = RefLib(), "reference"
    except Exception:
        lib, kind = OracleLib(), "port"
    colors = uniform_colors(side, side, WORKLOAD["seed"])
    kw = dict(dither=WORKLOAD["dither"], color_space=WORKLOAD["color_space"], kmeans_niter=WORKLOAD["kmeans_niter"])
    for _ in range(warmup):
        lib.quantize(side, side, colors, WORKLOAD["K"], **kw)
    t0 = time.perf_counter()
    for _ in range(steps):
        code, _, _ = lib.quantize(side, side, colors, WORKLOAD["K"], **kw)
        assert code == 0
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dict(value=side * side / dt / 1e6, unit="Mpixels/s", cores=cores, kind=kind,
                sample=f"{side}x{side} crop-sized image of the same synthetic workload, {steps} step(s) of "
                       f"{dt:.2f} s; FLANN absent -> exact brute-force NN stand-in (OpenMP); OpenBLAS 1 thread"), dt


class QuietStdout:
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner at
    communicator creation, for one): while the bench runs, file descriptor 1 points at stderr."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = args.steps + args.warmup
    side = 1536 if total <= 8 else (1024 if total <= 20 else 768)
    with QuietStdout():
        cb, dt = time_reference(side, args.steps, args.warmup)
    line = {"metric": "Mpixels/s end-to-end quantize() at K=256", "value": cb["value"], "unit": "Mpixels/s",
            "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD["name"] + f" - bounded sample {side}x{side}", "K": WORKLOAD["K"]},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_ours(args):
    with QuietStdout():
        line = measure_ours(args)
    if line is not None:
        print(json.dumps(line))


def measure_ours(args):
    import numpy as np
    import torch
    import patolette_b200 as pb
    from patolette_b200 import _lib
    from synth import uniform_colors

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    lib = _lib.load()
    assert lib.patolette_b200_set_device(local) == 0
    side = args.side or WORKLOAD["w"]
    w = h = side
    n = w * h
    K = WORKLOAD["K"]
    shard = bool(args.shard and world > 1)
    if shard:  # every rank the SAME image; moment rows exchanged over a gloo group (host bytes)
        pb.set_sharding(rank, world, pb.torch_allgather(dist.new_group(backend="gloo")))
    colors = uniform_colors(w, h, WORKLOAD["seed"] + (0 if shard else rank))  # replicas: every rank its own image
    planar = np.asfortranarray(colors)                               # [N,3] F-order == 3 planes
    opts = _lib.QuantizationOptions(False, False, WORKLOAD["color_space"], 0, 512 ** 2, False)
    code = C.c_int(0)
    palette = np.zeros((K, 3), order="F")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm (value) ----------------
    d_in = torch.from_numpy(planar.T.copy()).cuda()                 # 3 x N planes, contiguous
    d_map = torch.empty(n, dtype=torch.int64, device="cuda")
    stream = torch.cuda.Stream()  # a non-blocking stream: the legacy NULL stream serialises against everything
    torch.cuda.set_stream(stream)
    lib.patolette_b200_set_stream(C.c_void_p(stream.cuda_stream), 1)

    def step_resident():
        lib.patolette_b200_device(w, h, d_in.data_ptr(), None, K, C.byref(opts), palette.ctypes.data,
                                  d_map.data_ptr(), C.byref(code))
        assert code.value == 0, code.value

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = pb.last_timings()["launches"]
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = int(pb.last_timings()["launches"])
    ms_step = reduce_max_ms(ms_total, dist) / args.steps
    stage = pb.last_timings()

    # ---------------- end-to-end arm: host buffers through the reference ABI ----------------
    h_in = torch.from_numpy(planar.T.copy()).pin_memory()
    h_map = torch.empty(n, dtype=torch.int64).pin_memory()
    lib.patolette_b200_set_stream(None, 0)

    def step_e2e():
        lib.patolette(w, h, h_in.data_ptr(), None, K, C.byref(opts), palette.ctypes.data, h_map.data_ptr(),
                      C.byref(code))
        assert code.value == 0

    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    e2e_s = reduce_max_ms(e2e_s * 1e3, dist) / 1e3
    e2e_stage = pb.last_timings()
    assert (h_map.numpy() == d_map.cpu().numpy()).all(), "host and device arms disagree"

    # ---------------- per-kernel profile (untimed extra step) -> roofline ----------------
    lib.patolette_b200_set_stream(C.c_void_p(stream.cuda_stream), 1)
    cnt = (C.c_ulonglong * 16)()
    lib.patolette_b200_ordered_counts(cnt, 1)
    lib.patolette_b200_profile_enable(1)
    step_resident()
    torch.cuda.synchronize()
    buf = C.create_string_buffer(1 << 18)
    lib.patolette_b200_profile_json(buf, len(buf))
    lib.patolette_b200_profile_enable(0)
    lib.patolette_b200_set_stream(None, 0)
    prof = json.loads(buf.value.decode())
    lib.patolette_b200_ordered_counts(cnt, 0)
    ord_acc, ord_rep = int(cnt[0]), int(cnt[1])
    peak, peak_src = measured_peaks()
    kernels = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])
    top_name, top = kernels[0]
    # Algorithmic bytes of the resolve kernels = what the sequential walk must read: one 32-byte record per
    # (block, chain) it walks + the 4 KB of terms of every block it replays.  The pixel
    # planes are attributed to the summary kernels, which are the ones that stream them.
    resolve_bytes = (ord_acc + ord_rep) * 32.0 + ord_rep * 4096.0
    resolve_ms = sum(v["ms"] for k, v in prof.items() if k.startswith("k_ord_resolve"))
    for k, v in prof.items():
        if k.startswith("k_ord_resolve") and resolve_ms > 0:
            v["bytes"] = resolve_bytes * v["ms"] / resolve_ms
    gbs = top["bytes"] / (top["ms"] * 1e-3) / 1e9 if top["ms"] > 0 else 0.0
    ncu = {}
    try:
        with open(os.path.join(ROOT, "profiles", "r01_ncu_summary.json")) as f:
            ncu = json.load(f)
    except Exception:
        pass

    def stage_roofline(names, bytes_per_step):
        ms = sum(v["ms"] for k, v in prof.items() if any(k.startswith(nm) for nm in names))
        g = bytes_per_step / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        return {"kernels": names, "ms_per_step": round(ms, 3), "algorithmic_bytes_per_step": bytes_per_step,
                "achieved": round(g, 1), "unit": "GB/s", "frac": round(g / peak, 4)}

    pass_bytes = sum(v["bytes"] for k, v in prof.items() if k.startswith("k_ord_summary_"))  # 24|32 B per pixel-visit per pass
    roofline = {"bound": "hbm", "kernel": top_name, "achieved": gbs, "peak": peak, "unit": "GB/s",
                "frac": gbs / peak, "traffic": ncu.get(top_name, {}).get("dram_bytes_per_launch"),
                "peak_source": peak_src,
                "launches_per_step": top["launches"], "ms_per_step_in_kernel": top["ms"],
                "algorithmic_bytes_per_step": top["bytes"],
                "note": "the step has no single dominant HBM kernel: bit-exact ordered sums are instruction-/latency-bound "
                        "(DESIGN.md sections 3-4); per-stage rooflines in `stages`",
                "top_kernel_launch_ms": [round(x, 3) for x in top.get("each", [])],
                "stages": {
                    "covariance (ordered mean + centred passes: k_ord_*)": stage_roofline(["k_ord_"], pass_bytes),
                    "assignment (k_nearest: exact f64 1-NN over per-cell candidate lists)": stage_roofline(["k_nearest"], 32.0 * n),
                    "projection + bucket sort + partition (k_dots_minmax, k_buckets, k_tile_*, k_scatter)":
                        stage_roofline(["k_dots_minmax", "k_buckets", "k_tile_", "k_scatter", "k_class_start"],
                              sum(v["bytes"] for k, v in prof.items() if k in ("k_dots_minmax", "k_buckets", "k_scatter"))),
                },
                "kernels": {k: {"ms": round(v["ms"], 3), "launches": v["launches"],
                                "GB/s": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 and v["bytes"] else None}
                            for k, v in kernels[:14]}}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return None
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_baseline, _ = time_reference(1536, 1, 0)
    line = {
        "metric": "Mpixels/s end-to-end quantize() at K=256",
        "value": aggregate_throughput(n, 1 if shard else world, ms_step), "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if shard else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD["name"] if not args.side else f"{side}x{side} debug size", "K": K,
                   "parallelism": "single GPU" if world == 1 else (
                       f"chain-sharded x{world} (one image, ordered sums split by chain, moment rows all-gathered)" if shard
                       else f"replicas x{world} (one image per rank, no collective)"),
                   "l2": "inputs (403 MB) larger than L2; no flush needed",
                   "mode": "exact (bit-identical to the reference CPU path)"},
        "e2e": {"value": (1 if shard else world) * n / e2e_s / 1e6, "unit": "Mpixels/s", "h2d_bytes_per_step": 24 * n,
                "d2h_bytes_per_step": 8 * n + 24 * K, "ms_per_step": e2e_s * 1e3, "host_buffers": "pinned",
                "stage_ms": {k: round(v, 3) for k, v in e2e_stage.items()}},
        "gpu_launches": launches * args.steps if launches else None,
        "gpu_launches_per_step": launches,
        "stage_ms": {k: round(v, 3) for k, v in stage.items()},
        "roofline": roofline, "clocks": clocks,
        "ordered_sums": {"blocks_accepted": ord_acc, "blocks_replayed": ord_rep,
                         "replay_frac": ord_rep / max(ord_acc + ord_rep, 1),
                         "replay_reasons": {"flag": int(cnt[2]), "binade_guess": int(cnt[3]), "bounds": int(cnt[4])},
                         "replay_rounds": int(cnt[5]), "elementwise_subchunks": int(cnt[6]),
                         "accepted_two_parity": int(cnt[7]),
                         "resolve_warp_Mcycles": {"scan_walk": round(int(cnt[8]) / 1e6, 2), "two_parity": round(int(cnt[9]) / 1e6, 2),
                                                  "replay": round(int(cnt[10]) / 1e6, 2), "slowest_warp": round(int(cnt[12]) / 1e6, 3)},
                         "records_walked_singly": int(cnt[11])},
    }
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    if dist is not None:
        dist.destroy_process_group()
    return line


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
