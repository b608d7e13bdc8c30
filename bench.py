#!/usr/bin/env python
"""bench.py - headline metric of the patolette pixel-array hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): Mpixels/s of one full quantisation at K=256.  A "step" is one pass of the whole
hot path (colour transform -> GQ -> LQ split loop -> palette -> KMeans refinement -> Riemersma dither)
over one synthetic image.  Workload = the configuration the metric is quoted on, BASELINE config[3] ("C4"):
16384 x 16384 uniform-random sRGB f64 (seed 3), K=256, ICtCp, kmeans_niter=10 (default 512^2 sample cap),
dither on.  It fits one GPU (~45 GB of 180 GB).

  value   : device-resident throughput - the image already in HBM (patolette_b200_device), CUDA events
            on the stream the library launches on, max over ranks.
  e2e     : same metric through the reference-facing C ABI patolette() with HOST buffers (pinned torch
            tensors), host->device and device->host copies inside the timed region;
            e2e.pageable = the same through the Python quantize() with plain numpy arrays (what a
            drop-in caller passes), e2e.u8 = the uint8 ingest extension (patolette_b200_u8).
  roofline: the kernel with the largest accumulated CUDA-event time of one profiled step (the library's
            per-kernel profiler, untimed extra step on one stream): algorithmic bytes / duration against
            MEASURED_PEAKS.json; `stages` = the same per stage, `fp64` = FP64-ALU fractions of the
            compute-bound kernels against a DFMA peak measured in this run, `dither_ns_per_px`.
  cpu_baseline / --impl reference: the reference's own code (oracle/_ref, else the oracle port) on this
            box's host cores, same options, on a bounded sample (4096^2; the CPU needs ~12 min and 45 GB
            for one 16384^2 step).
  other_configs: BASELINE config[1] (C2) and config[2] (C3), device-resident, a few steps each.

N > 1 (torchrun, one rank per GPU): ONE 16384^2 image, sharded (strong scaling) - see DESIGN.md section 7.
`--replicas` runs one image per rank instead (weak scaling, no data-path collective).
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOAD = dict(name="16384x16384 uniform sRGB f64, K=256, ICtCp, kmeans_niter=10 (512^2 sample cap), dither on "
                     "(BASELINE config[3], the configuration the metric is quoted on)",
                golden="c4_16384_k256_ictcp_kmeans10_dither",
                w=16384, h=16384, K=256, seed=3, color_space=2, dither=True, kmeans_niter=10)
OTHER = {
    "C2": dict(name="4096x4096, K=256, ICtCp, dither off, kmeans off (BASELINE config[1])", golden="c2_4096_k256_ictcp",
               w=4096, h=4096, K=256, seed=1, color_space=2, dither=False, kmeans_niter=0),
    "C3": dict(name="8192x8192, K=256, CIELuv, dither on, kmeans off (BASELINE config[2])", golden="c3_8192_k256_cieluv_dither",
               w=8192, h=8192, K=256, seed=2, color_space=1, dither=True, kmeans_niter=0),
}
REFERENCE_SIDE = 4096  # the CPU arm's bounded sample: same options, 4096 x 4096
METRIC = "Mpixels/s end-to-end quantize() at K=256"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--side", type=int, default=0, help="override the image side (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip other_configs / pageable / u8 legs")
    ap.add_argument("--replicas", action="store_true", help="N > 1: one image per rank (weak scaling) instead of sharding one image")
    ap.add_argument("--ref-budget-s", type=float, default=240.0, help="--impl reference: wall-clock budget for the timed steps")
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
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(sm)}


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
    """Whole-job Mpixels/s when every rank processes its own n_pixels image per step (replicas, weak scaling);
    a sharded job passes world=1: all ranks together process ONE image per step."""
    return world * n_pixels / (ms_per_step * 1e-3) / 1e6


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def workload_kwargs(wl):
    return dict(dither=wl["dither"], color_space=wl["color_space"], kmeans_niter=wl["kmeans_niter"])


def time_reference(side: int, steps: int, warmup: int, budget_s: float, wl=WORKLOAD):
    """The reference's CPU implementation (oracle/_ref, else the oracle port) on a side x side image of the
    same synthetic workload with the same options.  Runs `warmup` untimed steps (capped at one: there is
    nothing to warm on the CPU beyond the page cache) and then as many of the `steps` as fit in budget_s
    (at least one)."""
    cores = host_cores()
    os.environ["OMP_NUM_THREADS"] = str(cores)      # faiss + the exact-NN stand-in scale; patolette's own C does not
    os.environ["OPENBLAS_NUM_THREADS"] = "1"        # see oracle/reflib.py
    from oracle.reflib import OracleLib, RefLib
    from synth import uniform_colors
    try:
        lib, kind = RefLib(), "reference"
    except Exception:
        lib, kind = OracleLib(), "port"
    colors = uniform_colors(side, side, wl["seed"])
    kw = workload_kwargs(wl)
    if warmup > 0 and side <= 1024:
        lib.quantize(side, side, colors, wl["K"], **kw)
    done, t0 = 0, time.perf_counter()
    while done < max(steps, 1):
        code, _, _ = lib.quantize(side, side, colors, wl["K"], **kw)
        assert code == 0
        done += 1
        if time.perf_counter() - t0 > budget_s * done / (done + 1):  # the next step would overrun the budget
            break
    dt = (time.perf_counter() - t0) / done
    return dict(value=side * side / dt / 1e6, unit="Mpixels/s", cores=cores, kind=kind, cpu=cpu_model(),
                steps_run=done, seconds_per_step=round(dt, 2),
                sample=f"{side}x{side} image of the same synthetic workload, same options (ICtCp, KMeans "
                       f"{wl['kmeans_niter']}, dither {'on' if wl['dither'] else 'off'}), {done} timed step(s) of {dt:.1f} s "
                       f"(the requested {steps} do not fit a few minutes); a 16384^2 step needs ~12 min and 45 GB on "
                       f"the CPU; FLANN absent -> exact brute-force NN stand-in (OpenMP over pixels for the map, "
                       f"serial inside the dither); faiss OpenMP on {cores} threads; OpenBLAS 1 thread"), dt


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
    side = args.side or REFERENCE_SIDE
    with QuietStdout():
        cb, dt = time_reference(side, args.steps, args.warmup, args.ref_budget_s)
    line = {"metric": METRIC, "value": cb["value"], "unit": "Mpixels/s",
            "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "steps_run": cb["steps_run"],
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD["name"] + f" - CPU arm on a bounded sample: {side}x{side}, same options", "K": WORKLOAD["K"],
                       "sample_side": side},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_ours(args):
    with QuietStdout():
        line = measure_ours(args)
    if line is not None:
        print(json.dumps(line))


def golden_entry(name):
    try:
        with open(os.path.join(ROOT, "tests", "golden", "golden_big.json")) as f:
            return json.load(f)["cases"].get(name)
    except Exception:
        return None


def sha_map(t) -> str:
    """sha256 of a (host) int64/uintp tensor's bytes."""
    import numpy as np
    return hashlib.sha256(np.ascontiguousarray(t.numpy() if hasattr(t, "numpy") else t).tobytes()).hexdigest()


def check_golden(wl, side, pal, map_sha):
    """Parity at full size: the frozen hashes of the reference's own output (tests/golden/golden_big.json)."""
    g = golden_entry(wl.get("golden", "")) if side == wl["w"] else None
    if g is None:
        return "no golden frozen for this size"
    import numpy as np
    pal_sha = hashlib.sha256(np.ascontiguousarray(pal.ravel(order="F")).tobytes()).hexdigest()
    assert pal_sha == g["palette_sha256"], "palette differs from the frozen reference output"
    assert map_sha == g["map_sha256"], "palette_map differs from the frozen reference output"
    return "bit-identical to the reference's frozen output (palette + map sha256, tests/golden/golden_big.json)"


SPLIT_COUNTS = None  # split-selection routes of the last profiled step (pb_certify.cu)


def profile_step(lib, run_step):
    """One extra, untimed step under the library's per-kernel CUDA-event profiler."""
    import torch
    cnt = (C.c_ulonglong * 16)()
    sc = (C.c_ulonglong * 4)()
    lib.patolette_b200_ordered_counts(cnt, 1)
    lib.patolette_b200_split_counts(sc, 1)
    lib.patolette_b200_profile_enable(1)
    run_step()
    torch.cuda.synchronize()
    buf = C.create_string_buffer(1 << 18)
    lib.patolette_b200_profile_json(buf, len(buf))
    lib.patolette_b200_profile_enable(0)
    lib.patolette_b200_ordered_counts(cnt, 0)
    lib.patolette_b200_split_counts(sc, 0)
    global SPLIT_COUNTS
    SPLIT_COUNTS = {"certified": int(sc[0]), "refused": int(sc[1]), "re_evaluated_exactly": int(sc[2])}
    return json.loads(buf.value.decode()), [int(c) for c in cnt]


def kernel_table(prof, top=18):
    ks = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])
    return {k: {"ms": round(v["ms"], 3), "launches": v["launches"],
                "GB/s": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 and v["bytes"] else None}
            for k, v in ks[:top]}


# FP64 operations per unit of the compute-bound kernels: k_color's ICtCp path = 9 glibc-exact
# pow() of ~75 FP64 instructions each + the matrices; the exact 1-NN = 8 flops per (pixel, candidate), K candidates
# for the reference's brute-force contract ("algorithmic") - the candidate lists evaluate ~15.
FP64_OPS = {"k_color_per_px": 9 * 75 + 60, "nn_per_px_candidate": 8}


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
    wl = WORKLOAD
    side = args.side or wl["w"]
    w = h = side
    n = w * h
    K = wl["K"]
    shard = world > 1 and not args.replicas and hasattr(pb, "init_sharding")
    if shard:
        pb.init_sharding(dist)  # NCCL communicator of the library (unique id broadcast through torch.distributed)
    # a sharded run: rank r brings pixels [first, first + count) of the ONE image (generated without the rest)
    first, count = (pb.shard_range(n, rank, world) if shard else (0, n))
    if shard:
        from synth import uniform_colors_slice
        colors = uniform_colors_slice(w, h, wl["seed"], first, count)
    else:
        colors = uniform_colors(w, h, wl["seed"] + (0 if world == 1 else rank))  # replicas: every rank its own image
    opts = _lib.QuantizationOptions(wl["dither"], False, wl["color_space"], wl["kmeans_niter"], 512 ** 2, False)
    code = C.c_int(0)
    palette = np.zeros((K, 3), order="F")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # host buffers of the end-to-end arm: 3 x count planes (= the reference's column-major layout), pinned
    h_in = torch.empty((3, count), dtype=torch.float64).pin_memory()
    for j in range(3):
        h_in[j].numpy()[:] = colors[:, j]
    h_map = torch.empty(count, dtype=torch.int64).pin_memory()

    # ---------------- device-resident arm (value) ----------------
    d_in = h_in.cuda()
    d_map = torch.empty(count, dtype=torch.int64, device="cuda")
    stream = torch.cuda.Stream()  # a non-blocking stream: the legacy NULL stream serialises against everything
    torch.cuda.set_stream(stream)
    lib.patolette_b200_set_stream(C.c_void_p(stream.cuda_stream), 1)

    def run_abi(in_ptr, map_ptr, device_io):
        if shard:
            lib.patolette_b200_sharded(w, h, in_ptr, None, K, C.byref(opts), palette.ctypes.data, map_ptr, device_io, C.byref(code))
        elif device_io:
            lib.patolette_b200_device(w, h, in_ptr, None, K, C.byref(opts), palette.ctypes.data, map_ptr, C.byref(code))
        else:
            lib.patolette(w, h, in_ptr, None, K, C.byref(opts), palette.ctypes.data, map_ptr, C.byref(code))
        assert code.value == 0, code.value

    def step_resident():
        run_abi(d_in.data_ptr(), d_map.data_ptr(), 1)

    def whole_map_sha(t_dev):
        """sha256 of the whole map on rank 0 (a sharded run gathers the slices over NCCL first)."""
        if not shard:
            return sha_map(t_dev.cpu()) if rank == 0 else None
        S = pb.shard_range(n, 0, world)[1] if world > 1 else n
        pad = torch.zeros(S, dtype=torch.int64, device="cuda")
        pad[:count] = t_dev
        allm = torch.empty(S * world, dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(allm, pad)
        return sha_map(allm[:n].cpu()) if rank == 0 else None

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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
    pal_resident = palette.copy()
    map_sha = whole_map_sha(d_map)

    # ---------------- end-to-end arm: host buffers through the reference ABI ----------------
    lib.patolette_b200_set_stream(None, 0)

    def step_e2e():
        run_abi(h_in.data_ptr(), h_map.data_ptr(), 0)

    for _ in range(min(max(args.warmup, 1), 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    e2e_s = reduce_max_ms(e2e_s * 1e3, dist) / 1e3
    e2e_stage = pb.last_timings()
    e2e_sha = whole_map_sha(h_map.cuda()) if shard else (sha_map(h_map) if rank == 0 else None)
    if rank == 0:
        assert e2e_sha == map_sha, "host and device arms disagree"
        assert np.array_equal(palette.view(np.uint64), pal_resident.view(np.uint64)), "host and device arms disagree (palette)"
    parity = check_golden(wl, side, pal_resident, map_sha) if rank == 0 else None

    e2e_extra = {}
    if world == 1 and not args.no_extras:
        # the same call a drop-in numpy user makes: pageable f64 [N,3] through quantize()
        reps = 2
        pb.quantize(w, h, colors, K, tile_size=0, **workload_kwargs(wl))
        t0 = time.perf_counter()
        for _ in range(reps):
            ok, pal2, map2, msg = pb.quantize(w, h, colors, K, tile_size=0, **workload_kwargs(wl))
            assert ok, msg
        dt = (time.perf_counter() - t0) / reps
        assert sha_map(map2) == map_sha, "quantize() and the device arm disagree"
        e2e_extra["pageable"] = {"value": n / dt / 1e6, "unit": "Mpixels/s", "ms_per_step": dt * 1e3, "steps": reps,
                                 "api": "patolette_b200.quantize(numpy f64 [N,3], pageable)",
                                 "stage_ms": {k: round(v, 3) for k, v in pb.last_timings().items()}}
        del pal2, map2
        if hasattr(pb, "quantize_u8"):
            # N1: uint8 RGB in, /255 on the device, u8 map out - palette and map must equal the f64 ABI's
            # on the image u8 / 255.0; timed on its own image (the uniform f64 workload is not 8-bit)
            rgb8 = np.random.default_rng(wl["seed"]).integers(0, 256, (n, 3), dtype=np.uint8)
            pb.quantize_u8(w, h, rgb8, K, **workload_kwargs(wl))
            t0 = time.perf_counter()
            for _ in range(reps):
                ok, pal8, map8, msg = pb.quantize_u8(w, h, rgb8, K, **workload_kwargs(wl))
                assert ok, msg
            dt8 = (time.perf_counter() - t0) / reps
            e2e_extra["u8"] = {"value": n / dt8 / 1e6, "unit": "Mpixels/s", "ms_per_step": dt8 * 1e3, "steps": reps,
                               "api": "patolette_b200.quantize_u8(numpy uint8 [N,3]) -> u8 map", "h2d_bytes_per_step": 3 * n,
                               "d2h_bytes_per_step": n + 24 * K,
                               "stage_ms": {k: round(v, 3) for k, v in pb.last_timings().items()}}
            # N3: the reference's DEFAULT call, tile_size = 512 - saliency weights on the device
            # (pb_saliency.cu), then the weighted pipeline; same 8-bit image
            pb.quantize_u8(w, h, rgb8, K, tile_size=512, **workload_kwargs(wl))
            t0 = time.perf_counter()
            for _ in range(reps):
                ok, pal8, map8, msg = pb.quantize_u8(w, h, rgb8, K, tile_size=512, **workload_kwargs(wl))
                assert ok, msg
            dts = (time.perf_counter() - t0) / reps
            e2e_extra["u8_saliency"] = {"value": n / dts / 1e6, "unit": "Mpixels/s", "ms_per_step": dts * 1e3, "steps": reps,
                                        "api": "patolette_b200.quantize_u8(numpy uint8 [N,3], tile_size=512) -> u8 map "
                                               "(saliency weights + weighted pipeline)",
                                        "note": "the dither stage of a saliency run is ~4x the unweighted one on this noise image "
                                                "(not so with caller-supplied weights of the same range); open question, "
                                                "profiles/r02_saliency.md",
                                        "stage_ms": {k: round(v, 3) for k, v in pb.last_timings().items()}}
            del rgb8, pal8, map8
    del colors

    # ---------------- per-kernel profile (untimed extra step) -> roofline ----------------
    lib.patolette_b200_set_stream(C.c_void_p(stream.cuda_stream), 1)
    prof, cnt = profile_step(lib, step_resident)
    split_routes = dict(SPLIT_COUNTS or {})
    fp64_peak = None
    if hasattr(lib, "patolette_b200_fp64_peak"):
        v = lib.patolette_b200_fp64_peak()
        fp64_peak = v if v > 0 else None
    lib.patolette_b200_set_stream(None, 0)
    ord_acc, ord_rep = cnt[0], cnt[1]
    peak, peak_src = measured_peaks()
    kernels = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])
    # `roofline` is stated for the dominant STREAMING kernel.  The dither recurrence and the ordered-sum resolve are
    # dependent chains ("latency ... report ns/pixel, no roofline claim"): they are reported as
    # `dither_ns_per_px` / inside the covariance stage and named in `dominant_overall` when they lead.
    latency_bound = ("k_riemersma", "k_ord_resolve")
    top_name, top = next((kv for kv in kernels if not kv[0].startswith(latency_bound)), kernels[0])
    # Algorithmic bytes of the resolve kernels = what the sequential walk must read: one 32-byte record per
    # (block, chain) it walks + the 4 KB of terms of every block it replays.  The pixel
    # planes are attributed to the kernels that stream them.
    resolve_bytes = (ord_acc + ord_rep) * 32.0 + ord_rep * 4096.0
    resolve_ms = sum(v["ms"] for k, v in prof.items() if k.startswith("k_ord_resolve"))
    for k, v in prof.items():
        if k.startswith("k_ord_resolve") and resolve_ms > 0:
            v["bytes"] = resolve_bytes * v["ms"] / resolve_ms
    gbs = top["bytes"] / (top["ms"] * 1e-3) / 1e9 if top["ms"] > 0 else 0.0
    ncu = {}
    for name in ("r02_ncu_summary.json", "r01_ncu_summary.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                ncu = json.load(f)
            ncu["_file"] = name
            break
        except Exception:
            pass

    # DRAM bytes per launch of the top kernel from the committed `ncu --set full` capture (tools/ncu_summarise.py); the
    # capture names the image side it was taken at - another size is scaled by the pixel ratio and says so
    traffic, traffic_src = None, None
    ent = ncu.get(top_name)
    if ent and ent.get("dram_bytes_per_launch"):
        cap_side = ent.get("side") or ncu.get("side")
        per_launch_px = ent.get("pixels_per_launch")
        traffic = ent["dram_bytes_per_launch"]
        traffic_src = f"profiles/{ncu.get('_file')}: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of {top_name}"
        my_px = (top["bytes"] / top["launches"]) / max(ent.get("algorithmic_bytes_per_pixel", 0) or 1e30, 1e-30) if ent.get("algorithmic_bytes_per_pixel") else None
        if per_launch_px and my_px and abs(my_px / per_launch_px - 1) > 0.02:
            traffic = traffic * my_px / per_launch_px
            traffic_src += f" (captured on a launch of {per_launch_px / 1e6:.1f} M pixels at side {cap_side}; scaled to this run's average launch of {my_px / 1e6:.1f} M pixels)"

    def stage_roofline(names, bytes_per_step, p=prof):
        ms = sum(v["ms"] for k, v in p.items() if any(k.startswith(nm) for nm in names))
        g = bytes_per_step / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        return {"kernels": names, "ms_per_step": round(ms, 3), "algorithmic_bytes_per_step": bytes_per_step,
                "achieved": round(g, 1), "unit": "GB/s", "frac": round(g / peak, 4)}

    def fp64_entry(name, ops, p=prof):
        ms = sum(v["ms"] for k, v in p.items() if k.startswith(name))
        if not ms:
            return None
        t = ops / (ms * 1e-3) / 1e12
        return {"ms_per_step": round(ms, 3), "fp64_ops_per_step": ops, "achieved_TFLOPs": round(t, 3),
                "peak_TFLOPs": fp64_peak, "frac": round(t / fp64_peak, 4) if fp64_peak else None}

    # covariance stage: every ordered-sum pass streams 24 (32 weighted) B per pixel-visit - the bytes of the
    # summary sweeps are the pass bytes; blocksum re-reads are NOT algorithmic
    pass_bytes = sum(v["bytes"] for k, v in prof.items() if k.startswith("k_ord_fast") or k.startswith("k_ord_summary_"))
    dither_ms = sum(v["ms"] for k, v in prof.items() if k.startswith("k_riemersma"))
    roofline = {"bound": "hbm", "kernel": top_name, "dominant_overall": kernels[0][0], "achieved": gbs, "peak": peak, "unit": "GB/s",
                "frac": gbs / peak, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peak_src,
                "launches_per_step": top["launches"], "ms_per_step_in_kernel": top["ms"],
                "algorithmic_bytes_per_step": top["bytes"],
                "top_kernel_launch_ms": [round(x, 3) for x in top.get("each", [])][:40],
                "stages": {
                    "covariance (ordered mean + centred passes: k_ord_*)": stage_roofline(["k_ord_"], pass_bytes),
                    "projection + bucket sort + partition (k_dots_minmax, k_buckets, k_tile_*, k_scatter)":
                        stage_roofline(["k_dots_minmax", "k_buckets", "k_tile_", "k_scatter", "k_class_start"],
                              sum(v["bytes"] for k, v in prof.items() if k in ("k_dots_minmax", "k_buckets", "k_buckets_hist") or k.startswith("k_scatter"))),
                    "per-bucket ordered sums (k_bucket_chains_*)": stage_roofline(["k_bucket_chains"], sum(v["bytes"] for k, v in prof.items() if k.startswith("k_bucket_chains"))),
                    "colour transforms (k_color)": stage_roofline(["k_color"], sum(v["bytes"] for k, v in prof.items() if k == "k_color")),
                    "dither (hilbert rank, permute, k_riemersma_*, unpermute)": stage_roofline(["k_hilbert", "k_permute", "k_riemersma", "k_unpermute"], (2 * (24 + 8) + 32.0) * n),
                },
                "fp64": {"peak_source": "DFMA throughput kernel run by this bench (patolette_b200_fp64_peak)" if fp64_peak else None,
                         "k_color": fp64_entry("k_color", FP64_OPS["k_color_per_px"] * 2.0 * n)},
                "dither_ns_per_px": round(dither_ms * 1e6 / n, 4) if dither_ms else None,
                "kernels": kernel_table(prof)}

    # ---------------- BASELINE's other single-GPU configurations (device-resident, a few steps) ----------------
    other = {}
    if world == 1 and not args.no_extras and not args.side:
        del d_in, d_map, h_in, h_map
        for tag, ow in OTHER.items():
            on = ow["w"] * ow["h"]
            oc = uniform_colors(ow["w"], ow["h"], ow["seed"])
            od_in = torch.from_numpy(np.ascontiguousarray(oc.T)).cuda()
            del oc
            od_map = torch.empty(on, dtype=torch.int64, device="cuda")
            oopts = _lib.QuantizationOptions(ow["dither"], False, ow["color_space"], ow["kmeans_niter"], 512 ** 2, False)
            opal = np.zeros((ow["K"], 3), order="F")
            lib.patolette_b200_set_stream(C.c_void_p(stream.cuda_stream), 1)

            def ostep():
                lib.patolette_b200_device(ow["w"], ow["h"], od_in.data_ptr(), None, ow["K"], C.byref(oopts), opal.ctypes.data,
                                          od_map.data_ptr(), C.byref(code))
                assert code.value == 0, code.value

            for _ in range(3):
                ostep()
            torch.cuda.synchronize()
            osteps = 10
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(osteps):
                ostep()
            b.record(stream)
            torch.cuda.synchronize()
            oms = a.elapsed_time(b) / osteps
            oprof, ocnt = profile_step(lib, ostep)
            lib.patolette_b200_set_stream(None, 0)
            rec = {"workload": ow["name"], "value": on / (oms * 1e-3) / 1e6, "unit": "Mpixels/s (device-resident)",
                   "ms_per_step": round(oms, 3), "steps": osteps,
                   "parity": check_golden(ow, ow["w"], opal, sha_map(od_map.cpu())),
                   "stage_ms": {k: round(v, 3) for k, v in pb.last_timings().items()},
                   "covariance_stage": stage_roofline(["k_ord_"], sum(v["bytes"] for k, v in oprof.items() if k.startswith("k_ord_fast") or k.startswith("k_ord_summary_")), oprof),
                   "kernels": kernel_table(oprof, 12)}
            if not ow["dither"]:
                nn_ms = sum(v["ms"] for k, v in oprof.items() if k.startswith("k_nearest"))
                rec["assignment (k_nearest: exact f64 1-NN over per-cell candidate lists)"] = {
                    "hbm": stage_roofline(["k_nearest"], 32.0 * on, oprof),
                    "fp64_algorithmic (brute-force contract: 8 flops x K candidates per pixel; the lists evaluate ~15)":
                        fp64_entry("k_nearest", FP64_OPS["nn_per_px_candidate"] * ow["K"] * 1.0 * on, oprof)}
            else:
                dms = sum(v["ms"] for k, v in oprof.items() if k.startswith("k_riemersma"))
                rec["dither_ns_per_px"] = round(dms * 1e6 / on, 4)
            other[tag] = rec
            del od_in, od_map

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return None
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_baseline, _ = time_reference(REFERENCE_SIDE if not args.side else min(args.side, REFERENCE_SIDE), 1, 0, 60.0)
    jobs = 1 if (shard or world == 1) else world
    line = {
        "metric": METRIC,
        "value": aggregate_throughput(n, jobs, ms_step), "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak" if (world > 1 and not shard) else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"] if not args.side else f"{side}x{side} debug size of the C4 options", "K": K,
                   "parallelism": "single GPU" if world == 1 else (
                       pb.sharding_description() if shard else f"replicas x{world} (one image per rank, no collective)"),
                   "l2": f"inputs ({24 * n / 1e6:.0f} MB) larger than L2; no flush needed",
                   "mode": "exact (bit-identical to the reference CPU path)",
                   "parity": parity},
        "e2e": {"value": jobs * n / e2e_s / 1e6, "unit": "Mpixels/s", "h2d_bytes_per_step": 24 * n * jobs,
                "d2h_bytes_per_step": (8 * n + 24 * K) * jobs, "ms_per_step": e2e_s * 1e3, "host_buffers": "pinned",
                "api": ("patolette_b200_sharded(): every rank passes its pixel slice (f64 planes) and receives the map of its slice; bytes are the job's totals"
                        if shard else "patolette() C ABI (lib/include/patolette.h:22-32), f64 planes in, size_t map out"),
                "stage_ms": {k: round(v, 3) for k, v in e2e_stage.items()}, **e2e_extra},
        "gpu_launches": launches * args.steps if launches else None,
        "gpu_launches_per_step": launches,
        "stage_ms": {k: round(v, 3) for k, v in stage.items()},
        "roofline": roofline, "clocks": clocks,
        "ordered_sums": {"blocks_accepted": ord_acc, "blocks_replayed": ord_rep,
                         "replay_frac": ord_rep / max(ord_acc + ord_rep, 1),
                         "replay_reasons": {"flag": cnt[2], "binade_guess": cnt[3], "bounds": cnt[4]},
                         "accepted_two_parity": cnt[7],
                         "resolve_warp_Mcycles": {"scan_walk": round(cnt[8] / 1e6, 2), "two_parity": round(cnt[9] / 1e6, 2),
                                                  "replay": round(cnt[10] / 1e6, 2), "slowest_warp": round(cnt[12] / 1e6, 3)},
                         "records_walked_singly": cnt[11],
                         "fast_pairs": cnt[13], "general_pairs": cnt[14]},
        "split_routes": split_routes,
        "other_configs": other,
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
