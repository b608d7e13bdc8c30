#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` export (tools/ncu_capture.sh) into profiles/<tag>_ncu_summary.json:
per captured launch the numbers the roofline discussion needs, and per kernel the average DRAM traffic per
launch (bench.py reads `dram_bytes_per_launch` for the `roofline.traffic` field).
   tools/ncu_summarise.py gpurun_out/r01g_ncu_raw.csv.gz profiles/r01_ncu_summary.json"""
import csv, gzip, io, json, re, sys

src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.DictReader(io.TextIOWrapper(gzip.open(src))))
units, rows = rows[0], rows[1:]
KEEP = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3,
         "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def prof_name(kernel: str) -> str:
    m = re.search(r"(k_[a-z0-9_]+)", kernel)
    if not m:
        return kernel
    base = m.group(1)
    t = re.search(r"<\D*(\d)", kernel)  # first template argument: KIND (0 mean, 1 centred)
    if base in ("k_ord_summary", "k_ord_resolve", "k_ord_blocksum") and t:
        return base + ("_mean" if t.group(1) == "0" else "_centered")
    return base


launches, per = [], {}
for r in rows:
    rec = {"kernel": r["Kernel Name"][:90], "name": prof_name(r["Kernel Name"])}
    for k, short in KEEP.items():
        if k in r and r[k] != "":
            v = float(r[k].replace(",", ""))
            u = units.get(k, "")
            if short.startswith("dram_r") or short.startswith("dram_w") or short == "time":
                v *= SCALE.get(u, 1.0)
            rec[short] = v
    if "dram_read" in rec:
        rec["dram_bytes"] = rec["dram_read"] + rec.get("dram_write", 0.0)
        if rec.get("time"):
            rec["dram_GBps"] = round(rec["dram_bytes"] / rec["time"] / 1e9, 1)
    launches.append(rec)
    per.setdefault(rec["name"], []).append(rec)
out = {"source": src, "note": "ncu --set full --clock-control none; times are cold-cache and serialised: compare shares, not absolutes",
       "launches": launches}
for name, L in per.items():
    out[name] = {"captured_launches": len(L),
                 "dram_bytes_per_launch": sum(x.get("dram_bytes", 0.0) for x in L) / len(L),
                 "time_s_per_launch": sum(x.get("time", 0.0) for x in L) / len(L),
                 "registers": L[0].get("registers"), "warps_active_pct": round(sum(x.get("warps_active_pct", 0) for x in L) / len(L), 1),
                 "issue_active_pct": round(sum(x.get("issue_active_pct", 0) for x in L) / len(L), 1),
                 "fp64_pipe_pct": round(sum(x.get("fp64_pipe_pct", 0) for x in L) / len(L), 1)}
json.dump(out, open(dst, "w"), indent=1)
for name in per:
    print(name, {k: (round(v, 6) if isinstance(v, float) else v) for k, v in out[name].items()})
