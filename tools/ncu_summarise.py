#!/usr/bin/env python
"""Condense `ncu --set full` captures into profiles/<tag>_ncu_summary.json: per captured launch the numbers the roofline
discussion needs, and per kernel the average DRAM traffic per launch (bench.py reads `dram_bytes_per_launch`,
`pixels_per_launch` and `algorithmic_bytes_per_pixel` for its `roofline.traffic` field).

   tools/ncu_summarise.py OUT.json REPORT[:side] [REPORT[:side] ...] [--bpp kernel=bytes ...]

REPORT is a .ncu-rep (read through `ncu -i ... --page raw --csv`) or an exported .csv / .csv.gz; `side` = image side the
capture was run at.  The FIRST report that holds a kernel defines its per-kernel entry (put the capture taken at the
bench's own size first).  --bpp gives a kernel's algorithmic bytes per pixel-visit, from which the launch's pixel count
is derived out of... nothing: pixels come from the grid where the kernel's launch geometry fixes them, else stay null."""
import csv, gzip, io, json, re, subprocess, sys

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
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "lsu_wavefronts_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared_bank_conflicts",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3,
         "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def prof_name(kernel: str) -> str:
    m = re.search(r"(k_[a-z0-9_]+)", kernel)
    if not m:
        return kernel
    base = m.group(1)
    t = re.search(r"<\D*(\d)", kernel)  # first template argument: KIND (0 mean, 1 centred)
    if base in ("k_ord_summary", "k_ord_resolve", "k_ord_blocksum", "k_ord_fast") and t:
        return base + ("_mean" if t.group(1) == "0" else "_centered")
    return base


def read_rows(path):
    if path.endswith(".ncu-rep"):
        text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        f = io.StringIO(text)
    elif path.endswith(".gz"):
        f = io.TextIOWrapper(gzip.open(path))
    else:
        f = open(path)
    rows = list(csv.DictReader(f))
    return rows[0], rows[1:]


def main():
    out_path, args = sys.argv[1], sys.argv[2:]
    bpp, reports = {}, []
    i = 0
    while i < len(args):
        if args[i] == "--bpp":
            k, v = args[i + 1].split("=")
            bpp[k] = float(v)
            i += 2
        else:
            reports.append(args[i])
            i += 1
    out = {"note": "ncu --set full --clock-control none; times are cold-cache and serialised: compare shares, not absolutes",
           "sources": [], "launches": []}
    per = {}
    for spec in reports:
        path, _, side = spec.partition(":")
        side = int(side) if side else None
        units, rows = read_rows(path)
        out["sources"].append({"file": path, "side": side, "launches": len(rows)})
        for r in rows:
            rec = {"kernel": r["Kernel Name"][:90], "name": prof_name(r["Kernel Name"]), "side": side, "source": path}
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
            out["launches"].append(rec)
            per.setdefault((rec["name"], path), []).append(rec)
    for (name, path), L in per.items():
        if name in out:
            continue  # an earlier report already defines this kernel
        ent = {"captured_launches": len(L), "side": L[0]["side"], "source": path,
               "dram_bytes_per_launch": sum(x.get("dram_bytes", 0.0) for x in L) / len(L),
               "time_s_per_launch": sum(x.get("time", 0.0) for x in L) / len(L),
               "registers": L[0].get("registers")}
        for key in ("warps_active_pct", "issue_active_pct", "fp64_pipe_pct", "dram_throughput_pct", "lsu_wavefronts_pct"):
            ent[key] = round(sum(x.get(key, 0) for x in L) / len(L), 1)
        if name in bpp:  # DRAM traffic ~ algorithmic bytes for the streaming kernels: the launch's pixel count
            ent["algorithmic_bytes_per_pixel"] = bpp[name]
        out[name] = ent
    json.dump(out, open(out_path, "w"), indent=1)
    for name, ent in out.items():
        if isinstance(ent, dict) and "captured_launches" in ent:
            print(name, {k: (round(v, 6) if isinstance(v, float) else v) for k, v in ent.items() if k != "source"})


if __name__ == "__main__":
    main()
