#!/usr/bin/env python
"""Top stall-sampled SASS instructions of one kernel section of an `ncu --page source --csv` export.
   tools/ncu_top_sass.py <file.csv.gz> <section-index> [top]"""
import csv, gzip, io, sys
path, sec, top = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 40
lines = io.TextIOWrapper(gzip.open(path)).read().split("\n")
starts = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
seen, secs = set(), []
for s in starts:  # every kernel appears twice (two views); keep the first
    secs.append(s)
s0 = secs[sec]
s1 = secs[sec + 1] if sec + 1 < len(secs) else len(lines)
print(lines[s0][:160])
rows = list(csv.reader(lines[s0 + 1:s1]))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[1:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]]) for r in body)
texec = sum(int(r[ix["Instructions Executed"]]) for r in body)
print("instructions", len(body), "samples", tot, "warp-instr executed", texec)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]]) for r in body) for h in stall_cols}
print({k[6:]: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]]))[:top]
for i in sorted(order):
    r = body[i]
    st = sorted(((int(r[ix[h]]), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {int(r[ix['# Samples']]):7d} {100*int(r[ix['# Samples']])/max(tot,1):5.1f}% exec {int(r[ix['Instructions Executed']]):9d}  {r[ix['Source']].strip()[:70]:70s} {st}")
