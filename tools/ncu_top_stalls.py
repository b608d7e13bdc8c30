#!/usr/bin/env python
"""Top stall locations of one kernel from an `ncu --page source --csv --print-source sass` export.
   ncu -i rep.ncu-rep --page source --csv --print-source sass --kernel-name regex:NAME --launch-count 1 > src.csv
   tools/ncu_top_stalls.py src.csv [n]"""
import csv, sys
f = open(sys.argv[1]); next(f)
rows = [r for r in csv.DictReader(f) if (r.get('# Samples') or '').isdigit()]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
tot = sum(int(r['# Samples']) for r in rows)
print('total samples', tot, 'instructions', len(rows))
stalls = [k for k in rows[0].keys() if k.startswith('stall_') and 'Not Issued' not in k]
for r in sorted(rows, key=lambda r: -int(r['# Samples']))[:n]:
    st = {k[6:]: int(r[k]) for k in stalls if r[k] and r[k].isdigit() and int(r[k]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(r['Address'][-5:], r['# Samples'].rjust(6), r['Instructions Executed'].rjust(9), r['Source'][:64].ljust(64), st)
agg = {}
for r in rows:
    for k in stalls:
        if r[k] and r[k].isdigit():
            agg[k[6:]] = agg.get(k[6:], 0) + int(r[k])
print([(k, round(100 * v / max(tot, 1), 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]])
