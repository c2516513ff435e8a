#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel from `ncu --page source --csv` output.
usage: ncu_top_stalls.py <source.csv> [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = None
data = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot, "instructions", len(data))
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stall_cols}
print("by reason:", {k: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))
for i in order[:top_n]:
    r = data[i]
    s = int(r[ix["# Samples"]] or 0)
    reasons = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"{100*s/tot:5.1f}%  #{i:4d} {r[ix['Source']].strip()[:70]:70s} exec={r[ix['Instructions Executed']]:>10s} shw={r[ix['L1 Wavefronts Shared']]:>10s}/{r[ix['L1 Wavefronts Shared Ideal']]:>10s} {reasons}")
