#!/usr/bin/env python
"""Summarise an ncu --page source --print-source sass CSV: stall-reason totals and the
hottest SASS instructions / regions. usage: scripts_stalls.py sass.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: 0 for h in stall_cols}
data = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: samples = int(r[col["# Samples"]] or 0)
    except ValueError: continue
    ex = int(r[col["Instructions Executed"]] or 0)
    st = {h: int(r[col[h]] or 0) for h in stall_cols}
    for h in stall_cols: tot[h] += st[h]
    data.append((samples, ex, r[col["Source"]], st, len(data)))
S = sum(d[0] for d in data)
print("total samples", S, "instructions executed", sum(d[1] for d in data))
for h, v in sorted(tot.items(), key=lambda x: -x[1])[:10]:
    print(f"  {h:28s} {v:8d} {100*v/max(1,S):5.1f}%")
print("--- hottest instructions")
for s, ex, src, st, idx in sorted(data, key=lambda d: -d[0])[:top]:
    main = max(st.items(), key=lambda x: x[1])
    print(f"{idx:5d} {s:7d} {100*s/S:5.1f}% ex={ex:10d} {main[0][6:]:12s} {src[:90]}")
