#!/usr/bin/env python
"""Per-source-line sample totals from `ncu --page source --csv --print-source cuda,sass`.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | scripts_lines.py [topN]"""
import csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
fname = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; si = hdr.index("# Samples"); ei = hdr.index("Instructions Executed"); continue
    if hdr and len(r) > ei and r[0].isdigit():
        try: s = int(r[si] or 0); ex = int(r[ei] or 0)
        except ValueError: continue
        st = {}
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h and i < len(r) and r[i]:
                try: st[h[6:]] = int(r[i])
                except ValueError: pass
        out.append((s, ex, fname, int(r[0]), r[1].strip()[:80], st))
S = sum(o[0] for o in out)
print("total samples", S, "instr", sum(o[1] for o in out))
for s, ex, f, ln, src, st in sorted(out, key=lambda x: -x[0])[:top]:
    tops = ",".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:3] if v)
    print(f"{s:7d} {100*s/max(S,1):5.1f}% ex={ex:10d} {f}:{ln:<4d} {src:80s} {tops}")
