#!/usr/bin/env python
"""Where a kernel's warp-stall samples fall: usage  tools/ncu_hot.py <ncu_source.csv> <kernel substring> [min pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]; thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "body": []}; sections.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None: cur["hdr"] = r; continue
    cur["body"].append(r)
sec = [s for s in sections if want in s["name"]][0]
idx = {x: i for i, x in enumerate(sec["hdr"])}
k, src = idx["Warp Stall Sampling (All Samples)"], idx["Source"]
ex = idx.get("Instructions Executed")
def val(r, c=k):
    try: return float(r[c])
    except Exception: return 0.0
body = sec["body"]; tot = sum(val(r) for r in body) or 1.0
print(sec["name"][:90], "samples", tot, "instructions", len(body))
step = max(1, len(body) // 24)
for a in range(0, len(body), step):
    print(f"  [{a:5d},{min(a+step,len(body)):5d})  {sum(val(r) for r in body[a:a+step])/tot*100:5.1f}%   inst_exec {sum(val(r, ex) for r in body[a:a+step]):.0f}")
for i, r in enumerate(body):
    if val(r) / tot * 100 >= thr: print(f"{i:5d} {val(r)/tot*100:5.1f}%  {r[src][:120]}")
