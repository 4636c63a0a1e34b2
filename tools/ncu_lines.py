#!/usr/bin/env python
"""Per CUDA source line: stall samples + instructions executed, from an .ncu-rep.  usage: ncu_lines.py rep kernel_regex [topN]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
cur_file = None; hdr = None; rows = []
seen_kernel = 0
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0] not in ("", "-") and r[0].isdigit():
        si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
        rows.append((cur_file, int(r[0]), r[1].strip(), int(r[si]) if r[si].isdigit() else 0, int(r[ii]) if r[ii].isdigit() else 0))
tot = sum(x[3] for x in rows) or 1
toti = sum(x[4] for x in rows) or 1
print(f"total samples {tot}, total warp-instructions {toti}")
agg = {}
for f, ln, src, s, i in rows:
    k = (f, ln)
    a = agg.setdefault(k, [src, 0, 0]); a[1] += s; a[2] += i
for (f, ln), (src, s, i) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"{100*s/tot:5.1f}% smp {100*i/toti:5.1f}% ins  {f}:{ln:<4d} {src[:110]}")
