#!/bin/bash
# one ncu --set full capture of the grouped SPR kernels.  usage: tools/gpu_ncu_spr.sh <tag> [kernel regex]
OUT=gpurun_out/${1:-ns}; mkdir -p $OUT
K=${2:-spr_gemit_kernel}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 2 -c 1 \
  -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-partitioned --no-mcmc --evals-per-step 2 --spr-batches-per-step 1 > $OUT/ncu.log 2>&1
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/ncu_raw.csv 2>/dev/null
ncu -i $OUT/prof.ncu-rep --page source --csv > $OUT/ncu_source.csv 2>/dev/null
python tools/summarize_ncu.py $OUT/ncu_raw.csv > $OUT/ncu_summary.txt 2>&1
cat $OUT/ncu_summary.txt
python - <<PY
import csv
rows=list(csv.reader(open("$OUT/ncu_source.csv")))
hdr=rows[0]
print(hdr[:12])
idx={h:i for i,h in enumerate(hdr)}
key=[h for h in hdr if 'Warp Stall Sampling (All' in h or h=='# Samples' or 'Samples' in h][:1]
print(key)
k=idx[key[0]] if key else None
src=idx.get('Source')
body=[r for r in rows[1:] if len(r)>max(k or 0, src or 0)]
def val(r):
    try: return float(r[k])
    except: return 0.0
body.sort(key=val, reverse=True)
tot=sum(val(r) for r in body)
print("total samples", tot)
for r in body[:40]:
    print(f"{val(r)/max(tot,1)*100:5.1f}%  {r[src][:150]}")
PY
