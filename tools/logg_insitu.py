"""Is a folded log-G evaluation slower back to back than alone?  usage: python tools/logg_insitu.py [chains] [cfg]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import delphy_b200 as db
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx = db.Context(0)
ems, tabs = [], []
for c in range(chains):
    e, s, info = db.synth_generate(db.synth_params(cfg, seed=20251017 + c))
    ems.append(e); tabs.append(db.DeviceSites(ctx, s))
fo = db.Forest(ctx, ems, tabs, sites_index=np.arange(chains))
st = torch.cuda.ExternalStream(ctx.stream)
def ev(): return torch.cuda.Event(enable_timing=True)
for _ in range(10): fo.eval_log_G()
ctx.synchronize()
# (1) back to back
a, b = ev(), ev()
a.record(st)
for _ in range(200): fo.eval_log_G()
b.record(st); ctx.synchronize()
print(f"back to back: {a.elapsed_time(b)/200*1e3:.1f} us per evaluation")
# (2) one at a time, queue drained in between (includes one launch latency)
ts = []
for _ in range(50):
    a, b = ev(), ev()
    a.record(st); fo.eval_log_G(); b.record(st); ctx.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
print(f"isolated: median {np.median(ts):.1f} us, min {np.min(ts):.1f} us")
# (3) groups of 4 with per-evaluation events inside a saturated queue
evs = [ev() for _ in range(41)]
evs[0].record(st)
for i in range(40):
    fo.eval_log_G(); evs[i + 1].record(st)
ctx.synchronize()
d = [evs[i].elapsed_time(evs[i + 1]) * 1e3 for i in range(40)]
print("per-evaluation inside a saturated queue:", " ".join(f"{x:.0f}" for x in d[:20]))
# (4) a large unrelated memset between evaluations (evicts L2): does the evaluation get faster or slower?
big = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
with torch.cuda.stream(st):
    ts = []
    for _ in range(20):
        big.zero_()
        a, b = ev(), ev()
        a.record(st); fo.eval_log_G(); b.record(st)
        ts.append((a, b))
ctx.synchronize()
print(f"after an L2 flush: median {np.median([x.elapsed_time(y)*1e3 for x, y in ts]):.1f} us")
