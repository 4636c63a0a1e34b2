import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import delphy_b200 as db
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx = db.Context(0)
ctx.set_log_G_path("general")   # the debug masks act on the general schedule
ems, tabs = [], []
for c in range(chains):
    e, s, info = db.synth_generate(db.synth_params(cfg, seed=20251017 + c))
    ems.append(e); tabs.append(db.DeviceSites(ctx, s))
fo = db.Forest(ctx, ems, tabs, sites_index=np.arange(chains))
st = torch.cuda.ExternalStream(ctx.stream)
print("alg bytes", fo.log_G_algorithmic_bytes, "device bytes", fo.device_bytes, "max_depth", info["max_depth"])
for mask in [0, 1, 2, 4, 8, 15, 32, 47]:
    os.environ["DPHY_DEBUG_MASK"] = str(mask)
    for _ in range(3): fo.eval_log_G()
    ctx.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(10): fo.eval_log_G()
    b.record(st); ctx.synchronize()
    print(f"mask={mask:2d} eval={a.elapsed_time(b)/10*1e3:8.1f} us")
