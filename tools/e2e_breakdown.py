"""Where the end-to-end time goes: upload (pinned / pageable), eval, download.  usage: python tools/e2e_breakdown.py [chains] [cfg]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import delphy_b200 as db
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx = db.Context(0)
ems, tabs = [], []
for c in range(chains):
    e, s, info = db.synth_generate(db.synth_params(cfg, seed=20251017 + c))
    ems.append(e); tabs.append(db.DeviceSites(ctx, s))
pinned = [e.pinned(ctx) for e in ems]
nbytes = sum(getattr(e, k).nbytes for e in ems for k in db.HostEmat.FIELDS_I32 + db.HostEmat.FIELDS_U8 + db.HostEmat.FIELDS_F64)
for name, src in (("pinned", pinned), ("pageable", ems)):
    for rep in range(4):
        ctx.synchronize()
        t0 = time.perf_counter()
        fo = db.Forest(ctx, src, tabs, sites_index=np.arange(chains))
        t1 = time.perf_counter()
        ctx.synchronize()
        t2 = time.perf_counter()
        fo.eval_log_G(); ctx.synchronize()
        t3 = time.perf_counter()
        out = fo.log_G()
        t4 = time.perf_counter()
        fo.close(); ctx.synchronize()
        t5 = time.perf_counter()
    print(f"{name:9s} chains={chains} bytes={nbytes/1e6:.1f}MB upload_call={1e3*(t1-t0):.3f} ms (+sync {1e3*(t2-t1):.3f}) -> {nbytes/(t2-t0)/1e9:.1f} GB/s  eval={1e3*(t3-t2):.3f}  get={1e3*(t4-t3):.3f}  close={1e3*(t5-t4):.3f}  total={1e3*(t5-t0):.3f} ms -> {chains/(t5-t0):.0f} evals/s")
# raw PCIe rate for reference: one big pinned -> device copy
import torch
a = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); b.copy_(a, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"raw pinned H2D of {nbytes/1e6:.1f} MB: {1e3*(t1-t0):.3f} ms -> {nbytes/(t1-t0)/1e9:.1f} GB/s")
