"""Where an edit + evaluation of a 16-chain forest goes: dphy_forest_apply_rows (one branch reform per chain), evaluation, read-back --
host wall clock per phase with a synchronize in between.  usage: python tools/edit_timing.py [chains] [steps]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import delphy_b200 as db
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ctx = db.Context(0)
ems, tabs = [], []
for c in range(chains):
    e, s, info = db.synth_generate(db.synth_params(4, seed=20251017 + c))
    ems.append(e); tabs.append(db.DeviceSites(ctx, s))
fo = db.Forest(ctx, ems, tabs, sites_index=np.arange(chains))
fo.eval_log_G(); fo.log_G()
rng = np.random.default_rng(5)
acc = {"rows (python)": 0.0, "apply_rows": 0.0, "eval": 0.0, "log_G": 0.0}
for it in range(steps + 2):
    t0 = time.perf_counter()
    rows = []
    for k, e in enumerate(ems):
        has = np.nonzero((np.diff(e.mut_off) > 0) & (e.parent >= 0))[0]
        v = int(has[rng.integers(len(has))]); m0, m1 = int(e.mut_off[v]), int(e.mut_off[v + 1])
        lo, hi = e.t[e.parent[v]], e.t[v]
        e.mut_t[m0:m1] = np.sort(lo + (hi - lo) * rng.random(m1 - m0))
        rows.append(db.node_row(k, e, v))
    t1 = time.perf_counter()
    fo.apply_rows(rows); ctx.synchronize()
    t2 = time.perf_counter()
    fo.eval_log_G(); ctx.synchronize()
    t3 = time.perf_counter()
    fo.log_G()
    t4 = time.perf_counter()
    if it >= 2:
        for k, v in zip(acc, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)): acc[k] += v
for k, v in acc.items(): print(f"{k:16s} {v / steps * 1e3:8.3f} ms per step")
