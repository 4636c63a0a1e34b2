"""Cost of an evo-model change followed by a log-G re-evaluation (the reference's global-move cycle): dphy_sites_set_evo on
every site table + one evaluation of the forest + scalar download.  usage: python tools/set_evo_timing.py [chains] [cfg]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import delphy_b200 as db
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx = db.Context(0)
ems, tabs, hs = [], [], []
for c in range(chains):
    e, s, info = db.synth_generate(db.synth_params(cfg, seed=20251017 + c))
    ems.append(e); tabs.append(db.DeviceSites(ctx, s)); hs.append(s)
fo = db.Forest(ctx, ems, tabs, sites_index=np.arange(chains))
fo.eval_log_G(); fo.log_G()
for rep in range(3):
    t0 = time.perf_counter()
    for k in range(chains):
        hs[k].mu = hs[k].mu * 1.01
        tabs[k].set_evo(mu=hs[k].mu)
    t1 = time.perf_counter()
    fo.eval_log_G(); out = fo.log_G()
    t2 = time.perf_counter()
    print(f"set_evo x{chains}: {1e3*(t1-t0):.3f} ms ({1e6*(t1-t0)/chains:.1f} us per table)   eval + get: {1e3*(t2-t1):.3f} ms   -> {chains/(t2-t0):.0f} evals/s after a model change")
