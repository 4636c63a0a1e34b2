"""Time of the per-cycle tallies of a 16-chain forest: site tallies (Ttwiddle_l + num_muts_l of every tree, one call),
Ttwiddle_beta_a per tree, integer tallies.  usage: python tools/tally_timing.py [chains] [cfg]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import delphy_b200 as db
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx = db.Context(0)
ems, tabs = [], []
for c in range(chains):
    e, s, info = db.synth_generate(db.synth_params(cfg, seed=20251017 + c))
    ems.append(e); tabs.append(db.DeviceSites(ctx, s))
fo = db.Forest(ctx, ems, tabs, sites_index=np.arange(chains))
fo.log_G()
for name, fn in (("site_tallies (all trees, one call)", lambda: fo.site_tallies()),
                 ("Ttwiddle_l per tree x%d" % chains, lambda: [fo.Ttwiddle_l(k, want_T_l_a=False) for k in range(chains)]),
                 ("Ttwiddle_beta_a per tree x%d" % chains, lambda: [fo.Ttwiddle_beta_a(k) for k in range(chains)]),
                 ("tallies (num_muts, num_muts_ab, T; first call runs the general pass)", lambda: fo.tallies()),
                 ("tallies again", lambda: fo.tallies())):
    fn()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    print(f"{name}: {(time.perf_counter()-t0)/5*1e3:.3f} ms")
