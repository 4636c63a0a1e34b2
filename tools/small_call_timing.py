"""Per-call latency of the drop-in's device round trips on small trees (what an MCMC cycle of the reference pays per hot call)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import delphy_b200 as db

for cfg, kw in ((1, {}), (3, dict(num_tips=2000)), (3, {})):
    emat, sites, info = db.synth_generate(db.synth_params(cfg, **kw))
    with db.Context(0) as ctx:
        ds = db.DeviceSites(ctx, sites)
        pe = emat.pinned(ctx)
        for src, name in ((emat, "pageable"), (pe, "pinned")):
            for _ in range(5):
                fo = db.Forest(ctx, [src], [ds]); fo.log_G(); fo.close()
            t0 = time.perf_counter(); n = 50
            tu = tg = 0.0
            for _ in range(n):
                a = time.perf_counter()
                fo = db.Forest(ctx, [src], [ds])
                b = time.perf_counter()
                fo.log_G()
                c = time.perf_counter()
                fo.close()
                tu += b - a; tg += c - b
            print(f"tips={(emat.num_nodes+1)//2} {name}: upload {tu/n*1e3:.3f} ms, get_log_G {tg/n*1e3:.3f} ms, total {(time.perf_counter()-t0)/n*1e3:.3f} ms")
        fo = db.Forest(ctx, [pe], [ds])
        lam = fo.lambda_i(0)
        X = int(next(v for v in range(emat.num_nodes) if v != emat.root and emat.parent[v] != emat.root))
        req = db.spr_requests_for_attached(emat, 0, [X], lam, info["t_max_tip"])
        for _ in range(3):
            b = fo.spr_study_batch(req); b.regions(0); b.close()
        n = 30; t0 = time.perf_counter()
        for _ in range(n):
            b = fo.spr_study_batch(req); r = b.regions(0); b.close()
        print(f"  one full SPR study + regions download: {(time.perf_counter()-t0)/n*1e3:.3f} ms ({len(r)} regions)")
        t0 = time.perf_counter()
        for _ in range(n):
            fo.nsmn = fo.num_sites_missing(0)
        print(f"  nsmn getter: {(time.perf_counter()-t0)/n*1e3:.3f} ms")
        fo.close(); ds.close()
