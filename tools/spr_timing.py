import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import delphy_b200 as db
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
nst = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ctx = db.Context(0)
e, s, info = db.synth_generate(db.synth_params(cfg))
ds = db.DeviceSites(ctx, s); fo = db.Forest(ctx, [e], [ds])
fo.eval_log_G(); lam = fo.lambda_i(0)
rng = np.random.default_rng(1234)
xs = [int(v) for v in rng.permutation(e.num_nodes)[:4*nst] if v != e.root and e.parent[v] != e.root][:nst]
reqs = db.spr_requests_for_attached(e, 0, xs, lam, info["t_max_tip"])
st = torch.cuda.ExternalStream(ctx.stream)
for keep in (False, True):
    held = []
    for it in range(6):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record(st); bt = fo.spr_study_batch(reqs); ctx.join_side_streams(); b.record(st)
        t1 = time.perf_counter()
        ctx.synchronize()
        t2 = time.perf_counter()
        print(f"keep={keep} it={it} gpu={a.elapsed_time(b):.3f} ms host_enqueue={1e3*(t1-t0):.3f} ms total={1e3*(t2-t0):.3f} ms regions={bt.total_regions()}")
        if keep: held.append(bt)
        else: bt.close()
    for h in held: h.close()
