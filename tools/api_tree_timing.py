"""Cost of loading a delphy.api.Tree buffer straight to the device vs a plain dphy_forest_upload of the flat arrays (100k tips)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import delphy_b200 as db
import oracle_lib as ol
from helpers import to_oracle

for cfg, tips in ((4, 100000), (5, 50000)):
    emat, sites, info = db.synth_generate(db.synth_params(cfg, num_tips=tips))
    e, s = to_oracle(emat, sites)
    data = ol.api_tree_write(e, s.ref)
    t0 = time.perf_counter(); want, _ = ol.api_tree_read(data); t_orc = time.perf_counter() - t0
    ctx = db.Context(0)
    ds = db.DeviceSites(ctx, sites)
    for rep in range(3):
        t0 = time.perf_counter(); fo = db.Forest.from_api_trees(ctx, [data], [ds]); ctx.synchronize(); t_api = time.perf_counter() - t0
        t0 = time.perf_counter(); out = fo.write_api_tree(0); t_wr = time.perf_counter() - t0
        if rep == 2:
            got = fo.download_tree(0)
            ok = all(np.array_equal(getattr(got, f), getattr(want, f)) for f in ("fs_off", "fs_site", "fs_from", "mut_off", "miss_off", "t", "mut_t"))
            back, _ = ol.api_tree_read(out)
            ok2 = all(np.array_equal(getattr(back, f), getattr(want, f)) for f in ("fs_off", "fs_site", "fs_from", "mut_site", "miss_start", "t"))
        fo.close()
        pin = ctx.host_array_like(np.frombuffer(data, np.uint8))
        t0 = time.perf_counter(); fo = db.Forest.from_api_trees(ctx, [pin], [ds]); ctx.synchronize(); t_pin = time.perf_counter() - t0
        fo.close()
        t0 = time.perf_counter(); fp = db.Forest(ctx, [emat], [ds]); ctx.synchronize(); t_up = time.perf_counter() - t0
        fp.close()
    print(f"cfg {cfg} tips {tips}: buffer {len(data)/1e6:.1f} MB, F={int(want.fs_off[-1])}; api->device {t_api*1e3:.2f} ms ({t_pin*1e3:.2f} ms from a page-locked buffer), plain upload {t_up*1e3:.2f} ms, "
          f"device->api {t_wr*1e3:.2f} ms, oracle CPU reader {t_orc*1e3:.0f} ms; arrays match {ok}, round trip {ok2}")
    ds.close(); ctx.close()
