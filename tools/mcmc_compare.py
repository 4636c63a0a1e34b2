"""GPU-box tool: the reference's CLI, stock vs link-time drop-in, on synthetic alignments.
usage: python tools/mcmc_compare.py <out.json> [what ...]   what: verify | posterior | speed2k | speed10k | speed30k"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import delphy_b200 as db
from delphy_b200 import mcmc
from delphy_b200.maple import write_maple

out_path = sys.argv[1]
what = sys.argv[2:] or ["verify", "posterior", "speed2k"]
os.makedirs("/tmp/mcmc", exist_ok=True)
res = {}


def maple(name, cfg, **kw):
    path = f"/tmp/mcmc/{name}.maple"
    if not os.path.exists(path):
        emat, sites, info = db.synth_generate(db.synth_params(cfg, **kw))
        write_maple(emat, sites, path, info["t_max_tip"])
    return path


def brief(r):
    return {k: r[k] for k in ("steps_per_s", "wall_s", "init_s", "mcmc_s", "returncode")} | {"last": r["samples"][-1] if r["samples"] else None,
                                                                                                "tail": r["stderr_tail"][-3:] if r["returncode"] else None}


if "verify" in what:
    # every study of the run on the device, each one re-run by the reference's builder and compared (aborts on a difference)
    for name, cfg, kw, steps in (("cfg1", 1, {}, 400000), ("t2k", 3, dict(num_tips=2000), 200000)):
        r = mcmc.run_cli(mcmc.DROPIN_CLI, maple(name, cfg, **kw), steps, threads=1, seed=3, log_every=steps // 4,
                         env=dict(DPHY_DROPIN_VERIFY=1, DPHY_DROPIN_BOUNDED_ON_DEVICE=1))
        res["verify_" + name] = brief(r)
        print("verify", name, res["verify_" + name], flush=True)

if "posterior" in what:
    steps = 4000000
    for arm, binary in (("stock", mcmc.STOCK_CLI), ("dropin", mcmc.DROPIN_CLI)):
        r = mcmc.run_cli(binary, maple("cfg1", 1), steps, threads=1, seed=11, log_every=20000)
        res["posterior_" + arm] = brief(r) | {"means": mcmc.posterior_means(r["samples"])}
        print("posterior", arm, res["posterior_" + arm], flush=True)

for tag, cfg, kw, steps in (("speed2k", 3, dict(num_tips=2000), 1000000), ("speed10k", 3, {}, 400000), ("speed30k", 4, dict(num_tips=30000), 200000)):
    if tag not in what:
        continue
    name = tag.replace("speed", "t")
    for arm, binary in (("stock", mcmc.STOCK_CLI), ("dropin", mcmc.DROPIN_CLI)):
        r = mcmc.run_cli(binary, maple(name, cfg, **kw), steps, threads=1, seed=5, log_every=steps // 10)
        res[f"{tag}_{arm}"] = brief(r)
        print(tag, arm, res[f"{tag}_{arm}"], flush=True)

json.dump(res, open(out_path, "w"), indent=1)
