"""GPU-box tool: MCMC steps/s of the stock CLI and the drop-in CLI vs host threads (and GPUs).  usage: mcmc_threads.py out.json tips steps gpus"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import delphy_b200 as db
from delphy_b200 import mcmc
from delphy_b200.maple import write_maple
out, tips, steps, gpus = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
os.makedirs("/tmp/mcmc", exist_ok=True)
path = f"/tmp/mcmc/t{tips}.maple"
emat, sites, info = db.synth_generate(db.synth_params(3, num_tips=tips))
write_maple(emat, sites, path, info["t_max_tip"])
res = {}
for threads in (1, 2, 4, 8):
    for arm, binary in (("stock", mcmc.STOCK_CLI), ("dropin", mcmc.DROPIN_CLI)):
        r = mcmc.run_cli(binary, path, steps, threads=threads, seed=5, log_every=steps // 10, env=dict(DPHY_DEVICES=gpus, DPHY_DROPIN_STATS=1), timeout=900)
        res[f"{arm}_t{threads}"] = dict(steps_per_s=r["steps_per_s"], mcmc_s=r["mcmc_s"], init_s=r["init_s"], rc=r["returncode"],
                                        stats=[l for l in r["stderr_tail"] if "drop-in" in l][-8:])
        print(arm, threads, res[f"{arm}_t{threads}"], flush=True)
json.dump(res, open(out, "w"), indent=1)
