#!/usr/bin/env python
"""Generates tests/golden/ref_synth_*.json by running the REFERENCE'S OWN CODE (oracle/_ref/libdelphy_ref.so, compiled
in place from /root/reference by oracle/Makefile) on small synthetic EMATs.  Run in the build container only:

    python tests/golden/make_golden.py

The inputs are regenerated deterministically from the recorded generator parameters (dphy_synth_generate), so only
the parameters and the reference's outputs are stored."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from helpers import synth, to_oracle  # noqa: E402
from oracle_lib import Oracle  # noqa: E402

CASES = {
    "ref_synth_small": (0, {}),
    "ref_synth_small_p2_rootmuts": (0, dict(num_root_mutations=6, num_partitions=2, site_rate_heterogeneity=1)),
    "ref_synth_200tips": (1, {}),
}


def main():
    r = Oracle("ref")
    for name, (cfg, ov) in CASES.items():
        emat, sites, info = synth(cfg, **ov)
        e, s = to_oracle(emat, sites)
        lam = r.lambda_i(e, s)
        out = dict(config=cfg, overrides=ov, num_nodes=emat.num_nodes, num_sites=sites.num_sites,
                   input_checksum=float(emat.t.sum() + emat.mut_t[emat.mut_t > -1e300].sum() + emat.mut_site.sum()),
                   log_root_prior=r.log_root_prior(e, s), log_G_below_root=r.log_G_below_root(e, s),
                   lambda_i=[x.hex() for x in lam.tolist()], nsmn=r.nsmn(e, s).tolist(),
                   num_muts=r.num_muts(e, s), num_muts_ab=r.num_muts_ab(e, s).tolist(),
                   num_muts_beta_ab=r.num_muts_beta_ab(e, s).tolist(), T=r.T(e, s),
                   Ttwiddle_beta_a=r.Ttwiddle_beta_a(e, s).tolist(),
                   Ttwiddle_l_sum=float(r.Ttwiddle_l(e, s).sum()), T_l_a_sum=r.T_l_a(e, s).sum(axis=0).tolist(),
                   num_muts_l_nonzero=int((r.num_muts_l(e, s) > 0).sum()), studies=[])
        rng = np.random.default_rng(3)
        xs = [int(v) for v in rng.permutation(emat.num_nodes) if v != emat.root][:6] + [int(emat.child0[emat.root])]
        for X in xs:
            for limit in (2**31 - 1, 1):
                regs, sm = r.spr_study_from_attached(e, s, X, lam, limit, True, 0.8, info["t_max_tip"])
                out["studies"].append(dict(X=X, limit=limit, branch=regs["branch"].tolist(), mut_idx=regs["mut_idx"].tolist(),
                                           min_muts=regs["min_muts"].tolist(), t_min=[x.hex() for x in regs["t_min"].tolist()],
                                           t_max=[x.hex() for x in regs["t_max"].tolist()],
                                           W_over_Wmax=regs["W_over_Wmax"].tolist(), sum_W_over_Wmax=sm.sum_W_over_Wmax,
                                           log_Wmax=sm.log_Wmax))
        with open(os.path.join(HERE, name + ".json"), "w") as fh:
            json.dump(out, fh)
        print(name, os.path.getsize(os.path.join(HERE, name + ".json")), "bytes")


if __name__ == "__main__":
    main()
