"""The oracle (CPU) and the CUDA path (GPU) against golden vectors produced by the reference's own code
(tests/golden/make_golden.py ran oracle/_ref in the build container; the .json files travel to the GPU box)."""
import glob
import json
import os

import numpy as np
import pytest

import delphy_b200 as db
from helpers import synth, to_oracle
from oracle_lib import Oracle

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_synth_*.json")))


def _load(path):
    g = json.load(open(path))
    emat, sites, info = synth(g["config"], **g["overrides"])
    chk = float(emat.t.sum() + emat.mut_t[emat.mut_t > -1e300].sum() + emat.mut_site.sum())
    assert chk == g["input_checksum"], "synthetic generator no longer reproduces the golden inputs"
    return g, emat, sites, info


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_reference_golden(path):
    g, emat, sites, info = _load(path)
    e, s = to_oracle(emat, sites)
    o = Oracle("oracle")
    lam = o.lambda_i(e, s)
    assert [x.hex() for x in lam.tolist()] == g["lambda_i"]                      # bit-identical
    assert o.log_root_prior(e, s) == g["log_root_prior"]
    assert o.log_G_below_root(e, s, lam) == g["log_G_below_root"]
    assert o.nsmn(e, s).tolist() == g["nsmn"]
    assert o.num_muts(e, s) == g["num_muts"]
    assert o.num_muts_ab(e, s).tolist() == g["num_muts_ab"]
    assert o.num_muts_beta_ab(e, s).tolist() == g["num_muts_beta_ab"]
    assert o.T(e, s) == g["T"]
    assert o.Ttwiddle_beta_a(e, s).tolist() == g["Ttwiddle_beta_a"]
    for st in g["studies"]:
        regs, sm = o.spr_study_from_attached(e, s, st["X"], lam, st["limit"], True, 0.8, info["t_max_tip"])
        assert regs["branch"].tolist() == st["branch"] and regs["mut_idx"].tolist() == st["mut_idx"]
        assert regs["min_muts"].tolist() == st["min_muts"]
        assert [x.hex() for x in regs["t_min"].tolist()] == st["t_min"]
        assert regs["W_over_Wmax"].tolist() == st["W_over_Wmax"]


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_cuda_matches_reference_golden(path):
    g, emat, sites, info = _load(path)
    with db.Context(0) as ctx:
        ds = db.DeviceSites(ctx, sites)
        fo = db.Forest(ctx, [emat], [ds])
        fo.eval_log_G()
        rp, br, _ = fo.log_G()
        assert rp[0] == pytest.approx(g["log_root_prior"], rel=1e-9)
        assert br[0] == pytest.approx(g["log_G_below_root"], rel=1e-9)
        lam = fo.lambda_i(0)
        want = np.array([float.fromhex(x) for x in g["lambda_i"]])
        np.testing.assert_allclose(lam, want, rtol=1e-9)
        assert fo.num_sites_missing(0).tolist() == g["nsmn"]
        tl = fo.tallies()[0]
        assert tl["num_muts"] == g["num_muts"] and tl["num_muts_ab"].tolist() == g["num_muts_ab"]
        assert fo.num_muts_beta_ab(0).tolist() == g["num_muts_beta_ab"]
        np.testing.assert_allclose(fo.Ttwiddle_beta_a(0), np.array(g["Ttwiddle_beta_a"]), rtol=1e-9)
        xs = sorted({st["X"] for st in g["studies"]}, key=[st["X"] for st in g["studies"]].index)
        for limit in (2**31 - 1, 1):
            reqs = db.spr_requests_for_attached(emat, 0, xs, want, info["t_max_tip"], limit, True)
            b = fo.spr_study_batch(reqs)
            for i, X in enumerate(xs):
                st = next(s_ for s_ in g["studies"] if s_["X"] == X and s_["limit"] == limit)
                regs = b.regions(i)
                assert regs["branch"].tolist() == st["branch"] and regs["mut_idx"].tolist() == st["mut_idx"]
                assert regs["min_muts"].tolist() == st["min_muts"]
                assert [x.hex() for x in regs["t_min"].tolist()] == st["t_min"]
                assert [x.hex() for x in regs["t_max"].tolist()] == st["t_max"]
                np.testing.assert_allclose(regs["W_over_Wmax"], np.array(st["W_over_Wmax"]), rtol=1e-9, atol=1e-300)
            b.close()
        fo.close(); ds.close()
