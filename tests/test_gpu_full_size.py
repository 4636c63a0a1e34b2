"""GPU parity at BASELINE.json's full shapes (configs[3]: 100k tips x 29,903 sites; configs[4]: 50k tips x 197,000 sites,
heavy missing data, 2 partitions) -- against the oracle where it finishes in seconds, plus size-independent properties:
both log-G schedules agree, evaluations are bit-reproducible, the parts of a partition sum to the whole tree, and every
SPR study's weights are normalised (max W/Wmax == 1, sum == sum_W_over_Wmax)."""
import numpy as np
import pytest

import delphy_b200 as db
from helpers import rel_err, synth, to_oracle
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def orc():
    return Oracle("oracle")


@pytest.mark.parametrize("cfg", [4, 5])
def test_full_size_log_G_and_spr(orc, cfg):
    emat, sites, info = synth(cfg)
    e, s = to_oracle(emat, sites)
    with db.Context(0) as ctx:
        ds = db.DeviceSites(ctx, sites)
        fo = db.Forest(ctx, [emat], [ds])
        # ---- log G, both schedules, vs the oracle
        lam_o = orc.lambda_i(e, s)
        want_rp, want_br = orc.log_root_prior(e, s), orc.log_G_below_root(e, s, lam_o)
        got = {}
        for path in ("general", "auto"):
            ctx.set_log_G_path(path)
            fo.eval_log_G()
            rp, br, _ = fo.log_G()
            assert rp[0] == pytest.approx(want_rp, rel=RTOL)
            assert br[0] == pytest.approx(want_br, rel=RTOL)
            assert rel_err(fo.lambda_i(0), lam_o) <= RTOL
            fo.eval_log_G()
            assert fo.log_G()[1][0] == br[0]                       # bit-reproducible run to run
            got[path] = br[0]
        assert got["auto"] == pytest.approx(got["general"], rel=1e-12)
        # ---- integer outputs: bit-exact
        np.testing.assert_array_equal(fo.num_sites_missing(0), orc.nsmn(e, s))
        tl = fo.tallies()[0]
        assert tl["num_muts"] == orc.num_muts(e, s)
        np.testing.assert_array_equal(tl["num_muts_ab"], orc.num_muts_ab(e, s))
        np.testing.assert_array_equal(fo.num_muts_l(0)[0], orc.num_muts_l(e, s))
        want = orc.Ttwiddle_beta_a(e, s)
        np.testing.assert_allclose(fo.Ttwiddle_beta_a(0), want, rtol=RTOL, atol=1e-9 * np.abs(want).max())
        # ---- SPR: full studies of a few nodes, region by region, + the normalisation properties
        rng = np.random.default_rng(cfg)
        xs = [int(v) for v in rng.permutation(emat.num_nodes)[:64] if v != emat.root and emat.parent[v] != emat.root][:5]
        reqs = db.spr_requests_for_attached(emat, 0, xs, fo.lambda_i(0), info["t_max_tip"])
        b = fo.spr_study_batch(reqs)
        summ = b.summaries()
        for i, X in enumerate(xs):
            regs = b.regions(i)
            want_regs, ws = orc.spr_study_from_attached(e, s, X, lam_o, t_max_tip=info["t_max_tip"])
            assert len(regs) == len(want_regs)
            for k in ("branch", "mut_idx", "min_muts", "t_min", "t_max"):
                assert np.array_equal(regs[k], want_regs[k]), (X, k)
            assert np.allclose(regs["W_over_Wmax"], want_regs["W_over_Wmax"], rtol=1e-9, atol=1e-300)
            assert regs["W_over_Wmax"].max() == 1.0
            assert summ[i].sum_W_over_Wmax == pytest.approx(float(np.sum(regs["W_over_Wmax"])), rel=1e-12)
            assert summ[i].sum_W_over_Wmax == pytest.approx(ws.sum_W_over_Wmax, rel=1e-9)
        b.close(); fo.close(); ds.close()


def test_full_size_partition_parts_sum_to_whole(orc):
    """configs[3] cut into 8 parts (Run::repartition): the parts evaluated as one forest sum to the whole tree's log G and
    mutation count (core/run.cpp:340-357 asserts the same on the host)."""
    emat, sites, _ = synth(4)
    with db.Context(0) as ctx:
        ds = db.DeviceSites(ctx, sites)
        whole = db.Forest(ctx, [emat], [ds])
        _, br_w, lg_w = whole.log_G()
        nm_w = whole.tallies()[0]["num_muts"]
        parts, origs, cuts = db.partition_emat(emat, sites, 8, seed=5)
        assert len(parts) >= 4
        fo = db.Forest(ctx, parts, [ds])
        _, _, lg = fo.log_G()
        assert float(np.sum(lg)) == pytest.approx(lg_w[0], rel=RTOL)
        assert sum(t["num_muts"] for t in fo.tallies()) == nm_w
        fo.close(); whole.close(); ds.close()
