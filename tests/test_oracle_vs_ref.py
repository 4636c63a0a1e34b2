"""CPU-only: the C restatement (oracle/emat_oracle.c) against the reference's own code compiled in place
(oracle/_ref/libdelphy_ref.so) on synthetic EMATs: bit-identical wherever the algorithm is deterministic."""
import ctypes as C

import numpy as np
import pytest

from helpers import synth, to_oracle
from oracle_lib import Oracle, ref, ref_available

pytestmark = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (needs /root/reference)")

CASES = [
    (0, {}),
    (0, dict(num_root_mutations=6, num_partitions=2, site_rate_heterogeneity=1)),
    (0, dict(caterpillar=1, num_tips=250)),
    (1, {}),
    (2, {}),
]


@pytest.mark.parametrize("cfg,ov", CASES)
def test_synthetic_emats_satisfy_reference_invariants(cfg, ov):
    emat, sites, _ = synth(cfg, **ov)
    e, s = to_oracle(emat, sites)
    # assert_phylo_tree_integrity(tree, force=true) CHECK-aborts on violation (core/phylo_tree.cpp:18-135)
    assert ref().ref_assert_integrity(C.byref(e.as_struct()), C.byref(s.as_struct())) == 0


@pytest.mark.parametrize("cfg,ov", CASES)
def test_calc_functions_bitwise(cfg, ov):
    emat, sites, _ = synth(cfg, **ov)
    e, s = to_oracle(emat, sites)
    o, r = Oracle("oracle"), Oracle("ref")
    assert np.array_equal(o.state_frequencies(s), r.state_frequencies(s))
    assert np.array_equal(o.cum_Q_l(s), r.cum_Q_l(s))
    assert np.array_equal(o.lambda_i(e, s), r.lambda_i(e, s))
    assert o.log_root_prior(e, s) == r.log_root_prior(e, s)
    assert o.log_G_below_root(e, s) == r.log_G_below_root(e, s)
    assert np.array_equal(o.nsmn(e, s), r.nsmn(e, s))
    assert o.num_muts(e, s) == r.num_muts(e, s)
    assert np.array_equal(o.num_muts_ab(e, s), r.num_muts_ab(e, s))
    assert np.array_equal(o.num_muts_beta_ab(e, s), r.num_muts_beta_ab(e, s))
    assert np.array_equal(o.num_muts_l(e, s), r.num_muts_l(e, s))
    assert np.array_equal(o.num_muts_l_ab(e, s), r.num_muts_l_ab(e, s))
    assert o.T(e, s) == r.T(e, s)
    assert np.array_equal(o.T_l_a(e, s), r.T_l_a(e, s))
    assert np.array_equal(o.Ttwiddle_l(e, s), r.Ttwiddle_l(e, s))
    assert np.array_equal(o.Ttwiddle_beta_a(e, s), r.Ttwiddle_beta_a(e, s))


@pytest.mark.parametrize("cfg,ov", CASES[:4])
@pytest.mark.parametrize("limit", [2**31 - 1, 1, 0])
def test_spr_studies_bitwise_including_order(cfg, ov, limit):
    """Region ORDER, min_muts, times and weights of the oracle == the reference's Spr_study_builder + Spr_study."""
    emat, sites, info = synth(cfg, **ov)
    e, s = to_oracle(emat, sites)
    o, r = Oracle("oracle"), Oracle("ref")
    lam = o.lambda_i(e, s)
    rng = np.random.default_rng(5)
    xs = [int(v) for v in rng.permutation(emat.num_nodes) if v != emat.root][:25]
    xs += [int(emat.child0[emat.root]), int(emat.child1[emat.root])]
    for ccr in (True, False):
        for X in xs:
            a, sa = o.spr_study_from_attached(e, s, X, lam, limit, ccr, 0.8, info["t_max_tip"])
            b, sb = r.spr_study_from_attached(e, s, X, lam, limit, ccr, 0.8, info["t_max_tip"])
            assert len(a) == len(b)
            for k in ("branch", "mut_idx", "min_muts", "t_min", "t_max"):
                assert np.array_equal(a[k], b[k]), (X, k)
            # weights: identical formulas and libm; only the above-root region goes through gamma_q (shimmed by the
            # same restatement in _ref, see oracle/gamma_q.h "parity unpinned")
            assert np.array_equal(a["log_W_over_Wmax"], b["log_W_over_Wmax"])
            assert np.array_equal(a["W_over_Wmax"], b["W_over_Wmax"])
            if len(a):
                assert sa.sum_W_over_Wmax == sb.sum_W_over_Wmax and sa.log_Wmax == sb.log_Wmax and sa.mu == sb.mu


def test_missing_sites_at():
    emat, sites, _ = synth(1)
    e, s = to_oracle(emat, sites)
    o, r = Oracle("oracle"), Oracle("ref")
    for node in (0, 5, emat.root, int(emat.child0[emat.root])):
        a, b = o.missing_sites_at(e, s, node), r.missing_sites_at(e, s, node)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
