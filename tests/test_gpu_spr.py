"""GPU parity: batched SPR regraft studies (through the C ABI) vs the oracle.

Integer outputs (branch, mut_idx, min_muts, region ORDER, region count) and copied doubles (t_min, t_max) must be
bit-exact; weights within 1e-9 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

import delphy_b200 as db
from emat_fixtures import A, Cc, G, T, DBL_MAX, complex_tree
from helpers import from_oracle, synth, to_oracle
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu
INF = 2**31 - 1


@pytest.fixture(scope="module")
def ctx():
    c = db.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def orc():
    return Oracle("oracle")


def _cmp_regions(got, want, weights=True):
    assert len(got) == len(want)
    for k in ("branch", "mut_idx", "min_muts"):
        np.testing.assert_array_equal(got[k], want[k], err_msg=k)
    for k in ("t_min", "t_max"):
        assert np.array_equal(got[k], want[k]), k          # copies of input doubles: bit-exact
    if weights and len(want):
        np.testing.assert_allclose(got["log_W_over_Wmax"], want["log_W_over_Wmax"], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(got["W_over_Wmax"], want["W_over_Wmax"], rtol=1e-9, atol=1e-300)


def _fixture_forest(ctx):
    e, s, n = complex_tree()
    emat, sites = from_oracle(e, s)
    ds = db.DeviceSites(ctx, sites)
    fo = db.Forest(ctx, [emat], [ds])
    return e, s, n, emat, fo, ds


def test_reference_region_sets(ctx, orc):
    """The six studies of tests/spr_study_tests.cpp:91-205 -- exact region lists in the reference's DFS order."""
    e, s, n, emat, fo, ds = _fixture_forest(ctx)
    r, x, a, b, c = n["r"], n["x"], n["a"], n["b"], n["c"]
    lam, f, tmt = 0.9, 0.8, 3.0
    cases = [
        # X, t_X, start, k0, init deltas (site, from, to), limit, can_change_root
        (a, 1.5, b, 0, [(0, T, Cc)], INF, True),
        (a, 1.5, b, 0, [(0, T, Cc)], INF, False),
        (a, 1.5, b, 0, [(0, T, Cc)], 1, True),
        (x, 0.0, c, 2, [(0, G, T)], INF, True),
        (c, 3.0, x, 1, [(0, T, G)], INF, True),
        (c, 3.0, x, 1, [(0, T, G)], INF, False),
        (b, 2.0, a, 1, [(0, Cc, T), (1, A, G)], 1, True),
        (b, 1.2, a, 0, [(1, A, G)], 0, True),
    ]
    reqs = [db.spr_request(0, X, tX, sb, k0, len(dl), lam, tmt, lim, ccr, f) for (X, tX, sb, k0, dl, lim, ccr) in cases]
    batch = fo.spr_study_batch(reqs)
    summ = batch.summaries()
    for i, (X, tX, sb, k0, dl, lim, ccr) in enumerate(cases):
        miss = orc.missing_sites_at(e, s, X)
        want, ws = orc.spr_study(e, s, X, tX, miss, sb, k0, dl, lim, ccr, weights=(lam, f, tmt))
        got = batch.regions(i)
        _cmp_regions(got, want)
        assert summ[i].num_regions == len(want)
        if len(want):
            assert summ[i].log_Wmax == pytest.approx(ws.log_Wmax, rel=1e-9, abs=1e-9)
            assert summ[i].sum_W_over_Wmax == pytest.approx(ws.sum_W_over_Wmax, rel=1e-9)
            assert summ[i].mu == pytest.approx(ws.mu, rel=1e-12)
            assert summ[i].num_missing_at_X == ws.num_missing_at_X
    # golden set from the reference's test (full_spr_study_a, :91-110)
    got0 = batch.regions(0)
    assert sorted(zip(got0["branch"], got0["mut_idx"], got0["t_min"], got0["t_max"], got0["min_muts"])) == sorted([
        (b, 1, -0.5, 1.0, 1), (b, 2, 1.0, 1.5, 2), (b, 0, -1.0, -0.5, 1), (r, 1, -DBL_MAX, -1.0, 1),
        (c, 0, -1.0, 0.0, 1), (c, 1, 0.0, 1.0, 1), (c, 2, 1.0, 1.5, 1)])
    batch.close(); fo.close(); ds.close()


def test_detached_new_sequence(ctx, orc):
    """full_study_new_seq (tests/spr_study_tests.cpp:183-205): X == k_no_node, as build_usher_like_tree uses it."""
    e, s, n, emat, fo, ds = _fixture_forest(ctx)
    r = n["r"]
    req = db.spr_request(0, -1, 1.5, r, 1, 1, 0.9, 3.0, INF, True, 0.8, x_deltas=[(0, T)], x_missing=([3], [4]))
    batch = fo.spr_study_batch([req])
    want, _ = orc.spr_study(e, s, -1, 1.5, ([3], [4]), r, 1, [(0, A, T)], INF, True, weights=(0.9, 0.8, 3.0))
    _cmp_regions(batch.regions(0), want)
    batch.close(); fo.close(); ds.close()


@pytest.mark.parametrize("cfg,ov,nx", [
    (0, {}, 40),
    (0, dict(num_root_mutations=6, num_partitions=2, site_rate_heterogeneity=1), 40),
    (0, dict(caterpillar=1, num_tips=400), 30),
    (0, dict(num_tips=2500, muts_per_tip=24.0), 10),   # ~12 mutations per branch: more than the emit kernel caches per CTA, many slot rounds
    (1, {}, 48),
    (2, {}, 32),
    (3, {}, 12),
])
@pytest.mark.parametrize("limit", [INF, 1])
def test_synthetic_studies(ctx, orc, cfg, ov, nx, limit):
    emat, sites, info = synth(cfg, **ov)
    e, s = to_oracle(emat, sites)
    ds = db.DeviceSites(ctx, sites)
    fo = db.Forest(ctx, [emat], [ds])
    lam = fo.lambda_i(0)
    rng = np.random.default_rng(99 + cfg)
    xs = [int(v) for v in rng.permutation(emat.num_nodes) if v != emat.root][:nx]
    # include the two children of the root (pruning changes the root) and can_change_root both ways
    xs += [int(emat.child0[emat.root]), int(emat.child1[emat.root])]
    for ccr in (True, False):
        reqs = db.spr_requests_for_attached(emat, 0, xs, lam, info["t_max_tip"], limit, ccr)
        batch = fo.spr_study_batch(reqs)
        summ = batch.summaries()
        for i, X in enumerate(xs):
            want, ws = orc.spr_study_from_attached(e, s, X, lam, limit, ccr, 0.8, info["t_max_tip"])
            got = batch.regions(i)
            _cmp_regions(got, want)
            if len(want):
                assert summ[i].sum_W_over_Wmax == pytest.approx(ws.sum_W_over_Wmax, rel=1e-9)
                assert summ[i].log_Wmax == pytest.approx(ws.log_Wmax, rel=1e-9, abs=1e-9)
        # chosen regraft under a fixed RNG stream: same uniform draws -> same region index as the reference's scan
        us = np.random.default_rng(7).random(len(xs))
        r = np.array([u * sm.sum_W_over_Wmax for u, sm in zip(us, summ)])
        picked = batch.pick_nexus_regions(r)
        for i, X in enumerate(xs):
            want, _ = orc.spr_study_from_attached(e, s, X, lam, limit, ccr, 0.8, info["t_max_tip"])
            if len(want) == 0:
                continue
            got = batch.regions(i)
            import ctypes as C
            from oracle_lib import OrcRegion
            w_ref = orc.lib.orc_spr_pick_nexus_region(want.ctypes.data_as(C.POINTER(OrcRegion)), len(want), float(r[i]))
            w_dev = orc.lib.orc_spr_pick_nexus_region(got.ctypes.data_as(C.POINTER(OrcRegion)), len(got), float(r[i]))
            assert w_dev == w_ref                       # the reference's own scan over the device's weights
            cum = np.cumsum(want["W_over_Wmax"])
            near_boundary = np.min(np.abs(cum - r[i])) < 1e-9 * max(1.0, cum[-1])
            assert picked[i] == w_ref or near_boundary  # device-side CDF search
        batch.close()
    fo.close(); ds.close()


def test_find_region(ctx, orc):
    emat, sites, info = synth(1)
    e, s = to_oracle(emat, sites)
    ds = db.DeviceSites(ctx, sites); fo = db.Forest(ctx, [emat], [ds])
    lam = fo.lambda_i(0)
    X = int(next(v for v in range(emat.num_nodes) if v != emat.root and emat.parent[v] != emat.root))
    batch = fo.spr_study_batch(db.spr_requests_for_attached(emat, 0, [X], lam, info["t_max_tip"]))
    regs = batch.regions(0)
    import ctypes as C
    from oracle_lib import OrcRegion
    for i in (0, len(regs) // 2, len(regs) - 1):
        r = regs[i]
        if r["t_min"] == -DBL_MAX:
            t = r["t_max"] - 1.0
        else:
            t = 0.5 * (r["t_min"] + r["t_max"])
        want = orc.lib.orc_spr_find_region(regs.ctypes.data_as(C.POINTER(OrcRegion)), len(regs), int(r["branch"]), float(t))
        assert batch.find_region(0, int(r["branch"]), float(t)) == want
    assert batch.find_region(0, X, float(emat.t[X])) == -1
    batch.close(); fo.close(); ds.close()


def test_request_validation(ctx):
    emat, sites, info = synth(0)
    ds = db.DeviceSites(ctx, sites); fo = db.Forest(ctx, [emat], [ds])
    with pytest.raises(db.DphyError):
        fo.spr_study_batch([db.spr_request(0, emat.root, 0.0, 0, 0, 0, 1.0, 0.0)])   # X == root
    with pytest.raises(db.DphyError):
        fo.spr_study_batch([db.spr_request(0, 1, 0.0, emat.num_nodes + 3, 0, 0, 1.0, 0.0)])
    fo.close(); ds.close()


def test_studies_on_every_tree_of_a_forest(ctx, orc):
    """Studies addressed to trees other than the first one of a forest (device positions are forest-global)."""
    items = [synth(0, seed=31), synth(0, seed=32, num_tips=150), synth(1, seed=33)]
    tables = [db.DeviceSites(ctx, it[1]) for it in items]
    fo = db.Forest(ctx, [it[0] for it in items], tables, sites_index=np.arange(len(items)))
    reqs, meta = [], []
    for k, (emat, sites, info) in enumerate(items):
        lam = fo.lambda_i(k)
        xs = [int(v) for v in np.random.default_rng(k).permutation(emat.num_nodes) if v != emat.root][:6]
        reqs += db.spr_requests_for_attached(emat, k, xs, lam, info["t_max_tip"], 1 if k == 1 else INF)
        meta += [(k, X, lam) for X in xs]
    batch = fo.spr_study_batch(reqs)
    for i, (k, X, lam) in enumerate(meta):
        e, s = to_oracle(items[k][0], items[k][1])
        want, _ = orc.spr_study_from_attached(e, s, X, lam, 1 if k == 1 else INF, True, 0.8, items[k][2]["t_max_tip"])
        _cmp_regions(batch.regions(i), want)
    batch.close(); fo.close()
    for t in tables:
        t.close()
