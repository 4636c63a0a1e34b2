"""GPU parity: the CUDA log-G / tallies path (through the C ABI) vs the oracle on the same inputs.

Tolerances (BASELINE.json north_star): integers bit-exact; fp64 log-likelihood pieces <= 1e-9 relative.
"""
import numpy as np
import pytest

import delphy_b200 as db
from emat_fixtures import complex_tree
from helpers import from_oracle, rel_err, synth, to_oracle
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu

RTOL = 1e-9


@pytest.fixture(scope="module", params=["auto", "general"])
def ctx(request):
    """Every test runs twice: with the folded fast path enabled (taken whenever all site rates are uniform) and with the
    general per-event kernels forced."""
    c = db.Context(0)
    c.set_log_G_path(request.param)
    yield c
    c.close()


@pytest.fixture(scope="module")
def orc():
    return Oracle("oracle")


def _check_tree(ctx, orc, emat, sites, check_lists=True):
    e, s = to_oracle(emat, sites)
    ds = db.DeviceSites(ctx, sites)
    fo = db.Forest(ctx, [emat], [ds])
    try:
        np.testing.assert_array_equal(ds.state_frequencies(), orc.state_frequencies(s))
        cq = orc.cum_Q_l(s)
        assert rel_err(ds.cum_Q_l()[1:], cq[1:]) <= 1e-10   # the reference's sequential scan drifts ~L*eps/2 (see adapter_parity_main.cpp)
        fo.eval_log_G()
        rp, br, lg = fo.log_G()
        lam_o = orc.lambda_i(e, s, cq)
        assert rel_err(fo.lambda_i(0), lam_o) <= RTOL
        np.testing.assert_array_equal(fo.num_sites_missing(0), orc.nsmn(e, s))
        want_rp = orc.log_root_prior(e, s)
        want_br = orc.log_G_below_root(e, s, lam_o)
        assert rp[0] == pytest.approx(want_rp, rel=RTOL)
        assert br[0] == pytest.approx(want_br, rel=RTOL)
        assert lg[0] == pytest.approx(want_rp + want_br, rel=RTOL)
        tl = fo.tallies()[0]
        assert tl["num_muts"] == orc.num_muts(e, s)
        np.testing.assert_array_equal(tl["num_muts_ab"], orc.num_muts_ab(e, s))
        assert tl["T"] == pytest.approx(orc.T(e, s), rel=RTOL)
        if check_lists:
            np.testing.assert_array_equal(fo.num_muts_beta_ab(0), orc.num_muts_beta_ab(e, s))
            nl, nlab = fo.num_muts_l(0)
            np.testing.assert_array_equal(nl, orc.num_muts_l(e, s))
            np.testing.assert_array_equal(nlab, orc.num_muts_l_ab(e, s))
            want = orc.Ttwiddle_beta_a(e, s)
            np.testing.assert_allclose(fo.Ttwiddle_beta_a(0), want, rtol=RTOL, atol=1e-9 * np.abs(want).max())
            tw, tla = fo.Ttwiddle_l(0)
            want_l, want_la = orc.Ttwiddle_l(e, s), orc.T_l_a(e, s)
            np.testing.assert_allclose(tw, want_l, rtol=RTOL, atol=1e-9 * np.abs(want_l).max())
            np.testing.assert_allclose(tla, want_la, rtol=RTOL, atol=1e-9 * np.abs(want_la).max())
    finally:
        fo.close(); ds.close()


def test_reference_kat_fixture(ctx, orc):
    """The reference's own 5-node fixture (tests/phylo_tree_calc_tests.cpp:14-116), P=2, non-reversible Q."""
    e, s, _ = complex_tree()
    emat, sites = from_oracle(e, s)
    _check_tree(ctx, orc, emat, sites)
    # absolute known answers from the reference's tests
    ds = db.DeviceSites(ctx, sites); fo = db.Forest(ctx, [emat], [ds])
    tl = fo.tallies()[0]
    assert tl["num_muts"] == 5                                      # :441
    np.testing.assert_array_equal(fo.num_sites_missing(0), [1, 2, 2, 2, 2])   # :497
    np.testing.assert_array_equal(fo.num_muts_l(0)[0], [4, 1, 0, 0])          # :471
    assert tl["T"] == pytest.approx(8.0, abs=1e-9)                   # :236
    fo.close(); ds.close()


def test_root_prior_zero_pi(ctx, orc):
    """pi == 0 edge cases (tests/phylo_tree_calc_tests.cpp:355-379)."""
    e, s, _ = complex_tree()
    s.pi_a[0] = [0.3, 0.7, 0.0, 0.0]; s.pi_a[1] = [0.3, 0.0, 0.7, 0.0]
    emat, sites = from_oracle(e, s)
    rp, _, _ = ctx.log_G_host(emat, sites)
    assert rp == pytest.approx(orc.log_root_prior(e, s), rel=RTOL)
    s.pi_a[0] = [0.0, 0.3, 0.7, 0.0]
    emat, sites = from_oracle(e, s)
    rp, _, _ = ctx.log_G_host(emat, sites)
    assert rp == -np.inf


@pytest.mark.parametrize("cfg,ov", [
    (0, {}),
    (0, dict(num_root_mutations=7, num_partitions=2, site_rate_heterogeneity=1)),
    (0, dict(caterpillar=1, num_tips=700)),
    (0, dict(num_tips=2, num_sites=50, missing_len_max=20.0)),
    (0, dict(num_tips=257, missing_mean_intervals_per_tip=0.0, end_gaps=0)),
    (0, dict(num_tips=2500, muts_per_tip=24.0, missing_mean_intervals_per_tip=12.0)),   # long per-branch lists
    (1, {}),
    (2, {}),
    (3, {}),
])
def test_synthetic_parity(ctx, orc, cfg, ov):
    emat, sites, _ = synth(cfg, **ov)
    _check_tree(ctx, orc, emat, sites)


def test_forest_batch_matches_single(ctx, orc):
    """Several EMATs (different trees, two site tables) evaluated in one launch == each evaluated alone."""
    items = [synth(0, seed=11), synth(0, seed=12, num_tips=300), synth(1, seed=13), synth(0, seed=14, num_partitions=2)]
    tables = [db.DeviceSites(ctx, it[1]) for it in items]
    fo = db.Forest(ctx, [it[0] for it in items], tables, sites_index=np.arange(len(items)))
    fo.eval_log_G()
    rp, br, lg = fo.log_G()
    for k, (emat, sites, _) in enumerate(items):
        e, s = to_oracle(emat, sites)
        lam = orc.lambda_i(e, s)
        assert rel_err(fo.lambda_i(k), lam) <= RTOL
        assert br[k] == pytest.approx(orc.log_G_below_root(e, s, lam), rel=RTOL)
        assert rp[k] == pytest.approx(orc.log_root_prior(e, s), rel=RTOL)
        np.testing.assert_array_equal(fo.num_sites_missing(k), orc.nsmn(e, s))
    # bit-reproducible run to run
    fo.eval_log_G()
    rp2, br2, _ = fo.log_G()
    assert np.array_equal(br, br2) and np.array_equal(rp, rp2)
    fo.close()
    for t in tables:
        t.close()


def test_set_evo_and_node_times(ctx, orc):
    emat, sites, _ = synth(1)
    ds = db.DeviceSites(ctx, sites); fo = db.Forest(ctx, [emat], [ds])
    # Subrun::set_evo: new mu / nu
    rng = np.random.default_rng(5)
    sites.mu = sites.mu * 1.7
    sites.nu_l = rng.gamma(2.0, 0.5, size=sites.num_sites)
    ds.set_evo(nu_l=sites.nu_l, mu=sites.mu)
    e, s = to_oracle(emat, sites)
    fo.eval_log_G()
    _, br, _ = fo.log_G()
    assert br[0] == pytest.approx(orc.log_G_below_root(e, s), rel=RTOL)
    # a displaced inner node (core/subrun.cpp:223-231): move the root a little earlier
    emat.t[emat.root] -= 0.01
    fo.set_node_times(0, [emat.root], [emat.t[emat.root]])
    e, s = to_oracle(emat, sites)
    fo.eval_log_G()
    _, br, _ = fo.log_G()
    assert br[0] == pytest.approx(orc.log_G_below_root(e, s), rel=RTOL)
    fo.close(); ds.close()


def test_getters_reevaluate_after_set_evo(ctx, orc):
    """A getter called after dphy_sites_set_evo must not return the previous model's values (the forest compares the
    sites versions of its last evaluation with the current ones) -- the Device_emat::set_evo use of the adapter."""
    emat, sites, _ = synth(1)
    ds = db.DeviceSites(ctx, sites); fo = db.Forest(ctx, [emat], [ds])
    _, br0, _ = fo.log_G()
    sites.mu = sites.mu * 1.7
    ds.set_evo(mu=sites.mu)
    e, s = to_oracle(emat, sites)
    _, br1, _ = fo.log_G()                      # no explicit eval_log_G in between
    assert br1[0] != br0[0]
    lam = orc.lambda_i(e, s)
    assert br1[0] == pytest.approx(orc.log_G_below_root(e, s, lam), rel=RTOL)
    assert rel_err(fo.lambda_i(0), lam) <= RTOL
    fo.close(); ds.close()


def test_folded_and_general_paths_agree(orc):
    """Same forest, both schedules: the folded path (per-branch state-count vectors) and the per-event path give the
    same lambda_i / log G to ~1e-13, and the structure-only outputs (nsmn, num_muts) survive folded evaluations,
    time tallies and evo changes in between."""
    items = [synth(3, seed=21), synth(0, seed=22, num_tips=700, caterpillar=1), synth(0, seed=23, num_partitions=2, num_root_mutations=5)]
    with db.Context(0) as c:
        tables = [db.DeviceSites(c, it[1]) for it in items]
        fo = db.Forest(c, [it[0] for it in items], tables, sites_index=np.arange(len(items)))
        c.set_log_G_path("general")
        fo.eval_log_G()
        rp_g, br_g, _ = fo.log_G()
        lam_g = [fo.lambda_i(k) for k in range(len(items))]
        c.set_log_G_path("auto")
        launches = c.launches
        fo.eval_log_G()
        assert c.launches - launches == 2            # folded tile kernel + per-tree fold
        rp_f, br_f, _ = fo.log_G()
        np.testing.assert_allclose(br_f, br_g, rtol=1e-12)
        np.testing.assert_allclose(rp_f, rp_g, rtol=1e-13)
        fo.Ttwiddle_beta_a(0); fo.Ttwiddle_l(1)      # scratch users must not disturb the lambda_i tile prefixes
        for k, (emat, sites, _) in enumerate(items):
            e, s = to_oracle(emat, sites)
            np.testing.assert_allclose(fo.lambda_i(k), lam_g[k], rtol=1e-12)
            np.testing.assert_array_equal(fo.num_sites_missing(k), orc.nsmn(e, s))
            assert fo.tallies()[k]["num_muts"] == orc.num_muts(e, s)
        # new mu / kappa-like change through set_evo, still uniform nu: folded again, vs the oracle
        emat, sites, _ = items[0]
        sites.mu = sites.mu * 0.6
        tables[0].set_evo(mu=sites.mu)
        fo.eval_log_G()
        e, s = to_oracle(emat, sites)
        lam = orc.lambda_i(e, s)
        assert rel_err(fo.lambda_i(0), lam) <= RTOL
        assert fo.log_G()[1][0] == pytest.approx(orc.log_G_below_root(e, s, lam), rel=RTOL)
        fo.close()
        for t in tables:
            t.close()


def test_error_behaviour(ctx):
    """Out-of-range sites are rejected like the reference's std::out_of_range (core/mutations.h:187-191)."""
    emat, sites, _ = synth(0)
    bad = db.HostEmat(emat.root, 1, **{k: getattr(emat, k).copy() for k in db.HostEmat.FIELDS_I32 + db.HostEmat.FIELDS_U8 + db.HostEmat.FIELDS_F64})
    bad.miss_end[0] = sites.num_sites + 5
    ds = db.DeviceSites(ctx, sites)
    with pytest.raises(db.DphyError) as ei:
        db.Forest(ctx, [bad], [ds])
    assert ei.value.status == db.ERR_OUT_OF_RANGE
    ds.close()


def test_partition_parts_as_forest(ctx, orc):
    """Run::repartition on the device: the parts of one tree evaluated as a forest in one launch; additive tallies
    sum to the whole tree's (Run::check_global_and_local_totals_match, core/run.cpp:340-357)."""
    emat, sites, _ = synth(3)
    parts, origs, cuts = db.partition_emat(emat, sites, 8, seed=5)
    assert len(parts) >= 4
    ds = db.DeviceSites(ctx, sites)
    whole = db.Forest(ctx, [emat], [ds]); pf = db.Forest(ctx, parts, [ds])
    whole.eval_log_G(); pf.eval_log_G()
    _, _, lg_w = whole.log_G(); _, _, lg_p = pf.log_G()
    assert lg_p.sum() == pytest.approx(lg_w[0], rel=1e-10)
    tw = whole.tallies()[0]; tp = pf.tallies()
    assert sum(t["num_muts"] for t in tp) == tw["num_muts"]
    assert np.array_equal(sum(t["num_muts_ab"] for t in tp), tw["num_muts_ab"])
    assert sum(t["T"] for t in tp) == pytest.approx(tw["T"], rel=1e-11)
    lam = whole.lambda_i(0)
    for k, og in enumerate(origs):
        np.testing.assert_allclose(pf.lambda_i(k), lam[og], rtol=1e-10)
        np.testing.assert_array_equal(pf.num_sites_missing(k), whole.num_sites_missing(0)[og])
    tb = sum(pf.Ttwiddle_beta_a(k) for k in range(len(parts)))
    np.testing.assert_allclose(tb, whole.Ttwiddle_beta_a(0), rtol=1e-9)
    whole.close(); pf.close(); ds.close()


def _copy_emat(emat):
    return db.HostEmat(emat.root, 1, **{k: getattr(emat, k).copy() for k in db.HostEmat.FIELDS_I32 + db.HostEmat.FIELDS_U8 + db.HostEmat.FIELDS_F64})


def test_flatten_rejects_bad_topology(ctx):
    """The device-side flattening (Euler tour + list ranking) validates what the reference CHECKs in
    assert_tree_integrity (core/tree.h) -- cycles, child/parent mismatches, unreachable nodes, bad CSR offsets."""
    emat, sites, _ = synth(0)
    ds = db.DeviceSites(ctx, sites)
    inner = int(next(v for v in range(emat.num_nodes) if emat.child0[v] >= 0 and v != emat.root))
    tip = int(next(v for v in range(emat.num_nodes) if emat.child0[v] < 0))

    bad = _copy_emat(emat); bad.child0[inner] = bad.child1[inner]                 # same child twice
    with pytest.raises(db.DphyError) as ei:
        db.Forest(ctx, [bad], [ds])
    assert ei.value.status == db.ERR_INVALID_ARGUMENT

    bad = _copy_emat(emat); bad.parent[tip] = (int(bad.parent[tip]) + 1) % emat.num_nodes   # parent does not own the child
    with pytest.raises(db.DphyError):
        db.Forest(ctx, [bad], [ds])

    bad = _copy_emat(emat)                                                        # a 2-cycle detached from the root
    a, b = inner, int(emat.child0[inner])
    if emat.child0[b] >= 0:
        bad.parent[a] = b; bad.child0[b] = a
        with pytest.raises(db.DphyError):
            db.Forest(ctx, [bad], [ds])

    bad = _copy_emat(emat); bad.t[tip] = bad.t[int(bad.parent[tip])] - 1e-6       # a tip earlier than its parent
    with pytest.raises(db.DphyError) as ei:
        db.Forest(ctx, [bad], [ds])
    assert ei.value.status == db.ERR_INVALID_ARGUMENT

    bad = _copy_emat(emat); bad.mut_off[3] = bad.mut_off[emat.num_nodes] + 7      # offsets not monotone
    with pytest.raises(db.DphyError):
        db.Forest(ctx, [bad], [ds])

    bad = _copy_emat(emat)
    if len(bad.mut_site):
        bad.mut_site[0] = sites.num_sites                                          # site out of range
        with pytest.raises(db.DphyError) as ei:
            db.Forest(ctx, [bad], [ds])
        assert ei.value.status == db.ERR_OUT_OF_RANGE
    # the context stays usable after a rejected upload
    fo = db.Forest(ctx, [emat], [ds]); fo.eval_log_G(); fo.log_G(); fo.close()
    ds.close()


def test_upload_order_independence(ctx, orc):
    """Relabelling the host node indices changes nothing: the device order comes from the topology alone."""
    emat, sites, _ = synth(0, seed=77)
    rng = np.random.default_rng(3)
    n = emat.num_nodes
    perm = rng.permutation(n).astype(np.int32)          # new index of old node v
    inv = np.argsort(perm).astype(np.int32)             # old index of new node w
    def remap(a):
        return np.where(a >= 0, perm[np.maximum(a, 0)], -1).astype(np.int32)
    def csr(off, *arrs):
        cnt = (off[1:] - off[:-1])[inv]
        noff = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
        idx = np.concatenate([np.arange(off[v], off[v + 1]) for v in inv]) if len(arrs[0]) else np.zeros(0, np.int64)
        return (noff,) + tuple(a[idx.astype(np.int64)] for a in arrs)
    moff, msite, mfrom, mto, mt = csr(emat.mut_off, emat.mut_site, emat.mut_from, emat.mut_to, emat.mut_t)
    ioff, istart, iend = csr(emat.miss_off, emat.miss_start, emat.miss_end)
    foff, fsite, ffrom = csr(emat.fs_off, emat.fs_site, emat.fs_from)
    shuffled = db.HostEmat(int(perm[emat.root]), 1, parent=remap(emat.parent)[inv], child0=remap(emat.child0)[inv],
                           child1=remap(emat.child1)[inv], t=emat.t[inv], mut_off=moff, mut_site=msite, mut_from=mfrom,
                           mut_to=mto, mut_t=mt, miss_off=ioff, miss_start=istart, miss_end=iend, fs_off=foff,
                           fs_site=fsite, fs_from=ffrom)
    ds = db.DeviceSites(ctx, sites)
    fo = db.Forest(ctx, [emat, shuffled], [ds])
    fo.eval_log_G()
    rp, br, lg = fo.log_G()
    # same device order; only the placement inside the forest (vector-load alignment of the event chunks) differs
    assert rp[0] == rp[1] and br[0] == pytest.approx(br[1], rel=1e-13)
    np.testing.assert_allclose(fo.lambda_i(0), fo.lambda_i(1)[perm], rtol=1e-12)
    np.testing.assert_array_equal(fo.num_sites_missing(0), fo.num_sites_missing(1)[perm])
    fo.close(); ds.close()


def test_pinned_direct_upload_matches_staged_upload(orc):
    """dphy_forest_upload from page-locked caller arrays (DMA in place over two copy streams, flatten stages released by
    events) gives bit-identical device state to the staged upload of pageable arrays -- several trees, two site tables."""
    items = [synth(3, seed=31), synth(0, seed=32, num_tips=300), synth(1, seed=33), synth(0, seed=34, num_partitions=2)]
    with db.Context(0) as c:
        tables = [db.DeviceSites(c, it[1]) for it in items]
        idx = np.arange(len(items))
        fo_a = db.Forest(c, [it[0] for it in items], tables, sites_index=idx)
        pinned = [it[0].pinned(c) for it in items]
        fo_b = db.Forest(c, pinned, tables, sites_index=idx)
        ra, rb = fo_a.log_G(), fo_b.log_G()
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y)
        for k, (emat, sites, _) in enumerate(items):
            assert np.array_equal(fo_a.lambda_i(k), fo_b.lambda_i(k))
            assert np.array_equal(fo_a.num_sites_missing(k), fo_b.num_sites_missing(k))
            e, s = to_oracle(emat, sites)
            np.testing.assert_array_equal(fo_b.num_sites_missing(k), orc.nsmn(e, s))
            assert fo_b.tallies()[k]["num_muts"] == orc.num_muts(e, s)
        # a bad topology is still rejected on the direct path
        bad = _copy_emat(items[1][0])
        inner = int(next(v for v in range(bad.num_nodes) if bad.child0[v] >= 0 and v != bad.root))
        bad.child0[inner] = bad.child1[inner]                      # same child twice
        with pytest.raises(db.DphyError):
            db.Forest(c, [bad.pinned(c)], [tables[1]])
        fo_a.close(); fo_b.close()
        for t in tables:
            t.close()


def test_site_tallies_of_a_whole_forest(ctx, orc):
    """dphy_forest_calc_site_tallies: the per-site Gibbs inputs (calc_Ttwiddle_l, calc_num_muts_l) of every tree in one call
    == the per-tree getters == the oracle; trees with different numbers of sites share the padded output."""
    items = [synth(1, seed=41), synth(0, seed=42, num_tips=300, site_rate_heterogeneity=1), synth(0, seed=43, num_partitions=2),
             synth(0, seed=44, num_tips=64, num_sites=500)]
    tables = [db.DeviceSites(ctx, it[1]) for it in items]
    fo = db.Forest(ctx, [it[0] for it in items], tables, sites_index=np.arange(len(items)))
    tw, nm = fo.site_tallies()
    for k, (emat, sites, _) in enumerate(items):
        L = sites.num_sites
        e, s = to_oracle(emat, sites)
        np.testing.assert_array_equal(nm[k, :L], orc.num_muts_l(e, s))
        assert not nm[k, L:].any() and not tw[k, L:].any()
        want = orc.Ttwiddle_l(e, s)
        np.testing.assert_allclose(tw[k, :L], want, rtol=RTOL, atol=1e-9 * np.abs(want).max())
        np.testing.assert_allclose(tw[k, :L], fo.Ttwiddle_l(k, want_T_l_a=False)[0], rtol=1e-12, atol=1e-12 * np.abs(want).max())
    fo.close()
    for t in tables:
        t.close()


def test_set_evo_rebuilds_the_nu_tables_only_when_nu_changes(ctx, orc):
    """dphy_sites_set_evo(nu_l = NULL) keeps the cumulative-nu tables (Ttwiddle_beta_a still right under the new mu);
    with a new nu_l they are rebuilt (site-rate heterogeneity switched on after the upload)."""
    emat, sites, _ = synth(0, seed=51, num_tips=500, num_partitions=2)
    ds = db.DeviceSites(ctx, sites); fo = db.Forest(ctx, [emat], [ds])
    sites.mu = sites.mu * 0.5
    ds.set_evo(mu=sites.mu)                               # nu_l untouched
    e, s = to_oracle(emat, sites)
    want = orc.Ttwiddle_beta_a(e, s)
    np.testing.assert_allclose(fo.Ttwiddle_beta_a(0), want, rtol=RTOL, atol=1e-9 * np.abs(want).max())
    assert fo.log_G()[1][0] == pytest.approx(orc.log_G_below_root(e, s), rel=RTOL)
    sites.nu_l = np.random.default_rng(3).gamma(0.5, 2.0, size=sites.num_sites)
    ds.set_evo(nu_l=sites.nu_l)
    e, s = to_oracle(emat, sites)
    want = orc.Ttwiddle_beta_a(e, s)
    np.testing.assert_allclose(fo.Ttwiddle_beta_a(0), want, rtol=RTOL, atol=1e-9 * np.abs(want).max())
    assert fo.log_G()[1][0] == pytest.approx(orc.log_G_below_root(e, s), rel=RTOL)
    assert rel_err(ds.cum_Q_l()[1:], orc.cum_Q_l(s)[1:]) <= 1e-10
    fo.close(); ds.close()


def test_site_tallies_are_bit_reproducible(ctx):
    """calc_Ttwiddle_l / calc_T_l_a scatter O(M + I + F) terms into per-site bins from all over the tree.  The bins are fixed-point
    integers (order-independent), so repeated evaluations -- and evaluations of the same tree at another position of another forest,
    where every block and atomic lands at a different time -- give the same bits."""
    emat, sites, info = synth(2)
    other, osites, _ = synth(1)
    ds = db.DeviceSites(ctx, sites); dso = db.DeviceSites(ctx, osites)
    fo = db.Forest(ctx, [emat], [ds])
    first = fo.Ttwiddle_l(0, want_T_l_a=True)
    for _ in range(6):
        again = fo.Ttwiddle_l(0, want_T_l_a=True)
        assert np.array_equal(again[0], first[0]) and np.array_equal(again[1], first[1])
    fo2 = db.Forest(ctx, [other, emat], [dso, ds], sites_index=[0, 1])
    moved = fo2.Ttwiddle_l(1, want_T_l_a=True)
    assert np.array_equal(moved[0], first[0]) and np.array_equal(moved[1], first[1])
    st = fo.site_tallies()
    assert np.array_equal(st[0][0], first[0])
    fo.close(); fo2.close(); ds.close(); dso.close()


def test_set_evo_on_many_tables_in_one_launch(ctx, orc):
    """dphy_sites_set_evo_many (Run::push_global_params_to_subruns, core/run.cpp:267-275) == dphy_sites_set_evo table by table: the
    derived tables bit for bit, and log G of a forest over them; P = 1 and P = 2 tables, uniform and varying site rates, in one call."""
    items = [synth(0, seed=3, num_tips=300), synth(2, seed=4, num_tips=200), synth(5, seed=5, num_tips=150), synth(0, seed=6, num_tips=120)]
    many = [db.DeviceSites(ctx, it[1]) for it in items]
    single = [db.DeviceSites(ctx, it[1]) for it in items]
    fa = db.Forest(ctx, [it[0] for it in items], many, sites_index=np.arange(len(items)))
    fb = db.Forest(ctx, [it[0] for it in items], single, sites_index=np.arange(len(items)))
    try:
        for rnd in range(3):
            mus = [it[1].mu * (1.0 + 0.07 * (rnd + 1) + 0.01 * k) for k, it in enumerate(items)]
            qs = None
            if rnd == 1:          # another rate matrix too (rows still sum to zero)
                qs = []
                for it in items:
                    q = it[1].q_ab.reshape(-1, 4, 4).copy()
                    q[:, 0, 1] *= 1.5; q[:, 0, 0] = 0.0; q[:, 0, 0] = -q[:, 0, :].sum(axis=1)
                    qs.append(q)
            db.sites_set_evo_many(ctx, many, mus=mus, qs=qs)
            for k, t in enumerate(single):
                t.set_evo(mu=mus[k], q_ab=None if qs is None else qs[k])
            for a, b in zip(many, single):
                np.testing.assert_array_equal(a.cum_Q_l(), b.cum_Q_l())
                np.testing.assert_array_equal(a.state_frequencies(), b.state_frequencies())
            fa.eval_log_G(); fb.eval_log_G()
            for x, y in zip(fa.log_G(), fb.log_G()):
                np.testing.assert_array_equal(x, y)
            for k in range(len(items)):
                np.testing.assert_array_equal(fa.lambda_i(k), fb.lambda_i(k))
        # and against the oracle once
        e, s = to_oracle(items[1][0], many[1].host)
        lam = orc.lambda_i(e, s, orc.cum_Q_l(s))
        assert fa.log_G()[1][1] == pytest.approx(orc.log_G_below_root(e, s, lam), rel=RTOL)
        # a bad entry anywhere: nothing is committed
        before = many[0].cum_Q_l()
        with pytest.raises(db.DphyError):
            db.sites_set_evo_many(ctx, many, mus=[it[1].mu for it in items[:-1]] + [np.array([-1.0] * items[-1][1].num_partitions)])
        for t, it in zip(many, items):      # (the wrapper's mirror was updated before the call: put it back)
            t.host.mu = it[1].mu.copy()
        np.testing.assert_array_equal(many[0].cum_Q_l(), before)
    finally:
        fa.close(); fb.close()
        for t in many + single:
            t.close()
