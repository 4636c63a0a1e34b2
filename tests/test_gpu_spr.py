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
            assert picked[i] == w_dev                   # the device replays the reference's scan in its own order: same index, always
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


def test_grouped_studies_on_several_trees_of_a_forest(ctx, orc):
    """Full studies in numbers that take the grouped (event-scan) kernels, on several trees of one forest in ONE batch: 70 studies
    of tree 0 (two groups sharing the tree's template records + a thin remainder on the per-study kernels), 40 of tree 1 and 33 of
    tree 2 (one group each, different site tables), interleaved in the request list; weights and picks included."""
    items = [synth(0, seed=41, num_tips=300), synth(0, seed=42, num_tips=150, num_root_mutations=4), synth(1, seed=43)]
    tables = [db.DeviceSites(ctx, it[1]) for it in items]
    fo = db.Forest(ctx, [it[0] for it in items], tables, sites_index=np.arange(len(items)))
    per_tree = []
    for k, (emat, sites, info) in enumerate(items):
        lam = fo.lambda_i(k)
        nx = (70, 40, 33)[k]
        xs = [int(v) for v in np.random.default_rng(100 + k).permutation(emat.num_nodes) if v != emat.root][:nx - 2]
        xs += [int(emat.child0[emat.root]), int(emat.child1[emat.root])]
        rq = db.spr_requests_for_attached(emat, k, xs, lam, info["t_max_tip"], INF, k != 1)
        per_tree.append([(r, k, X, lam) for r, X in zip(rq, xs)])
    order = [t for trio in zip(*[pt[:33] for pt in per_tree]) for t in trio] + per_tree[0][33:] + per_tree[1][33:]
    batch = fo.spr_study_batch([o[0] for o in order])
    summ = batch.summaries()
    for i, (_, k, X, lam) in enumerate(order):
        e, s = to_oracle(items[k][0], items[k][1])
        want, ws = orc.spr_study_from_attached(e, s, X, lam, INF, k != 1, 0.8, items[k][2]["t_max_tip"])
        _cmp_regions(batch.regions(i), want)
        if len(want):
            assert summ[i].sum_W_over_Wmax == pytest.approx(ws.sum_W_over_Wmax, rel=1e-9)
            assert summ[i].log_Wmax == pytest.approx(ws.log_Wmax, rel=1e-9, abs=1e-9)
    batch.close(); fo.close()
    for t in tables:
        t.close()


@pytest.mark.parametrize("limit", [2, 3])
def test_bounded_studies_in_a_large_batch(ctx, orc, limit):
    """Radius-2 / 3 studies take the ball walk (kernels_spr_frontier.cuh) only in batches of 256 or more (a walk lasts as long as the
    largest ball of the batch): 300 of them on one tree, root children and both can_change_root settings included, against the
    reference's builder region by region."""
    emat, sites, info = synth(1, seed=77, num_root_mutations=3)
    e, s = to_oracle(emat, sites)
    ds = db.DeviceSites(ctx, sites)
    fo = db.Forest(ctx, [emat], [ds])
    lam = fo.lambda_i(0)
    xs = [int(v) for v in np.random.default_rng(5).permutation(emat.num_nodes) if v != emat.root][:298]
    xs += [int(emat.child0[emat.root]), int(emat.child1[emat.root])]
    for ccr in (True, False):
        reqs = db.spr_requests_for_attached(emat, 0, xs, lam, info["t_max_tip"], limit, ccr)
        batch = fo.spr_study_batch(reqs)
        for i, X in enumerate(xs):
            want, _ = orc.spr_study_from_attached(e, s, X, lam, limit, ccr, 0.8, info["t_max_tip"])
            _cmp_regions(batch.regions(i), want)
        batch.close()
    fo.close(); ds.close()


def _state_at_region(emat, sites, branch, k):
    """Sequence at region (branch, k): the reference overlaid with the mutations from the root down to the k-th of `branch`."""
    path = []
    v = branch
    while v >= 0:
        path.append(v)
        v = int(emat.parent[v])
    state = sites.ref.copy()
    for v in reversed(path):
        m0, m1 = int(emat.mut_off[v]), int(emat.mut_off[v + 1])
        if v == branch and v != emat.root:
            m1 = m0 + k
        for i in range(m0, m1):
            state[emat.mut_site[i]] = emat.mut_to[i]
    return state


@pytest.mark.parametrize("cfg,ov", [(0, {}), (0, dict(num_root_mutations=5, num_partitions=2)), (1, {})])
@pytest.mark.parametrize("limit", [INF, 2])
def test_builder_inputs_taken_as_given(ctx, orc, cfg, ov, limit):
    """DPHY_SPR_X_REL_START: X's state = state at the start region + the caller's deltas, missing_at_X = the caller's intervals;
    nothing about X is read from the tree (the contract of Spr_study_builder::seed_fill_from, core/spr_study.cpp:9-24, which
    Subrun::spr1_move uses on a tree mid-move, core/subrun.cpp:539-599).  Checked for (a) inputs consistent with the tree, where
    the result must also equal the FROM_TREE mode, (b) deltas / missing sets the tree knows nothing about, (c) start regions in
    the middle of a branch and away from X's sibling, (d) a detached X."""
    emat, sites, info = synth(cfg, **ov)
    e, s = to_oracle(emat, sites)
    ds = db.DeviceSites(ctx, sites)
    fo = db.Forest(ctx, [emat], [ds])
    lam = fo.lambda_i(0)
    rng = np.random.default_rng(5 + cfg)
    L = sites.num_sites
    nodes = [int(v) for v in rng.permutation(emat.num_nodes) if v != emat.root]
    reqs, wants = [], []

    def add(X, t_X, start, k0, deltas, missing, ccr, lam_X):
        reqs.append(db.spr_request(0, X, t_X, start, k0, len(deltas), lam_X, info["t_max_tip"], limit, ccr, 0.8,
                                   x_deltas=[(l, to) for (l, fr, to) in deltas], x_missing=missing,
                                   x_state_mode=db.SPR_X_REL_START))
        wants.append(orc.spr_study(e, s, X, t_X, missing, start, k0, deltas, limit, ccr, weights=(lam_X, 0.8, info["t_max_tip"])))

    for X in nodes[:12]:
        P = int(emat.parent[X])
        S = int(emat.child1[P]) if int(emat.child0[P]) == X else int(emat.child0[P])
        miss = orc.missing_sites_at(e, s, X)
        missing_mask = np.zeros(L, bool)
        for a, b in zip(*miss):
            missing_mask[a:b] = True
        # (a) consistent inputs: deltas P -> X at sites not missing at X
        dl = [(l, fr, to) for l, (fr, to) in db.net_branch_deltas(emat, X).items() if not missing_mask[l]]
        add(X, float(emat.t[X]), S, 0, dl, miss, True, float(lam[X]))
        # (b) deltas and missing set unrelated to what the tree says about X
        st0 = _state_at_region(emat, sites, S, 0)
        sites_b = rng.choice(L, size=7, replace=False)
        dl_b = [(int(l), int(st0[l]), int((st0[l] + 1 + rng.integers(3)) % 4)) for l in sites_b]
        a0 = int(rng.integers(0, L - 200))
        miss_b = ([a0, min(L - 50, a0 + 300)], [a0 + 120, min(L, a0 + 420)])
        mm = np.zeros(L, bool)
        for a, b in zip(*miss_b):
            mm[a:b] = True
        dl_b = [d for d in dl_b if not mm[d[0]]]
        add(X, float(emat.t[X]), S, 0, dl_b, miss_b, bool(rng.integers(2)), float(lam[X]))
    # (c) start regions in the middle of a branch, anywhere outside X's subtree
    with_muts = [v for v in nodes if emat.mut_off[v + 1] - emat.mut_off[v] >= 2][:6]
    for B in with_muts:
        X = next(v for v in nodes if v != B and not _is_ancestor(emat, v, B) and int(emat.parent[v]) != emat.root)
        k0 = int(rng.integers(1, emat.mut_off[B + 1] - emat.mut_off[B] + 1))
        stB = _state_at_region(emat, sites, B, k0)
        sites_c = rng.choice(L, size=3, replace=False)
        dl_c = [(int(l), int(stB[l]), int((stB[l] + 1) % 4)) for l in sites_c]
        add(X, float(emat.t[X]), B, k0, dl_c, ([], []), True, float(lam[X]))
    # (d) a sequence that is not in the tree, seeded from the root region
    nroot = int(emat.mut_off[emat.root + 1] - emat.mut_off[emat.root])
    st_root = _state_at_region(emat, sites, emat.root, nroot)
    dl_d = [(int(l), int(st_root[l]), int((st_root[l] + 2) % 4)) for l in rng.choice(L, size=4, replace=False)]
    add(-1, info["t_max_tip"], emat.root, nroot, dl_d, ([10], [60]), True, 0.7)

    batch = fo.spr_study_batch(reqs)
    summ = batch.summaries()
    for i, (want, ws) in enumerate(wants):
        _cmp_regions(batch.regions(i), want)
        if len(want):
            assert summ[i].sum_W_over_Wmax == pytest.approx(ws.sum_W_over_Wmax, rel=1e-9)
            assert summ[i].num_missing_at_X == ws.num_missing_at_X
    # (a) again through the tree-derived mode
    xs = nodes[:12]
    b2 = fo.spr_study_batch(db.spr_requests_for_attached(emat, 0, xs, lam, info["t_max_tip"], limit, True))
    for j in range(len(xs)):
        _cmp_regions(b2.regions(j), batch.regions(2 * j), weights=True)
    b2.close(); batch.close(); fo.close(); ds.close()


def _is_ancestor(emat, a, v):
    """True if a is v or an ancestor of v."""
    while v >= 0:
        if v == a:
            return True
        v = int(emat.parent[v])
    return False


def test_enumerate_then_weigh(ctx, orc):
    """The reference's two steps as two calls: Spr_study_builder::seed_fill_from (lambda_X == 0: regions, zero weights), then the
    Spr_study constructor (dphy_spr_batch_set_weights) -- also re-weighing with other parameters -- and log_alpha_in_region."""
    import ctypes as C
    from oracle_lib import OrcRegion
    emat, sites, info = synth(1)
    e, s = to_oracle(emat, sites)
    ds = db.DeviceSites(ctx, sites); fo = db.Forest(ctx, [emat], [ds])
    lam = fo.lambda_i(0)
    xs = [int(v) for v in np.random.default_rng(3).permutation(emat.num_nodes) if v != emat.root][:10]
    xs += [int(emat.child0[emat.root]), int(emat.child1[emat.root])]     # above-root regions with a real incomplete-gamma factor
    reqs = db.spr_requests_for_attached(emat, 0, xs, np.zeros_like(lam), info["t_max_tip"])
    batch = fo.spr_study_batch(reqs)
    plain = [batch.regions(i) for i in range(len(xs))]
    for i, X in enumerate(xs):
        want, _ = orc.spr_study_from_attached(e, s, X, lam, INF, True, 0.8, info["t_max_tip"])
        _cmp_regions(plain[i], want, weights=False)
        assert not plain[i]["W_over_Wmax"].any() and not plain[i]["log_W_over_Wmax"].any()
    with pytest.raises(db.DphyError):
        batch.pick_nexus_regions(np.zeros(len(xs)))
    for f, scale in ((0.8, 1.0), (0.6, 2.5)):
        batch.set_weights([(float(lam[X]) * scale, f, info["t_max_tip"]) for X in xs])
        summ = batch.summaries()
        for i, X in enumerate(xs):
            want, ws = orc.spr_study_from_attached(e, s, X, lam * scale, INF, True, f, info["t_max_tip"])
            got = batch.region_weights(i, plain[i].copy())
            _cmp_regions(got, want)
            assert summ[i].sum_W_over_Wmax == pytest.approx(ws.sum_W_over_Wmax, rel=1e-9)
            assert summ[i].log_Wmax == pytest.approx(ws.log_Wmax, rel=1e-9, abs=1e-9)
            assert summ[i].mu == pytest.approx(ws.mu, rel=1e-12)
            # log_alpha_in_region at a few regions (incl. the above-root one when there is one)
            idxs = {0, len(want) // 2, len(want) - 1} | {int(k) for k in np.nonzero(want["t_min"] == -DBL_MAX)[0]}
            for k in idxs:
                r = want[k]
                t = r["t_max"] - 0.37 * ((r["t_max"] - r["t_min"]) if r["t_min"] != -DBL_MAX else 0.05)
                la_ref = orc.lib.orc_spr_log_alpha_in_region(C.byref(e.as_struct()), want.ctypes.data_as(C.POINTER(OrcRegion)), len(want), int(k),
                                                             float(t), float(lam[X]) * scale, f, float(emat.t[X]), info["t_max_tip"],
                                                             ws.sum_W_over_Wmax)
                la = batch.log_alpha_in_region(i, int(k), float(t))
                assert la == pytest.approx(la_ref, rel=1e-9, abs=1e-9), (X, k)
    batch.close(); fo.close(); ds.close()


def test_errors_are_sticky_and_name_the_request(ctx):
    emat, sites, info = synth(0)
    ds = db.DeviceSites(ctx, sites); fo = db.Forest(ctx, [emat], [ds])
    lam = fo.lambda_i(0)
    X = int(next(v for v in range(emat.num_nodes) if v != emat.root and emat.parent[v] != emat.root and emat.child0[v] >= 0))
    good = db.spr_requests_for_attached(emat, 0, [X], lam, info["t_max_tip"])[0]
    inside = int(emat.child0[X])                               # start region inside X's subtree
    bad = db.spr_request(0, X, float(emat.t[X]), inside, 0, 0, float(lam[X]), info["t_max_tip"])
    batch = fo.spr_study_batch([good, bad])
    for call in (batch.summaries, batch.total_regions_checked, lambda: batch.regions(0), lambda: batch.pick_nexus_regions([0.1, 0.1]),
                 lambda: batch.find_region(0, 1, 0.0), batch.summaries):
        with pytest.raises(db.DphyError) as ei:
            call()
        assert "request 1" in str(ei.value)
    batch.close(); fo.close(); ds.close()


def test_batches_in_flight_share_parked_blocks(ctx):
    """Two batches are in flight at a time: the emit / normalisation passes of one run on the context's tail stream next to the
    set-up of the next, and the block of a batch destroyed before its tail has finished is parked and handed to a later batch
    (dphy_ctx::spr_blocks).  Results must not depend on any of it: a batch read after a burst of create / destroy pairs (different
    requests in between, so a reused block holds another batch's leftovers), after a forest edit that joins the tail, and two batches
    alive at once, are bit for bit the first batch's."""
    emat, sites, info = synth(0, seed=77, num_tips=1500)
    ds = db.DeviceSites(ctx, sites)
    fo = db.Forest(ctx, [emat], [ds])
    lam = fo.lambda_i(0)
    rng = np.random.default_rng(9)
    nodes = [int(v) for v in rng.permutation(emat.num_nodes) if v != emat.root]
    reqs_a = db.spr_requests_for_attached(emat, 0, nodes[:40], lam, info["t_max_tip"])
    reqs_b = db.spr_requests_for_attached(emat, 0, nodes[40:75], lam, info["t_max_tip"])

    def snapshot(reqs):
        b = fo.spr_study_batch(reqs)
        out = (b.regions(-1).copy(), [(s.num_regions, s.log_Wmax, s.sum_W_over_Wmax) for s in b.summaries()])
        b.close()
        return out

    def same(x, y):
        assert x[1] == y[1]
        assert x[0].tobytes() == y[0].tobytes()

    want_a, want_b = snapshot(reqs_a), snapshot(reqs_b)
    for _ in range(7):                                   # nothing read: every destroy finds its tail pending
        fo.spr_study_batch(reqs_b).close()
        fo.spr_study_batch(reqs_a).close()
    same(snapshot(reqs_a), want_a)
    same(snapshot(reqs_b), want_b)
    # two batches alive at once, read in the opposite order of their creation
    b1, b2 = fo.spr_study_batch(reqs_a), fo.spr_study_batch(reqs_b)
    got2 = (b2.regions(-1).copy(), [(s.num_regions, s.log_Wmax, s.sum_W_over_Wmax) for s in b2.summaries()])
    got1 = (b1.regions(-1).copy(), [(s.num_regions, s.log_Wmax, s.sum_W_over_Wmax) for s in b1.summaries()])
    b1.close(); b2.close()
    same(got1, want_a); same(got2, want_b)
    # an edit of the forest right behind an unread batch: set_node_times joins the tail before it touches the times
    fo.spr_study_batch(reqs_a).close()
    v = next(x for x in range(emat.num_nodes) if emat.child0[x] >= 0 and x != emat.root)
    fo.set_node_times(0, [v], [float(emat.t[v])])
    same(snapshot(reqs_a), want_a)
    ctx.join_side_streams(); ctx.synchronize()
    fo.close(); ds.close()
