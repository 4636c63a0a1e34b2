"""Pins the oracle (oracle/emat_oracle.c) AND the compiled reference (oracle/_ref) against the reference's own
known-answer tests, transcribed from /root/reference/tests (file:line cited per test).  CPU only."""
import math

import numpy as np
import pytest

from emat_fixtures import A, Cc, G, T, DBL_MAX, complex_tree
from oracle_lib import Oracle, ref_available

IMPLS = ["oracle"] + (["ref"] if ref_available() else [])


@pytest.fixture(params=IMPLS)
def orc(request):
    return Oracle(request.param)


def _evo(sites):
    mu = lambda l: sites.mu[sites.partition_for_site[l]]
    nu = lambda l: sites.nu_l[l]
    qa = lambda l, a: -sites.q_ab[sites.partition_for_site[l], a, a]
    qab = lambda l, a, b: sites.q_ab[sites.partition_for_site[l], a, b]
    pi = lambda l, a: sites.pi_a[sites.partition_for_site[l], a]
    return mu, nu, qa, qab, pi


def test_calc_T(orc):
    # tests/phylo_tree_calc_tests.cpp:236-246
    e, s, _ = complex_tree()
    assert orc.T(e, s) == pytest.approx(1.0 + 1.0 + 2.0 + 4.0, abs=1e-6)


def test_calc_Ttwiddle_beta_a(orc):
    # tests/phylo_tree_calc_tests.cpp:248-283
    e, s, _ = complex_tree()
    nu = s.nu_l
    exp = np.zeros((2, 4))
    exp[0][A] += 0.5 * nu[0]; exp[0][T] += 0.5 * nu[0]; exp[0][T] += 0.5 * nu[0]; exp[0][Cc] += 0.5 * nu[0]
    exp[0][T] += 2.0 * nu[0]; exp[0][A] += 1.0 * nu[0]; exp[0][T] += 1.0 * nu[0]; exp[0][G] += 2.0 * nu[0]
    exp[1][A] += 1.0 * nu[1]; exp[1][A] += 1.0 * nu[1]; exp[1][A] += 1.0 * nu[1]; exp[1][G] += 1.0 * nu[1]
    exp[0][A] += 4.0 * nu[2]
    np.testing.assert_allclose(orc.Ttwiddle_beta_a(e, s), exp, atol=1e-6)


def _expected_T_l_a():
    # tests/phylo_tree_calc_tests.cpp:285-313
    exp = np.zeros((4, 4))
    exp[0][A] += 0.5; exp[0][T] += 0.5; exp[0][T] += 0.5; exp[0][Cc] += 0.5; exp[0][T] += 2.0
    exp[0][A] += 1.0; exp[0][T] += 1.0; exp[0][G] += 2.0
    exp[1][A] += 1.0; exp[1][A] += 1.0; exp[1][A] += 1.0; exp[1][G] += 1.0
    exp[2][A] += 4.0
    return exp


def test_calc_T_l_a(orc):
    e, s, _ = complex_tree()
    np.testing.assert_allclose(orc.T_l_a(e, s), _expected_T_l_a(), atol=1e-6)


def test_calc_Ttwiddle_l(orc):
    # tests/phylo_tree_calc_tests.cpp:315-327
    e, s, _ = complex_tree()
    _, _, qa, _, _ = _evo(s)
    Tla = _expected_T_l_a()
    exp = [sum(qa(l, a) * Tla[l][a] for a in range(4)) for l in range(4)]
    np.testing.assert_allclose(orc.Ttwiddle_l(e, s), exp, atol=1e-6)


def test_calc_log_root_prior(orc):
    # tests/phylo_tree_calc_tests.cpp:355-379 (pi=0 edge cases)
    e, s, _ = complex_tree()
    s.pi_a[0] = [0.3, 0.7, 0.0, 0.0]
    s.pi_a[1] = [0.3, 0.0, 0.7, 0.0]
    _, _, _, _, pi = _evo(s)
    exp = math.log(pi(0, A)) + math.log(pi(1, A)) + math.log(pi(2, A))
    assert orc.log_root_prior(e, s) == pytest.approx(exp, abs=1e-6)
    s.pi_a[0] = [0.0, 0.3, 0.7, 0.0]
    assert orc.log_root_prior(e, s) == -math.inf


def test_calc_log_G_below_root(orc):
    # tests/phylo_tree_calc_tests.cpp:381-439
    e, s, _ = complex_tree()
    mu, nu, qa, qab, _ = _evo(s)
    exp = 0.0
    exp += (-mu(0) * nu(0) * qa(0, A) * 0.5 + math.log(mu(0) * nu(0) * qab(0, A, T)) + -mu(0) * nu(0) * qa(0, T) * 0.5
            + -mu(0) * nu(0) * qa(0, T) * 0.5 + math.log(mu(0) * nu(0) * qab(0, T, Cc)) + -mu(0) * nu(0) * qa(0, Cc) * 0.5
            + -mu(0) * nu(0) * qa(0, T) * 2.0
            + -mu(0) * nu(0) * qa(0, A) * 1.0 + math.log(mu(0) * nu(0) * qab(0, A, T)) + -mu(0) * nu(0) * qa(0, T) * 1.0
            + math.log(mu(0) * nu(0) * qab(0, T, G)) + -mu(0) * nu(0) * qa(0, G) * 2.0)
    exp += (-mu(1) * nu(1) * qa(1, A) * 1.0 + -mu(1) * nu(1) * qa(1, A) * 1.0
            + -mu(1) * nu(1) * qa(1, A) * 1.0 + math.log(mu(1) * nu(1) * qab(1, A, G)) + -mu(1) * nu(1) * qa(1, G) * 1.0)
    exp += -mu(2) * nu(2) * qa(2, A) * 4.0
    assert orc.log_G_below_root(e, s) == pytest.approx(exp, abs=1e-6)


def test_integer_tallies(orc):
    # tests/phylo_tree_calc_tests.cpp:441-505
    e, s, n = complex_tree()
    assert orc.num_muts(e, s) == 5
    ab = np.zeros((4, 4), int); ab[A][T] += 1; ab[T][Cc] += 1; ab[A][G] += 1; ab[A][T] += 1; ab[T][G] += 1
    np.testing.assert_array_equal(orc.num_muts_ab(e, s), ab)
    bab = np.zeros((2, 4, 4), int); bab[0][A][T] += 1; bab[0][T][Cc] += 1; bab[1][A][G] += 1; bab[0][A][T] += 1; bab[0][T][G] += 1
    np.testing.assert_array_equal(orc.num_muts_beta_ab(e, s), bab)
    np.testing.assert_array_equal(orc.num_muts_l(e, s), [4, 1, 0, 0])
    lab = np.zeros((4, 4, 4), int); lab[0][A][T] += 1; lab[0][T][Cc] += 1; lab[1][A][G] += 1; lab[0][A][T] += 1; lab[0][T][G] += 1
    np.testing.assert_array_equal(orc.num_muts_l_ab(e, s), lab)
    np.testing.assert_array_equal(orc.nsmn(e, s), [1, 2, 2, 2, 2])


def test_calc_cum_Q_l(orc):
    # tests/phylo_tree_calc_tests.cpp:531-545 (sequence ACGT)
    e, s, _ = complex_tree()
    s.ref[:] = [A, Cc, G, T]
    mu, nu, qa, _, _ = _evo(s)
    exp = np.cumsum([0.0] + [mu(l) * nu(l) * qa(l, [A, Cc, G, T][l]) for l in range(4)])
    np.testing.assert_allclose(orc.cum_Q_l(s), exp, atol=1e-6)


def test_calc_lambda_i(orc):
    # tests/phylo_tree_calc_tests.cpp:557-607
    e, s, n = complex_tree()
    mu, nu, qa, _, _ = _evo(s)
    lam = orc.lambda_i(e, s)
    m = lambda l, a: mu(l) * nu(l) * qa(l, a)
    assert lam[n["r"]] == pytest.approx(m(0, A) + m(1, A) + m(2, A), abs=1e-6)
    assert lam[n["x"]] == pytest.approx(m(0, T) + m(1, A), abs=1e-6)
    assert lam[n["a"]] == pytest.approx(m(0, Cc) + m(1, A), abs=1e-6)
    assert lam[n["b"]] == pytest.approx(m(0, T) + m(1, G), abs=1e-6)
    assert lam[n["c"]] == pytest.approx(m(0, G) + m(2, A), abs=1e-6)


def test_state_frequencies(orc):
    e, s, _ = complex_tree()
    np.testing.assert_array_equal(orc.state_frequencies(s), [[1, 1, 0, 0], [2, 0, 0, 0]])


def test_missing_sites_at(orc):
    e, s, n = complex_tree()
    st, en = orc.missing_sites_at(e, s, n["a"])
    assert list(zip(st, en)) == [(2, 4)]          # x's [2,3) and r's [3,4) coalesce
    st, en = orc.missing_sites_at(e, s, n["c"])
    assert list(zip(st, en)) == [(1, 2), (3, 4)]


def _regions(arr):
    return sorted((int(r["branch"]), int(r["mut_idx"]), float(r["t_min"]), float(r["t_max"]), int(r["min_muts"])) for r in arr)


def test_spr_study_sets(orc):
    """tests/spr_study_tests.cpp:91-205: region SETS with exact integer min_muts."""
    e, s, n = complex_tree()
    r, x, a, b, c = n["r"], n["x"], n["a"], n["b"], n["c"]
    INF = 2**31 - 1
    # full_spr_study_a (:91-110): deltas x->a = {T0C}; missing_at(a) = [2,4)
    miss_a = orc.missing_sites_at(e, s, a)
    d_xa = [(0, T, Cc)]
    regs, _ = orc.spr_study(e, s, a, 1.5, miss_a, b, 0, d_xa, INF, True)
    assert _regions(regs) == sorted([
        (b, 1, -0.5, 1.0, 1), (b, 2, 1.0, 1.5, 2), (b, 0, -1.0, -0.5, 1), (r, 1, -DBL_MAX, -1.0, 1),
        (c, 0, -1.0, 0.0, 1), (c, 1, 0.0, 1.0, 1), (c, 2, 1.0, 1.5, 1)])
    # full_spr_study_a_no_root (:112-130)
    regs, _ = orc.spr_study(e, s, a, 1.5, miss_a, b, 0, d_xa, INF, False)
    assert _regions(regs) == sorted([
        (b, 1, -0.5, 1.0, 1), (b, 2, 1.0, 1.5, 2), (b, 0, -1.0, -0.5, 1),
        (c, 0, -1.0, 0.0, 1), (c, 1, 0.0, 1.0, 1), (c, 2, 1.0, 1.5, 1)])
    # study_a_up_to_1_mut_away (:132-153)
    regs, _ = orc.spr_study(e, s, a, 1.5, miss_a, b, 0, d_xa, 1, True)
    assert _regions(regs) == sorted([
        (b, 1, -0.5, 1.0, 1), (b, 2, 1.0, 1.5, 2), (b, 0, -1.0, -0.5, 1), (r, 1, -DBL_MAX, -1.0, 1),
        (c, 0, -1.0, 0.0, 1)])
    # full_spr_study_x (:155-165): deltas c->x: c is GNAN (site 0: G), x is TANN => {G0T}; missing_at(x) = [2,4)
    miss_x = orc.missing_sites_at(e, s, x)
    regs, _ = orc.spr_study(e, s, x, 0.0, miss_x, c, 2, [(0, G, T)], INF, True)
    assert _regions(regs) == [(c, 3, -DBL_MAX, 0.0, 1)]
    # full_spr_study_c (:167-181): deltas x->c = {T0G}; missing_at(c) = {1,3}
    miss_c = orc.missing_sites_at(e, s, c)
    regs, _ = orc.spr_study(e, s, c, 3.0, miss_c, x, 1, [(0, T, G)], INF, True)
    assert _regions(regs) == sorted([
        (x, 2, -DBL_MAX, 0.0, 1), (b, 0, 0.0, 1.0, 1), (b, 1, 1.0, 2.0, 1), (a, 0, 0.0, 0.5, 1), (a, 1, 0.5, 1.0, 1)])
    # full_study_new_seq (:183-205): X detached (k_no_node), TAAN, start above the root
    regs, _ = orc.spr_study(e, s, -1, 1.5, ([3], [4]), r, 1, [(0, A, T)], INF, True)
    assert _regions(regs) == sorted([
        (a, 0, 0.0, 0.5, 0), (a, 1, 0.5, 1.0, 1), (b, 0, 0.0, 1.0, 0), (b, 1, 1.0, 1.5, 1),
        (x, 1, -0.5, 0.0, 0), (x, 0, -1.0, -0.5, 1), (r, 1, -DBL_MAX, -1.0, 1),
        (c, 0, -1.0, 0.0, 1), (c, 1, 0.0, 1.0, 0), (c, 2, 1.0, 1.5, 1)])


def test_gamma_q_absolutes():
    # tests/safe_gamma_math_tests.cpp:35-63 (the three absolute values the reference pins) + scipy cross-check
    from oracle_lib import oracle
    lib = oracle()
    assert lib.orc_gamma_q_export(271.4, 6601.0) == 0.0
    assert lib.orc_gamma_q_export(1000.0, 100.0) == 1.0
    assert lib.orc_gamma_q_export(3.5, 0.0) == 1.0
    sp = pytest.importorskip("scipy.special")
    rng = np.random.default_rng(7)
    for _ in range(300):
        a = float(rng.uniform(1.0, 40.0)); x = float(rng.uniform(0.0, 80.0))
        want = float(sp.gammaincc(a, x))
        got = lib.orc_gamma_q_export(a, x)
        assert got == pytest.approx(want, rel=1e-11, abs=1e-300)
