"""Hand-built EMAT fixtures transcribed from the reference's own unit tests (golden vectors).

Each fixture cites the reference test source (relative to /root/reference) it was transcribed from.
Letters: A=0, C=1, G=2, T=3 (core/sequence.h:155).
"""
from __future__ import annotations

import numpy as np

from oracle_lib import Emat, Sites, emat_from_lists

A, Cc, G, T = 0, 1, 2, 3
DBL_MAX = np.finfo(np.float64).max


def complex_tree():
    """tests/phylo_tree_calc_tests.cpp:14-116 (Phylo_tree_calc_complex_test) == tests/spr_study_tests.cpp:14-82.

    Time:             -1.0          0.0        1.0        2.0        3.0
                                     +-- T0C -- a (CANN)
                        +A2N- A0T ---+ x (TANN)
    (AACA) A3N- C2A --+ r (AAAN)     +-------- A1G ------- b (TGNN)
                        +A1N--------A0T------- T0G ------------------ c (GNAN)
    """
    r, x, a, b, c = 0, 1, 2, 3, 4
    ref = [A, A, Cc, A]
    parent = [-1, r, x, x, r]
    children = [[x, c], [a, b], [], [], []]
    t = [-1.0, 0.0, 1.0, 2.0, 3.0]
    mutations = [
        [(Cc, 2, A, -DBL_MAX)],                 # r: root "mutation" C2A
        [(A, 0, T, -0.5)],                      # x
        [(T, 0, Cc, 0.5)],                      # a
        [(A, 1, G, 1.0)],                       # b
        [(A, 0, T, 0.0), (T, 0, G, 1.0)],       # c
    ]
    # Missation{3, rA}: site 3 missing from r with from-state A == ref => no from_states entry.
    # Missation{2, rA} on x: ref[2] = C, from = A => from_states entry (2, A).
    miss = [[(3, 4)], [(2, 3)], [], [], [(1, 2)]]
    from_states = [[], [(2, A)], [], [], []]
    emat = emat_from_lists(r, parent, children, t, mutations, miss, from_states)

    def qmat(base):
        q = np.zeros((4, 4))
        k = base
        for i in range(4):
            for j in range(4):
                if i != j:
                    q[i, j] = k
                    k = round(k + 0.1, 10)
        for i in range(4):
            q[i, i] = -(q[i].sum())
        return q
    # tests/phylo_tree_calc_tests.cpp:48-72
    q0 = qmat(0.6)
    q1 = qmat(2.6)
    sites = Sites(ref=ref, partition_for_site=[0, 1, 0, 1], nu_l=[0.2, 0.3, 0.4, 0.5], mu=[0.1, 1.1],
                  pi_a=[[0.05, 0.15, 0.25, 0.55], [0.07, 0.17, 0.23, 0.53]], q_ab=[q0, q1])
    names = dict(r=r, x=x, a=a, b=b, c=c)
    return emat, sites, names


def complex_tree_single_partition():
    """Same tree with the default single-partition JC-like model used by tests/spr_study_tests.cpp (no evo needed
    there); we attach a simple HKY-free model so weights can be computed."""
    emat, sites, names = complex_tree()
    q = np.full((4, 4), 1.0 / 3.0)
    np.fill_diagonal(q, -1.0)
    s1 = Sites(ref=sites.ref, partition_for_site=[0, 0, 0, 0], nu_l=[1.0, 1.0, 1.0, 1.0], mu=[0.25],
               pi_a=[[0.25, 0.25, 0.25, 0.25]], q_ab=[q])
    return emat, s1, names
