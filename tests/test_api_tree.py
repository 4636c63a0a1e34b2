"""CPU-only: the oracle's restatement of the reference's tree wire format (delphy.api.Tree, core/api.fbs:13-49; reader
api_tree_and_tree_info_to_phylo_tree + fix_up_missations, core/api.cpp:127-186, core/phylo_tree.cpp:379-478; writer phylo_tree_to_api_tree,
core/api.cpp:34-98) against (1) the committed fixtures -- bytes written and re-read by the reference's own code
(tests/golden/make_api_tree_fixtures.py) -- and (2) the compiled reference itself when it is present; plus the product's host-only
header parser (dphy_api_tree_parse) on good and on damaged buffers."""
import os

import numpy as np
import pytest

import delphy_b200 as db
import oracle_lib as ol
from helpers import API_TREE_CASES, EMAT_FIELDS, assert_same_emat, free_site_for, synth, to_oracle, with_extra_intervals

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fixture(name):
    data = open(os.path.join(GOLDEN, f"api_tree_{name}.bin"), "rb").read()
    z = np.load(os.path.join(GOLDEN, f"api_tree_{name}.npz"))
    want = ol.Emat(int(z["root"]), *[z[k] for k in EMAT_FIELDS])
    return data, want


@pytest.mark.parametrize("name", sorted(API_TREE_CASES))
def test_oracle_reads_the_reference_written_fixture(name):
    data, want = _fixture(name)
    got, ref_seq = ol.api_tree_read(data)
    assert_same_emat(got, want)
    cfg, ov = API_TREE_CASES[name]
    emat, sites, _ = synth(cfg, **ov)
    assert np.array_equal(ref_seq, sites.ref)
    # the tree behind the fixture: identical apart from the times, which the format stores as float32
    assert_same_emat(got, to_oracle(emat, sites)[0], float32_times_of_b=True)
    assert int(got.fs_off[-1]) > 0          # the from_states, which the format does not carry, were reconstructed


@pytest.mark.parametrize("name", sorted(API_TREE_CASES))
def test_product_header_parser_on_the_fixture(name):
    data, want = _fixture(name)
    v = db.api_tree_parse(data)
    assert v["num_nodes"] == want.num_nodes and v["root"] == want.root
    assert v["num_mutations"] == int(want.mut_off[-1]) and v["num_missation_intervals"] == int(want.miss_off[-1])
    cfg, ov = API_TREE_CASES[name]
    assert np.array_equal(v["ref_seq"], synth(cfg, **ov)[1].ref)


def test_oracle_writer_round_trips():
    for name in sorted(API_TREE_CASES):
        data, want = _fixture(name)
        _, ref_seq = ol.api_tree_read(data)
        again = ol.api_tree_write(want, ref_seq)
        got, ref2 = ol.api_tree_read(again)
        assert_same_emat(got, want)
        assert np.array_equal(ref2, ref_seq)
        assert db.api_tree_parse(again)["num_nodes"] == want.num_nodes


@pytest.mark.skipif(not ol.ref_available(), reason="oracle/_ref/libdelphy_ref.so not built")
@pytest.mark.parametrize("cfg,tips", [(0, 300), (2, 400), (5, 250)])
def test_oracle_matches_the_compiled_reference_both_ways(cfg, tips):
    emat, sites, _ = synth(cfg, num_tips=tips)
    e, s = to_oracle(emat, sites)
    by_ref = ol.api_tree_write(e, s.ref, "ref", s)
    by_orc = ol.api_tree_write(e, s.ref)
    reads = [ol.api_tree_read(by_ref, "ref"), ol.api_tree_read(by_ref), ol.api_tree_read(by_orc, "ref"), ol.api_tree_read(by_orc)]
    for got, ref_seq in reads:
        assert_same_emat(got, reads[0][0])
        assert np.array_equal(ref_seq, s.ref)
    assert_same_emat(reads[0][0], e, float32_times_of_b=True)


def _touches(e, v, a, b):
    """does [a, b) overlap or touch a missation interval of node v?"""
    return any(e.miss_start[k] <= b and a <= e.miss_end[k] for k in range(e.miss_off[v], e.miss_off[v + 1]))


def _broken_variants(e, L):
    """Trees that fix_up_missations would rewrite (-2) or CHECK-fail on (-3).  The added intervals touch none of the node's own, so
    the lists stay ascending and apart (a buffer that is merely unsorted is another refusal)."""
    out = {}
    # a mutation on a site that is missing at its own node
    x = next(v for v in range(e.num_nodes) if v != e.root and e.mut_off[v + 1] > e.mut_off[v]
             and not _touches(e, v, int(e.mut_site[e.mut_off[v]]), int(e.mut_site[e.mut_off[v]]) + 1))
    l = int(e.mut_site[e.mut_off[x]])
    out["mutation_on_missing_site"] = (with_extra_intervals(e, {x: [(l, l + 1)]}), -2)
    # a site missing at a node and again at its child
    tip = next(v for v in range(e.num_nodes) if e.child0[v] < 0 and e.miss_off[v + 1] > e.miss_off[v]
               and not _touches(e, int(e.parent[v]), int(e.miss_start[e.miss_off[v]]), int(e.miss_start[e.miss_off[v]]) + 1))
    s0 = int(e.miss_start[e.miss_off[tip]])
    out["missing_twice_along_a_path"] = (with_extra_intervals(e, {int(e.parent[tip]): [(s0, s0 + 1)]}), -2)
    # the same site missing at both children of an inner node (the reference factors it up to the parent)
    p = next(v for v in range(e.num_nodes) if e.child0[v] >= 0 and e.child0[e.child0[v]] < 0 and e.child0[e.child1[v]] < 0)
    c0, c1 = int(e.child0[p]), int(e.child1[p])
    sites_mut = [int(e.mut_site[k]) for c in (c0, c1) for k in range(e.mut_off[c], e.mut_off[c + 1])]
    f = free_site_for(e, [c0, c1], L, sites_mut)
    out["common_missation_of_siblings"] = (with_extra_intervals(e, {c0: [(f, f + 1)], c1: [(f, f + 1)]}), -2)
    # a mutation whose `from` is not the state above it
    k = int(e.mut_off[x])
    bad = ol.Emat(e.root, e.parent, e.child0, e.child1, e.t, e.mut_off, e.mut_site, e.mut_from.copy(), e.mut_to, e.mut_t,
                  e.miss_off, e.miss_start, e.miss_end, e.fs_off, e.fs_site, e.fs_from)
    bad.mut_from[k] = next(a for a in range(4) if a != e.mut_from[k] and a != e.mut_to[k])
    out["from_state_contradiction"] = (bad, -3)
    return out


def test_oracle_refuses_trees_the_reference_would_rewrite():
    emat, sites, _ = synth(0, num_tips=120, num_sites=4000)
    e, s = to_oracle(emat, sites)
    for name, (tree, code) in _broken_variants(e, s.num_sites).items():
        data = ol.api_tree_write(tree, s.ref)
        with pytest.raises(ValueError) as ei:
            ol.api_tree_read(data)
        assert ei.value.args[0] == code, name
        if code == -2 and ol.ref_available():
            # the reference does load it -- and hands back a different tree: exactly what the refusal is about
            back, _ = ol.api_tree_read(data, "ref")
            same = all(np.array_equal(getattr(back, f), getattr(tree, f)) for f in ("mut_off", "mut_site", "miss_off", "miss_start", "miss_end"))
            assert not same, name


def test_parsers_survive_damaged_buffers():
    data, want = _fixture("small")
    rng = np.random.default_rng(5)
    # truncations: never a view that reaches past the end
    for cut in [0, 3, 8, 11, 12, 20, 47, 48, 100, len(data) // 2, len(data) - 1]:
        with pytest.raises(db.DphyError):
            db.api_tree_parse(data[:cut])
        with pytest.raises(ValueError):
            ol.api_tree_read(data[:cut])
    # random 32-bit words in the header region (size prefix, root offset, vtable, table): either refused, or a view that stays inside
    for _ in range(2000):
        b = bytearray(data)
        at = int(rng.integers(0, len(data) - 4))
        b[at:at + 4] = rng.integers(0, 256, 4, dtype=np.uint8).tobytes()
        try:
            v = db.api_tree_parse(bytes(b))
        except db.DphyError:
            continue
        # every vector lies inside the buffer (a damaged table may make two of them overlap: still memory-safe)
        assert 0 <= v["num_nodes"] and max(16 * v["num_nodes"], 16 * v["num_mutations"], 12 * v["num_missation_intervals"], v["num_sites"]) <= len(data)
        try:
            ol.api_tree_read(bytes(b))
        except ValueError:
            pass
