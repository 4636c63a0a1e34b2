"""CPU-only: the C-ABI library loads and exports every symbol include/delphy_b200.h declares; without a GPU the
product path fails loudly (no CPU fallback); the synthetic generator is deterministic."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import delphy_b200 as db

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header="delphy_b200.h"):
    hdr = open(os.path.join(ROOT, "include", header)).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(dphy_[a-zA-Z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(db.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/delphy_b200.h but not exported"
    assert db.lib().dphy_version().decode().startswith("delphy_b200")


def test_input_generator_is_a_separate_library():
    """include/dphy_synth.h -> libdphy_synth.so: the product library exports none of it, and the generator links no CUDA."""
    synth = C.CDLL(db.SYNTH_LIB_PATH)
    names = _declared_symbols("dphy_synth.h")
    assert names == ["dphy_synth_default_params", "dphy_synth_free", "dphy_synth_generate"]
    prod = C.CDLL(db.LIB_PATH)
    for n in names:
        assert hasattr(synth, n)
        assert not hasattr(prod, n)
    # it carries its own copy of the host-only tree cutter (input preparation for the reference arm's partitioned CPU baseline)
    assert hasattr(synth, "dphy_partition_split") and not hasattr(synth, "dphy_ctx_create")


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(db.DphyError) as ei:
        db.Context(0)
    assert ei.value.status == db.ERR_CUDA


def test_struct_layouts_match_header():
    assert C.sizeof(db.CandidateRegion) == 48          # == sizeof(delphy::Candidate_region), core/spr_study.h:17-32
    assert C.sizeof(db.Tallies) == 8 + 64 + 24
    assert C.sizeof(db.SprSummary) == 40


def test_synth_generator_deterministic_and_shaped():
    a = db.synth_generate(db.synth_params(1))
    b = db.synth_generate(db.synth_params(1))
    for k in db.HostEmat.FIELDS_I32 + db.HostEmat.FIELDS_U8 + db.HostEmat.FIELDS_F64:
        assert np.array_equal(getattr(a[0], k), getattr(b[0], k)), k
    e, s, info = a
    assert e.num_nodes == 2 * 200 - 1 and s.num_sites == 29903
    # mutations sorted by (t, site) per branch, inside [t_parent, t_node] (core/phylo_tree.cpp:18-56)
    for v in range(e.num_nodes):
        lo, hi = e.mut_off[v], e.mut_off[v + 1]
        if v == e.root or hi - lo == 0:
            continue
        tt = e.mut_t[lo:hi]
        assert np.all(np.diff(tt) >= 0)
        assert tt[0] >= e.t[e.parent[v]] and tt[-1] <= e.t[v]
