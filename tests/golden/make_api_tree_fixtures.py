"""Writes tests/golden/api_tree_*.bin / .npz -- run HERE (needs /root/reference compiled into oracle/_ref/libdelphy_ref.so).

For a few small synthetic EMATs: the bytes the REFERENCE's own writer produces (phylo_tree_to_api_tree, core/api.cpp:34-98, through
its FlatBufferBuilder) and the tree the REFERENCE's own reader makes of them (api_tree_and_tree_info_to_phylo_tree, core/api.cpp:127-186,
including fix_up_missations), flattened to the arrays of include/delphy_b200.h.  The oracle's restatement (oracle/emat_oracle.c) and the
device loader (dphy_forest_upload_api_trees) are both checked against these.

    python tests/golden/make_api_tree_fixtures.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import delphy_b200 as db           # noqa: E402
import oracle_lib as ol            # noqa: E402
from helpers import to_oracle, API_TREE_CASES as CASES, EMAT_FIELDS as FIELDS      # noqa: E402


def main():
    for name, (cfg, ov) in CASES.items():
        emat, sites, info = db.synth_generate(db.synth_params(cfg, **ov))
        e, s = to_oracle(emat, sites)
        data = ol.api_tree_write(e, s.ref, "ref", s)
        back, ref_seq = ol.api_tree_read(data, "ref")
        assert np.array_equal(ref_seq, s.ref)
        with open(os.path.join(HERE, f"api_tree_{name}.bin"), "wb") as f:
            f.write(data)
        np.savez_compressed(os.path.join(HERE, f"api_tree_{name}.npz"), root=np.int32(back.root), **{k: getattr(back, k) for k in FIELDS})
        print(name, "bytes", len(data), "nodes", back.num_nodes, "M", int(back.mut_off[-1]), "I", int(back.miss_off[-1]), "F", int(back.fs_off[-1]))


if __name__ == "__main__":
    main()
