"""Tree partitioning (the reference's Run::repartition): additivity of the tallies over parts, on CPU with the oracle,
and the N>1 sharding logic under torch.distributed (gloo, world_size 2)."""
import os
import sys

import numpy as np
import pytest

import delphy_b200 as db
from helpers import synth, to_oracle
from oracle_lib import Oracle, ref, ref_available


@pytest.mark.parametrize("cfg,ov,nparts", [(0, {}, 3), (1, {}, 4), (2, {}, 8), (0, dict(num_root_mutations=5), 2)])
def test_parts_are_valid_and_additive(cfg, ov, nparts):
    emat, sites, _ = synth(cfg, **ov)
    parts, origs, cuts = db.partition_emat(emat, sites, nparts, seed=7)
    assert 1 <= len(parts) <= nparts
    # every original node appears in exactly one part as a non-root (cut points: non-root in the parent part)
    seen = np.zeros(emat.num_nodes, int)
    for p, og in zip(parts, origs):
        nonroot = np.ones(p.num_nodes, bool); nonroot[p.root] = False
        np.add.at(seen, og[nonroot], 1)
    seen[emat.root] += 1
    assert np.all(seen == 1)
    o = Oracle("oracle")
    e, s = to_oracle(emat, sites)
    lam = o.lambda_i(e, s)
    want = o.log_root_prior(e, s) + o.log_G_below_root(e, s, lam)
    tot = 0.0; nm = 0; T = 0.0
    ab = np.zeros((4, 4), int)
    for p, og in zip(parts, origs):
        pe, _ = to_oracle(p, sites)
        if ref_available():
            import ctypes as C
            assert ref().ref_assert_integrity(C.byref(pe.as_struct()), C.byref(s.as_struct())) == 0
        pl = o.lambda_i(pe, s)
        np.testing.assert_allclose(pl, lam[og], rtol=1e-11)        # lambda at every node is unchanged by the cut
        tot += (o.log_root_prior(pe, s) if p.includes_run_root else 0.0) + o.log_G_below_root(pe, s, pl)
        nm += o.num_muts(pe, s); ab += o.num_muts_ab(pe, s); T += o.T(pe, s)
        np.testing.assert_array_equal(o.nsmn(pe, s), o.nsmn(e, s)[og])
    assert tot == pytest.approx(want, rel=1e-11)                   # Run::check_global_and_local_totals_match
    assert nm == o.num_muts(e, s) and np.array_equal(ab, o.num_muts_ab(e, s))
    assert T == pytest.approx(o.T(e, s), rel=1e-12)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    emat, sites, _ = synth(1)
    parts, origs, _ = db.partition_emat(emat, sites, 4, seed=3)
    o = Oracle("oracle")
    _, s = to_oracle(emat, sites)
    mine = [i for i in range(len(parts)) if i % world == rank]          # parts -> ranks, round robin
    acc = np.zeros(2 + 16)
    for i in mine:
        pe, _ = to_oracle(parts[i], sites)
        lam = o.lambda_i(pe, s)
        acc[0] += (o.log_root_prior(pe, s) if parts[i].includes_run_root else 0.0) + o.log_G_below_root(pe, s, lam)
        acc[1] += o.num_muts(pe, s)
        acc[2:] += o.num_muts_ab(pe, s).reshape(-1)
    t = torch.from_numpy(acc)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)                            # the per-cycle exchange
    tm = torch.tensor([float(rank + 1)])
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)                           # bench.py's max-over-ranks timing
    q.put((rank, t.numpy().copy(), float(tm.item()), len(mine)))
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    emat, sites, _ = synth(1)
    o = Oracle("oracle")
    e, s = to_oracle(emat, sites)
    want = o.log_root_prior(e, s) + o.log_G_below_root(e, s)
    for rank, acc, tmax, nmine in res:
        assert acc[0] == pytest.approx(want, rel=1e-11)
        assert int(acc[1]) == o.num_muts(e, s)
        assert np.array_equal(acc[2:].astype(int).reshape(4, 4), o.num_muts_ab(e, s))
        assert tmax == 2.0 and nmine >= 1


def _ref_stencil(e, s, nparts, seed):
    import ctypes as C
    L = ref()
    cuts = np.zeros(nparts + 2, np.int32)
    L.ref_partition_stencil.restype = C.c_int32
    n = L.ref_partition_stencil(C.byref(e.as_struct()), C.byref(s.as_struct()), nparts, C.c_uint32(seed),
                                cuts.ctypes.data_as(C.POINTER(C.c_int32)), len(cuts))
    assert n >= 0
    return cuts[:n].copy()


@pytest.mark.skipif(not ref_available(), reason="needs oracle/_ref (the reference compiled in place)")
@pytest.mark.parametrize("cfg,ov", [(0, {}), (1, {}), (2, {}), (0, dict(caterpillar=1, num_tips=300))])
def test_stencil_and_parts_match_the_compiled_reference(cfg, ov):
    """Same seed => the reference's own cut points (generate_random_partition_stencil over std::mt19937{seed}) and the same
    parts node for node (partition_tree: numbering, orig_tree_index, topology)."""
    import ctypes as C
    emat, sites, _ = synth(cfg, **ov)
    e, s = to_oracle(emat, sites)
    L = ref()
    L.ref_partition_part.restype = C.c_int32
    for nparts in (2, 3, 8):
        for seed in (1, 7, 12345, 2**31 + 5):
            want = _ref_stencil(e, s, nparts, seed)
            part = db.Partition(emat, sites, nparts, seed=seed)
            np.testing.assert_array_equal(part.cut_points, want)
            n = emat.num_nodes
            i32 = lambda: np.zeros(n, np.int32)
            for i in range(len(part.parts)):
                og, pa, c0, c1 = i32(), i32(), i32(), i32()
                rpi = C.c_int32(-1)
                cuts = np.ascontiguousarray(want, np.int32)
                m = L.ref_partition_part(C.byref(e.as_struct()), C.byref(s.as_struct()), cuts.ctypes.data_as(C.POINTER(C.c_int32)), len(cuts), i,
                                         *[a.ctypes.data_as(C.POINTER(C.c_int32)) for a in (og, pa, c0, c1)], n, C.byref(rpi))
                assert m == part.parts[i].num_nodes and rpi.value == len(part.parts) - 1
                np.testing.assert_array_equal(part.origs[i], og[:m])
                np.testing.assert_array_equal(part.parts[i].parent, pa[:m])
                np.testing.assert_array_equal(part.parts[i].child0, c0[:m])
                np.testing.assert_array_equal(part.parts[i].child1, c1[:m])
            part.close()


def _same_emat(a, b):
    assert a.root == b.root
    for k in db.HostEmat.FIELDS_I32 + db.HostEmat.FIELDS_U8 + db.HostEmat.FIELDS_F64:
        np.testing.assert_array_equal(getattr(a, k), getattr(b, k), err_msg=k)


@pytest.mark.parametrize("cfg,ov,nparts", [(0, {}, 3), (1, {}, 5), (0, dict(num_root_mutations=4), 2)])
def test_reassemble(cfg, ov, nparts):
    """Run::reassemble (core/run.cpp:195-256): untouched parts give back the tree bit for bit; edits made inside the parts
    (node times, one branch's mutation list, a subtree swap) land on the right nodes of the whole tree."""
    emat, sites, _ = synth(cfg, **ov)
    part = db.Partition(emat, sites, nparts, seed=5)
    _same_emat(part.reassemble(), emat)
    rng = np.random.default_rng(1)
    want_t = emat.t.copy()
    edited = []
    swapped = None
    for i, (p, og) in enumerate(zip(part.parts, part.origs)):
        q = db.HostEmat(p.root, p.includes_run_root, **{k: getattr(p, k).copy() for k in db.HostEmat.FIELDS_I32 + db.HostEmat.FIELDS_U8 + db.HostEmat.FIELDS_F64})
        inner = [v for v in range(q.num_nodes) if q.child0[v] >= 0 and v != q.root]
        for v in inner[:5]:                       # displace inner nodes slightly (stay above the children)
            lo = q.t[q.parent[v]]
            q.t[v] = lo + 0.5 * (q.t[v] - lo)
            want_t[og[v]] = q.t[v]
        if swapped is None and len(inner) > 0:    # flip the two children of one inner node
            v = inner[0]
            q.child0[v], q.child1[v] = q.child1[v], q.child0[v]
            swapped = (int(og[v]), int(og[q.child0[v]]), int(og[q.child1[v]]))
        edited.append(q)
    merged = part.reassemble(edited)
    np.testing.assert_array_equal(merged.t, want_t)
    v, c0, c1 = swapped
    assert merged.child0[v] == c0 and merged.child1[v] == c1 and merged.parent[c0] == v and merged.parent[c1] == v
    # everything else is untouched
    for k in ("mut_off", "mut_site", "mut_from", "mut_to", "mut_t", "miss_off", "miss_start", "miss_end", "fs_off", "fs_site", "fs_from", "parent"):
        np.testing.assert_array_equal(getattr(merged, k), getattr(emat, k), err_msg=k)
    # a part of the wrong shape is rejected
    with pytest.raises(db.DphyError):
        part.reassemble(edited[:-1])
    part.close()
