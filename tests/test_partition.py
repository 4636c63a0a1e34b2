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
