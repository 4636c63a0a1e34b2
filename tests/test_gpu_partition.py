"""Partition parts evaluated by the CUDA path and combined across ranks (Run::check_global_and_local_totals_match,
core/run.cpp:340-357): the per-cycle tallies of the parts, packed on the device, must sum to the whole tree's -- on one GPU, and
over NCCL with the parts spread over two GPUs (skipped on a one-GPU box)."""
import os

import numpy as np
import pytest

import delphy_b200 as db
from helpers import synth, to_oracle
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg,ov,ppr", [(1, {}, 4), (2, {}, 3), (0, dict(num_partitions=2, num_root_mutations=4), 2), (3, {}, 8)])
def test_parts_sum_to_whole_on_one_gpu(cfg, ov, ppr):
    import torch
    from delphy_b200 import partitioned as pp
    emat, sites, info = synth(cfg, **ov)
    with db.Context(0) as ctx:
        pt = pp.PartitionedTree(ctx, emat, sites, world=1, rank=0, parts_per_rank=ppr, torch=torch, device="cuda:0")
        assert pt.num_parts >= 2
        for scale in (1.0, 1.37):
            pt.cycle(scale)
            got = pt.totals()
            want = pp.whole_tree_totals(ctx, emat, sites, torch, "cuda:0", scale)
            pp.check_totals(got, want)
        # and against the oracle on the whole tree
        orc = Oracle("oracle")
        e, s = to_oracle(emat, sites)
        pt.cycle(1.0)
        got = pt.totals()
        lam = orc.lambda_i(e, s)
        assert got[0] == pytest.approx(orc.log_root_prior(e, s) + orc.log_G_below_root(e, s, lam), rel=1e-9)
        assert got[1] == pytest.approx(orc.T(e, s), rel=1e-9)
        assert int(got[2]) == orc.num_muts(e, s)
        np.testing.assert_array_equal(got[3:19].astype(int).reshape(4, 4), orc.num_muts_ab(e, s))
        np.testing.assert_allclose(got[19:].reshape(-1, 4), orc.Ttwiddle_beta_a(e, s).reshape(-1, 4), rtol=1e-9)
        # Run::reassemble of the untouched parts gives the tree back
        merged = pt.partition.reassemble()
        for k in db.HostEmat.FIELDS_I32 + db.HostEmat.FIELDS_U8 + db.HostEmat.FIELDS_F64:
            np.testing.assert_array_equal(getattr(merged, k), getattr(emat, k), err_msg=k)
        pt.close()


def _nccl_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    from delphy_b200 import partitioned as pp
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    emat, sites, info = synth(3)
    with db.Context(rank) as ctx:
        pt = pp.PartitionedTree(ctx, emat, sites, world, rank, parts_per_rank=2, dist=dist, torch=torch, device=f"cuda:{rank}")
        out = []
        for scale in (1.0, 0.8):
            pt.cycle(scale)
            out.append(pt.totals())
        want = [pp.whole_tree_totals(ctx, emat, sites, torch, f"cuda:{rank}", sc) for sc in (1.0, 0.8)] if rank == 0 else None
        q.put((rank, out, want, len(pt.mine), pt.num_parts))
        pt.close()
    dist.destroy_process_group()


def test_parts_on_two_gpus_all_reduced_over_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from delphy_b200 import partitioned as pp
    mctx = mp.get_context("spawn")
    q = mctx.Queue()
    port = 29700 + (os.getpid() % 200)
    procs = [mctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = res[0][2]
    for rank, out, _, nmine, nparts in res:
        assert nmine >= 1 and nparts >= 3
        for got, w in zip(out, want):
            pp.check_totals(got, w)             # every rank holds the whole-tree totals after the one all-reduce
    np.testing.assert_array_equal(res[0][1][0], res[1][1][0])
