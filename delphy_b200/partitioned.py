"""One tree cut into partition parts that live on different GPUs (BASELINE.json configs[3]).

The reference cuts its tree into `num_parts` sub-trees once per cycle (Run::repartition, core/run.cpp:110-193), lets one Subrun
per part run local moves, pushes the new global parameters to every Subrun (core/run.cpp:267-275), checks that the parts'
log G sums to the whole tree's (core/run.cpp:340-357) and merges the parts back (Run::reassemble, core/run.cpp:195-256).
Here part i lives on rank i % world.  Per cycle every rank applies the new model to its site table, evaluates its parts and
packs the additive tallies on the device; ONE all-reduce (NCCL over NVLink, 19 + 4P doubles) gives every rank the whole-tree
totals the global moves need (log G, T, num_muts, num_muts_ab, Ttwiddle_beta_a).  There is no other collective on the path.
"""
import time

import numpy as np

import delphy_b200 as db


def tallies_len(num_partitions: int) -> int:
    return 19 + 4 * num_partitions


class PartitionedTree:
    """The parts of `emat` owned by this rank, resident on `ctx`'s device."""

    def __init__(self, ctx, emat, sites, world, rank, parts_per_rank=2, seed=20251017, dist=None, torch=None, device=None):
        self.ctx, self.world, self.rank, self.dist, self.torch = ctx, world, rank, dist, torch
        self.partition = db.Partition(emat, sites, world * parts_per_rank, seed=seed)
        self.num_parts = len(self.partition.parts)
        self.mine = [i for i in range(self.num_parts) if i % world == rank]
        self.table = db.DeviceSites(ctx, sites)
        # a rank may own no part when the tree is too small to be cut that often: it then contributes zeros
        self.forest = db.Forest(ctx, [self.partition.parts[i] for i in self.mine], [self.table]) if self.mine else None
        self.n = tallies_len(sites.num_partitions)
        self.packed = torch.zeros(self.n, dtype=torch.float64, device=device)
        self.stream = torch.cuda.ExternalStream(ctx.stream, device=device)
        self.mu0 = sites.mu.copy()

    def cycle(self, mu_scale=1.0, reduce=True):
        """New global parameters -> every part re-evaluated -> packed tallies all-reduced.  Asynchronous on the ctx stream."""
        self.table.set_evo(mu=self.mu0 * mu_scale)                       # push_global_params_to_subruns
        if self.forest is not None:
            self.forest.cycle_tallies_device(self.packed.data_ptr(), self.n)
        if reduce and self.world > 1:
            with self.torch.cuda.stream(self.stream):
                self.dist.all_reduce(self.packed, op=self.dist.ReduceOp.SUM)

    def totals(self):
        self.ctx.synchronize()
        self.torch.cuda.synchronize()
        return self.packed.cpu().numpy().copy()

    def close(self):
        if self.forest is not None:
            self.forest.close()
        self.table.close()
        self.partition.close()


def whole_tree_totals(ctx, emat, sites, torch, device, mu_scale=1.0):
    """The same packed tallies computed on the uncut tree (the value the parts must sum to)."""
    table = db.DeviceSites(ctx, sites)
    table.set_evo(mu=sites.mu * mu_scale)
    fo = db.Forest(ctx, [emat], [table])
    out = torch.zeros(tallies_len(sites.num_partitions), dtype=torch.float64, device=device)
    fo.cycle_tallies_device(out.data_ptr(), len(out))
    ctx.synchronize()
    v = out.cpu().numpy().copy()
    fo.close(); table.close()
    return v


def check_totals(got, want):
    """Integers (num_muts, num_muts_ab) bit-exact; doubles within 1e-9 relative (BASELINE.json north_star)."""
    np.testing.assert_array_equal(got[2:19], want[2:19])
    np.testing.assert_allclose(got[:2], want[:2], rtol=1e-9)
    np.testing.assert_allclose(got[19:], want[19:], rtol=1e-9)


def timed_cycles(pt: PartitionedTree, steps, torch, reduce=True, barrier=None):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(3):
        pt.cycle(1.0 + 1e-3 * (i + 1), reduce)
    if barrier:
        barrier()
    ev0.record(pt.stream)
    h0 = time.perf_counter()
    for i in range(steps):
        pt.cycle(1.0 + 1e-3 * ((i % 7) + 1), reduce)
    host_ms = (time.perf_counter() - h0) * 1e3 / steps
    ev1.record(pt.stream)
    if barrier:
        barrier()
    else:
        torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / steps, host_ms
