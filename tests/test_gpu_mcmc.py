"""End to end: the reference's own MCMC driver (tools/delphy.cpp + run.cpp + subrun.cpp, compiled unmodified) with the hot path
substituted at link time by this repository (delphy_b200/adapter/_build/delphy_b200_cli), against the stock build of the same
sources (oracle/_ref/delphy).  Both binaries are built where the reference checkout exists and travel prebuilt.

  * every SPR study of a run -- full and bounded, on trees mid-move (after Spr_move::peel_graft / move, including prunings at
    the root) -- is re-run by the reference's builder on the same inputs and compared region by region (DPHY_DROPIN_VERIFY);
    the reference's own CHECKs (core/subrun.cpp:605-608: min_muts of the chosen regions == analyze_graft's counts) run too;
  * posterior summaries (t_MRCA, mu, n0; BASELINE.json north_star) of the two builds agree within Monte-Carlo error -- in fact
    the chains coincide, since every quantity the device returns matches the host's to ~1e-12 and the RNG stream is shared."""
import os

import numpy as np
import pytest

import delphy_b200 as db
from delphy_b200 import mcmc
from delphy_b200.maple import write_maple

pytestmark = pytest.mark.gpu


def _need_binaries():
    if not (os.path.exists(mcmc.STOCK_CLI) and os.path.exists(mcmc.DROPIN_CLI)):
        pytest.skip("the reference CLI builds are absent (they need the reference checkout at build time)")


@pytest.fixture(scope="module")
def maple_cfg1(tmp_path_factory):
    emat, sites, info = db.synth_generate(db.synth_params(1))        # BASELINE.json configs[0]: 200 tips x 29,903 sites
    path = str(tmp_path_factory.mktemp("mcmc") / "cfg1.maple")
    assert write_maple(emat, sites, path, info["t_max_tip"]) == 200
    return path


def test_every_study_of_a_run_matches_the_reference_builder(maple_cfg1):
    _need_binaries()
    r = mcmc.run_cli(mcmc.DROPIN_CLI, maple_cfg1, 150000, threads=1, seed=3, log_every=50000,
                     env=dict(DPHY_DROPIN_VERIFY=1, DPHY_DROPIN_BOUNDED_ON_DEVICE=1), timeout=600)
    assert r["returncode"] == 0, "\n".join(r["stderr_tail"])
    assert r["samples"][-1]["step"] == 150000


def test_partitioned_run_on_the_drop_in(maple_cfg1):
    """Four subruns on four host threads, each with its own device context (core/run.cpp:682-693)."""
    _need_binaries()
    r = mcmc.run_cli(mcmc.DROPIN_CLI, maple_cfg1, 200000, threads=4, seed=5, log_every=100000, env=dict(DPHY_DROPIN_VERIFY=1), timeout=600)
    assert r["returncode"] == 0, "\n".join(r["stderr_tail"])
    s = mcmc.run_cli(mcmc.STOCK_CLI, maple_cfg1, 200000, threads=4, seed=5, log_every=100000, timeout=600)
    assert s["returncode"] == 0
    # same seed, same thread count: the reference is deterministic, and so is the substituted build
    for a, b in zip(r["samples"], s["samples"]):
        assert a["step"] == b["step"] and a["num_muts"] == b["num_muts"]
        assert a["log_G"] == pytest.approx(b["log_G"], abs=0.02)


def test_posterior_summaries_agree_with_the_stock_build(maple_cfg1):
    _need_binaries()
    steps = 1500000
    runs = {}
    for arm, binary in (("stock", mcmc.STOCK_CLI), ("dropin", mcmc.DROPIN_CLI)):
        r = mcmc.run_cli(binary, maple_cfg1, steps, threads=1, seed=11, log_every=15000, timeout=900)
        assert r["returncode"] == 0, "\n".join(r["stderr_tail"])
        runs[arm] = mcmc.posterior_means(r["samples"])
    for key in ("t_MRCA", "mu", "n0"):
        a, b = runs["stock"][key], runs["dropin"][key]
        err = np.hypot(a["sem"], b["sem"])
        assert abs(a["mean"] - b["mean"]) <= 4.0 * err + 1e-12, (key, a, b)
    # and the truth the alignment was simulated with is recovered: mu = 1.39e-3 /site/yr, tips span 2020-07..2021-01
    assert 1.0 < runs["dropin"]["mu"]["mean"] < 2.2


def test_usher_like_initial_tree_through_the_drop_in(tmp_path):
    """build_usher_like_tree (core/phylo_tree.cpp:796-1047): one full-tree SPR study of a sequence that is not in the tree
    (X == k_no_node) per tip, over the tree built so far -- every one of them on the device in the substituted build, each re-run
    by the reference's builder (DPHY_DROPIN_VERIFY).  The resulting initial trees must coincide with the stock build's."""
    _need_binaries()
    emat, sites, info = db.synth_generate(db.synth_params(3, num_tips=1200))
    path = str(tmp_path / "t1200.maple")
    write_maple(emat, sites, path, info["t_max_tip"])
    args = ["--v0-init", "old-usher-like"]
    a = mcmc.run_cli(mcmc.DROPIN_CLI, path, 2000, threads=1, seed=9, log_every=1000, extra_args=args, env=dict(DPHY_DROPIN_VERIFY=1), timeout=900)
    b = mcmc.run_cli(mcmc.STOCK_CLI, path, 2000, threads=1, seed=9, log_every=1000, extra_args=args, timeout=900)
    assert a["returncode"] == 0, "\n".join(a["stderr_tail"])
    assert b["returncode"] == 0, "\n".join(b["stderr_tail"])
    s0a, s0b = a["samples"][0], b["samples"][0]
    assert s0a["step"] == 0 and s0a["num_muts"] == s0b["num_muts"]
    assert s0a["log_G"] == pytest.approx(s0b["log_G"], abs=0.02) and s0a["T"] == pytest.approx(s0b["T"], abs=0.02)
