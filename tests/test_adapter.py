"""The reference-side binding (delphy_b200/adapter/): Delphy's own C++ signatures over the C ABI.

oracle/_ref/adapter_parity links the reference's own translation units (compiled in place), the adapter and
libdelphy_b200.so, and calls delphy::f(...) and delphy::b200::f(...) side by side on a synthetic EMAT (see
delphy_b200/adapter/adapter_parity_main.cpp).  It is built where the reference checkout exists and travels to the GPU box
as a prebuilt binary; nothing here reads /root/reference at run time."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "adapter_parity")
HDR = os.path.join(ROOT, "delphy_b200", "adapter", "delphy_b200_adapter.h")


def test_adapter_declares_the_reference_signatures():
    h = open(HDR).read()
    # every hot-path function of core/phylo_tree_calc.h that SURVEY.md section 8(a) lists, same name
    for name in ["calc_num_sites_missing_at_every_node", "calc_state_frequencies_per_partition_of", "calc_T", "calc_T_l_a",
                 "calc_Ttwiddle_l", "calc_Ttwiddle_beta_a", "calc_cum_Q_l_for_sequence", "calc_lambda_for_sequence",
                 "calc_lambda_i", "calc_log_root_prior", "calc_log_G_below_root", "calc_num_muts", "calc_num_muts_ab",
                 "calc_num_muts_beta_ab", "calc_num_muts_l", "calc_num_muts_l_ab", "count_mutations"]:
        assert re.search(r"\bauto\s+%s\s*\(" % name, h), name
    for name in ["struct Spr_study_builder", "struct Spr_study", "seed_fill_from", "pick_nexus_region", "find_region",
                 "max_muts_from_start", "candidate_regions", "log_Wmax", "sum_W_over_Wmax"]:
        assert name in h, name
    assert "namespace delphy::b200" in h


def test_adapter_binary_fails_loudly_without_a_gpu():
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/adapter_parity not built (needs the reference checkout)")
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([BIN, "1"], capture_output=True, text=True, timeout=120)
    assert "no usable CUDA device (there is no CPU fallback)" in r.stdout
    assert r.returncode != 0


@pytest.mark.gpu
@pytest.mark.parametrize("config,tips,seed,studies", [
    (1, 0, 0, 6),          # 200 tips x 29,903 sites, uniform site rates
    (2, 400, 7, 6),        # Ebola-like: site-rate heterogeneity + missations
    (5, 300, 11, 4),       # mpox-like: 2 partitions, heavy missing data, 197k sites
    (3, 3000, 3, 4),       # 3,000 tips (several device tiles)
])
def test_adapter_matches_reference_functions(config, tips, seed, studies):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/adapter_parity not built (needs the reference checkout)")
    r = subprocess.run([BIN, str(config), str(tips), str(seed), str(studies)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "0 failures" in r.stdout
    assert r.stdout.count("\nok ") + r.stdout.startswith("ok ") >= 20
