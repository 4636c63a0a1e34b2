// adapter_parity_main.cpp -- TEST DRIVER (built by oracle/Makefile into oracle/_ref/adapter_parity; needs a GPU to run).
//
// Links the reference's OWN translation units (compiled in place from the reference checkout), the adapter
// (delphy::b200::*, same signatures) and libdelphy_b200.so into one binary, builds a delphy::Phylo_tree +
// Global_evo_model from a synthetic EMAT, and calls every hot-path function twice -- delphy::f(...) on the CPU and
// delphy::b200::f(...) on the B200 -- the way a Delphy maintainer would after applying INTEGRATION.md.
// Integer results must be identical; doubles within 1e-9 relative (BASELINE.json north_star).
// usage: adapter_parity [config 0..5] [num_tips] [seed] [num_spr_studies]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>

#include "absl/random/random.h"

#include "phylo_tree.h"
#include "phylo_tree_calc.h"
#include "site_deltas.h"
#include "spr_study.h"

#include "delphy_b200_adapter.h"
#include "dphy_synth.h"

using namespace delphy;

namespace {

int g_failures = 0;
int g_checks = 0;

auto close_rel(double a, double b, double rtol = 1e-9) -> bool {
  if (a == b) { return true; }
  if (std::isinf(a) || std::isinf(b)) { return false; }
  return std::abs(a - b) <= rtol * std::max(std::abs(a), std::abs(b)) + 1e-300;
}
auto expect(bool ok, const std::string& what) -> void {
  ++g_checks;
  if (not ok) { ++g_failures; std::printf("FAIL %s\n", what.c_str()); } else { std::printf("ok   %s\n", what.c_str()); }
}

auto tree_from(const dphy_emat_host& e, const dphy_sites_host& s) -> Phylo_tree {
  auto tree = Phylo_tree{e.num_nodes};
  tree.root = e.root;
  tree.ref_sequence.resize(s.num_sites);
  for (auto l = 0; l != s.num_sites; ++l) { tree.ref_sequence[l] = static_cast<Real_seq_letter>(s.ref[l]); }
  for (auto v = 0; v != e.num_nodes; ++v) {
    auto& node = tree.at(v);
    node.parent = e.parent[v];
    if (e.child0[v] >= 0) { node.children = {e.child0[v], e.child1[v]}; } else { node.children = {}; }
    node.t = e.t[v];
    node.t_min = e.child0[v] >= 0 ? -std::numeric_limits<float>::max() : static_cast<float>(e.t[v]);
    node.t_max = e.child0[v] >= 0 ? +std::numeric_limits<float>::max() : static_cast<float>(e.t[v]);
    for (auto i = e.mut_off[v]; i != e.mut_off[v + 1]; ++i) {
      node.mutations.push_back(Mutation{static_cast<Real_seq_letter>(e.mut_from[i]), e.mut_site[i],
                                        static_cast<Real_seq_letter>(e.mut_to[i]), e.mut_t[i]});
    }
    for (auto i = e.miss_off[v]; i != e.miss_off[v + 1]; ++i) {
      node.missations.intervals.insert(Site_interval{e.miss_start[i], e.miss_end[i]});
    }
    for (auto i = e.fs_off[v]; i != e.fs_off[v + 1]; ++i) {
      node.missations.from_states.insert_or_assign(e.fs_site[i], static_cast<Real_seq_letter>(e.fs_from[i]));
    }
  }
  return tree;
}

auto evo_from(const dphy_sites_host& s) -> Global_evo_model {
  auto part = Site_vector<Partition_index>(s.partition_for_site, s.partition_for_site + s.num_sites);
  auto nu = std::vector<double>(s.nu_l, s.nu_l + s.num_sites);
  auto models = Partition_vector<Site_evo_model>(s.num_partitions);
  for (auto p = 0; p != s.num_partitions; ++p) {
    models[p].mu = s.mu[p];
    for (auto a = 0; a != 4; ++a) {
      models[p].pi_a[static_cast<Real_seq_letter>(a)] = s.pi_a[p * 4 + a];
      for (auto b = 0; b != 4; ++b) {
        models[p].q_ab[static_cast<Real_seq_letter>(a)][static_cast<Real_seq_letter>(b)] = s.q_ab[p * 16 + a * 4 + b];
      }
    }
  }
  return Global_evo_model{std::move(part), std::move(nu), std::move(models)};
}

}  // namespace

int main(int argc, char** argv) {
  const auto config = argc > 1 ? std::atoi(argv[1]) : 2;
  const auto num_tips = argc > 2 ? std::atoi(argv[2]) : 0;
  const auto seed = argc > 3 ? std::strtoull(argv[3], nullptr, 10) : 0ull;
  const auto num_studies = argc > 4 ? std::atoi(argv[4]) : 6;

  auto params = dphy_synth_params{};
  dphy_synth_default_params(&params, config);
  if (num_tips > 0) { params.num_tips = num_tips; }
  if (seed != 0) { params.seed = seed; }
  dphy_synth_emat* synth = nullptr;
  if (dphy_synth_generate(&params, &synth) != DPHY_OK) { std::printf("synthetic EMAT generation failed\n"); return 2; }
  auto tree = tree_from(synth->emat, synth->sites);
  auto evo = evo_from(synth->sites);
  const auto t_max_tip = synth->t_max_tip;
  std::printf("EMAT: %d nodes, %d sites, %d partitions, %lld mutations, %lld missation intervals\n", (int)std::ssize(tree),
              (int)tree.num_sites(), evo.num_partitions(), (long long)synth->num_mutations, (long long)synth->num_intervals);
  dphy_synth_free(synth);
  auto scope = Local_arena_scope{};

  try {
    // ---- phylo_tree_calc.h: stateless drop-ins, one call each --------------------------------------------------------------
    expect(calc_num_sites_missing_at_every_node(tree) == b200::calc_num_sites_missing_at_every_node(tree),
           "calc_num_sites_missing_at_every_node (int, exact)");
    {
      auto want = calc_state_frequencies_per_partition_of(tree.ref_sequence, evo);
      auto got = b200::calc_state_frequencies_per_partition_of(tree.ref_sequence, evo);
      auto same = want.size() == got.size();
      for (auto p = 0; same && p != std::ssize(want); ++p) { for (auto a : k_all_real_seq_letters) { same = same && want[p][a] == got[p][a]; } }
      expect(same, "calc_state_frequencies_per_partition_of (int, exact)");
    }
    auto cum_Q_l = calc_cum_Q_l_for_sequence(tree.ref_sequence, evo);
    {
      auto got = b200::calc_cum_Q_l_for_sequence(tree.ref_sequence, evo);
      auto same = cum_Q_l.size() == got.size();
      // The reference's strictly left-to-right fp64 scan drifts by up to ~L * eps/2 relative (a few 1e-13 at L = 29,903,
      // above 1e-12 at L = 197,000: the addends take only 4P distinct values, so the rounding errors do not average out);
      // the device's chunked scan stays within ~1e-14 of the exact sums.  1e-10 covers the reference's own drift.
      auto worst = 0.0;
      for (auto l = size_t{0}; same && l != cum_Q_l.size(); ++l) {
        same = close_rel(cum_Q_l[l], got[l], 1e-10);
        if (cum_Q_l[l] != 0.0) { worst = std::max(worst, std::abs(cum_Q_l[l] - got[l]) / std::abs(cum_Q_l[l])); }
      }
      expect(same, "calc_cum_Q_l_for_sequence (1e-10; worst rel " + std::to_string(worst * 1e12) + "e-12)");
      expect(close_rel(calc_lambda_for_sequence(tree.ref_sequence, evo), b200::calc_lambda_for_sequence(tree.ref_sequence, evo), 1e-10),
             "calc_lambda_for_sequence (1e-10)");
    }
    auto lambda_i = calc_lambda_i(tree, evo, cum_Q_l);
    {
      auto got = b200::calc_lambda_i(tree, evo, cum_Q_l);
      auto same = lambda_i.size() == got.size();
      for (auto v = size_t{0}; same && v != lambda_i.size(); ++v) { same = close_rel(lambda_i[v], got[v]); }
      expect(same, "calc_lambda_i (1e-9)");
    }
    auto freqs = calc_state_frequencies_per_partition_of(tree.ref_sequence, evo);
    expect(close_rel(calc_log_root_prior(tree, evo), b200::calc_log_root_prior(tree, evo)), "calc_log_root_prior (1e-9)");
    expect(close_rel(calc_log_root_prior(tree, evo, freqs), b200::calc_log_root_prior(tree, evo, freqs)), "calc_log_root_prior/freqs (1e-9)");
    expect(close_rel(calc_log_G_below_root(tree, evo), b200::calc_log_G_below_root(tree, evo)), "calc_log_G_below_root (1e-9)");
    expect(close_rel(calc_log_G_below_root(tree, evo, lambda_i, freqs), b200::calc_log_G_below_root(tree, evo, lambda_i, freqs)),
           "calc_log_G_below_root/lambda_i (1e-9)");
    expect(count_mutations(tree) == b200::count_mutations(tree), "count_mutations (int, exact)");
    expect(calc_num_muts(tree) == b200::calc_num_muts(tree), "calc_num_muts (int, exact)");
    {
      auto want = calc_num_muts_ab(tree); auto got = b200::calc_num_muts_ab(tree);
      auto same = true;
      for (auto a : k_all_real_seq_letters) { for (auto b : k_all_real_seq_letters) { same = same && want[a][b] == got[a][b]; } }
      expect(same, "calc_num_muts_ab (int, exact)");
    }
    {
      auto want = calc_num_muts_beta_ab(tree, evo); auto got = b200::calc_num_muts_beta_ab(tree, evo);
      auto same = want.size() == got.size();
      for (auto p = 0; same && p != std::ssize(want); ++p) {
        for (auto a : k_all_real_seq_letters) { for (auto b : k_all_real_seq_letters) { same = same && want[p][a][b] == got[p][a][b]; } }
      }
      expect(same, "calc_num_muts_beta_ab (int, exact)");
    }
    expect(calc_num_muts_l(tree) == b200::calc_num_muts_l(tree), "calc_num_muts_l (int, exact)");
    {
      auto want = calc_num_muts_l_ab(tree); auto got = b200::calc_num_muts_l_ab(tree);
      auto same = want.size() == got.size();
      for (auto l = size_t{0}; same && l != want.size(); ++l) {
        for (auto a : k_all_real_seq_letters) { for (auto b : k_all_real_seq_letters) { same = same && want[l][a][b] == got[l][a][b]; } }
      }
      expect(same, "calc_num_muts_l_ab (int, exact)");
    }
    expect(close_rel(calc_T(tree), b200::calc_T(tree)), "calc_T (1e-9)");
    {
      // each entry is a sum of +-T_below terms of size up to T: entries that cancel to ~0 carry an absolute residue of
      // ~1e-16 * T that depends on the summation order => relative 1e-9 plus an absolute floor of 1e-9 * max|want|
      auto want = calc_T_l_a(tree); auto got = b200::calc_T_l_a(tree);
      auto same = want.size() == got.size();
      auto scale = 0.0;
      for (const auto& w : want) { for (auto a : k_all_real_seq_letters) { scale = std::max(scale, std::abs(w[a])); } }
      for (auto l = size_t{0}; same && l != want.size(); ++l) {
        for (auto a : k_all_real_seq_letters) { same = same && (close_rel(want[l][a], got[l][a]) || std::abs(want[l][a] - got[l][a]) <= 1e-9 * scale); }
      }
      expect(same, "calc_T_l_a (1e-9)");
    }
    {
      auto want = calc_Ttwiddle_l(tree, evo); auto got = b200::calc_Ttwiddle_l(tree, evo);
      auto same = want.size() == got.size();
      auto scale = 0.0;
      for (auto w : want) { scale = std::max(scale, std::abs(w)); }
      for (auto l = size_t{0}; same && l != want.size(); ++l) { same = close_rel(want[l], got[l]) || std::abs(want[l] - got[l]) <= 1e-9 * scale; }
      expect(same, "calc_Ttwiddle_l (1e-9)");
    }
    {
      auto want = calc_Ttwiddle_beta_a(tree, evo); auto got = b200::calc_Ttwiddle_beta_a(tree, evo);
      auto same = want.size() == got.size();
      for (auto p = 0; same && p != std::ssize(want); ++p) { for (auto a : k_all_real_seq_letters) { same = same && close_rel(want[p][a], got[p][a]); } }
      expect(same, "calc_Ttwiddle_beta_a (1e-9)");
    }

    // ---- resident use: one upload, many evaluations, evo change, node-time change ------------------------------------------
    {
      auto dev = b200::Device_emat{tree, evo};
      expect(close_rel(calc_log_root_prior(tree, evo) + calc_log_G_below_root(tree, evo), dev.calc_cur_log_G()), "Device_emat::calc_cur_log_G");
      auto evo2 = evo;
      for (auto& model : evo2.partition_evo_model) { model.mu *= 1.7; }
      dev.set_evo(evo2);
      expect(close_rel(calc_log_G_below_root(tree, evo2), dev.calc_log_G_below_root()), "Device_emat::set_evo -> calc_log_G_below_root");
      // move one inner node half-way towards its parent (keeps every mutation time inside its branch? no: pick a node
      // whose branch and child branches hold no mutations)
      auto moved = k_no_node;
      for (auto v = 0; v != std::ssize(tree) && moved == k_no_node; ++v) {
        const auto& node = tree.at(v);
        if (node.is_inner_node() && v != tree.root && node.mutations.empty() && tree.at(node.children[0]).mutations.empty() &&
            tree.at(node.children[1]).mutations.empty()) { moved = v; }
      }
      if (moved != k_no_node) {
        auto tree2 = tree;
        tree2.at(moved).t = 0.5 * (tree.at(moved).t + tree.at_parent_of(moved).t);
        dev.set_node_times({moved}, {tree2.at(moved).t});
        expect(close_rel(calc_log_G_below_root(tree2, evo2), dev.calc_log_G_below_root()), "Device_emat::set_node_times -> calc_log_G_below_root");
      }
    }

    // ---- spr_study.h: the call sequence of Subrun::spr1_move (core/subrun.cpp:548-553) ----------------------------------------
    auto rng = std::mt19937_64{12345};
    auto dev = b200::Device_emat{tree, evo};
    for (auto k = 0; k != num_studies; ++k) {
      auto X = k_no_node;
      do { X = static_cast<Node_index>(rng() % std::ssize(tree)); } while (X == tree.root || tree.at(X).parent == tree.root);
      const auto limit = (k % 3 == 0) ? std::numeric_limits<int>::max() : (k % 3);
      const auto t_X = tree.at(X).t;
      const auto P = tree.at(X).parent;
      const auto S = tree.at(P).sibling_of(X);
      auto missing_at_X = reconstruct_missing_sites_at(tree, X);

      auto ref_builder = Spr_study_builder{tree, X, t_X, missing_at_X};
      ref_builder.max_muts_from_start = limit;
      ref_builder.seed_fill_from(S, 0, calc_site_deltas_between(tree, P, X), true);
      auto ref_study = Spr_study{std::move(ref_builder), lambda_i.at(X), 0.8, t_X, t_max_tip};

      auto gpu_builder = b200::Spr_study_builder{tree, X, t_X, missing_at_X};
      gpu_builder.resident = &dev;
      gpu_builder.max_muts_from_start = limit;
      gpu_builder.seed_fill_from(S, 0, calc_site_deltas_between(tree, P, X), true);
      auto gpu_study = b200::Spr_study{std::move(gpu_builder), lambda_i.at(X), 0.8, t_X, t_max_tip};

      auto same = std::ssize(ref_study.candidate_regions) == std::ssize(gpu_study.candidate_regions);
      for (auto i = 0; same && i != std::ssize(ref_study.candidate_regions); ++i) {
        const auto& a = ref_study.candidate_regions[i]; const auto& b = gpu_study.candidate_regions[i];
        same = a.branch == b.branch && a.mut_idx == b.mut_idx && a.min_muts == b.min_muts && a.t_min == b.t_min && a.t_max == b.t_max &&
               close_rel(a.W_over_Wmax, b.W_over_Wmax) && std::abs(a.log_W_over_Wmax - b.log_W_over_Wmax) <= 1e-9 * (1.0 + std::abs(a.log_W_over_Wmax));
      }
      same = same && close_rel(ref_study.log_Wmax, gpu_study.log_Wmax) && close_rel(ref_study.sum_W_over_Wmax, gpu_study.sum_W_over_Wmax) &&
             close_rel(ref_study.mu, gpu_study.mu);
      // same RNG stream -> same chosen regraft (BASELINE.json: "chosen regraft under a fixed RNG stream")
      auto bitgen_a = std::mt19937_64{777u + k}; auto bitgen_b = std::mt19937_64{777u + k};
      same = same && ref_study.pick_nexus_region(bitgen_a) == gpu_study.pick_nexus_region(bitgen_b);
      same = same && ref_study.find_region(S, tree.at(P).t) == gpu_study.find_region(S, tree.at(P).t);
      expect(same, "Spr_study X=" + std::to_string(X) + " limit=" + std::to_string(limit) + " regions=" +
                       std::to_string(std::ssize(ref_study.candidate_regions)) + " (order, ints exact, weights 1e-9, picked region)");
    }

    // ---- error behaviour: the reference throws std::out_of_range for a missation outside the sequence ----------------------
    {
      auto bad = tree;
      bad.at(bad.root).mutations.push_back(Mutation{Real_seq_letter::A, static_cast<Site_index>(tree.num_sites()) + 5, Real_seq_letter::C,
                                                    -std::numeric_limits<double>::max()});
      auto threw = false;
      try { (void)b200::calc_log_G_below_root(bad, evo); } catch (const std::out_of_range&) { threw = true; }
      expect(threw, "mutation site out of range -> std::out_of_range");
    }
  } catch (const std::exception& ex) {
    std::printf("FAIL exception: %s\n", ex.what());
    return 1;
  }
  std::printf("%d checks, %d failures\n", g_checks, g_failures);
  return g_failures == 0 ? 0 : 1;
}
