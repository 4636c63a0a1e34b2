// oracle/ref_capi.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Thin extern "C" wrapper around the reference's OWN hot-path functions (compiled in place from
// /root/reference/core by oracle/Makefile into oracle/_ref/libdelphy_ref.so).  It converts the flat
// arrays of emat_oracle.h into a delphy::Phylo_tree / Global_evo_model and calls the reference's
// functions unchanged.  Used (a) to validate oracle/emat_oracle.c, (b) to generate tests/golden/*,
// (c) as the "reference" CPU baseline in bench.py.  The product never links this.
#include <chrono>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "phylo_tree.h"
#include "phylo_tree_calc.h"
#include "spr_study.h"
#include "site_deltas.h"
#include "evo_model.h"
#include "tree_partitioning.h"
#include "api.h"
#include "io.h"
#include <sstream>
#include <random>

#include "emat_oracle.h"

using namespace delphy;

// core/io.cpp reads these two globals of core/cmdline.cpp (the CLI's own argv); defined here so that linking read_maple does not drag
// the flag parser -- and the static initializers of its regex-based option library -- into a library loaded by a Python process
namespace delphy {
bool delphy_invoked_via_cli{false};
std::vector<std::string> delphy_cli_args{};
}

namespace {

auto make_tree(const orc_emat* e, const orc_sites* s) -> Phylo_tree {
  auto tree = Phylo_tree{e->num_nodes};
  tree.root = e->root;
  tree.ref_sequence.resize(s->num_sites);
  for (auto l = 0; l != s->num_sites; ++l) { tree.ref_sequence[l] = static_cast<Real_seq_letter>(s->ref[l]); }
  for (auto v = 0; v != e->num_nodes; ++v) {
    auto& node = tree.at(v);
    node.parent = e->parent[v];
    if (e->child0[v] >= 0) { node.children = {e->child0[v], e->child1[v]}; } else { node.children = {}; }
    node.t = e->t[v];
    if (e->child0[v] >= 0) {
      node.t_min = -std::numeric_limits<float>::max();
      node.t_max = +std::numeric_limits<float>::max();
    } else {
      node.t_min = node.t_max = static_cast<float>(e->t[v]);
    }
    for (auto i = e->mut_off[v]; i != e->mut_off[v + 1]; ++i) {
      node.mutations.push_back(Mutation{static_cast<Real_seq_letter>(e->mut_from[i]), e->mut_site[i],
                                        static_cast<Real_seq_letter>(e->mut_to[i]), e->mut_t[i]});
    }
    for (auto i = e->miss_off[v]; i != e->miss_off[v + 1]; ++i) {
      node.missations.intervals.insert(Site_interval{e->miss_start[i], e->miss_end[i]});
    }
    for (auto i = e->fs_off[v]; i != e->fs_off[v + 1]; ++i) {
      node.missations.from_states.insert_or_assign(e->fs_site[i], static_cast<Real_seq_letter>(e->fs_from[i]));
    }
  }
  return tree;
}

auto make_evo(const orc_sites* s) -> Global_evo_model {
  auto part = Site_vector<Partition_index>(s->partition_for_site, s->partition_for_site + s->num_sites);
  auto nu = std::vector<double>(s->nu_l, s->nu_l + s->num_sites);
  auto models = Partition_vector<Site_evo_model>(s->num_partitions);
  for (auto p = 0; p != s->num_partitions; ++p) {
    models[p].mu = s->mu[p];
    for (auto a = 0; a != 4; ++a) {
      models[p].pi_a[static_cast<Real_seq_letter>(a)] = s->pi_a[p * 4 + a];
      for (auto b = 0; b != 4; ++b) {
        models[p].q_ab[static_cast<Real_seq_letter>(a)][static_cast<Real_seq_letter>(b)] = s->q_ab[p * 16 + a * 4 + b];
      }
    }
  }
  return Global_evo_model{std::move(part), std::move(nu), std::move(models)};
}

auto make_seq(const orc_sites* s) -> Real_sequence {
  auto seq = Real_sequence(s->num_sites);
  for (auto l = 0; l != s->num_sites; ++l) { seq[l] = static_cast<Real_seq_letter>(s->ref[l]); }
  return seq;
}

auto copy_regions(const Scratch_vector<Candidate_region>& src, orc_region* out, int cap) -> int {
  static_assert(sizeof(Candidate_region) == sizeof(orc_region));
  if (std::ssize(src) > cap) { return -1; }
  for (auto i = 0; i != std::ssize(src); ++i) {
    out[i].branch = src[i].branch; out[i].mut_idx = src[i].mut_idx;
    out[i].t_min = src[i].t_min; out[i].t_max = src[i].t_max;
    out[i].min_muts = src[i].min_muts; out[i].pad_ = 0;
    out[i].log_W_over_Wmax = src[i].log_W_over_Wmax; out[i].W_over_Wmax = src[i].W_over_Wmax;
  }
  return static_cast<int>(std::ssize(src));
}

}  // namespace

extern "C" {

int ref_assert_integrity(const orc_emat* e, const orc_sites* s) {
  auto tree = make_tree(e, s);
  assert_phylo_tree_integrity(tree, true);   // CHECK-fails (aborts) on violation
  return 0;
}

void ref_state_frequencies_per_partition(const orc_sites* s, int32_t* out) {
  auto evo = make_evo(s);
  auto r = calc_state_frequencies_per_partition_of(make_seq(s), evo);
  for (auto p = 0; p != s->num_partitions; ++p) for (auto a = 0; a != 4; ++a) out[p * 4 + a] = r[p][static_cast<Real_seq_letter>(a)];
}

void ref_cum_Q_l(const orc_sites* s, double* out) {
  auto evo = make_evo(s);
  auto r = calc_cum_Q_l_for_sequence(make_seq(s), evo);
  std::memcpy(out, r.data(), sizeof(double) * r.size());
}

void ref_lambda_i(const orc_emat* e, const orc_sites* s, const double* cum_Q_l, double* out) {
  auto tree = make_tree(e, s); auto evo = make_evo(s);
  auto cq = std::vector<double>(cum_Q_l, cum_Q_l + s->num_sites + 1);
  auto r = calc_lambda_i(tree, evo, cq);
  std::memcpy(out, r.data(), sizeof(double) * r.size());
}

double ref_log_root_prior(const orc_emat* e, const orc_sites* s) {
  auto tree = make_tree(e, s); auto evo = make_evo(s);
  return calc_log_root_prior(tree, evo);
}

double ref_log_G_below_root(const orc_emat* e, const orc_sites* s) {
  auto tree = make_tree(e, s); auto evo = make_evo(s);
  return calc_log_G_below_root(tree, evo);
}

double ref_path_log_G(const orc_emat* e, const orc_sites* s, int32_t A, int32_t B) {
  auto tree = make_tree(e, s); auto evo = make_evo(s);
  auto cq = calc_cum_Q_l_for_sequence(tree.ref_sequence, evo);
  auto lambda_i = calc_lambda_i(tree, evo, cq);
  auto freqs = calc_state_frequencies_per_partition_of(tree.ref_sequence, evo);
  return calc_path_log_G(tree, A, B, evo, lambda_i, freqs);
}

void ref_num_sites_missing_at_every_node(const orc_emat* e, const orc_sites* s, int32_t* out) {
  auto tree = make_tree(e, s);
  auto r = calc_num_sites_missing_at_every_node(tree);
  std::memcpy(out, r.data(), sizeof(int32_t) * r.size());
}

int32_t ref_num_muts(const orc_emat* e, const orc_sites* s) { auto tree = make_tree(e, s); return calc_num_muts(tree); }

void ref_num_muts_ab(const orc_emat* e, const orc_sites* s, int32_t* out) {
  auto tree = make_tree(e, s);
  auto r = calc_num_muts_ab(tree);
  for (auto a = 0; a != 4; ++a) for (auto b = 0; b != 4; ++b)
    out[a * 4 + b] = r[static_cast<Real_seq_letter>(a)][static_cast<Real_seq_letter>(b)];
}

void ref_num_muts_beta_ab(const orc_emat* e, const orc_sites* s, int32_t* out) {
  auto tree = make_tree(e, s); auto evo = make_evo(s);
  auto r = calc_num_muts_beta_ab(tree, evo);
  for (auto p = 0; p != s->num_partitions; ++p) for (auto a = 0; a != 4; ++a) for (auto b = 0; b != 4; ++b)
    out[p * 16 + a * 4 + b] = r[p][static_cast<Real_seq_letter>(a)][static_cast<Real_seq_letter>(b)];
}

void ref_num_muts_l(const orc_emat* e, const orc_sites* s, int32_t* out) {
  auto tree = make_tree(e, s);
  auto r = calc_num_muts_l(tree);
  std::memcpy(out, r.data(), sizeof(int32_t) * r.size());
}

void ref_num_muts_l_ab(const orc_emat* e, const orc_sites* s, int32_t* out) {
  auto tree = make_tree(e, s);
  auto r = calc_num_muts_l_ab(tree);
  for (auto l = 0; l != s->num_sites; ++l) for (auto a = 0; a != 4; ++a) for (auto b = 0; b != 4; ++b)
    out[l * 16 + a * 4 + b] = r[l][static_cast<Real_seq_letter>(a)][static_cast<Real_seq_letter>(b)];
}

double ref_T(const orc_emat* e, const orc_sites* s) { auto tree = make_tree(e, s); return calc_T(tree); }

void ref_T_l_a(const orc_emat* e, const orc_sites* s, double* out) {
  auto tree = make_tree(e, s);
  auto r = calc_T_l_a(tree);
  for (auto l = 0; l != s->num_sites; ++l) for (auto a = 0; a != 4; ++a) out[l * 4 + a] = r[l][static_cast<Real_seq_letter>(a)];
}

void ref_Ttwiddle_l(const orc_emat* e, const orc_sites* s, double* out) {
  auto tree = make_tree(e, s); auto evo = make_evo(s);
  auto r = calc_Ttwiddle_l(tree, evo);
  std::memcpy(out, r.data(), sizeof(double) * r.size());
}

void ref_Ttwiddle_beta_a(const orc_emat* e, const orc_sites* s, double* out) {
  auto tree = make_tree(e, s); auto evo = make_evo(s);
  auto r = calc_Ttwiddle_beta_a(tree, evo);
  for (auto p = 0; p != s->num_partitions; ++p) for (auto a = 0; a != 4; ++a) out[p * 4 + a] = r[p][static_cast<Real_seq_letter>(a)];
}

int32_t ref_missing_sites_at(const orc_emat* e, const orc_sites* s, int32_t node, int32_t* starts, int32_t* ends, int32_t cap) {
  auto scope = Local_arena_scope{};
  auto tree = make_tree(e, s);
  auto r = reconstruct_missing_sites_at(tree, node);
  auto n = 0;
  for (const auto& [st, en] : r) { if (n >= cap) return -1; starts[n] = st; ends[n] = en; ++n; }
  return n;
}

// Spr_study_builder::seed_fill_from + (optionally) Spr_study ctor, exactly as emat_oracle.h describes.
int32_t ref_spr_study_build(const orc_emat* e, const orc_sites* s,
                            int32_t X, double t_X,
                            const int32_t* missing_starts, const int32_t* missing_ends, int32_t n_missing,
                            int32_t start_branch, int32_t start_mut_idx,
                            const int32_t* init_site, const uint8_t* init_from, const uint8_t* init_to, int32_t n_init,
                            int32_t max_muts_from_start, int32_t can_change_root,
                            int32_t with_weights, double lambda_X, double annealing_factor, double t_max_tip,
                            orc_region* out, int32_t cap, orc_study_summary* summary) {
  auto scope = Local_arena_scope{};
  auto tree = make_tree(e, s);
  auto missing_at_X = Scratch_interval_set{};
  for (auto i = 0; i != n_missing; ++i) { missing_at_X.insert(Site_interval{missing_starts[i], missing_ends[i]}); }
  auto deltas = Site_deltas{};
  for (auto i = 0; i != n_init; ++i) {
    deltas.insert(site_deltas_entry(init_site[i], static_cast<Real_seq_letter>(init_from[i]),
                                    static_cast<Real_seq_letter>(init_to[i])));
  }
  auto builder = Spr_study_builder{tree, X, t_X, missing_at_X};
  builder.max_muts_from_start = max_muts_from_start;
  builder.seed_fill_from(start_branch, start_mut_idx, std::move(deltas), can_change_root != 0);
  if (!with_weights || builder.result.empty()) {
    if (summary) { std::memset(summary, 0, sizeof(*summary)); summary->num_regions = (int)std::ssize(builder.result); }
    return copy_regions(builder.result, out, cap);
  }
  auto study = Spr_study{std::move(builder), lambda_X, annealing_factor, t_X, t_max_tip};
  if (summary) {
    summary->mu = study.mu; summary->log_Wmax = study.log_Wmax; summary->sum_W_over_Wmax = study.sum_W_over_Wmax;
    summary->num_regions = (int)std::ssize(study.candidate_regions);
    summary->num_missing_at_X = missing_at_X.num_sites();
  }
  return copy_regions(study.candidate_regions, out, cap);
}

// A study of an attached X seeded as Subrun::spr1_move does (core/subrun.cpp:540-553), without peel_graft.
int32_t ref_spr_study_from_attached(const orc_emat* e, const orc_sites* s, int32_t X,
                                    int32_t max_muts_from_start, int32_t can_change_root,
                                    double annealing_factor, double t_max_tip, const double* lambda_i,
                                    orc_region* out, int32_t cap, orc_study_summary* summary) {
  auto scope = Local_arena_scope{};
  auto tree = make_tree(e, s);
  auto P = tree.at(X).parent;
  auto S = tree.at(P).sibling_of(X);
  auto missing_at_X = reconstruct_missing_sites_at(tree, X);
  auto deltas = calc_site_deltas_between(tree, P, X);
  auto builder = Spr_study_builder{tree, X, tree.at(X).t, missing_at_X};
  builder.max_muts_from_start = max_muts_from_start;
  builder.seed_fill_from(S, 0, std::move(deltas), can_change_root != 0);
  if (builder.result.empty()) {
    if (summary) { std::memset(summary, 0, sizeof(*summary)); summary->num_missing_at_X = missing_at_X.num_sites(); }
    return 0;
  }
  auto study = Spr_study{std::move(builder), lambda_i[X], annealing_factor, tree.at(X).t, t_max_tip};
  if (summary) {
    summary->mu = study.mu; summary->log_Wmax = study.log_Wmax; summary->sum_W_over_Wmax = study.sum_W_over_Wmax;
    summary->num_regions = (int)std::ssize(study.candidate_regions);
    summary->num_missing_at_X = missing_at_X.num_sites();
  }
  return copy_regions(study.candidate_regions, out, cap);
}

// ---- CPU baseline timers (bench.py cpu_baseline / --impl reference) --------------------------------
// One "log-lik eval" = calc_log_root_prior + calc_lambda_i (from a cached cum_Q_l) + calc_log_G_below_root
// (BASELINE.md section 2).  Trees are built once outside the timed region.  Returns seconds for `reps` evals
// per thread; n_threads independent replicas of the same tree, one per thread (the reference's own
// parallelism is one Subrun per thread, core/run.cpp:682-693).
double ref_bench_log_G(const orc_emat* e, const orc_sites* s, int32_t reps, int32_t n_threads, double* out_log_G) {
  auto evo = make_evo(s);
  auto trees = std::vector<Phylo_tree>{};
  for (auto i = 0; i != n_threads; ++i) { trees.push_back(make_tree(e, s)); }
  auto cq = calc_cum_Q_l_for_sequence(trees[0].ref_sequence, evo);
  auto freqs = calc_state_frequencies_per_partition_of(trees[0].ref_sequence, evo);
  auto results = std::vector<double>(n_threads, 0.0);
  auto work = [&](int tid) {
    auto acc = 0.0;
    for (auto r = 0; r != reps; ++r) {
      auto lambda_i = calc_lambda_i(trees[tid], evo, cq);
      acc = calc_log_root_prior(trees[tid], evo, freqs) + calc_log_G_below_root(trees[tid], evo, lambda_i, freqs);
    }
    results[tid] = acc;
  };
  auto t0 = std::chrono::steady_clock::now();
  auto threads = std::vector<std::thread>{};
  for (auto i = 1; i < n_threads; ++i) { threads.emplace_back(work, i); }
  work(0);
  for (auto& th : threads) { th.join(); }
  auto t1 = std::chrono::steady_clock::now();
  if (out_log_G) { *out_log_G = results[0]; }
  return std::chrono::duration<double>(t1 - t0).count();
}

// Full (unbounded) SPR studies: reconstruct_missing_sites_at + seed_fill_from + Spr_study ctor for each X in
// xs[].  Thread tid handles xs[tid], xs[tid + n_threads], ...; returns seconds and the
// total number of candidate regions scored.
double ref_bench_spr(const orc_emat* e, const orc_sites* s, const int32_t* xs, int32_t n_x, int32_t n_threads,
                     double annealing_factor, double t_max_tip, int64_t* out_regions) {
  auto evo = make_evo(s);
  auto tree = make_tree(e, s);
  auto cq = calc_cum_Q_l_for_sequence(tree.ref_sequence, evo);
  auto lambda_i = calc_lambda_i(tree, evo, cq);
  auto counts = std::vector<int64_t>(n_threads, 0);
  auto work = [&](int tid) {
    for (auto k = tid; k < n_x; k += n_threads) {
      auto scope = Local_arena_scope{};
      auto X = xs[k];
      auto P = tree.at(X).parent;
      auto S = tree.at(P).sibling_of(X);
      auto missing_at_X = reconstruct_missing_sites_at(tree, X);
      auto deltas = calc_site_deltas_between(tree, P, X);
      auto builder = Spr_study_builder{tree, X, tree.at(X).t, missing_at_X};
      builder.seed_fill_from(S, 0, std::move(deltas), true);
      if (builder.result.empty()) { continue; }
      auto study = Spr_study{std::move(builder), lambda_i[X], annealing_factor, tree.at(X).t, t_max_tip};
      counts[tid] += std::ssize(study.candidate_regions);
    }
  };
  auto t0 = std::chrono::steady_clock::now();
  auto threads = std::vector<std::thread>{};
  for (auto i = 1; i < n_threads; ++i) { threads.emplace_back(work, i); }
  work(0);
  for (auto& th : threads) { th.join(); }
  auto t1 = std::chrono::steady_clock::now();
  auto total = int64_t{0};
  for (auto c : counts) { total += c; }
  if (out_regions) { *out_regions = total; }
  return std::chrono::duration<double>(t1 - t0).count();
}

// ---- tree partitioning: the reference's own stencil and parts (core/tree_partitioning.h:139-239) ------------------------------
// generate_random_partition_stencil with bitgen = std::mt19937{seed}, the generator type Run holds (core/run.h:20).
int32_t ref_partition_stencil(const orc_emat* e, const orc_sites* s, int32_t num_parts, uint32_t seed, int32_t* out_cuts, int32_t cap) {
  auto scope = Local_arena_scope{};
  auto tree = make_tree(e, s);
  auto bitgen = std::mt19937{seed};
  auto stencil = generate_random_partition_stencil(tree, num_parts, bitgen);
  if ((int32_t)stencil.size() > cap) return -1;
  for (auto i = 0; i != std::ssize(stencil); ++i) out_cuts[i] = stencil[i].cut_point;
  return (int32_t)stencil.size();
}

// partition_tree: topology + orig_tree_index of part `part_index`; returns its node count (-1 if cap is too small)
int32_t ref_partition_part(const orc_emat* e, const orc_sites* s, const int32_t* cuts, int32_t n_cuts, int32_t part_index,
                           int32_t* orig, int32_t* parent, int32_t* child0, int32_t* child1, int32_t cap, int32_t* root_part_index) {
  auto scope = Local_arena_scope{};
  auto tree = make_tree(e, s);
  auto stencil = std::vector<Partition_part_info>{};
  for (auto i = 0; i != n_cuts; ++i) stencil.push_back({cuts[i]});
  auto partition = partition_tree(tree, stencil);
  if (root_part_index) *root_part_index = partition.root_part_index();
  if (part_index < 0 || part_index >= std::ssize(partition.parts())) return -2;
  const auto& part = partition.parts()[part_index];
  if (std::ssize(part) > cap) return -1;
  for (auto v = 0; v != std::ssize(part); ++v) {
    orig[v] = part.at(v).orig_tree_index();
    parent[v] = part.at(v).parent;
    if (part.at(v).is_tip()) { child0[v] = -1; child1[v] = -1; }
    else { child0[v] = part.at(v).children[0]; child1[v] = part.at(v).children[1]; }
  }
  return (int32_t)std::ssize(part);
}

// The reference's own multithreaded scheme (core/run.cpp:682-693): the tree cut into parts, one part per worker thread, each
// thread evaluating its part's log G (Subrun::calc_cur_log_G: calc_lambda_i + [root prior] + calc_log_G_below_root) `reps` times.
// The parts are passed in already cut (Run::repartition's per-part Phylo_trees).  Returns seconds; *out_log_G = sum over parts.
double ref_bench_log_G_parts(const orc_emat* const* parts, int32_t n_parts, const orc_sites* s, int32_t reps, double* out_log_G) {
  auto evo = make_evo(s);
  auto trees = std::vector<Phylo_tree>{};
  for (auto i = 0; i != n_parts; ++i) { trees.push_back(make_tree(parts[i], s)); }
  auto cq = calc_cum_Q_l_for_sequence(trees[0].ref_sequence, evo);
  auto freqs = calc_state_frequencies_per_partition_of(trees[0].ref_sequence, evo);
  auto results = std::vector<double>(n_parts, 0.0);
  auto work = [&](int tid) {
    auto acc = 0.0;
    for (auto r = 0; r != reps; ++r) {
      auto lambda_i = calc_lambda_i(trees[tid], evo, cq);
      acc = (parts[tid]->includes_run_root ? calc_log_root_prior(trees[tid], evo, freqs) : 0.0)
          + calc_log_G_below_root(trees[tid], evo, lambda_i, freqs);
    }
    results[tid] = acc;
  };
  auto t0 = std::chrono::steady_clock::now();
  auto threads = std::vector<std::thread>{};
  for (auto i = 1; i < n_parts; ++i) { threads.emplace_back(work, i); }
  work(0);
  for (auto& th : threads) { th.join(); }
  auto t1 = std::chrono::steady_clock::now();
  if (out_log_G) { *out_log_G = 0.0; for (auto v : results) { *out_log_G += v; } }
  return std::chrono::duration<double>(t1 - t0).count();
}

// ---- the FlatBuffers wire format of a tree (core/api.fbs:13-49) -------------------------------------------------------------------
// phylo_tree_to_api_tree (core/api.cpp:34-98): the bytes the reference writes for this EMAT (size-prefixed buffer).
// Returns the length, or -1 if cap is too small.
int64_t ref_api_tree_write(const orc_emat* e, const orc_sites* s, uint8_t* out, int64_t cap) {
  auto scope = Local_arena_scope{};
  auto tree = make_tree(e, s);
  auto fb = phylo_tree_to_api_tree(tree);
  if ((int64_t)fb.size() > cap) return -1;
  std::memcpy(out, fb.data(), fb.size());
  return (int64_t)fb.size();
}

// api_tree_and_tree_info_to_phylo_tree (core/api.cpp:127-186, which ends in fix_up_missations, core/phylo_tree.cpp:379-478)
// on `buf`, flattened to the arrays of orc_emat.  Two calls: with arrays == NULL it returns the totals in counts[5] =
// {num_nodes, root, M, I, F}; with arrays it fills them (caller-allocated to those sizes; ref_out[L] = ref_sequence).
int32_t ref_api_tree_read(const uint8_t* buf, int32_t* counts, int32_t* parent, int32_t* child0, int32_t* child1, double* t,
                          int32_t* mut_off, int32_t* mut_site, uint8_t* mut_from, uint8_t* mut_to, double* mut_t,
                          int32_t* miss_off, int32_t* miss_start, int32_t* miss_end,
                          int32_t* fs_off, int32_t* fs_site, uint8_t* fs_from, uint8_t* ref_out) {
  auto scope = Local_arena_scope{};
  // the reader wants a TreeInfo buffer next to the tree (names, tip-date ranges): one with empty names and no uncertain dates
  auto api_tree = flatbuffers::GetSizePrefixedRoot<api::Tree>(buf);
  auto n = static_cast<int32_t>(api_tree->nodes()->size());
  auto blank = Phylo_tree{n};
  for (auto v = 0; v != n; ++v) {
    auto api_node = api_tree->nodes()->Get(v);
    if (api_node->left_child() != k_no_node) { blank.at(v).children = {api_node->left_child(), api_node->right_child()}; }
    blank.at(v).t_min = blank.at(v).t_max = 0.0;
  }
  auto info = phylo_tree_to_api_tree_info(blank);
  auto tree = api_tree_and_tree_info_to_phylo_tree(buf, info.data());
  auto M = 0, I = 0, F = 0;
  for (auto v = 0; v != n; ++v) {
    M += (int)std::ssize(tree.at(v).mutations);
    I += (int)tree.at(v).missations.intervals.num_intervals();
    F += (int)std::ssize(tree.at(v).missations.from_states);
  }
  counts[0] = n; counts[1] = tree.root; counts[2] = M; counts[3] = I; counts[4] = F;
  if (!parent) return 0;
  auto m = 0, i = 0, f = 0;
  for (auto v = 0; v != n; ++v) {
    const auto& node = tree.at(v);
    parent[v] = node.parent;
    if (node.is_tip()) { child0[v] = -1; child1[v] = -1; } else { child0[v] = node.children[0]; child1[v] = node.children[1]; }
    t[v] = node.t;
    mut_off[v] = m; miss_off[v] = i; fs_off[v] = f;
    for (const auto& mu : node.mutations) {
      mut_site[m] = mu.site; mut_from[m] = static_cast<uint8_t>(mu.from); mut_to[m] = static_cast<uint8_t>(mu.to); mut_t[m] = mu.t; ++m;
    }
    for (const auto& [a, b] : node.missations.intervals) { miss_start[i] = a; miss_end[i] = b; ++i; }
    for (const auto& [l, st] : node.missations.from_states) { fs_site[f] = l; fs_from[f] = static_cast<uint8_t>(st); ++f; }
  }
  mut_off[n] = m; miss_off[n] = i; fs_off[n] = f;
  for (auto l = 0; l != std::ssize(tree.ref_sequence); ++l) { ref_out[l] = static_cast<uint8_t>(tree.ref_sequence[l]); }
  return 0;
}

// ---- MAPLE input (core/io.cpp:98-254) ----------------------------------------------------------------------------------------------
// read_maple on `text`.  Two calls: with ref_out == NULL only counts[6] = {L, num_tips, total deltas, total intervals, total name
// bytes, number of warnings} is filled; with the arrays (sized from the first call) everything.  Returns 0, or -1 if read_maple throws.
int32_t ref_maple_read(const char* text, int64_t len, int64_t* counts, uint8_t* ref_out, double* t_min, double* t_max,
                       int64_t* name_off, char* names, int32_t* delta_off, int32_t* delta_site, uint8_t* delta_from, uint8_t* delta_to,
                       int32_t* miss_off, int32_t* miss_start, int32_t* miss_end) {
  auto scope = Local_arena_scope{};
  auto is = std::istringstream{std::string(text, static_cast<size_t>(len))};
  auto num_warnings = int64_t{0};
  auto mf = Maple_file{};
  try {
    mf = read_maple(is, [](int, std::size_t) {}, [&](const std::string&, const Sequence_warning&) { ++num_warnings; });
  } catch (const std::exception&) {
    return -1;
  }
  auto nd = int64_t{0}, ni = int64_t{0}, nb = int64_t{0};
  for (const auto& tip : mf.tip_descs) {
    nd += std::ssize(tip.seq_deltas); ni += tip.missations.intervals.num_intervals(); nb += std::ssize(tip.name);
  }
  counts[0] = std::ssize(mf.ref_sequence); counts[1] = std::ssize(mf.tip_descs); counts[2] = nd; counts[3] = ni; counts[4] = nb;
  counts[5] = num_warnings;
  if (!ref_out) return 0;
  for (auto l = 0; l != std::ssize(mf.ref_sequence); ++l) { ref_out[l] = static_cast<uint8_t>(mf.ref_sequence[l]); }
  auto d = 0, i = 0; auto b = int64_t{0};
  for (auto k = 0; k != std::ssize(mf.tip_descs); ++k) {
    const auto& tip = mf.tip_descs[k];
    t_min[k] = tip.t_min; t_max[k] = tip.t_max;
    name_off[k] = b; std::memcpy(names + b, tip.name.data(), tip.name.size()); b += std::ssize(tip.name);
    delta_off[k] = d; miss_off[k] = i;
    for (const auto& sd : tip.seq_deltas) {
      delta_site[d] = sd.site; delta_from[d] = static_cast<uint8_t>(sd.from); delta_to[d] = static_cast<uint8_t>(sd.to); ++d;
    }
    for (const auto& [a, e] : tip.missations.intervals) { miss_start[i] = a; miss_end[i] = e; ++i; }
  }
  name_off[mf.tip_descs.size()] = b; delta_off[mf.tip_descs.size()] = d; miss_off[mf.tip_descs.size()] = i;
  return 0;
}

}  // extern "C"
