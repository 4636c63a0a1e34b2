// delphy_b200_adapter.cpp -- Delphy's C++ signatures implemented over the C ABI of libdelphy_b200.so.
// See delphy_b200_adapter.h.  Nothing here computes: it flattens, calls include/delphy_b200.h, and re-shapes results.
#include "delphy_b200_adapter.h"

#include <atomic>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "absl/random/distributions.h"

namespace delphy::b200 {

namespace {

thread_local dphy_ctx* tl_ctx = nullptr;
thread_local int tl_device = -1;   // -1: not chosen yet

// Host threads are spread round-robin over the first DPHY_DEVICES CUDA devices (default 1): with one Subrun per worker thread
// (core/run.cpp:682-693) that places the partition parts on the GPUs of the box.
auto next_device() -> int {
  static std::atomic<int> counter{0};
  const char* e = std::getenv("DPHY_DEVICES");
  const int n = e != nullptr ? std::max(1, std::atoi(e)) : 1;
  return counter.fetch_add(1) % n;
}

struct Ctx_reaper {   // destroys the thread's ctx at thread exit
  ~Ctx_reaper() { if (tl_ctx) { dphy_ctx_destroy(tl_ctx); tl_ctx = nullptr; } }
};
thread_local Ctx_reaper tl_reaper;

// dphy_status -> the exception the reference would have thrown
auto check(dphy_ctx* ctx, int status, const char* what) -> void {
  if (status == DPHY_OK) { return; }
  auto msg = std::string{what} + ": " + dphy_last_error(ctx);
  switch (status) {
    case DPHY_ERR_OUT_OF_RANGE: throw std::out_of_range(msg);
    case DPHY_ERR_INVALID_ARGUMENT: throw std::invalid_argument(msg);
    default: throw std::runtime_error(msg);
  }
}

auto letter(int a) -> Real_seq_letter { return static_cast<Real_seq_letter>(a); }

}  // namespace

auto set_thread_device(int device) -> void { tl_device = device; }

auto thread_ctx() -> dphy_ctx* {
  if (tl_ctx == nullptr) {
    (void)&tl_reaper;
    if (tl_device < 0) { tl_device = next_device(); }
    auto st = dphy_ctx_create(tl_device, &tl_ctx);
    if (st != DPHY_OK) {
      tl_ctx = nullptr;
      throw std::runtime_error("delphy_b200: no usable CUDA device (there is no CPU fallback)");
    }
  }
  return tl_ctx;
}

// ---- flattening -----------------------------------------------------------------------------------------------------------
auto Flat_emat::view(bool includes_run_root) const -> dphy_emat_host {
  auto e = dphy_emat_host{};
  e.num_nodes = static_cast<int32_t>(parent.size());
  e.root = root;
  e.includes_run_root = includes_run_root ? 1 : 0;
  e.parent = parent.data(); e.child0 = child0.data(); e.child1 = child1.data(); e.t = t.data();
  e.mut_off = mut_off.data(); e.mut_site = mut_site.data(); e.mut_from = mut_from.data(); e.mut_to = mut_to.data();
  e.mut_t = mut_t.data();
  e.miss_off = miss_off.data(); e.miss_start = miss_start.data(); e.miss_end = miss_end.data();
  e.fs_off = fs_off.data(); e.fs_site = fs_site.data(); e.fs_from = fs_from.data();
  return e;
}

auto Flat_sites::view() const -> dphy_sites_host {
  auto s = dphy_sites_host{};
  s.num_sites = static_cast<int32_t>(ref.size());
  s.num_partitions = static_cast<int32_t>(mu.size());
  s.ref = ref.data(); s.partition_for_site = partition_for_site.data(); s.nu_l = nu_l.data();
  s.mu = mu.data(); s.pi_a = pi_a.data(); s.q_ab = q_ab.data();
  return s;
}

auto flatten(const Phylo_tree& tree) -> Flat_emat {
  auto f = Flat_emat{};
  const auto n = static_cast<size_t>(std::ssize(tree));
  f.root = tree.root;
  f.parent.resize(n); f.child0.resize(n); f.child1.resize(n); f.t.resize(n);
  f.mut_off.resize(n + 1); f.miss_off.resize(n + 1); f.fs_off.resize(n + 1);
  auto m = size_t{0}, iv = size_t{0}, fs = size_t{0};
  for (const auto& node : tree.nodes) {
    m += node.mutations.size(); iv += node.missations.intervals.num_intervals(); fs += node.missations.from_states.size();
  }
  f.mut_site.reserve(m + 1); f.mut_from.reserve(m + 1); f.mut_to.reserve(m + 1); f.mut_t.reserve(m + 1);
  f.miss_start.reserve(iv + 1); f.miss_end.reserve(iv + 1); f.fs_site.reserve(fs + 1); f.fs_from.reserve(fs + 1);
  for (auto v = size_t{0}; v != n; ++v) {
    const auto& node = tree.nodes[v];
    f.parent[v] = node.parent;
    if (node.is_tip()) { f.child0[v] = -1; f.child1[v] = -1; }
    else { f.child0[v] = node.children[0]; f.child1[v] = node.children[1]; }
    f.t[v] = node.t;
    f.mut_off[v] = static_cast<int32_t>(f.mut_site.size());
    for (const auto& mut : node.mutations) {
      f.mut_site.push_back(mut.site); f.mut_from.push_back(static_cast<uint8_t>(mut.from));
      f.mut_to.push_back(static_cast<uint8_t>(mut.to)); f.mut_t.push_back(mut.t);
    }
    f.miss_off[v] = static_cast<int32_t>(f.miss_start.size());
    for (const auto& [start, end] : node.missations.intervals) { f.miss_start.push_back(start); f.miss_end.push_back(end); }
    f.fs_off[v] = static_cast<int32_t>(f.fs_site.size());
    for (const auto& [site, from] : node.missations.from_states) {
      f.fs_site.push_back(site); f.fs_from.push_back(static_cast<uint8_t>(from));
    }
  }
  f.mut_off[n] = static_cast<int32_t>(f.mut_site.size());
  f.miss_off[n] = static_cast<int32_t>(f.miss_start.size());
  f.fs_off[n] = static_cast<int32_t>(f.fs_site.size());
  // keep .data() non-null for empty lists
  f.mut_site.push_back(0); f.mut_from.push_back(0); f.mut_to.push_back(0); f.mut_t.push_back(0.0);
  f.miss_start.push_back(0); f.miss_end.push_back(0); f.fs_site.push_back(0); f.fs_from.push_back(0);
  return f;
}

auto flatten(const Real_sequence& ref_sequence, const Global_evo_model& evo) -> Flat_sites {
  auto s = Flat_sites{};
  const auto L = static_cast<size_t>(std::ssize(ref_sequence));
  const auto P = static_cast<size_t>(evo.num_partitions());
  if (evo.partition_for_site.size() != L || evo.nu_l.size() != L) {
    throw std::invalid_argument("delphy_b200: evo model and reference sequence disagree on the number of sites");
  }
  s.ref.resize(L); s.partition_for_site.resize(L); s.nu_l.resize(L);
  for (auto l = size_t{0}; l != L; ++l) {
    s.ref[l] = static_cast<uint8_t>(ref_sequence[l]);
    s.partition_for_site[l] = evo.partition_for_site[l];
    s.nu_l[l] = evo.nu_l[l];
  }
  s.mu.resize(P); s.pi_a.resize(P * 4); s.q_ab.resize(P * 16);
  for (auto p = size_t{0}; p != P; ++p) {
    const auto& model = evo.partition_evo_model[p];
    s.mu[p] = model.mu;
    for (auto a = 0; a != 4; ++a) {
      s.pi_a[p * 4 + a] = model.pi_a[letter(a)];
      for (auto b = 0; b != 4; ++b) { s.q_ab[p * 16 + a * 4 + b] = model.q_ab[letter(a)][letter(b)]; }
    }
  }
  return s;
}

// ---- Device_emat --------------------------------------------------------------------------------------------------------------
Device_emat::Device_emat(const Phylo_tree& tree, const Global_evo_model& evo, bool includes_run_root)
    : ctx_{thread_ctx()}, num_sites_{static_cast<int>(tree.num_sites())}, num_partitions_{evo.num_partitions()},
      num_nodes_{static_cast<int>(std::ssize(tree))}, ref_sequence_{tree.ref_sequence} {
  auto fs = flatten(tree.ref_sequence, evo);
  auto hs = fs.view();
  check(ctx_, dphy_sites_upload(ctx_, &hs, &sites_), "dphy_sites_upload");
  auto fe = flatten(tree);
  auto he = fe.view(includes_run_root);
  auto zero = int32_t{0};
  auto st = dphy_forest_upload(ctx_, 1, &he, &zero, 1, &sites_, &forest_);
  if (st != DPHY_OK) {
    auto msg = std::string{dphy_last_error(ctx_)};   // destroying the sites table may overwrite the message
    dphy_sites_destroy(ctx_, sites_);
    sites_ = nullptr;
    switch (st) {
      case DPHY_ERR_OUT_OF_RANGE: throw std::out_of_range(msg);
      case DPHY_ERR_INVALID_ARGUMENT: throw std::invalid_argument(msg);
      default: throw std::runtime_error(msg);
    }
  }
}

Device_emat::~Device_emat() {
  if (forest_) { dphy_forest_destroy(ctx_, forest_); }
  if (sites_) { dphy_sites_destroy(ctx_, sites_); }
}

auto Device_emat::set_evo(const Global_evo_model& evo) -> void {
  if (evo.num_partitions() != num_partitions_) { throw std::invalid_argument("delphy_b200: number of partitions changed"); }
  auto fs = flatten(ref_sequence_, evo);
  check(ctx_, dphy_sites_set_evo(ctx_, sites_, fs.nu_l.data(), fs.mu.data(), fs.pi_a.data(), fs.q_ab.data()), "dphy_sites_set_evo");
}

auto Device_emat::set_node_times(const std::vector<Node_index>& nodes, const std::vector<double>& t) -> void {
  if (nodes.size() != t.size()) { throw std::invalid_argument("delphy_b200: nodes and t differ in length"); }
  auto ids = std::vector<int32_t>(nodes.begin(), nodes.end());
  check(ctx_, dphy_forest_set_node_times(ctx_, forest_, 0, static_cast<int32_t>(ids.size()), ids.data(), t.data()),
        "dphy_forest_set_node_times");
}

auto Device_emat::calc_lambda_i() -> Node_vector<double> {
  auto out = Node_vector<double>(num_nodes_, 0.0);
  check(ctx_, dphy_forest_get_lambda_i(ctx_, forest_, 0, out.data()), "dphy_forest_get_lambda_i");
  return out;
}

auto Device_emat::calc_num_sites_missing_at_every_node() -> Node_vector<int> {
  static_assert(sizeof(int) == sizeof(int32_t));
  auto out = Node_vector<int>(num_nodes_, 0);
  check(ctx_, dphy_forest_get_num_sites_missing(ctx_, forest_, 0, out.data()), "dphy_forest_get_num_sites_missing");
  return out;
}

auto Device_emat::calc_log_root_prior() -> double {
  auto v = 0.0;
  check(ctx_, dphy_forest_get_log_G(ctx_, forest_, &v, nullptr, nullptr), "dphy_forest_get_log_G");
  return v;
}

auto Device_emat::calc_log_G_below_root() -> double {
  auto v = 0.0;
  check(ctx_, dphy_forest_get_log_G(ctx_, forest_, nullptr, &v, nullptr), "dphy_forest_get_log_G");
  return v;
}

auto Device_emat::calc_cur_log_G() -> double {
  auto v = 0.0;
  check(ctx_, dphy_forest_get_log_G(ctx_, forest_, nullptr, nullptr, &v), "dphy_forest_get_log_G");
  return v;
}

namespace {
auto tallies_of(dphy_ctx* ctx, dphy_forest* forest) -> dphy_tallies {
  auto t = dphy_tallies{};
  check(ctx, dphy_forest_calc_tallies(ctx, forest, &t), "dphy_forest_calc_tallies");
  return t;
}
}  // namespace

auto Device_emat::calc_num_muts() -> int { return tallies_of(ctx_, forest_).num_muts; }
auto Device_emat::calc_T() -> double { return tallies_of(ctx_, forest_).T; }

auto Device_emat::calc_num_muts_ab() -> Seq_matrix<int> {
  auto t = tallies_of(ctx_, forest_);
  auto out = Seq_matrix<int>{0};
  for (auto a = 0; a != 4; ++a) { for (auto b = 0; b != 4; ++b) { out[letter(a)][letter(b)] = t.num_muts_ab[a * 4 + b]; } }
  return out;
}

auto Device_emat::calc_num_muts_beta_ab() -> Partition_vector<Seq_matrix<int>> {
  auto flat = std::vector<int32_t>(static_cast<size_t>(num_partitions_) * 16, 0);
  check(ctx_, dphy_forest_calc_num_muts_beta_ab(ctx_, forest_, 0, flat.data()), "dphy_forest_calc_num_muts_beta_ab");
  auto out = Partition_vector<Seq_matrix<int>>(num_partitions_, Seq_matrix<int>{0});
  for (auto p = 0; p != num_partitions_; ++p) {
    for (auto a = 0; a != 4; ++a) { for (auto b = 0; b != 4; ++b) { out[p][letter(a)][letter(b)] = flat[p * 16 + a * 4 + b]; } }
  }
  return out;
}

auto Device_emat::calc_num_muts_l() -> Node_vector<int> {
  auto out = Node_vector<int>(num_sites_, 0);
  check(ctx_, dphy_forest_calc_num_muts_l(ctx_, forest_, 0, out.data(), nullptr), "dphy_forest_calc_num_muts_l");
  return out;
}

auto Device_emat::calc_num_muts_l_ab() -> Node_vector<Seq_matrix<int>> {
  auto flat = std::vector<int32_t>(static_cast<size_t>(num_sites_) * 16, 0);
  check(ctx_, dphy_forest_calc_num_muts_l(ctx_, forest_, 0, nullptr, flat.data()), "dphy_forest_calc_num_muts_l");
  auto out = Node_vector<Seq_matrix<int>>(num_sites_, Seq_matrix<int>{0});
  for (auto l = 0; l != num_sites_; ++l) {
    for (auto a = 0; a != 4; ++a) { for (auto b = 0; b != 4; ++b) { out[l][letter(a)][letter(b)] = flat[static_cast<size_t>(l) * 16 + a * 4 + b]; } }
  }
  return out;
}

auto Device_emat::calc_T_l_a() -> std::vector<Seq_vector<double>> {
  auto flat = std::vector<double>(static_cast<size_t>(num_sites_) * 4, 0.0);
  check(ctx_, dphy_forest_calc_Ttwiddle_l(ctx_, forest_, 0, nullptr, flat.data()), "dphy_forest_calc_Ttwiddle_l");
  auto out = std::vector<Seq_vector<double>>(num_sites_, Seq_vector<double>{0.0});
  for (auto l = 0; l != num_sites_; ++l) { for (auto a = 0; a != 4; ++a) { out[l][letter(a)] = flat[static_cast<size_t>(l) * 4 + a]; } }
  return out;
}

auto Device_emat::calc_Ttwiddle_l() -> std::vector<double> {
  auto out = std::vector<double>(num_sites_, 0.0);
  check(ctx_, dphy_forest_calc_Ttwiddle_l(ctx_, forest_, 0, out.data(), nullptr), "dphy_forest_calc_Ttwiddle_l");
  return out;
}

auto Device_emat::calc_Ttwiddle_beta_a() -> Partition_vector<Seq_vector<double>> {
  auto flat = std::vector<double>(static_cast<size_t>(num_partitions_) * 4, 0.0);
  check(ctx_, dphy_forest_calc_Ttwiddle_beta_a(ctx_, forest_, 0, flat.data()), "dphy_forest_calc_Ttwiddle_beta_a");
  auto out = Partition_vector<Seq_vector<double>>(num_partitions_, Seq_vector<double>{0.0});
  for (auto p = 0; p != num_partitions_; ++p) { for (auto a = 0; a != 4; ++a) { out[p][letter(a)] = flat[p * 4 + a]; } }
  return out;
}

auto Device_emat::calc_state_frequencies_per_partition() -> Partition_vector<Seq_vector<int>> {
  auto flat = std::vector<int32_t>(static_cast<size_t>(num_partitions_) * 4, 0);
  check(ctx_, dphy_calc_state_frequencies_per_partition(ctx_, sites_, flat.data()), "dphy_calc_state_frequencies_per_partition");
  auto out = Partition_vector<Seq_vector<int>>(num_partitions_, Seq_vector<int>{0});
  for (auto p = 0; p != num_partitions_; ++p) { for (auto a = 0; a != 4; ++a) { out[p][letter(a)] = flat[p * 4 + a]; } }
  return out;
}

auto Device_emat::calc_cum_Q_l() -> std::vector<double> {
  auto out = std::vector<double>(static_cast<size_t>(num_sites_) + 1, 0.0);
  check(ctx_, dphy_calc_cum_Q_l(ctx_, sites_, out.data()), "dphy_calc_cum_Q_l");
  return out;
}

// ---- stateless drop-ins ----------------------------------------------------------------------------------------------------------
namespace {

// the reference functions that take no evo model (counts, branch lengths, missing-site counts) do not depend on one
auto neutral_evo(const Phylo_tree& tree) -> Global_evo_model {
  auto evo = make_single_partition_global_evo_model(tree.num_sites());
  auto& model = evo.partition_evo_model[0];
  model.mu = 1.0;
  for (auto a = 0; a != 4; ++a) {
    model.pi_a[letter(a)] = 0.25;
    for (auto b = 0; b != 4; ++b) { model.q_ab[letter(a)][letter(b)] = a == b ? -1.0 : 1.0 / 3.0; }
  }
  return evo;
}

// a tree that only carries a sequence (for the *_for_sequence functions): a single root tip
auto lone_root(const Real_sequence& seq) -> Phylo_tree {
  auto tree = Phylo_tree{1};
  tree.root = 0;
  tree.ref_sequence = seq;
  tree.at(0).parent = k_no_node;
  tree.at(0).t = 0.0;
  tree.at(0).t_min = tree.at(0).t_max = 0.0f;
  return tree;
}

}  // namespace

auto count_mutations(const Phylo_tree& tree) -> int { return b200::calc_num_muts(tree); }

auto calc_num_sites_missing_at_every_node(const Phylo_tree& tree) -> Node_vector<int> {
  return Device_emat{tree, neutral_evo(tree)}.calc_num_sites_missing_at_every_node();
}
auto calc_state_frequencies_per_partition_of(const Real_sequence& seq, const Global_evo_model& evo)
    -> Partition_vector<Seq_vector<int>> {
  return Device_emat{lone_root(seq), evo}.calc_state_frequencies_per_partition();
}
auto calc_T(const Phylo_tree& tree) -> double { return Device_emat{tree, neutral_evo(tree)}.calc_T(); }
auto calc_T_l_a(const Phylo_tree& tree) -> std::vector<Seq_vector<double>> { return Device_emat{tree, neutral_evo(tree)}.calc_T_l_a(); }
auto calc_Ttwiddle_l(const Phylo_tree& tree, const Global_evo_model& evo) -> std::vector<double> {
  return Device_emat{tree, evo}.calc_Ttwiddle_l();
}
auto calc_Ttwiddle_beta_a(const Phylo_tree& tree, const Global_evo_model& evo) -> Partition_vector<Seq_vector<double>> {
  return Device_emat{tree, evo}.calc_Ttwiddle_beta_a();
}
auto calc_cum_Q_l_for_sequence(const Real_sequence& seq, const Global_evo_model& evo) -> std::vector<double> {
  return Device_emat{lone_root(seq), evo}.calc_cum_Q_l();
}
auto calc_lambda_for_sequence(const Real_sequence& seq, const Global_evo_model& evo) -> double {
  return b200::calc_cum_Q_l_for_sequence(seq, evo).back();
}
auto calc_lambda_i(const Phylo_tree& tree, const Global_evo_model& evo, const std::vector<double>& ref_cum_Q_l) -> Node_vector<double> {
  // ref_cum_Q_l is a pure function of (tree.ref_sequence, evo): the device derives its own copy at upload
  if (std::ssize(ref_cum_Q_l) != tree.num_sites() + 1) { throw std::invalid_argument("delphy_b200: ref_cum_Q_l must have L+1 entries"); }
  return Device_emat{tree, evo}.calc_lambda_i();
}
auto calc_log_root_prior(const Phylo_tree& tree, const Global_evo_model& evo) -> double {
  return Device_emat{tree, evo}.calc_log_root_prior();
}
auto calc_log_root_prior(const Phylo_tree& tree, const Global_evo_model& evo, const Partition_vector<Seq_vector<int>>&) -> double {
  return Device_emat{tree, evo}.calc_log_root_prior();
}
auto calc_log_G_below_root(const Phylo_tree& tree, const Global_evo_model& evo) -> double {
  return Device_emat{tree, evo}.calc_log_G_below_root();
}
auto calc_log_G_below_root(const Phylo_tree& tree, const Global_evo_model& evo, const Node_vector<double>&,
                           const Partition_vector<Seq_vector<int>>&) -> double {
  // lambda_i and the state frequencies are recomputed on the device in the same pass (they are functions of (tree, evo))
  return Device_emat{tree, evo}.calc_log_G_below_root();
}
auto calc_num_muts(const Phylo_tree& tree) -> int { return Device_emat{tree, neutral_evo(tree)}.calc_num_muts(); }
auto calc_num_muts_ab(const Phylo_tree& tree) -> Seq_matrix<int> { return Device_emat{tree, neutral_evo(tree)}.calc_num_muts_ab(); }
auto calc_num_muts_beta_ab(const Phylo_tree& tree, const Global_evo_model& evo) -> Partition_vector<Seq_matrix<int>> {
  return Device_emat{tree, evo}.calc_num_muts_beta_ab();
}
auto calc_num_muts_l(const Phylo_tree& tree) -> Node_vector<int> { return Device_emat{tree, neutral_evo(tree)}.calc_num_muts_l(); }
auto calc_num_muts_l_ab(const Phylo_tree& tree) -> Node_vector<Seq_matrix<int>> {
  return Device_emat{tree, neutral_evo(tree)}.calc_num_muts_l_ab();
}

// ---- SPR study ------------------------------------------------------------------------------------------------------------------
auto Spr_study_builder::seed_fill_from(Branch_index cur_branch, int cur_mut_idx, Site_deltas cur_to_X_deltas, bool can_change_root_in)
    -> void {
  start_branch = cur_branch;
  start_mut_idx = cur_mut_idx;
  init_min_muts = static_cast<int>(std::ssize(cur_to_X_deltas));
  can_change_root = can_change_root_in;
  seeded = true;
  result_valid = false;
  x_delta_site.clear(); x_delta_to.clear(); x_missing_start.clear(); x_missing_end.clear();
  if (X == k_no_node) {
    // X is not in the tree (build_usher_like_tree, core/phylo_tree.cpp:918-932): hand the device X's sequence as deltas
    // from the REFERENCE sequence.  X's state is the start region's state with cur_to_X_deltas applied; the caller
    // seeds from the root region, whose state is the reference sequence with the root's "mutations" applied.
    if (cur_branch != tree->root || cur_mut_idx != std::ssize(tree->at_root().mutations)) {
      throw std::invalid_argument("delphy_b200: a study of a detached X must be seeded from the root region");
    }
    auto root_state = absl::flat_hash_map<Site_index, Real_seq_letter>{};
    for (const auto& m : tree->at_root().mutations) { root_state.insert_or_assign(m.site, m.to); }
    for (const auto& [l, delta] : cur_to_X_deltas) { root_state.insert_or_assign(l, delta.to); }
    for (const auto& [l, state] : root_state) {
      if (state != tree->ref_sequence.at(l)) { x_delta_site.push_back(l); x_delta_to.push_back(static_cast<uint8_t>(state)); }
    }
    for (const auto& [start, end] : *missing_at_X) { x_missing_start.push_back(start); x_missing_end.push_back(end); }
  }
}

auto run_spr_study(const Spr_study_builder& builder, double lambda_X, double annealing_factor, double t_max_tip,
                   std::vector<Candidate_region>& regions, dphy_spr_summary& summary) -> void {
  static_assert(sizeof(Candidate_region) == sizeof(dphy_candidate_region), "Candidate_region must stay the 48-byte record");
  if (not builder.seeded) { throw std::invalid_argument("delphy_b200: Spr_study_builder::seed_fill_from was not called"); }
  auto owned = std::unique_ptr<Device_emat>{};
  auto* dev = builder.resident;
  if (dev == nullptr) {
    owned = builder.evo != nullptr ? std::make_unique<Device_emat>(*builder.tree, *builder.evo)
                                   : std::make_unique<Device_emat>(*builder.tree, neutral_evo(*builder.tree));
    dev = owned.get();
  }
  auto req = dphy_spr_request{};
  req.tree = 0;
  req.X = builder.X;
  req.t_X = builder.t_X;
  req.start_branch = builder.start_branch;
  req.start_mut_idx = builder.start_mut_idx;
  req.init_min_muts = builder.init_min_muts;
  req.max_muts_from_start = builder.max_muts_from_start;
  req.can_change_root = builder.can_change_root ? 1 : 0;
  req.lambda_X = lambda_X;
  req.annealing_factor = annealing_factor;
  req.t_max_tip = t_max_tip;
  req.n_x_deltas = static_cast<int32_t>(builder.x_delta_site.size());
  req.x_delta_site = builder.x_delta_site.data(); req.x_delta_to = builder.x_delta_to.data();
  req.n_x_missing = static_cast<int32_t>(builder.x_missing_start.size());
  req.x_missing_start = builder.x_missing_start.data(); req.x_missing_end = builder.x_missing_end.data();

  auto* ctx = dev->ctx();
  dphy_spr_batch* batch = nullptr;
  check(ctx, dphy_spr_study_batch(ctx, dev->forest(), 1, &req, &batch), "dphy_spr_study_batch");
  auto st = dphy_spr_batch_get_summaries(ctx, batch, &summary);
  if (st == DPHY_OK) {
    regions.assign(static_cast<size_t>(summary.num_regions), Candidate_region{});
    auto got = dphy_spr_batch_get_regions(ctx, batch, 0, reinterpret_cast<dphy_candidate_region*>(regions.data()),
                                          summary.num_regions);
    if (got < 0) { st = static_cast<int>(got); }
  }
  auto msg = std::string{st == DPHY_OK ? "" : dphy_last_error(ctx)};
  dphy_spr_batch_destroy(ctx, batch);
  if (st != DPHY_OK) {
    switch (st) {
      case DPHY_ERR_OUT_OF_RANGE: throw std::out_of_range(msg);
      case DPHY_ERR_INVALID_ARGUMENT: throw std::invalid_argument(msg);
      default: throw std::runtime_error(msg);
    }
  }
}

auto Spr_study_builder::regions() -> const std::vector<Candidate_region>& {
  if (not result_valid) {
    // weights need lambda_X; the builder's own result has none (the reference leaves them at 0.0 until Spr_study runs)
    auto summary = dphy_spr_summary{};
    run_spr_study(*this, 1.0, 1.0, t_X, result, summary);
    for (auto& region : result) { region.log_W_over_Wmax = 0.0; region.W_over_Wmax = 0.0; }
    result_valid = true;
  }
  return result;
}

Spr_study::Spr_study(Spr_study_builder&& builder, double lambda_X, double annealing_factor, double t_X, double t_max_tip)
    : tree{builder.tree}, lambda_X{lambda_X}, mu{0.0}, annealing_factor{annealing_factor}, t_X{t_X}, t_max_tip{t_max_tip},
      log_Wmax{0.0}, sum_W_over_Wmax{0.0} {
  auto summary = dphy_spr_summary{};
  run_spr_study(builder, lambda_X, annealing_factor, t_max_tip, candidate_regions, summary);
  mu = summary.mu;
  log_Wmax = summary.log_Wmax;
  sum_W_over_Wmax = summary.sum_W_over_Wmax;
}

auto Spr_study::pick_nexus_region(absl::BitGenRef bitgen) const -> int {
  // one uniform draw, exactly as the reference consumes it; the CDF walk over the downloaded weights
  auto r = absl::Uniform<double>(bitgen, 0.0, sum_W_over_Wmax);
  for (auto i = 0; i != std::ssize(candidate_regions); ++i) {
    const auto& region = candidate_regions[i];
    if (region.W_over_Wmax >= r) { return i; }
    r -= region.W_over_Wmax;
  }
  return 0;
}

auto Spr_study::find_region(Branch_index branch, double t) const -> int {
  for (auto i = 0; i != std::ssize(candidate_regions); ++i) {
    const auto& region = candidate_regions[i];
    if (region.branch == branch && region.t_min < t && t <= region.t_max) { return i; }
  }
  return -1;
}

}  // namespace delphy::b200
