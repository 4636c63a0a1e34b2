// dropin_spr_study.cpp -- LINK-TIME drop-in for core/spr_study.cpp.
//
// Linked into the reference's unmodified driver INSTEAD of the reference's spr_study.o.  It defines delphy::Spr_study_builder's
// and delphy::Spr_study's out-of-line members against the reference's own header (core/spr_study.h), so the object layouts and
// every call site (core/subrun.cpp:543-603, core/phylo_tree.cpp:918-963) stay as they are.
//
// What runs where:
//   * seed_fill_from with max_muts_from_start == INT_MAX -- the FULL-tree studies: 1 % of spr1_moves (core/subrun.cpp:495-499)
//     and every study of build_usher_like_tree -- is one device pass (dphy_spr_study_batch in its "builder inputs as given"
//     mode: start-relative deltas + missing_at_X, valid on a tree mid-move), and the Spr_study constructor that follows is the
//     device's weight pass over the still-resident regions (dphy_spr_batch_set_weights).  On a 100k-tip tree such a study is
//     ~350k regions, tens of milliseconds of hash-map work on the host.
//   * BOUNDED studies (max_muts_from_start == 1, the other 99 %: a ball of a few dozen regions, microseconds on the host) are NOT
//     shipped: the driver mutates the tree in place before every study and offers no hook, so using the device means shipping
//     O(N) bytes to save O(ball) work.  They run through the reference's own builder, compiled unmodified from
//     core/spr_study.cpp under the class names Ref_spr_study_builder / Ref_spr_study (see the Makefile: -DSpr_study_builder=...).
//     DPHY_DROPIN_BOUNDED_ON_DEVICE=1 sends them to the device as well (used by the parity tests).
//   * DPHY_DROPIN_VERIFY=1 (tests): every device study is re-run by the reference's builder / constructor on the same inputs and
//     compared region by region -- order, branch, mut_idx, t_min, t_max, min_muts bit-exact, weights to 1e-9 -- aborting on the
//     first difference.  With BOUNDED_ON_DEVICE this checks every study an MCMC run makes, on trees mid-move.
//   * pick_nexus_region, pick_time_in_region, find_region, log_alpha_in_region consume the host RNG / read one region
//     ("RNG on host"): they forward to the reference's own code on the layout-identical object.
//     The device versions (dphy_spr_batch_pick_nexus_regions, _find_region, _log_alpha_in_region) serve batched callers.
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <type_traits>

#include "spr_study.h"

// the reference's own implementation, same header under other class names (objects are layout-identical by construction)
#define Spr_study_builder Ref_spr_study_builder
#define Spr_study Ref_spr_study
#define Candidate_region Ref_candidate_region
#undef DELPHY_SPR_STUDY_H_
#include "spr_study.h"
#undef Spr_study_builder
#undef Spr_study
#undef Candidate_region

#include "dropin_resident.h"

namespace delphy {

static_assert(sizeof(Spr_study_builder) == sizeof(Ref_spr_study_builder) && sizeof(Spr_study) == sizeof(Ref_spr_study) &&
              sizeof(Candidate_region) == sizeof(Ref_candidate_region) && sizeof(Candidate_region) == sizeof(dphy_candidate_region),
              "the renamed reference classes must mirror core/spr_study.h");
static_assert(std::is_trivially_copyable_v<Candidate_region>);

namespace {

using b200::Resident;
using b200::throw_on_error;

auto as_ref(Spr_study_builder& b) -> Ref_spr_study_builder& { return reinterpret_cast<Ref_spr_study_builder&>(b); }
auto as_ref(const Spr_study& s) -> const Ref_spr_study& { return reinterpret_cast<const Ref_spr_study&>(s); }

auto bounded_on_device() -> bool {
  static const bool v = [] { const char* e = std::getenv("DPHY_DROPIN_BOUNDED_ON_DEVICE"); return e != nullptr && std::atoi(e) != 0; }();
  return v;
}

auto verify_enabled() -> bool {
  static const bool v = [] { const char* e = std::getenv("DPHY_DROPIN_VERIFY"); return e != nullptr && std::atoi(e) != 0; }();
  return v;
}

auto close_rel(double a, double b, double rel, double abs_tol) -> bool {
  if (a == b) { return true; }
  return std::abs(a - b) <= abs_tol + rel * std::max(std::abs(a), std::abs(b));
}

// The regions of the calling thread's last device study stay on the device until the Spr_study constructor (or the next study).
struct Pending_study {
  const Spr_study_builder* builder = nullptr;
  dphy_ctx* ctx = nullptr;
  dphy_spr_batch* batch = nullptr;
  auto drop() -> void {
    if (batch != nullptr) { dphy_spr_batch_destroy(ctx, batch); batch = nullptr; }
    builder = nullptr;
  }
  ~Pending_study() { drop(); }
};

auto pending() -> Pending_study& {
  Resident::get();                       // construct the ctx owner first so that it is destroyed after this object
  thread_local Pending_study p;
  return p;
}

}  // namespace

// ---- Spr_study_builder --------------------------------------------------------------------------------------------------------------------------
auto Spr_study_builder::seed_fill_from(Branch_index init_branch, int init_mut_idx, Site_deltas init_to_X_deltas, bool can_change_root)
    -> void {
  CHECK(work_stack.empty());
  CHECK_EQ(cur_branch, k_no_node);

  if (max_muts_from_start != std::numeric_limits<int>::max() && not bounded_on_device()) {
    pending().drop();                      // a stale record must never match this builder's address
    b200::count_call("Spr_study_builder::seed_fill_from (bounded, host)");
    as_ref(*this).seed_fill_from(init_branch, init_mut_idx, std::move(init_to_X_deltas), can_change_root);
    return;
  }

  auto& p = pending();
  p.drop();
  b200::count_call(max_muts_from_start == std::numeric_limits<int>::max() ? "Spr_study_builder::seed_fill_from (full, device)"
                                                                           : "Spr_study_builder::seed_fill_from (bounded, device)");
  auto& r = Resident::get();
  auto* ctx = r.ctx();
  // a sequence that is not in the tree yet (build_usher_like_tree): the node vector also holds the tips still to be attached,
  // so only the part hanging from the root is shipped, under compact node indices
  const auto compact = X == k_no_node;
  auto* forest = compact ? r.sync_reachable_tree(*tree) : r.sync_tree(*tree, nullptr);
  auto verify_deltas = Site_deltas{};
  if (verify_enabled()) { verify_deltas = init_to_X_deltas; }

  auto d_site = std::vector<int32_t>{}; auto d_to = std::vector<uint8_t>{};
  d_site.reserve(init_to_X_deltas.size()); d_to.reserve(init_to_X_deltas.size());
  for (const auto& [l, delta] : init_to_X_deltas) { d_site.push_back(l); d_to.push_back(static_cast<uint8_t>(delta.to)); }
  auto m_start = std::vector<int32_t>{}, m_end = std::vector<int32_t>{};
  for (const auto& [start, end] : *missing_at_X) { m_start.push_back(start); m_end.push_back(end); }

  auto req = dphy_spr_request{};
  req.tree = 0;
  req.X = X;
  req.t_X = t_X;
  req.start_branch = compact ? r.of_orig().at(init_branch) : init_branch;
  req.start_mut_idx = init_mut_idx;
  req.init_min_muts = static_cast<int32_t>(std::ssize(init_to_X_deltas));
  req.max_muts_from_start = max_muts_from_start;
  req.can_change_root = can_change_root ? 1 : 0;
  req.x_state_mode = DPHY_SPR_X_REL_START;
  req.lambda_X = 0.0;                       // enumerate only: the weights are the Spr_study constructor's business
  req.annealing_factor = 1.0;
  req.t_max_tip = t_X;
  req.n_x_deltas = static_cast<int32_t>(d_site.size()); req.x_delta_site = d_site.data(); req.x_delta_to = d_to.data();
  req.n_x_missing = static_cast<int32_t>(m_start.size()); req.x_missing_start = m_start.data(); req.x_missing_end = m_end.data();

  throw_on_error(ctx, dphy_spr_study_batch(ctx, forest, 1, &req, &p.batch), "dphy_spr_study_batch");
  p.ctx = ctx;
  auto summary = dphy_spr_summary{};
  auto st = dphy_spr_batch_get_summaries(ctx, p.batch, &summary);
  if (st == DPHY_OK) {
    result.resize(static_cast<size_t>(summary.num_regions));
    const auto got = dphy_spr_batch_get_regions(ctx, p.batch, 0, reinterpret_cast<dphy_candidate_region*>(result.data()), summary.num_regions);
    if (got < 0) { st = static_cast<int>(got); }
  }
  if (st != DPHY_OK) {
    auto msg = std::string{dphy_last_error(ctx)};
    p.drop();
    if (st == DPHY_ERR_OUT_OF_RANGE) { throw std::out_of_range(msg); }
    if (st == DPHY_ERR_INVALID_ARGUMENT) { throw std::invalid_argument(msg); }
    throw std::runtime_error(msg);
  }
  if (compact) { for (auto& region : result) { region.branch = r.to_orig().at(region.branch); } }
  p.builder = this;
  cur_to_X_deltas = std::move(init_to_X_deltas);

  if (verify_enabled()) {
    auto ref = Ref_spr_study_builder{*tree, X, t_X, *missing_at_X};
    ref.max_muts_from_start = max_muts_from_start;
    ref.seed_fill_from(init_branch, init_mut_idx, std::move(verify_deltas), can_change_root);
    CHECK_EQ(std::ssize(ref.result), std::ssize(result)) << "device study: number of regions (X=" << X << ")";
    for (auto i = 0; i != std::ssize(result); ++i) {
      const auto& a = result[i]; const auto& b = ref.result[i];
      CHECK(a.branch == b.branch && a.mut_idx == b.mut_idx && a.t_min == b.t_min && a.t_max == b.t_max && a.min_muts == b.min_muts)
          << "device study differs at region " << i << " of " << std::ssize(result) << " (X=" << X << ", limit=" << max_muts_from_start
          << "): got " << a << " want {branch=" << b.branch << ", mut_idx=" << b.mut_idx << ", " << b.t_min << "<t<=" << b.t_max
          << ", min_muts=" << b.min_muts << "}";
    }
  }
}

// The builder's step-wise interface (core/spr_study.h:123-166) has no caller outside core/spr_study.cpp; it keeps working on the
// host through the reference's own code.
auto Spr_study_builder::do_pending_work() -> void { as_ref(*this).do_pending_work(); }
auto Spr_study_builder::move_to_neighbor(Branch_index target_branch, int target_mut_idx, bool is_backtracking) -> void {
  as_ref(*this).move_to_neighbor(target_branch, target_mut_idx, is_backtracking);
}
auto Spr_study_builder::visit_cur_region() -> void { as_ref(*this).visit_cur_region(); }
auto Spr_study_builder::seed_neighbors_except(Branch_index old_branch, int old_mut_idx) -> void {
  as_ref(*this).seed_neighbors_except(old_branch, old_mut_idx);
}
auto Spr_study_builder::account_for_Xs_detachment(bool can_change_root) -> void { as_ref(*this).account_for_Xs_detachment(can_change_root); }
auto Spr_study_builder::remove_regions_in_Xs_future() -> void { as_ref(*this).remove_regions_in_Xs_future(); }

// ---- Spr_study -----------------------------------------------------------------------------------------------------------------------------------------
Spr_study::Spr_study(Spr_study_builder&& builder, double lambda_X, double annealing_factor, double t_X, double t_max_tip)
    : tree{builder.tree}, lambda_X{lambda_X}, annealing_factor{annealing_factor}, t_X{t_X}, t_max_tip{t_max_tip},
      candidate_regions{} {
  auto& p = pending();
  if (p.builder != &builder || p.batch == nullptr) {
    // regions enumerated on the host (a bounded study): the reference's own constructor, then adopt its fields
    auto ref = Ref_spr_study{std::move(as_ref(builder)), lambda_X, annealing_factor, t_X, t_max_tip};
    mu = ref.mu;
    candidate_regions.resize(ref.candidate_regions.size());
    if (not ref.candidate_regions.empty()) {
      std::memcpy(static_cast<void*>(candidate_regions.data()), ref.candidate_regions.data(), ref.candidate_regions.size() * sizeof(Candidate_region));
    }
    log_Wmax = ref.log_Wmax;
    sum_W_over_Wmax = ref.sum_W_over_Wmax;
    return;
  }
  candidate_regions = std::move(builder.result);
  mu = lambda_X / (tree->num_sites() - builder.missing_at_X->num_sites());                        // core/spr_study.cpp:239
  CHECK(not candidate_regions.empty());
  auto* ctx = p.ctx;
  auto wp = dphy_spr_weight_params{lambda_X, annealing_factor, t_max_tip};
  auto st = dphy_spr_batch_set_weights(ctx, p.batch, &wp);
  auto summary = dphy_spr_summary{};
  if (st == DPHY_OK) { st = dphy_spr_batch_get_summaries(ctx, p.batch, &summary); }
  if (st == DPHY_OK) {
    const auto got = dphy_spr_batch_get_region_weights(ctx, p.batch, 0, reinterpret_cast<dphy_candidate_region*>(candidate_regions.data()),
                                                       static_cast<int64_t>(candidate_regions.size()));
    if (got < 0) { st = static_cast<int>(got); }
  }
  auto msg = std::string{st == DPHY_OK ? "" : dphy_last_error(ctx)};
  p.drop();
  if (st != DPHY_OK) { throw std::runtime_error(msg); }
  log_Wmax = summary.log_Wmax;
  sum_W_over_Wmax = summary.sum_W_over_Wmax;

  if (verify_enabled()) {
    auto rb = Ref_spr_study_builder{*tree, builder.X, builder.t_X, *builder.missing_at_X};
    rb.result.resize(candidate_regions.size());
    std::memcpy(static_cast<void*>(rb.result.data()), candidate_regions.data(), candidate_regions.size() * sizeof(Candidate_region));
    for (auto& region : rb.result) { region.log_W_over_Wmax = 0.0; region.W_over_Wmax = 0.0; }
    auto ref = Ref_spr_study{std::move(rb), lambda_X, annealing_factor, t_X, t_max_tip};
    CHECK(close_rel(mu, ref.mu, 1e-12, 0.0)) << mu << " != " << ref.mu;
    CHECK(close_rel(log_Wmax, ref.log_Wmax, 1e-9, 1e-9)) << log_Wmax << " != " << ref.log_Wmax;
    CHECK(close_rel(sum_W_over_Wmax, ref.sum_W_over_Wmax, 1e-9, 0.0)) << sum_W_over_Wmax << " != " << ref.sum_W_over_Wmax;
    for (auto i = 0; i != std::ssize(candidate_regions); ++i) {
      const auto& a = candidate_regions[i]; const auto& b = ref.candidate_regions[i];
      CHECK(close_rel(a.log_W_over_Wmax, b.log_W_over_Wmax, 1e-9, 1e-9) && close_rel(a.W_over_Wmax, b.W_over_Wmax, 1e-9, 1e-300))
          << "device weights differ at region " << i << ": " << a.log_W_over_Wmax << " / " << a.W_over_Wmax << " vs "
          << b.log_W_over_Wmax << " / " << b.W_over_Wmax;
    }
  }
}

auto Spr_study::dump() -> void { const_cast<Ref_spr_study&>(as_ref(*this)).dump(); }
auto Spr_study::pick_nexus_region(absl::BitGenRef bitgen) const -> int { return as_ref(*this).pick_nexus_region(bitgen); }
auto Spr_study::pick_time_in_region(int region_idx, absl::BitGenRef bitgen) const -> double {
  return as_ref(*this).pick_time_in_region(region_idx, bitgen);
}
auto Spr_study::find_region(Branch_index branch, double t) const -> int { return as_ref(*this).find_region(branch, t); }
auto Spr_study::log_alpha_in_region(int region_idx, double t) const -> double { return as_ref(*this).log_alpha_in_region(region_idx, t); }

}  // namespace delphy
