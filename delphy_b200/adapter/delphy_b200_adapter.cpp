// delphy_b200_adapter.cpp -- Delphy's C++ signatures implemented over the C ABI of libdelphy_b200.so.
// See delphy_b200_adapter.h.  Nothing here computes: it flattens, calls include/delphy_b200.h, and re-shapes results.
#include "delphy_b200_adapter.h"

#include <stdexcept>
#include <string>

#include "absl/random/distributions.h"

namespace delphy::b200 {

namespace {

thread_local dphy_ctx* tl_ctx = nullptr;
thread_local int tl_device = 0;

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
    auto msg = std::string{dphy_last_error(ctx_)};
    dphy_sites_destroy(ctx_, sites_); sites_ = nullptr;
    ctx_->~dphy_ctx == nullptr ? void() : void();   // (no-op: keep ctx alive for the thread)
    check(ctx_, st, msg.c_str());
  }
}
