// dropin_resident.cpp -- see dropin_resident.h.
#include "dropin_resident.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>

#include "delphy_b200_adapter.h"   // thread_ctx()

namespace delphy::b200 {

auto throw_on_error(dphy_ctx* ctx, int status, const char* what) -> void {
  if (status == DPHY_OK) { return; }
  auto msg = std::string{what} + ": " + dphy_last_error(ctx);
  switch (status) {
    case DPHY_ERR_OUT_OF_RANGE: throw std::out_of_range(msg);
    case DPHY_ERR_INVALID_ARGUMENT: throw std::invalid_argument(msg);
    default: throw std::runtime_error(msg);
  }
}

namespace {
struct Call_counters {
  std::mutex mu;
  std::map<std::string, long long> counts;
  ~Call_counters() {
    if (const char* e = std::getenv("DPHY_DROPIN_STATS"); e != nullptr && std::atoi(e) != 0) {
      for (const auto& [name, n] : counts) { std::fprintf(stderr, "[delphy_b200 drop-in] calls: %-48s %lld\n", name.c_str(), n); }
    }
  }
};
Call_counters g_calls;
}  // namespace

auto count_call(const char* name) -> void {
  static const bool on = [] { const char* e = std::getenv("DPHY_DROPIN_STATS"); return e != nullptr && std::atoi(e) != 0; }();
  if (!on) { return; }
  auto lock = std::lock_guard<std::mutex>{g_calls.mu};
  ++g_calls.counts[name];
}

template <typename T>
auto Pinned_array<T>::ensure(dphy_ctx* c, size_t n) -> T* {
  if (n + 1 > capacity) {      // + 1: a non-null pointer even for empty lists
    release();
    ctx = c;
    auto want = std::max<size_t>(n + 1, capacity + capacity / 2);
    void* p = nullptr;
    throw_on_error(c, dphy_host_alloc(c, want * sizeof(T), &p), "dphy_host_alloc");
    data = static_cast<T*>(p);
    capacity = want;
  }
  return data;
}

template <typename T>
auto Pinned_array<T>::release() -> void {
  if (data != nullptr) { dphy_host_free(ctx, data); data = nullptr; capacity = 0; }
}

template struct Pinned_array<int32_t>;
template struct Pinned_array<uint8_t>;
template struct Pinned_array<double>;

auto Pinned_flat_emat::view(bool includes_run_root) const -> dphy_emat_host {
  auto e = dphy_emat_host{};
  e.num_nodes = num_nodes;
  e.root = root;
  e.includes_run_root = includes_run_root ? 1 : 0;
  e.parent = parent.data; e.child0 = child0.data; e.child1 = child1.data; e.t = t.data;
  e.mut_off = mut_off.data; e.mut_site = mut_site.data; e.mut_from = mut_from.data; e.mut_to = mut_to.data; e.mut_t = mut_t.data;
  e.miss_off = miss_off.data; e.miss_start = miss_start.data; e.miss_end = miss_end.data;
  e.fs_off = fs_off.data; e.fs_site = fs_site.data; e.fs_from = fs_from.data;
  return e;
}

auto Pinned_flat_emat::release() -> void {
  parent.release(); child0.release(); child1.release(); mut_off.release(); mut_site.release(); miss_off.release();
  miss_start.release(); miss_end.release(); fs_off.release(); fs_site.release(); mut_from.release(); mut_to.release();
  fs_from.release(); t.release(); mut_t.release();
}

namespace {

// A small persistent pool for the flatten passes: starting threads per call costs more than flattening a 20k-node tree.
// One flatten at a time may use it (try_lock); a concurrent caller -- another Subrun's thread -- simply runs its passes alone.
class Flatten_pool {
 public:
  explicit Flatten_pool(int workers) {
    for (auto k = 0; k < workers; ++k) { threads_.emplace_back([this, k] { worker(k); }); }
  }
  ~Flatten_pool() {
    { auto lock = std::unique_lock<std::mutex>{mu_}; stop_ = true; ++epoch_; }
    cv_.notify_all();
    for (auto& th : threads_) { th.join(); }
  }
  auto workers() const -> int { return static_cast<int>(threads_.size()); }
  // fn(part) for part = 1 .. workers() on the pool; part 0 runs on the caller; returns when all are done
  auto run(const std::function<void(int)>& fn) -> void {
    { auto lock = std::unique_lock<std::mutex>{mu_}; fn_ = &fn; pending_ = workers(); ++epoch_; }
    cv_.notify_all();
    fn(0);
    auto lock = std::unique_lock<std::mutex>{mu_};
    done_.wait(lock, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }
  std::mutex user;      // held by the flatten that owns the pool
 private:
  auto worker(int k) -> void {
    auto seen = uint64_t{0};
    for (;;) {
      const std::function<void(int)>* fn = nullptr;
      {
        auto lock = std::unique_lock<std::mutex>{mu_};
        cv_.wait(lock, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (stop_) { return; }
        fn = fn_;
      }
      if (fn != nullptr) { (*fn)(k + 1); }
      { auto lock = std::unique_lock<std::mutex>{mu_}; if (--pending_ == 0) { done_.notify_one(); } }
    }
  }
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  int pending_ = 0;
  uint64_t epoch_ = 0;
  bool stop_ = false;
};

auto flatten_threads() -> int;

auto flatten_pool() -> Flatten_pool& {
  static Flatten_pool pool{std::max(0, flatten_threads() - 1)};
  return pool;
}

// run fn(lo, hi) over [0, n) split into equal ranges over the pool (the calling thread takes the first range)
template <typename F>
auto parallel_ranges(size_t n, int threads, F&& fn) -> void {
  if (threads <= 1 || n < 8192) { fn(size_t{0}, n); return; }
  auto& pool = flatten_pool();
  auto own = std::unique_lock<std::mutex>{pool.user, std::try_to_lock};
  if (!own.owns_lock() || pool.workers() == 0) { fn(size_t{0}, n); return; }
  const auto parts = static_cast<size_t>(pool.workers() + 1);
  const auto chunk = (n + parts - 1) / parts;
  const auto body = std::function<void(int)>{[&](int part) {
    const auto lo = std::min(n, static_cast<size_t>(part) * chunk), hi = std::min(n, lo + chunk);
    if (lo < hi) { fn(lo, hi); }
  }};
  pool.run(body);
}

auto flatten_threads() -> int {
  static const int n = [] {
    const char* e = std::getenv("DPHY_FLATTEN_THREADS");
    if (e != nullptr) { return std::max(1, std::atoi(e)); }
    return static_cast<int>(std::clamp(std::thread::hardware_concurrency() / 2, 1u, 8u));
  }();
  return n;
}

}  // namespace

auto flatten_into(dphy_ctx* ctx, const Phylo_tree& tree, Pinned_flat_emat& out) -> void {
  const auto n = static_cast<size_t>(std::ssize(tree));
  out.num_nodes = static_cast<int32_t>(n);
  out.root = tree.root;
  auto* parent = out.parent.ensure(ctx, n); auto* child0 = out.child0.ensure(ctx, n); auto* child1 = out.child1.ensure(ctx, n);
  auto* t = out.t.ensure(ctx, n);
  auto* mut_off = out.mut_off.ensure(ctx, n + 1); auto* miss_off = out.miss_off.ensure(ctx, n + 1); auto* fs_off = out.fs_off.ensure(ctx, n + 1);
  const auto threads = flatten_threads();

  // pass 1: node scalars + list sizes (written into the offset arrays, shifted by one)
  mut_off[0] = 0; miss_off[0] = 0; fs_off[0] = 0;
  parallel_ranges(n, threads, [&](size_t lo, size_t hi) {
    for (auto v = lo; v != hi; ++v) {
      const auto& node = tree.nodes[v];
      parent[v] = node.parent;
      if (node.is_tip()) { child0[v] = -1; child1[v] = -1; }
      else { child0[v] = node.children[0]; child1[v] = node.children[1]; }
      t[v] = node.t;
      mut_off[v + 1] = static_cast<int32_t>(node.mutations.size());
      miss_off[v + 1] = static_cast<int32_t>(node.missations.intervals.num_intervals());
      fs_off[v + 1] = static_cast<int32_t>(node.missations.from_states.size());
    }
  });
  for (auto v = size_t{0}; v != n; ++v) { mut_off[v + 1] += mut_off[v]; miss_off[v + 1] += miss_off[v]; fs_off[v + 1] += fs_off[v]; }
  out.num_muts = mut_off[n]; out.num_ivls = miss_off[n]; out.num_fs = fs_off[n];

  auto* mut_site = out.mut_site.ensure(ctx, out.num_muts); auto* mut_from = out.mut_from.ensure(ctx, out.num_muts);
  auto* mut_to = out.mut_to.ensure(ctx, out.num_muts); auto* mut_t = out.mut_t.ensure(ctx, out.num_muts);
  auto* miss_start = out.miss_start.ensure(ctx, out.num_ivls); auto* miss_end = out.miss_end.ensure(ctx, out.num_ivls);
  auto* fs_site = out.fs_site.ensure(ctx, out.num_fs); auto* fs_from = out.fs_from.ensure(ctx, out.num_fs);

  // pass 2: every thread copies the lists of its own node range to their final places
  parallel_ranges(n, threads, [&](size_t lo, size_t hi) {
    for (auto v = lo; v != hi; ++v) {
      const auto& node = tree.nodes[v];
      auto m = static_cast<size_t>(mut_off[v]);
      for (const auto& mut : node.mutations) {
        mut_site[m] = mut.site; mut_from[m] = static_cast<uint8_t>(mut.from); mut_to[m] = static_cast<uint8_t>(mut.to); mut_t[m] = mut.t;
        ++m;
      }
      auto i = static_cast<size_t>(miss_off[v]);
      for (const auto& [start, end] : node.missations.intervals) { miss_start[i] = start; miss_end[i] = end; ++i; }
      auto f = static_cast<size_t>(fs_off[v]);
      for (const auto& [site, from] : node.missations.from_states) { fs_site[f] = site; fs_from[f] = static_cast<uint8_t>(from); ++f; }
    }
  });
}

auto flatten_reachable_into(dphy_ctx* ctx, const Phylo_tree& tree, Pinned_flat_emat& out, std::vector<int32_t>& to_orig,
                            std::vector<int32_t>& of_orig) -> void {
  const auto n_all = static_cast<size_t>(std::ssize(tree));
  to_orig.clear();
  of_orig.assign(n_all, -1);
  if (tree.root != k_no_node) {
    auto stack = std::vector<Node_index>{tree.root};
    while (not stack.empty()) {
      const auto v = stack.back();
      stack.pop_back();
      of_orig[v] = static_cast<int32_t>(to_orig.size());
      to_orig.push_back(v);
      if (not tree.at(v).is_tip()) { stack.push_back(tree.at(v).children[1]); stack.push_back(tree.at(v).children[0]); }
    }
  }
  const auto n = to_orig.size();
  out.num_nodes = static_cast<int32_t>(n);
  out.root = n != 0 ? 0 : -1;
  auto* parent = out.parent.ensure(ctx, n); auto* child0 = out.child0.ensure(ctx, n); auto* child1 = out.child1.ensure(ctx, n);
  auto* t = out.t.ensure(ctx, n);
  auto* mut_off = out.mut_off.ensure(ctx, n + 1); auto* miss_off = out.miss_off.ensure(ctx, n + 1); auto* fs_off = out.fs_off.ensure(ctx, n + 1);
  mut_off[0] = 0; miss_off[0] = 0; fs_off[0] = 0;
  for (auto i = size_t{0}; i != n; ++i) {
    const auto& node = tree.nodes[to_orig[i]];
    mut_off[i + 1] = mut_off[i] + static_cast<int32_t>(node.mutations.size());
    miss_off[i + 1] = miss_off[i] + static_cast<int32_t>(node.missations.intervals.num_intervals());
    fs_off[i + 1] = fs_off[i] + static_cast<int32_t>(node.missations.from_states.size());
  }
  out.num_muts = mut_off[n]; out.num_ivls = miss_off[n]; out.num_fs = fs_off[n];
  auto* mut_site = out.mut_site.ensure(ctx, out.num_muts); auto* mut_from = out.mut_from.ensure(ctx, out.num_muts);
  auto* mut_to = out.mut_to.ensure(ctx, out.num_muts); auto* mut_t = out.mut_t.ensure(ctx, out.num_muts);
  auto* miss_start = out.miss_start.ensure(ctx, out.num_ivls); auto* miss_end = out.miss_end.ensure(ctx, out.num_ivls);
  auto* fs_site = out.fs_site.ensure(ctx, out.num_fs); auto* fs_from = out.fs_from.ensure(ctx, out.num_fs);
  for (auto i = size_t{0}; i != n; ++i) {
    const auto& node = tree.nodes[to_orig[i]];
    parent[i] = static_cast<size_t>(i) == 0 ? -1 : of_orig[node.parent];
    if (node.is_tip()) { child0[i] = -1; child1[i] = -1; }
    else { child0[i] = of_orig[node.children[0]]; child1[i] = of_orig[node.children[1]]; }
    t[i] = node.t;
    auto m = static_cast<size_t>(mut_off[i]);
    for (const auto& mut : node.mutations) {
      mut_site[m] = mut.site; mut_from[m] = static_cast<uint8_t>(mut.from); mut_to[m] = static_cast<uint8_t>(mut.to); mut_t[m] = mut.t;
      ++m;
    }
    auto iv = static_cast<size_t>(miss_off[i]);
    for (const auto& [start, end] : node.missations.intervals) { miss_start[iv] = start; miss_end[iv] = end; ++iv; }
    auto f = static_cast<size_t>(fs_off[i]);
    for (const auto& [site, from] : node.missations.from_states) { fs_site[f] = site; fs_from[f] = static_cast<uint8_t>(from); ++f; }
  }
}

namespace {
template <typename T>
auto same_array(const Pinned_array<T>& a, const Pinned_array<T>& b, size_t n) -> bool {
  return n == 0 || std::memcmp(a.data, b.data, n * sizeof(T)) == 0;
}
auto same_flat(const Pinned_flat_emat& a, const Pinned_flat_emat& b) -> bool {
  if (a.num_nodes != b.num_nodes || a.root != b.root || a.num_muts != b.num_muts || a.num_ivls != b.num_ivls || a.num_fs != b.num_fs) { return false; }
  const auto n = static_cast<size_t>(a.num_nodes), m = static_cast<size_t>(a.num_muts), iv = static_cast<size_t>(a.num_ivls), fs = static_cast<size_t>(a.num_fs);
  return same_array(a.t, b.t, n) && same_array(a.parent, b.parent, n) && same_array(a.child0, b.child0, n) && same_array(a.child1, b.child1, n) &&
         same_array(a.mut_off, b.mut_off, n + 1) && same_array(a.mut_t, b.mut_t, m) && same_array(a.mut_site, b.mut_site, m) &&
         same_array(a.mut_from, b.mut_from, m) && same_array(a.mut_to, b.mut_to, m) &&
         same_array(a.miss_off, b.miss_off, n + 1) && same_array(a.miss_start, b.miss_start, iv) && same_array(a.miss_end, b.miss_end, iv) &&
         same_array(a.fs_off, b.fs_off, n + 1) && same_array(a.fs_site, b.fs_site, fs) && same_array(a.fs_from, b.fs_from, fs);
}
}  // namespace

// ---- Resident ------------------------------------------------------------------------------------------------------------------------------------
Resident::Resident() : ctx_{thread_ctx()} {}

Resident::~Resident() {
  // thread_local: runs before the adapter's ctx reaper (constructed earlier, by thread_ctx() above)
  if (const char* e = std::getenv("DPHY_DROPIN_STATS"); e != nullptr && std::atoi(e) != 0) {
    std::fprintf(stderr, "[delphy_b200 drop-in] thread stats: %lld tree uploads (flatten %.3f s, upload+device flatten %.3f s), "
                 "sites: %lld uploads, %lld set_evo, %lld reused (%.3f s); %lld calls answered by the resident forest\n", (long long)uploads, flatten_seconds, upload_seconds,
                 (long long)sites_uploads, (long long)set_evos, (long long)sites_reused, sites_seconds, (long long)uploads_skipped);
  }
  drop_forest();
  if (sites_ != nullptr) { dphy_sites_destroy(ctx_, sites_); sites_ = nullptr; }
  flat_.release(); prev_.release();
}

auto Resident::get() -> Resident& {
  thread_local Resident instance;
  return instance;
}

auto Resident::drop_forest() -> void {
  if (forest_ != nullptr) { dphy_forest_destroy(ctx_, forest_); forest_ = nullptr; }
}

auto Resident::sync_sites(const Real_sequence& seq, const Global_evo_model* evo) -> dphy_sites* {
  struct Timer {
    double& acc; std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    ~Timer() { acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
  } timer{sites_seconds};
  const auto L = static_cast<size_t>(std::ssize(seq));
  static_assert(sizeof(Real_seq_letter) == 1);
  const auto* seq_bytes = reinterpret_cast<const uint8_t*>(seq.data());
  const auto same_seq = sites_ != nullptr && ref_.size() == L && (L == 0 || std::memcmp(ref_.data(), seq_bytes, L) == 0);
  if (evo == nullptr && same_seq) { ++sites_reused; return sites_; }

  // the model the table should hold
  auto P = size_t{1};
  auto mu = std::vector<double>{}, pi = std::vector<double>{}, q = std::vector<double>{};
  const int32_t* part = nullptr;
  const double* nu = nullptr;
  auto neutral_part = std::vector<int32_t>{};
  auto neutral_nu = std::vector<double>{};
  if (evo != nullptr) {
    if (evo->partition_for_site.size() != L || evo->nu_l.size() != L) {
      throw std::invalid_argument("delphy_b200: evo model and reference sequence disagree on the number of sites");
    }
    P = static_cast<size_t>(evo->num_partitions());
    mu.resize(P); pi.resize(P * 4); q.resize(P * 16);
    for (auto p = size_t{0}; p != P; ++p) {
      const auto& model = evo->partition_evo_model[p];
      mu[p] = model.mu;
      for (auto a = 0; a != 4; ++a) {
        pi[p * 4 + a] = model.pi_a[static_cast<Real_seq_letter>(a)];
        for (auto b = 0; b != 4; ++b) { q[p * 16 + a * 4 + b] = model.q_ab[static_cast<Real_seq_letter>(a)][static_cast<Real_seq_letter>(b)]; }
      }
    }
    static_assert(sizeof(Partition_index) == sizeof(int32_t));
    part = reinterpret_cast<const int32_t*>(evo->partition_for_site.data());
    nu = evo->nu_l.data();
  } else {
    // no model given and nothing usable resident: a neutral one (the caller reads only structure: counts, missing sites, regions)
    mu = {1.0}; pi = {0.25, 0.25, 0.25, 0.25};
    q.assign(16, 1.0 / 3.0);
    for (auto a = 0; a != 4; ++a) { q[a * 4 + a] = -1.0; }
    neutral_part.assign(L, 0); neutral_nu.assign(L, 1.0);
    part = neutral_part.data(); nu = neutral_nu.data();
  }

  const auto same_structure = same_seq && mu_.size() == P && (L == 0 || std::memcmp(part_.data(), part, L * sizeof(int32_t)) == 0);
  if (same_structure) {
    const auto same_nu = L == 0 || std::memcmp(nu_.data(), nu, L * sizeof(double)) == 0;
    const auto same_model = mu_ == mu && pi_ == pi && q_ == q;
    if (same_nu && same_model) { ++sites_reused; return sites_; }
    ++set_evos;
    // Subrun::set_evo (core/subrun.h:29-30): same sequence, new parameters
    throw_on_error(ctx_, dphy_sites_set_evo(ctx_, sites_, same_nu ? nullptr : nu, mu.data(), pi.data(), q.data()), "dphy_sites_set_evo");
    if (!same_nu) { nu_.assign(nu, nu + L); }
    mu_ = std::move(mu); pi_ = std::move(pi); q_ = std::move(q);
    return sites_;
  }

  ++sites_uploads;
  drop_forest();     // a forest refers to its sites table (and its folded weights to the reference sequence)
  const auto same_shape = sites_ != nullptr && ref_.size() == L && mu_.size() == P;
  ref_.assign(seq_bytes, seq_bytes + L);
  part_.assign(part, part + L);
  nu_.assign(nu, nu + L);
  mu_ = std::move(mu); pi_ = std::move(pi); q_ = std::move(q);
  auto hs = dphy_sites_host{};
  hs.num_sites = static_cast<int32_t>(L); hs.num_partitions = static_cast<int32_t>(P);
  hs.ref = ref_.data(); hs.partition_for_site = part_.data(); hs.nu_l = nu_.data();
  hs.mu = mu_.data(); hs.pi_a = pi_.data(); hs.q_ab = q_.data();
  if (same_shape) {
    // Run::normalize_root re-references the sequence every cycle (core/run.cpp:258-265): same table, new contents
    throw_on_error(ctx_, dphy_sites_update(ctx_, sites_, &hs), "dphy_sites_update");
    return sites_;
  }
  if (sites_ != nullptr) { dphy_sites_destroy(ctx_, sites_); sites_ = nullptr; }
  throw_on_error(ctx_, dphy_sites_upload(ctx_, &hs, &sites_), "dphy_sites_upload");
  return sites_;
}

auto Resident::sync_tree(const Phylo_tree& tree, const Global_evo_model* evo) -> dphy_forest* {
  using clock = std::chrono::steady_clock;
  sync_sites(tree.ref_sequence, evo);
  // (flat_ is never a DMA source: uploads go out of prev_, which is only rewritten by the swap below after a synchronize)
  const auto t0 = clock::now();
  flatten_into(ctx_, tree, flat_);
  // The driver recomputes in bursts on an unchanged tree (Run::recalc_derived_quantities makes five hot calls in a row, and every
  // pop-model proposal of a global move revalidates everything, core/run.cpp:302-314,760-779): if the flattened arrays equal the
  // ones the resident forest was built from, that forest -- already evaluated -- answers the call as it is.
  if (prev_valid_ && forest_ != nullptr && same_flat(flat_, prev_)) {
    ++uploads_skipped;
    flatten_seconds += std::chrono::duration<double>(clock::now() - t0).count();
    return forest_;
  }
  const auto t1 = clock::now();
  throw_on_error(ctx_, dphy_ctx_synchronize(ctx_), "dphy_ctx_synchronize");    // the previous upload has left prev_
  drop_forest();
  std::swap(flat_, prev_);        // prev_ now holds the arrays being uploaded (DMA'd in place: they stay untouched until the next upload)
  prev_valid_ = true;
  auto he = prev_.view(true);     // both pieces of log G are computed; callers pick (Subrun::calc_cur_log_G, core/subrun.cpp:58-68)
  auto zero = int32_t{0};
  throw_on_error(ctx_, dphy_forest_upload(ctx_, 1, &he, &zero, 1, &sites_, &forest_), "dphy_forest_upload");
  const auto t2 = clock::now();
  ++uploads;
  flatten_seconds += std::chrono::duration<double>(t1 - t0).count();
  upload_seconds += std::chrono::duration<double>(t2 - t1).count();
  return forest_;
}

auto Resident::sync_reachable_tree(const Phylo_tree& tree) -> dphy_forest* {
  using clock = std::chrono::steady_clock;
  sync_sites(tree.ref_sequence, nullptr);
  throw_on_error(ctx_, dphy_ctx_synchronize(ctx_), "dphy_ctx_synchronize");
  const auto t0 = clock::now();
  flatten_reachable_into(ctx_, tree, flat_, to_orig_, of_orig_);
  const auto t1 = clock::now();
  drop_forest();
  std::swap(flat_, prev_);
  prev_valid_ = false;            // compact node indices: never reused for a whole-tree call
  auto he = prev_.view(true);
  auto zero = int32_t{0};
  throw_on_error(ctx_, dphy_forest_upload(ctx_, 1, &he, &zero, 1, &sites_, &forest_), "dphy_forest_upload");
  const auto t2 = clock::now();
  ++uploads;
  flatten_seconds += std::chrono::duration<double>(t1 - t0).count();
  upload_seconds += std::chrono::duration<double>(t2 - t1).count();
  return forest_;
}

}  // namespace delphy::b200
