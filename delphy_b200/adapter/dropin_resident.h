// dropin_resident.h -- per-thread device residency behind the link-time drop-in (dropin_phylo_tree_calc.cpp,
// dropin_spr_study.cpp).
//
// The reference's hot functions are stateless (`calc_lambda_i(tree, evo, ref_cum_Q_l)` ...), its driver mutates the Phylo_tree
// in place between calls (core/subrun.cpp:223-231,276-284,316-319; core/spr_move.cpp:838-1156) and offers no hook, so every
// drop-in call ships the tree it is handed: a multi-threaded two-pass flatten straight into page-locked buffers that
// dphy_forest_upload DMAs from (no staging copy), then the device-side flattening.  What IS kept per host thread (the reference
// runs one Subrun per ctpl worker thread, core/run.cpp:682-693) is everything that does not depend on the tree's content: the
// dphy_ctx (stream, arena, memory pool), the page-locked buffers, and the sites table (reference sequence + Global_evo_model
// tables), which is compared with the caller's evo on every call and refreshed through dphy_sites_set_evo only when it changed.
#ifndef DELPHY_B200_DROPIN_RESIDENT_H_
#define DELPHY_B200_DROPIN_RESIDENT_H_

#include <cstdint>
#include <vector>

#include "evo_model.h"
#include "phylo_tree.h"

#include "delphy_b200.h"

namespace delphy::b200 {

// One reusable page-locked array (dphy_host_alloc); grows geometrically, never shrinks.
template <typename T>
struct Pinned_array {
  T* data = nullptr;
  size_t capacity = 0;
  dphy_ctx* ctx = nullptr;
  auto ensure(dphy_ctx* c, size_t n) -> T*;
  auto release() -> void;
};

struct Pinned_flat_emat {
  Pinned_array<int32_t> parent, child0, child1, mut_off, mut_site, miss_off, miss_start, miss_end, fs_off, fs_site;
  Pinned_array<uint8_t> mut_from, mut_to, fs_from;
  Pinned_array<double> t, mut_t;
  int32_t num_nodes = 0, root = -1;
  int64_t num_muts = 0, num_ivls = 0, num_fs = 0;
  auto view(bool includes_run_root) const -> dphy_emat_host;
  auto release() -> void;
};

// Phylo_tree -> SoA + CSR (include/delphy_b200.h, dphy_emat_host), in two passes over the nodes split across host threads:
// list sizes -> offsets (one sequential prefix sum over N), then every thread copies its own node range.
auto flatten_into(dphy_ctx* ctx, const Phylo_tree& tree, Pinned_flat_emat& out) -> void;
// The same for the part of `tree` reachable from its root, renumbered in DFS order: build_usher_like_tree studies a tree whose
// node vector already holds every future tip, most of them not attached yet (core/phylo_tree.cpp:905-932).  to_orig[i] is the
// tree's index of compact node i; of_orig[v] the compact index of node v (-1: not attached).
auto flatten_reachable_into(dphy_ctx* ctx, const Phylo_tree& tree, Pinned_flat_emat& out, std::vector<int32_t>& to_orig,
                            std::vector<int32_t>& of_orig) -> void;

class Resident {
 public:
  static auto get() -> Resident&;            // the calling thread's instance (created on first use; throws without a device)
  ~Resident();

  auto ctx() const -> dphy_ctx* { return ctx_; }

  // Make the sites table match (seq, evo).  evo == nullptr: any model will do (the caller reads only structure).
  auto sync_sites(const Real_sequence& seq, const Global_evo_model* evo) -> dphy_sites*;
  // Ship `tree` (always) against the synced sites table; the previous forest of this thread is dropped.
  auto sync_tree(const Phylo_tree& tree, const Global_evo_model* evo) -> dphy_forest*;
  // Ship only the nodes reachable from the root (see flatten_reachable_into); node indices on the device are compact ones.
  auto sync_reachable_tree(const Phylo_tree& tree) -> dphy_forest*;
  auto to_orig() const -> const std::vector<int32_t>& { return to_orig_; }
  auto of_orig() const -> const std::vector<int32_t>& { return of_orig_; }

  auto num_sites() const -> int { return static_cast<int>(ref_.size()); }
  auto num_partitions() const -> int { return static_cast<int>(mu_.size()); }
  auto num_nodes() const -> int { return flat_.num_nodes; }

  // counters (bench / tests): how many trees were shipped and how long the host-side flatten took in total
  int64_t uploads = 0, uploads_skipped = 0, sites_uploads = 0, set_evos = 0, sites_reused = 0;
  double flatten_seconds = 0.0, upload_seconds = 0.0, sites_seconds = 0.0;

 private:
  Resident();
  auto drop_forest() -> void;
  dphy_ctx* ctx_ = nullptr;
  dphy_sites* sites_ = nullptr;
  dphy_forest* forest_ = nullptr;
  Pinned_flat_emat flat_, prev_;         // the tree being shipped / the tree the resident forest was built from
  bool prev_valid_ = false;
  std::vector<int32_t> to_orig_, of_orig_;
  // host shadow of what the sites table holds
  std::vector<uint8_t> ref_;
  std::vector<int32_t> part_;
  std::vector<double> nu_, mu_, pi_, q_;
};

// Process-wide call counters of the substituted entry points, printed at exit when DPHY_DROPIN_STATS=1.
auto count_call(const char* name) -> void;

// status -> the exception the reference would have thrown (std::out_of_range / std::invalid_argument; CHECK -> runtime_error)
auto throw_on_error(dphy_ctx* ctx, int status, const char* what) -> void;

}  // namespace delphy::b200

#endif  // DELPHY_B200_DROPIN_RESIDENT_H_
