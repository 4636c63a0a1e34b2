// delphy_b200_adapter.h -- the reference-side binding: Delphy's own C++ signatures over the B200 C ABI.
//
// This header is compiled INSIDE the reference's source tree (it includes core/phylo_tree.h etc.); it is the file a
// Delphy maintainer adds next to core/phylo_tree_calc.h.  Every function below has the name, argument meaning, return
// type and error behaviour of the reference function it replaces (cited per declaration, paths relative to the
// reference checkout), lives in namespace delphy::b200, and is implemented in delphy_b200_adapter.cpp purely in terms
// of include/delphy_b200.h (the C ABI of libdelphy_b200.so).  There is no CPU fallback: if no CUDA device is usable
// the first call throws std::runtime_error.
//
// Two levels:
//   * stateless free functions (drop-in: `using namespace delphy::b200;` or the one-line call-site edits listed in
//     INTEGRATION.md).  Each call flattens the Phylo_tree, uploads it, evaluates, downloads.
//   * Device_emat: a resident device copy of one (tree, evo) pair for callers that evaluate repeatedly between tree
//     edits (Subrun::calc_cur_log_G, Run::recalc_derived_quantities, the SPR study of every move).
//
// Error mapping (include/delphy_b200.h, dphy_status): DPHY_ERR_OUT_OF_RANGE -> std::out_of_range,
// DPHY_ERR_INVALID_ARGUMENT -> std::invalid_argument, everything else -> std::runtime_error (the reference CHECK-aborts
// there; an exception is the closest recoverable equivalent).
#ifndef DELPHY_B200_ADAPTER_H_
#define DELPHY_B200_ADAPTER_H_

#include <memory>
#include <vector>

#include "absl/random/bit_gen_ref.h"

#include "evo_model.h"
#include "phylo_tree.h"
#include "phylo_tree_calc.h"   // only for the shared vocabulary types (Node_vector, Seq_vector, ...)
#include "spr_study.h"         // delphy::Candidate_region (the 48-byte record the device writes verbatim)

#include "delphy_b200.h"

namespace delphy::b200 {

// ---- flattening: delphy::Phylo_tree / Global_evo_model -> the SoA + CSR arrays of dphy_emat_host / dphy_sites_host ----
struct Flat_emat {
  std::vector<int32_t> parent, child0, child1, mut_off, mut_site, miss_off, miss_start, miss_end, fs_off, fs_site;
  std::vector<uint8_t> mut_from, mut_to, fs_from;
  std::vector<double> t, mut_t;
  int32_t root = -1;
  auto view(bool includes_run_root = true) const -> dphy_emat_host;
};
struct Flat_sites {
  std::vector<uint8_t> ref;
  std::vector<int32_t> partition_for_site;
  std::vector<double> nu_l, mu, pi_a, q_ab;
  auto view() const -> dphy_sites_host;
};
auto flatten(const Phylo_tree& tree) -> Flat_emat;
auto flatten(const Real_sequence& ref_sequence, const Global_evo_model& evo) -> Flat_sites;

// One CUDA device + stream + arena per host thread (the reference runs one Subrun per ctpl worker, core/run.cpp:682-693,
// with a thread_local arena, core/scratch_space.cpp:9-10).  Throws std::runtime_error when no device is usable.
auto thread_ctx() -> dphy_ctx*;
auto set_thread_device(int device) -> void;   // call before the first use on this thread (default: device 0)

// ---- resident device copy of one (tree, evo) pair -----------------------------------------------------------------------
class Device_emat {
 public:
  Device_emat(const Phylo_tree& tree, const Global_evo_model& evo, bool includes_run_root = true);
  ~Device_emat();
  Device_emat(const Device_emat&) = delete;
  auto operator=(const Device_emat&) -> Device_emat& = delete;

  // Subrun::set_evo / Run::set_mu etc.: new evo parameters, same tree
  auto set_evo(const Global_evo_model& evo) -> void;
  // accepted inner_node_displace / tip-date moves (core/subrun.cpp:223-231,276-284)
  auto set_node_times(const std::vector<Node_index>& nodes, const std::vector<double>& t) -> void;

  auto calc_lambda_i() -> Node_vector<double>;                            // core/phylo_tree_calc.cpp:420-436
  auto calc_num_sites_missing_at_every_node() -> Node_vector<int>;        // :67-76
  auto calc_log_root_prior() -> double;                                   // :467-504
  auto calc_log_G_below_root() -> double;                                 // :515-543
  auto calc_cur_log_G() -> double;                                        // Subrun::calc_cur_log_G core/subrun.cpp:58-68
  auto calc_num_muts() -> int;                                            // :577-585
  auto calc_num_muts_ab() -> Seq_matrix<int>;                             // :587-597
  auto calc_num_muts_beta_ab() -> Partition_vector<Seq_matrix<int>>;      // :599-610
  auto calc_num_muts_l() -> Node_vector<int>;                             // :612-622
  auto calc_num_muts_l_ab() -> Node_vector<Seq_matrix<int>>;              // :624-634
  auto calc_T() -> double;                                                // :120-128
  auto calc_T_l_a() -> std::vector<Seq_vector<double>>;                   // :130-174
  auto calc_Ttwiddle_l() -> std::vector<double>;                          // :176-222
  auto calc_Ttwiddle_beta_a() -> Partition_vector<Seq_vector<double>>;    // :288-369
  auto calc_state_frequencies_per_partition() -> Partition_vector<Seq_vector<int>>;   // :95-106
  auto calc_cum_Q_l() -> std::vector<double>;                             // :379-388

  auto ctx() const -> dphy_ctx* { return ctx_; }
  auto forest() const -> dphy_forest* { return forest_; }
  auto num_sites() const -> int { return num_sites_; }
  auto num_partitions() const -> int { return num_partitions_; }
  auto num_nodes() const -> int { return num_nodes_; }

 private:
  dphy_ctx* ctx_ = nullptr;
  dphy_sites* sites_ = nullptr;
  dphy_forest* forest_ = nullptr;
  int num_sites_ = 0, num_partitions_ = 0, num_nodes_ = 0;
  Real_sequence ref_sequence_;
};

// ---- stateless drop-ins: the signatures of core/phylo_tree_calc.h:37-223 -----------------------------------------------------
auto count_mutations(const Phylo_tree& tree) -> int;                                                     // .cpp:9-17
auto calc_num_sites_missing_at_every_node(const Phylo_tree& tree) -> Node_vector<int>;                   // .cpp:67-76
auto calc_state_frequencies_per_partition_of(const Real_sequence& seq, const Global_evo_model& evo)
    -> Partition_vector<Seq_vector<int>>;                                                                // .cpp:95-106
auto calc_T(const Phylo_tree& tree) -> double;                                                           // .cpp:120-128
auto calc_T_l_a(const Phylo_tree& tree) -> std::vector<Seq_vector<double>>;                              // .cpp:130-174
auto calc_Ttwiddle_l(const Phylo_tree& tree, const Global_evo_model& evo) -> std::vector<double>;        // .cpp:176-222
auto calc_Ttwiddle_beta_a(const Phylo_tree& tree, const Global_evo_model& evo)
    -> Partition_vector<Seq_vector<double>>;                                                             // .cpp:288-369
auto calc_cum_Q_l_for_sequence(const Real_sequence& seq, const Global_evo_model& evo) -> std::vector<double>;   // .cpp:379-388
auto calc_lambda_for_sequence(const Real_sequence& seq, const Global_evo_model& evo) -> double;          // .cpp:390-399
auto calc_lambda_i(const Phylo_tree& tree, const Global_evo_model& evo, const std::vector<double>& ref_cum_Q_l)
    -> Node_vector<double>;                                                                              // .cpp:420-436
auto calc_log_root_prior(const Phylo_tree& tree, const Global_evo_model& evo) -> double;                 // .cpp:458-465
auto calc_log_root_prior(const Phylo_tree& tree, const Global_evo_model& evo,
                         const Partition_vector<Seq_vector<int>>& state_frequencies_of_ref_sequence_per_partition)
    -> double;                                                                                           // .cpp:467-504
auto calc_log_G_below_root(const Phylo_tree& tree, const Global_evo_model& evo) -> double;               // .cpp:506-513
auto calc_log_G_below_root(const Phylo_tree& tree, const Global_evo_model& evo, const Node_vector<double>& lambda_i,
                           const Partition_vector<Seq_vector<int>>& state_frequencies_of_ref_sequence_per_partition)
    -> double;                                                                                           // .cpp:515-543
auto calc_num_muts(const Phylo_tree& tree) -> int;                                                       // .cpp:577-585
auto calc_num_muts_ab(const Phylo_tree& tree) -> Seq_matrix<int>;                                        // .cpp:587-597
auto calc_num_muts_beta_ab(const Phylo_tree& tree, const Global_evo_model& evo)
    -> Partition_vector<Seq_matrix<int>>;                                                                // .cpp:599-610
auto calc_num_muts_l(const Phylo_tree& tree) -> Node_vector<int>;                                        // .cpp:612-622
auto calc_num_muts_l_ab(const Phylo_tree& tree) -> Node_vector<Seq_matrix<int>>;                         // .cpp:624-634

// ---- SPR regraft study: the shapes of core/spr_study.h:69-205 ---------------------------------------------------------------
// Usage is the reference's (core/subrun.cpp:548-553):
//     auto builder = b200::Spr_study_builder{tree, X, t_X, missing_at_X};
//     builder.max_muts_from_start = limit;
//     builder.seed_fill_from(S, 0, std::move(deltas_P_to_X), includes_run_root);
//     auto study = b200::Spr_study{std::move(builder), lambda_X, annealing_factor, t_X, t_max_tip};
// The enumeration and the weights are ONE device pass, so seed_fill_from only records the request; the pass runs
// in the Spr_study constructor (or when somebody reads builder.regions()).
struct Spr_study_builder {
  const Phylo_tree* tree;
  const Global_evo_model* evo = nullptr;      // optional: only the site count / partition table are read; nullptr => 1 partition
  Device_emat* resident = nullptr;            // optional: reuse a resident copy that mirrors *tree (else one-shot upload)
  const Scratch_interval_set* missing_at_X;   // kept for signature parity; the device reconstructs it from the EMAT
  Node_index X = k_no_node;
  double t_X = std::numeric_limits<double>::max();
  int max_muts_from_start = std::numeric_limits<int>::max();

  Spr_study_builder(const Phylo_tree& tree, Node_index X, double t_X, const Scratch_interval_set& missing_at_X)
      : tree{&tree}, missing_at_X{&missing_at_X}, X{X}, t_X{t_X} {}

  auto seed_fill_from(Branch_index cur_branch, int cur_mut_idx, Site_deltas cur_to_X_deltas, bool can_change_root) -> void;

  // Spr_study_builder::result of the reference (regions without weights), materialised on demand
  auto regions() -> const std::vector<Candidate_region>&;

  // the recorded request
  Branch_index start_branch = k_no_node;
  int start_mut_idx = 0;
  int init_min_muts = 0;
  bool can_change_root = true;
  bool seeded = false;
  std::vector<int32_t> x_delta_site; std::vector<uint8_t> x_delta_to;            // X == k_no_node mode only
  std::vector<int32_t> x_missing_start, x_missing_end;
  std::vector<Candidate_region> result;
  bool result_valid = false;
};

struct Spr_study {
  Spr_study(Spr_study_builder&& builder, double lambda_X, double annealing_factor, double t_X, double t_max_tip);

  const Phylo_tree* tree;
  double lambda_X;
  double mu;
  double annealing_factor;
  double t_X;
  double t_max_tip;
  std::vector<Candidate_region> candidate_regions;

  double log_Wmax;
  double sum_W_over_Wmax;

  auto pick_nexus_region(absl::BitGenRef bitgen) const -> int;            // core/spr_study.cpp:404-422
  auto find_region(Branch_index branch, double t) const -> int;           // core/spr_study.cpp:474-484
};

// The shared engine of both: runs one study on the device and returns regions (+weights) and the summary.
auto run_spr_study(const Spr_study_builder& builder, double lambda_X, double annealing_factor, double t_max_tip,
                   std::vector<Candidate_region>& regions, dphy_spr_summary& summary) -> void;

}  // namespace delphy::b200

#endif  // DELPHY_B200_ADAPTER_H_
