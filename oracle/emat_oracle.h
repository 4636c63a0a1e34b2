/* oracle/emat_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference's EMAT log-G / tallies / SPR-study path, operating on
 * the same flat host arrays the product C-ABI (include/delphy_b200.h) accepts.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the product path
 * (delphy_b200/csrc) never links, imports or calls it.
 *
 * Parity pin: every function below is checked (tests/test_oracle_*.py) against
 *   (1) the reference's own known-answer fixtures, transcribed into tests/golden/ (file:line cited there), and
 *   (2) the reference's own sources compiled in place into oracle/_ref/libdelphy_ref.so (see Makefile).
 * Exception: Q(a,x) values inside the above-root SPR region -- see gamma_q.h ("parity unpinned").
 *
 * All citations are relative to /root/reference.
 */
#ifndef DPHY_EMAT_ORACLE_H_
#define DPHY_EMAT_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One EMAT (= one delphy::Phylo_tree, core/phylo_tree.h:14-63) flattened to SoA + CSR, host node order. */
typedef struct orc_emat {
  int32_t num_nodes;
  int32_t root;
  int32_t includes_run_root;   /* Subrun::includes_run_root_ (core/subrun.cpp:62-64) */
  int32_t reserved;
  const int32_t* parent;       /* [N]  -1 == k_no_node */
  const int32_t* child0;       /* [N]  -1 for tips (Binary_node::children[0]) */
  const int32_t* child1;       /* [N] */
  const double*  t;            /* [N]  Phylo_node::t */
  const int32_t* mut_off;      /* [N+1] CSR into the mutation arrays; the root's list is included */
  const int32_t* mut_site;     /* Mutation::site  (core/mutations.h:21-29) */
  const uint8_t* mut_from;     /* Real_seq_letter A=0,C=1,G=2,T=3 (core/sequence.h:155) */
  const uint8_t* mut_to;
  const double*  mut_t;
  const int32_t* miss_off;     /* [N+1] CSR into Missation_map::intervals */
  const int32_t* miss_start;
  const int32_t* miss_end;
  const int32_t* fs_off;       /* [N+1] CSR into Missation_map::from_states (sorted by site) */
  const int32_t* fs_site;
  const uint8_t* fs_from;
} orc_emat;

/* Reference sequence + Global_evo_model (core/evo_model.h:12-48). */
typedef struct orc_sites {
  int32_t num_sites;
  int32_t num_partitions;
  const uint8_t* ref;                  /* [L] */
  const int32_t* partition_for_site;   /* [L] */
  const double*  nu_l;                 /* [L] */
  const double*  mu;                   /* [P] */
  const double*  pi_a;                 /* [P][4] */
  const double*  q_ab;                 /* [P][4][4], q_a(a) = -q_ab[a][a] */
} orc_sites;

/* Same layout as delphy::Candidate_region (core/spr_study.h:17-32): 48 bytes. */
typedef struct orc_region {
  int32_t branch;
  int32_t mut_idx;
  double  t_min;
  double  t_max;
  int32_t min_muts;
  int32_t pad_;
  double  log_W_over_Wmax;
  double  W_over_Wmax;
} orc_region;

typedef struct orc_study_summary {
  double mu;                 /* Spr_study::mu (core/spr_study.cpp:239) */
  double log_Wmax;
  double sum_W_over_Wmax;
  int32_t num_regions;
  int32_t num_missing_at_X;
} orc_study_summary;

/* ---- phylo_tree_calc ---------------------------------------------------------------------------- */
void   orc_state_frequencies_per_partition(const orc_sites* s, int32_t* out /*[P*4]*/);       /* phylo_tree_calc.cpp:95-106 */
void   orc_cum_Q_l(const orc_sites* s, double* out /*[L+1]*/);                                /* :379-388 */
double orc_lambda_for_sequence(const orc_sites* s);                                           /* :390-399 */
void   orc_lambda_i(const orc_emat* e, const orc_sites* s, const double* cum_Q_l, double* out /*[N]*/);  /* :420-436 */
double orc_log_root_prior(const orc_emat* e, const orc_sites* s, const int32_t* ref_freqs /*[P*4]*/);   /* :467-504 */
double orc_log_G_below_root(const orc_emat* e, const orc_sites* s, const double* lambda_i);  /* :515-543 */
double orc_branch_log_G(const orc_emat* e, const orc_sites* s, int32_t X, double lambda_X);  /* phylo_tree_calc.h:185-206 */
double orc_path_log_G(const orc_emat* e, const orc_sites* s, int32_t A, int32_t B, const double* lambda_i,
                      const int32_t* ref_freqs);                                              /* :560-575 */
void   orc_num_sites_missing_at_every_node(const orc_emat* e, int32_t* out /*[N]*/);          /* :67-76 */
int32_t orc_num_muts(const orc_emat* e);                                                      /* :577-585 */
void   orc_num_muts_ab(const orc_emat* e, int32_t* out /*[16]*/);                             /* :587-597 */
void   orc_num_muts_beta_ab(const orc_emat* e, const orc_sites* s, int32_t* out /*[P*16]*/);  /* :599-610 */
void   orc_num_muts_l(const orc_emat* e, int32_t L, int32_t* out /*[L]*/);                    /* :612-622 */
void   orc_num_muts_l_ab(const orc_emat* e, int32_t L, int32_t* out /*[L*16]*/);              /* :624-634 */
double orc_T(const orc_emat* e);                                                              /* :120-128 */
void   orc_T_l_a(const orc_emat* e, const orc_sites* s, double* out /*[L*4]*/);               /* :130-174 */
void   orc_Ttwiddle_l(const orc_emat* e, const orc_sites* s, double* out /*[L]*/);            /* :176-222 */
void   orc_Ttwiddle_beta_a(const orc_emat* e, const orc_sites* s, double* out /*[P*4]*/);     /* :288-369 */
/* reconstruct_missing_sites_at (:41-56): returns #intervals written (merged, sorted); cap = capacity */
int32_t orc_missing_sites_at(const orc_emat* e, int32_t node, int32_t* starts, int32_t* ends, int32_t cap);
/* calc_site_state_at (:108-118) at the END of branch `node` */
uint8_t orc_site_state_at_node(const orc_emat* e, const orc_sites* s, int32_t node, int32_t site);

/* ---- spr_study ---------------------------------------------------------------------------------- */
/* Spr_study_builder::seed_fill_from (core/spr_study.cpp:9-224).  X may be -1 (k_no_node).
 * init deltas: n_init triples (site, from, to) = deltas from the start region to X.
 * Returns the number of regions (emission order == the reference's DFS order), or -1 if cap is too small. */
int32_t orc_spr_study_build(const orc_emat* e, int32_t num_sites,
                            int32_t X, double t_X,
                            const int32_t* missing_starts, const int32_t* missing_ends, int32_t n_missing,
                            int32_t start_branch, int32_t start_mut_idx,
                            const int32_t* init_site, const uint8_t* init_from, const uint8_t* init_to, int32_t n_init,
                            int32_t max_muts_from_start, int32_t can_change_root,
                            orc_region* out, int32_t cap);
/* Spr_study::Spr_study (core/spr_study.cpp:226-385): fills log_W_over_Wmax / W_over_Wmax in place. */
void orc_spr_study_weights(const orc_emat* e, int32_t num_sites, int32_t num_missing_at_X,
                           orc_region* regions, int32_t n_regions,
                           double lambda_X, double annealing_factor, double t_X, double t_max_tip,
                           orc_study_summary* out);
/* Spr_study::pick_nexus_region (:404-422) with the uniform draw r in [0, sum_W_over_Wmax) supplied by the caller */
int32_t orc_spr_pick_nexus_region(const orc_region* regions, int32_t n_regions, double r);
int32_t orc_spr_find_region(const orc_region* regions, int32_t n_regions, int32_t branch, double t);   /* :474-484 */
double  orc_spr_log_alpha_in_region(const orc_emat* e, const orc_region* regions, int32_t n_regions,
                                    int32_t region_idx, double t,
                                    double lambda_X, double annealing_factor, double t_X, double t_max_tip,
                                    double sum_W_over_Wmax);                                            /* :486-549 */
/* Helper mirroring the call pattern of Subrun::spr1_move (core/subrun.cpp:540-553) WITHOUT peel_graft:
 * start region = (sibling of X, 0), init deltas = net mutations on branch P->X at sites not missing at X. */
int32_t orc_spr_study_from_attached(const orc_emat* e, const orc_sites* s, int32_t X,
                                    int32_t max_muts_from_start, int32_t can_change_root,
                                    double annealing_factor, double t_max_tip, const double* lambda_i,
                                    orc_region* out, int32_t cap, orc_study_summary* summary);

/* ---- the FlatBuffers wire format of a tree: delphy.api.Tree (core/api.fbs:13-49) ------------------------------------------------
 * A view of the four vectors of a size-prefixed Tree buffer (as phylo_tree_to_api_tree writes it, core/api.cpp:34-98).  Records are
 * FlatBuffers structs, little endian (core/api_generated.h:157-265): Node 16 B {parent i32, left_child i32, right_child i32, t f32},
 * Mutation 16 B {branch i32, site i32, from u8, to u8, 2 B padding, t f32}, MissationInterval 12 B {branch, start_site, end_site}. */
typedef struct orc_api_tree_view {
  int32_t num_nodes, root, num_sites, pad_;
  int64_t num_muts, num_ivls;
  const uint8_t* nodes;
  const uint8_t* muts;
  const uint8_t* ivls;
  const uint8_t* ref_seq;
} orc_api_tree_view;
/* flatbuffers::GetSizePrefixedRoot<api::Tree> + the field accessors (core/api_generated.h:267-304), with every offset bounds-checked.
 * 0, or -1 if the buffer is malformed. */
int32_t orc_api_tree_parse(const uint8_t* buf, int64_t len, orc_api_tree_view* out);
/* api_tree_and_tree_info_to_phylo_tree (core/api.cpp:127-186) for a tree in normal form -- one that fix_up_missations
 * (core/phylo_tree.cpp:379-478) changes only by reconstructing the from_states.  counts[5] = {num_nodes, root, M, I, F} is always
 * filled; the arrays (sized from a first call with parent == NULL) are filled when parent != NULL.  Returns 0; -2 if the tree is not in
 * normal form (common missations not factored up, a site missing twice along a root path, a mutation on a missing site), -3 if a
 * mutation's `from` contradicts the sequence above it (the reference CHECKs), -1 for out-of-range indices. */
int32_t orc_api_tree_to_emat(const orc_api_tree_view* v, int32_t* counts, int32_t* parent, int32_t* child0, int32_t* child1, double* t,
                             int32_t* mut_off, int32_t* mut_site, uint8_t* mut_from, uint8_t* mut_to, double* mut_t,
                             int32_t* miss_off, int32_t* miss_start, int32_t* miss_end,
                             int32_t* fs_off, int32_t* fs_site, uint8_t* fs_from);
/* phylo_tree_to_api_tree (core/api.cpp:34-98): same content, a layout of our own (a FlatBuffers reader follows offsets, it does not
 * care where the vectors lie).  Returns the length, or -1 if cap is too small. */
int64_t orc_api_tree_write(const orc_emat* e, const uint8_t* ref, int32_t num_sites, uint8_t* out, int64_t cap);

double orc_gamma_q_export(double a, double x);

#ifdef __cplusplus
}
#endif
#endif /* DPHY_EMAT_ORACLE_H_ */
