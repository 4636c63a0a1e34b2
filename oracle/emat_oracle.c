/* oracle/emat_oracle.c -- TEST INFRASTRUCTURE ONLY (see emat_oracle.h).
 *
 * Plain-C, single-threaded restatement of the reference's algorithms for the EMAT log-G path and
 * the SPR regraft study, on flat arrays.  Floating-point expressions keep the reference's order of
 * operations (build with -ffp-contract=off) so that results are bit-identical to oracle/_ref
 * wherever the algorithm is deterministic.  Citations are relative to /root/reference.
 */
#include "emat_oracle.h"
#include "gamma_q.h"

#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

/* ---- evo-model accessors (core/evo_model.h:33-48) ---------------------------------------------- */
static inline double mu_l(const orc_sites* s, int l) { return s->mu[s->partition_for_site[l]]; }
static inline double q_l_a(const orc_sites* s, int l, int a) {
  return -s->q_ab[s->partition_for_site[l] * 16 + a * 4 + a];
}
static inline double q_l_ab(const orc_sites* s, int l, int a, int b) {
  return s->q_ab[s->partition_for_site[l] * 16 + a * 4 + b];
}

/* ---- traversal helpers (core/tree.h:243-318) ---------------------------------------------------- */
/* pre-order with children visited in order child0, child1 */
static int32_t* make_pre_order(const orc_emat* e) {
  int32_t n = e->num_nodes;
  int32_t* order = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  int32_t* stack = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  int32_t sp = 0, k = 0;
  if (n > 0) stack[sp++] = e->root;
  while (sp > 0) {
    int32_t v = stack[--sp];
    order[k++] = v;
    if (e->child0[v] >= 0) {       /* push child1 first so child0 is popped first */
      stack[sp++] = e->child1[v];
      stack[sp++] = e->child0[v];
    }
  }
  free(stack);
  return order;
}

/* ---- phylo_tree_calc.cpp:95-106 ----------------------------------------------------------------- */
void orc_state_frequencies_per_partition(const orc_sites* s, int32_t* out) {
  memset(out, 0, sizeof(int32_t) * 4u * (size_t)s->num_partitions);
  for (int l = 0; l != s->num_sites; ++l) {
    ++out[s->partition_for_site[l] * 4 + s->ref[l]];
  }
}

/* ---- phylo_tree_calc.cpp:379-388 ---------------------------------------------------------------- */
void orc_cum_Q_l(const orc_sites* s, double* out) {
  double so_far = 0.0;
  out[0] = 0.0;
  for (int l = 0; l != s->num_sites; ++l) {
    so_far += mu_l(s, l) * s->nu_l[l] * q_l_a(s, l, s->ref[l]);
    out[l + 1] = so_far;
  }
}

/* ---- phylo_tree_calc.cpp:390-399 ---------------------------------------------------------------- */
double orc_lambda_for_sequence(const orc_sites* s) {
  double lambda = 0.0;
  for (int l = 0; l != s->num_sites; ++l) {
    lambda += mu_l(s, l) * s->nu_l[l] * q_l_a(s, l, s->ref[l]);
  }
  return lambda;
}

/* ---- phylo_tree_calc.h:121-155 ------------------------------------------------------------------ */
static double delta_lambda_across_missations(const orc_emat* e, const orc_sites* s, const double* cumQ, int v) {
  double result = 0.0;
  for (int i = e->miss_off[v]; i != e->miss_off[v + 1]; ++i) {
    result -= cumQ[e->miss_end[i]] - cumQ[e->miss_start[i]];
  }
  for (int i = e->fs_off[v]; i != e->fs_off[v + 1]; ++i) {
    int l = e->fs_site[i];
    int ref_from = s->ref[l];
    result -= mu_l(s, l) * s->nu_l[l] * (q_l_a(s, l, e->fs_from[i]) - q_l_a(s, l, ref_from));
  }
  return result;
}
static double delta_lambda_across_branch(const orc_emat* e, const orc_sites* s, const double* cumQ, int v) {
  double result = 0.0;
  for (int i = e->mut_off[v]; i != e->mut_off[v + 1]; ++i) {
    int l = e->mut_site[i];
    result += mu_l(s, l) * s->nu_l[l] * (q_l_a(s, l, e->mut_to[i]) - q_l_a(s, l, e->mut_from[i]));
  }
  result += delta_lambda_across_missations(e, s, cumQ, v);
  return result;
}

/* ---- phylo_tree_calc.cpp:420-436 ---------------------------------------------------------------- */
void orc_lambda_i(const orc_emat* e, const orc_sites* s, const double* cumQ, double* out) {
  double lambda_ref = cumQ[s->num_sites];
  int32_t* order = make_pre_order(e);
  for (int k = 0; k != e->num_nodes; ++k) {
    int v = order[k];
    double lambda_parent = (v == e->root) ? lambda_ref : out[e->parent[v]];
    out[v] = lambda_parent + delta_lambda_across_branch(e, s, cumQ, v);
  }
  free(order);
}

/* ---- phylo_tree_calc.cpp:467-504 ---------------------------------------------------------------- */
double orc_log_root_prior(const orc_emat* e, const orc_sites* s, const int32_t* ref_freqs) {
  int P = s->num_partitions;
  int32_t* f = (int32_t*)malloc(sizeof(int32_t) * 4u * (size_t)P);
  memcpy(f, ref_freqs, sizeof(int32_t) * 4u * (size_t)P);
  int r = e->root;
  for (int i = e->mut_off[r]; i != e->mut_off[r + 1]; ++i) {
    int p = s->partition_for_site[e->mut_site[i]];
    --f[p * 4 + e->mut_from[i]];
    ++f[p * 4 + e->mut_to[i]];
  }
  for (int i = e->miss_off[r]; i != e->miss_off[r + 1]; ++i) {
    for (int l = e->miss_start[i]; l != e->miss_end[i]; ++l) {
      --f[s->partition_for_site[l] * 4 + s->ref[l]];
    }
  }
  for (int i = e->fs_off[r]; i != e->fs_off[r + 1]; ++i) {
    int l = e->fs_site[i];
    int p = s->partition_for_site[l];
    ++f[p * 4 + s->ref[l]];
    --f[p * 4 + e->fs_from[i]];
  }
  double result = 0.0;
  for (int p = 0; p != P; ++p) {
    for (int a = 0; a != 4; ++a) {
      double pi = s->pi_a[p * 4 + a];
      if (pi != 0.0) {
        result += f[p * 4 + a] * log(pi);
      } else if (f[p * 4 + a] != 0) {
        free(f);
        return -INFINITY;
      }
    }
  }
  free(f);
  return result;
}

/* ---- phylo_tree_calc.h:185-206 ------------------------------------------------------------------ */
double orc_branch_log_G(const orc_emat* e, const orc_sites* s, int32_t X, double lambda_X) {
  double t_P = e->t[e->parent[X]];
  double t_X = e->t[X];
  double result = -lambda_X * (t_X - t_P);
  for (int i = e->mut_off[X + 1] - 1; i >= e->mut_off[X]; --i) {   /* reverse order */
    int l = e->mut_site[i];
    int from = e->mut_from[i], to = e->mut_to[i];
    result -= mu_l(s, l) * s->nu_l[l] * (q_l_a(s, l, from) - q_l_a(s, l, to)) * (e->mut_t[i] - t_P);
    result += log(mu_l(s, l) * s->nu_l[l] * q_l_ab(s, l, from, to));
  }
  return result;
}

/* ---- phylo_tree_calc.cpp:515-543 ---------------------------------------------------------------- */
double orc_log_G_below_root(const orc_emat* e, const orc_sites* s, const double* lambda_i) {
  double result = 0.0;
  for (int v = 0; v != e->num_nodes; ++v) {
    if (v != e->root) {
      result += orc_branch_log_G(e, s, v, lambda_i[v]);
    }
  }
  return result;
}

/* ---- phylo_tree_calc.cpp:545-575 ---------------------------------------------------------------- */
double orc_path_log_G(const orc_emat* e, const orc_sites* s, int32_t A, int32_t B, const double* lambda_i,
                      const int32_t* ref_freqs) {
  double result = 0.0;
  for (int cur = B; cur != A; cur = e->parent[cur]) {
    if (cur == e->root) result += orc_log_root_prior(e, s, ref_freqs);
    else result += orc_branch_log_G(e, s, cur, lambda_i[cur]);
  }
  return result;
}

/* ---- phylo_tree_calc.cpp:67-76 ------------------------------------------------------------------ */
void orc_num_sites_missing_at_every_node(const orc_emat* e, int32_t* out) {
  int32_t* order = make_pre_order(e);
  for (int k = 0; k != e->num_nodes; ++k) {
    int v = order[k];
    int at_parent = (v == e->root) ? 0 : out[e->parent[v]];
    int here = 0;
    for (int i = e->miss_off[v]; i != e->miss_off[v + 1]; ++i) here += e->miss_end[i] - e->miss_start[i];
    out[v] = at_parent + here;
  }
  free(order);
}

/* ---- phylo_tree_calc.cpp:577-634 ---------------------------------------------------------------- */
int32_t orc_num_muts(const orc_emat* e) {
  int32_t n = 0;
  for (int v = 0; v != e->num_nodes; ++v) if (v != e->root) n += e->mut_off[v + 1] - e->mut_off[v];
  return n;
}
void orc_num_muts_ab(const orc_emat* e, int32_t* out) {
  memset(out, 0, sizeof(int32_t) * 16);
  for (int v = 0; v != e->num_nodes; ++v) if (v != e->root)
    for (int i = e->mut_off[v]; i != e->mut_off[v + 1]; ++i) ++out[e->mut_from[i] * 4 + e->mut_to[i]];
}
void orc_num_muts_beta_ab(const orc_emat* e, const orc_sites* s, int32_t* out) {
  memset(out, 0, sizeof(int32_t) * 16u * (size_t)s->num_partitions);
  for (int v = 0; v != e->num_nodes; ++v) if (v != e->root)
    for (int i = e->mut_off[v]; i != e->mut_off[v + 1]; ++i)
      ++out[s->partition_for_site[e->mut_site[i]] * 16 + e->mut_from[i] * 4 + e->mut_to[i]];
}
void orc_num_muts_l(const orc_emat* e, int32_t L, int32_t* out) {
  memset(out, 0, sizeof(int32_t) * (size_t)L);
  for (int v = 0; v != e->num_nodes; ++v) if (v != e->root)
    for (int i = e->mut_off[v]; i != e->mut_off[v + 1]; ++i) ++out[e->mut_site[i]];
}
void orc_num_muts_l_ab(const orc_emat* e, int32_t L, int32_t* out) {
  memset(out, 0, sizeof(int32_t) * 16u * (size_t)L);
  for (int v = 0; v != e->num_nodes; ++v) if (v != e->root)
    for (int i = e->mut_off[v]; i != e->mut_off[v + 1]; ++i)
      ++out[(size_t)e->mut_site[i] * 16 + e->mut_from[i] * 4 + e->mut_to[i]];
}

/* ---- phylo_tree_calc.cpp:120-128 ---------------------------------------------------------------- */
double orc_T(const orc_emat* e) {
  double T = 0.0;
  for (int v = 0; v != e->num_nodes; ++v) if (v != e->root) T += e->t[v] - e->t[e->parent[v]];
  return T;
}

/* total branch length below every node, post-order (phylo_tree_calc.cpp:131-141, :180-189) */
static double* make_T_below_node(const orc_emat* e) {
  int n = e->num_nodes;
  double* Tb = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  int32_t* order = make_pre_order(e);
  for (int k = n - 1; k >= 0; --k) {        /* reverse pre-order visits children before parents */
    int v = order[k];
    double T = 0.0;
    if (e->child0[v] >= 0) {
      int c0 = e->child0[v], c1 = e->child1[v];
      T += (e->t[c0] - e->t[v]) + Tb[c0];
      T += (e->t[c1] - e->t[v]) + Tb[c1];
    }
    Tb[v] = T;
  }
  free(order);
  return Tb;
}

/* ---- phylo_tree_calc.cpp:130-174 ---------------------------------------------------------------- */
void orc_T_l_a(const orc_emat* e, const orc_sites* s, double* out) {
  int L = s->num_sites;
  double* Tb = make_T_below_node(e);
  double T = Tb[e->root];
  for (int l = 0; l != L; ++l) {
    for (int a = 0; a != 4; ++a) out[l * 4 + a] = 0.0;
    out[l * 4 + s->ref[l]] = T;
  }
  for (int v = 0; v != e->num_nodes; ++v) {
    for (int i = e->mut_off[v]; i != e->mut_off[v + 1]; ++i) {
      double T_below_mut = Tb[v] + (v == e->root ? 0.0 : e->t[v] - e->mut_t[i]);
      out[e->mut_site[i] * 4 + e->mut_from[i]] -= T_below_mut;
      out[e->mut_site[i] * 4 + e->mut_to[i]] += T_below_mut;
    }
    double T_below_miss = Tb[v] + (v == e->root ? 0.0 : e->t[v] - e->t[e->parent[v]]);
    for (int i = e->miss_off[v]; i != e->miss_off[v + 1]; ++i)
      for (int l = e->miss_start[i]; l != e->miss_end[i]; ++l) out[l * 4 + s->ref[l]] -= T_below_miss;
    for (int i = e->fs_off[v]; i != e->fs_off[v + 1]; ++i) {
      int l = e->fs_site[i];
      out[l * 4 + s->ref[l]] += T_below_miss;
      out[l * 4 + e->fs_from[i]] -= T_below_miss;
    }
  }
  free(Tb);
}

/* ---- phylo_tree_calc.cpp:176-222 ---------------------------------------------------------------- */
void orc_Ttwiddle_l(const orc_emat* e, const orc_sites* s, double* out) {
  int L = s->num_sites;
  double* Tb = make_T_below_node(e);
  double T = Tb[e->root];
  for (int l = 0; l != L; ++l) out[l] = q_l_a(s, l, s->ref[l]) * T;
  for (int v = 0; v != e->num_nodes; ++v) {
    for (int i = e->mut_off[v]; i != e->mut_off[v + 1]; ++i) {
      int l = e->mut_site[i];
      double T_below_mut = Tb[v] + (v == e->root ? 0.0 : e->t[v] - e->mut_t[i]);
      out[l] -= q_l_a(s, l, e->mut_from[i]) * T_below_mut;
      out[l] += q_l_a(s, l, e->mut_to[i]) * T_below_mut;
    }
    double T_below_miss = Tb[v] + (v == e->root ? 0.0 : e->t[v] - e->t[e->parent[v]]);
    for (int i = e->miss_off[v]; i != e->miss_off[v + 1]; ++i)
      for (int l = e->miss_start[i]; l != e->miss_end[i]; ++l) out[l] -= q_l_a(s, l, s->ref[l]) * T_below_miss;
    for (int i = e->fs_off[v]; i != e->fs_off[v + 1]; ++i) {
      int l = e->fs_site[i];
      out[l] += q_l_a(s, l, s->ref[l]) * T_below_miss;
      out[l] -= q_l_a(s, l, e->fs_from[i]) * T_below_miss;
    }
  }
  free(Tb);
}

/* ---- phylo_tree_calc.cpp:288-369 ---------------------------------------------------------------- */
void orc_Ttwiddle_beta_a(const orc_emat* e, const orc_sites* s, double* out) {
  int P = s->num_partitions, L = s->num_sites, n = e->num_nodes;
  double* nt = (double*)calloc(4u * (size_t)P, sizeof(double));   /* ntwiddle_beta_a */
  for (int i = 0; i != 4 * P; ++i) out[i] = 0.0;
  for (int l = 0; l != L; ++l) nt[s->partition_for_site[l] * 4 + s->ref[l]] += s->nu_l[l];

  /* generic DFS "traversal" (core/tree.h:243-268): (node, children_so_far) visitations */
  typedef struct { int32_t node; int32_t phase; } item;   /* phase: 0 = enter, 1 = exit */
  item* stack = (item*)malloc(sizeof(item) * (size_t)(2 * n + 2));
  int sp = 0;
  if (n > 0) { stack[sp].node = e->root; stack[sp].phase = 0; ++sp; }
  while (sp > 0) {
    item it = stack[--sp];
    int v = it.node;
    if (it.phase == 0) {
      int parent = e->parent[v];
      for (int i = e->miss_off[v]; i != e->miss_off[v + 1]; ++i)
        for (int l = e->miss_start[i]; l != e->miss_end[i]; ++l)
          nt[s->partition_for_site[l] * 4 + s->ref[l]] -= s->nu_l[l];
      for (int i = e->fs_off[v]; i != e->fs_off[v + 1]; ++i) {
        int l = e->fs_site[i], b = s->partition_for_site[l];
        nt[b * 4 + s->ref[l]] += s->nu_l[l];
        nt[b * 4 + e->fs_from[i]] -= s->nu_l[l];
      }
      for (int i = e->mut_off[v]; i != e->mut_off[v + 1]; ++i) {
        int l = e->mut_site[i], b = s->partition_for_site[l];
        nt[b * 4 + e->mut_from[i]] -= s->nu_l[l];
        nt[b * 4 + e->mut_to[i]] += s->nu_l[l];
      }
      if (v != e->root) {
        double branch_length = e->t[v] - e->t[parent];
        for (int b = 0; b != P; ++b)
          for (int a = 0; a != 4; ++a) out[b * 4 + a] += nt[b * 4 + a] * branch_length;
        for (int i = e->mut_off[v + 1] - 1; i >= e->mut_off[v]; --i) {
          int l = e->mut_site[i], b = s->partition_for_site[l];
          out[b * 4 + e->mut_to[i]] -= s->nu_l[l] * (e->mut_t[i] - e->t[parent]);
          out[b * 4 + e->mut_from[i]] += s->nu_l[l] * (e->mut_t[i] - e->t[parent]);
        }
      }
      /* exit after children; children visited child0 first */
      stack[sp].node = v; stack[sp].phase = 1; ++sp;
      if (e->child0[v] >= 0) {
        stack[sp].node = e->child1[v]; stack[sp].phase = 0; ++sp;
        stack[sp].node = e->child0[v]; stack[sp].phase = 0; ++sp;
      }
    } else {
      for (int i = e->mut_off[v]; i != e->mut_off[v + 1]; ++i) {
        int l = e->mut_site[i], b = s->partition_for_site[l];
        nt[b * 4 + e->mut_to[i]] -= s->nu_l[l];
        nt[b * 4 + e->mut_from[i]] += s->nu_l[l];
      }
      for (int i = e->miss_off[v]; i != e->miss_off[v + 1]; ++i)
        for (int l = e->miss_start[i]; l != e->miss_end[i]; ++l)
          nt[s->partition_for_site[l] * 4 + s->ref[l]] += s->nu_l[l];
      for (int i = e->fs_off[v]; i != e->fs_off[v + 1]; ++i) {
        int l = e->fs_site[i], b = s->partition_for_site[l];
        nt[b * 4 + s->ref[l]] -= s->nu_l[l];
        nt[b * 4 + e->fs_from[i]] += s->nu_l[l];
      }
    }
  }
  free(stack);
  free(nt);
}

/* ---- phylo_tree_calc.cpp:41-56 (+ interval_set.h:238-288 merge semantics) ----------------------- */
typedef struct { int32_t s, e; } ivl;
static int ivl_cmp(const void* a, const void* b) {
  const ivl* x = (const ivl*)a; const ivl* y = (const ivl*)b;
  if (x->s != y->s) return x->s < y->s ? -1 : 1;
  return (x->e > y->e) - (x->e < y->e);
}
int32_t orc_missing_sites_at(const orc_emat* e, int32_t node, int32_t* starts, int32_t* ends, int32_t cap) {
  int total = 0;
  for (int cur = node; cur != -1; cur = e->parent[cur]) total += e->miss_off[cur + 1] - e->miss_off[cur];
  ivl* all = (ivl*)malloc(sizeof(ivl) * (size_t)(total > 0 ? total : 1));
  int k = 0;
  for (int cur = node; cur != -1; cur = e->parent[cur])
    for (int i = e->miss_off[cur]; i != e->miss_off[cur + 1]; ++i) { all[k].s = e->miss_start[i]; all[k].e = e->miss_end[i]; ++k; }
  qsort(all, (size_t)k, sizeof(ivl), ivl_cmp);
  int n_out = 0;
  for (int i = 0; i < k; ) {
    int cs = all[i].s, ce = all[i].e; ++i;
    while (i < k && all[i].s <= ce) { if (all[i].e > ce) ce = all[i].e; ++i; }   /* touching intervals coalesce */
    if (n_out >= cap) { free(all); return -1; }
    starts[n_out] = cs; ends[n_out] = ce; ++n_out;
  }
  free(all);
  return n_out;
}

/* ---- phylo_tree_calc.cpp:108-118 ---------------------------------------------------------------- */
uint8_t orc_site_state_at_node(const orc_emat* e, const orc_sites* s, int32_t node, int32_t site) {
  for (int cur = node; cur != -1; cur = e->parent[cur])
    for (int i = e->mut_off[cur + 1] - 1; i >= e->mut_off[cur]; --i)
      if (e->mut_site[i] == site) return e->mut_to[i];
  return s->ref[site];
}

/* =================================================================================================
 * SPR study (core/spr_study.h, core/spr_study.cpp)
 * ================================================================================================= */

/* Site_deltas (core/site_deltas.h:13-83) as a dense table: present/from/to per site + a counter. */
typedef struct {
  uint8_t* present; uint8_t* from; uint8_t* to; int32_t count;
} site_deltas;

static void sd_push_front(site_deltas* d, int site, int from, int to) {       /* site_deltas.h:42-65 */
  if (!d->present[site]) {
    d->present[site] = 1; d->from[site] = (uint8_t)from; d->to[site] = (uint8_t)to; ++d->count;
  } else {
    /* CHECK_EQ(delta_z.to, delta_post_z_to_x.from) in the reference */
    d->from[site] = (uint8_t)from;
    if (d->from[site] == d->to[site]) { d->present[site] = 0; --d->count; }
  }
}
static void sd_pop_front(site_deltas* d, int site, int from, int to) {        /* site_deltas.h:69-83 */
  sd_push_front(d, site, to, from);
}

static int missing_contains(const int32_t* ms, const int32_t* me, int n, int l) {   /* interval_set.h:128-135 */
  int lo = 0, hi = n;                 /* upper_bound on start */
  while (lo < hi) { int mid = (lo + hi) / 2; if (ms[mid] > l) hi = mid; else lo = mid + 1; }
  if (lo == 0) return 0;
  return l < me[lo - 1];
}

typedef struct {
  const orc_emat* e;
  int cur_branch, cur_mut_idx, cur_muts_from_start;
  site_deltas d;
  const int32_t* ms; const int32_t* me; int n_missing;
  int X; double t_X; int max_muts_from_start;
  orc_region* out; int cap; int n_out; int overflow;
} builder;

static int nmuts(const orc_emat* e, int b) { return e->mut_off[b + 1] - e->mut_off[b]; }

static double region_t_min(const orc_emat* e, int branch, int mut_idx) {      /* spr_study.h:90-95 */
  if (branch == e->root) return -DBL_MAX;
  if (mut_idx == 0) return e->t[e->parent[branch]];
  return e->mut_t[e->mut_off[branch] + mut_idx - 1];
}
static double region_t_max(const orc_emat* e, int branch, int mut_idx) {      /* spr_study.h:96-101 */
  if (branch == e->root) return e->t[branch];
  if (mut_idx == nmuts(e, branch)) return e->t[branch];
  return e->mut_t[e->mut_off[branch] + mut_idx];
}

static void move_to_neighbor(builder* b, int target_branch, int target_mut_idx, int is_backtracking) {  /* spr_study.cpp:43-91 */
  const orc_emat* e = b->e;
  if (b->cur_branch != -1 && target_branch == b->cur_branch) {
    if (target_mut_idx == b->cur_mut_idx + 1) {
      int i = e->mut_off[b->cur_branch] + b->cur_mut_idx;
      int l = e->mut_site[i];
      if (!missing_contains(b->ms, b->me, b->n_missing, l)) {
        sd_pop_front(&b->d, l, e->mut_from[i], e->mut_to[i]);
        b->cur_muts_from_start += (!is_backtracking ? +1 : -1);
      }
    } else if (target_mut_idx == b->cur_mut_idx - 1) {
      int i = e->mut_off[b->cur_branch] + target_mut_idx;
      int l = e->mut_site[i];
      if (!missing_contains(b->ms, b->me, b->n_missing, l)) {
        sd_push_front(&b->d, l, e->mut_from[i], e->mut_to[i]);
        b->cur_muts_from_start += (!is_backtracking ? +1 : -1);
      }
    }
  }
  b->cur_branch = target_branch;
  b->cur_mut_idx = target_mut_idx;
}

typedef struct { int32_t branch; int32_t mut_idx; int32_t back; } work_item;

int32_t orc_spr_study_build(const orc_emat* e, int32_t num_sites,
                            int32_t X, double t_X,
                            const int32_t* missing_starts, const int32_t* missing_ends, int32_t n_missing,
                            int32_t start_branch, int32_t start_mut_idx,
                            const int32_t* init_site, const uint8_t* init_from, const uint8_t* init_to, int32_t n_init,
                            int32_t max_muts_from_start, int32_t can_change_root,
                            orc_region* out, int32_t cap) {
  builder b;
  b.e = e; b.cur_branch = -1; b.cur_mut_idx = -1; b.cur_muts_from_start = 0;
  b.d.present = (uint8_t*)calloc((size_t)num_sites, 1);
  b.d.from = (uint8_t*)calloc((size_t)num_sites, 1);
  b.d.to = (uint8_t*)calloc((size_t)num_sites, 1);
  b.d.count = 0;
  for (int i = 0; i != n_init; ++i) {
    b.d.present[init_site[i]] = 1; b.d.from[init_site[i]] = init_from[i]; b.d.to[init_site[i]] = init_to[i];
    ++b.d.count;
  }
  b.ms = missing_starts; b.me = missing_ends; b.n_missing = n_missing;
  b.X = X; b.t_X = t_X; b.max_muts_from_start = max_muts_from_start;
  b.out = out; b.cap = cap; b.n_out = 0; b.overflow = 0;

  /* work stack: every forward move pushes (backtrack, forward); each region is entered once, so
   * 2 * (#regions + slack) bounds the depth. */
  size_t total_muts = (size_t)e->mut_off[e->num_nodes];
  size_t ws_cap = 4 * ((size_t)e->num_nodes + total_muts) + 16;
  work_item* ws = (work_item*)malloc(sizeof(work_item) * ws_cap);
  size_t sp = 0;

  /* seed_fill_from (spr_study.cpp:9-24): add_forward_movement(init_branch, init_mut_idx) */
  ws[sp].branch = b.cur_branch; ws[sp].mut_idx = b.cur_mut_idx; ws[sp].back = 1; ++sp;
  ws[sp].branch = start_branch; ws[sp].mut_idx = start_mut_idx; ws[sp].back = 0; ++sp;

  /* do_pending_work (spr_study.cpp:26-41) */
  while (sp > 0) {
    work_item w = ws[--sp];
    int old_branch = b.cur_branch, old_mut_idx = b.cur_mut_idx;
    move_to_neighbor(&b, w.branch, w.mut_idx, w.back);
    int in_scope = (b.cur_branch != b.X) && (b.cur_muts_from_start <= b.max_muts_from_start);   /* spr_study.h:86-89 */
    if (!w.back && in_scope) {
      /* visit_cur_region (spr_study.cpp:93-101) */
      if (b.n_out >= b.cap) { b.overflow = 1; break; }
      orc_region* r = &b.out[b.n_out++];
      memset(r, 0, sizeof(*r));
      r->branch = b.cur_branch; r->mut_idx = b.cur_mut_idx;
      r->t_min = region_t_min(e, b.cur_branch, b.cur_mut_idx);
      r->t_max = region_t_max(e, b.cur_branch, b.cur_mut_idx);
      r->min_muts = b.d.count;
      /* seed_neighbors_except (spr_study.cpp:103-128) */
#define MAYBE_MOVE_TO(nb, nm) do { \
        if (!((nb) == old_branch && (nm) == old_mut_idx)) { \
          ws[sp].branch = b.cur_branch; ws[sp].mut_idx = b.cur_mut_idx; ws[sp].back = 1; ++sp; \
          ws[sp].branch = (nb); ws[sp].mut_idx = (nm); ws[sp].back = 0; ++sp; \
        } } while (0)
      if (b.cur_branch != e->root) {
        if (b.cur_mut_idx > 0) { MAYBE_MOVE_TO(b.cur_branch, b.cur_mut_idx - 1); }
        else { int pb = e->parent[b.cur_branch]; MAYBE_MOVE_TO(pb, nmuts(e, pb)); }
      }
      if (b.cur_mut_idx < nmuts(e, b.cur_branch)) { MAYBE_MOVE_TO(b.cur_branch, b.cur_mut_idx + 1); }
      else if (e->child0[b.cur_branch] >= 0) {
        MAYBE_MOVE_TO(e->child0[b.cur_branch], 0);
        MAYBE_MOVE_TO(e->child1[b.cur_branch], 0);
      }
#undef MAYBE_MOVE_TO
    }
  }
  free(ws);
  free(b.d.present); free(b.d.from); free(b.d.to);
  if (b.overflow) return -1;

  int n = b.n_out;
  /* account_for_Xs_detachment (spr_study.cpp:130-209) */
  if (X == -1) {
    if (!can_change_root) {
      int w = 0;
      for (int i = 0; i != n; ++i) if (out[i].branch != e->root) out[w++] = out[i];
      n = w;
    }
  } else {
    int P = e->parent[X];
    int S = (e->child0[P] == X) ? e->child1[P] : e->child0[P];
    int num_muts_G_to_P = nmuts(e, P);
    for (int i = 0; i != n; ++i) {
      orc_region* r = &out[i];
      if (!can_change_root) {
        if (r->branch == e->root) { r->branch = -1; continue; }
      }
      if (r->branch != S && r->branch != P) continue;
      if (P != e->root) {
        if (r->branch == S) {
          if (r->mut_idx == 0) r->t_min = region_t_min(e, P, num_muts_G_to_P);
          r->mut_idx += num_muts_G_to_P;
        } else if (r->branch == P) {
          if (r->mut_idx == num_muts_G_to_P) r->branch = -1;
          else r->branch = S;
        }
      } else {
        if (!can_change_root) {
          if (r->branch == P) r->branch = -1;
        } else {
          if (r->branch == S && r->mut_idx == nmuts(e, S)) {
            r->mut_idx += num_muts_G_to_P;
            r->t_min = -DBL_MAX;
          } else {
            r->branch = -1;
          }
        }
      }
    }
    int w = 0;
    for (int i = 0; i != n; ++i) if (out[i].branch != -1) out[w++] = out[i];
    n = w;
  }
  /* remove_regions_in_Xs_future (spr_study.cpp:211-224) */
  {
    int w = 0;
    for (int i = 0; i != n; ++i) {
      if (out[i].t_min >= t_X) continue;
      if (out[i].t_max > t_X) out[i].t_max = t_X;
      out[w++] = out[i];
    }
    n = w;
  }
  return n;
}

/* ---- spr_study.cpp:226-385 ---------------------------------------------------------------------- */
void orc_spr_study_weights(const orc_emat* e, int32_t num_sites, int32_t num_missing_at_X,
                           orc_region* regions, int32_t n_regions,
                           double lambda_X, double annealing_factor, double t_X, double t_max_tip,
                           orc_study_summary* out) {
  double mu = lambda_X / (num_sites - num_missing_at_X);
  double f = annealing_factor;
  for (int i = 0; i != n_regions; ++i) {
    orc_region* r = &regions[i];
    double t_min = r->t_min, t_max = r->t_max;
    int m = r->min_muts;
    if (!(t_min == -DBL_MAX)) {
      double t_prime = 0.5 * (t_min + t_max);
      r->log_W_over_Wmax =
          log(f * lambda_X * (t_max - t_min)) +
          f * (-lambda_X * (t_X - t_prime) + m * log(mu * (t_X - t_prime) / 3));
    } else {
      double t_S = e->t[r->branch];
      double s_min = fabs(t_X - t_S);
      double t_early = t_X < t_S ? t_X : t_S;
      double tree_span = t_max_tip - t_early;
      double s_max = s_min + 20.0 * tree_span;
      double x_min = lambda_X * f * s_min;
      double x_max = lambda_X * f * s_max;
      if (x_max < 0.01) {
        double alpha = f * m + 1;
        r->log_W_over_Wmax =
            -M_LN2
            + log(f * lambda_X)
            + f * m * log(mu / 3)
            + alpha * log(s_max) + log1p(-pow(s_min / s_max, alpha))
            - log(alpha);
      } else {
        r->log_W_over_Wmax =
            -M_LN2
            + f * m * log(mu / (3 * lambda_X * f))
            + lgamma(f * m + 1)
            + orc_safe_log_gamma_integral(f * m + 1, x_min, x_max);
      }
    }
  }
  double log_Wmax = n_regions > 0 ? regions[0].log_W_over_Wmax : 0.0;
  for (int i = 0; i != n_regions; ++i)
    if (regions[i].log_W_over_Wmax > log_Wmax) log_Wmax = regions[i].log_W_over_Wmax;    /* std::max(a,b) */
  double sum = 0.0;
  for (int i = 0; i != n_regions; ++i) {
    regions[i].log_W_over_Wmax -= log_Wmax;
    regions[i].W_over_Wmax = exp(regions[i].log_W_over_Wmax);
    sum += regions[i].W_over_Wmax;
  }
  out->mu = mu; out->log_Wmax = log_Wmax; out->sum_W_over_Wmax = sum;
  out->num_regions = n_regions; out->num_missing_at_X = num_missing_at_X;
}

/* ---- spr_study.cpp:404-422 ---------------------------------------------------------------------- */
int32_t orc_spr_pick_nexus_region(const orc_region* regions, int32_t n_regions, double r) {
  int chosen = 0;
  for (int i = 0; i != n_regions; ++i) {
    if (regions[i].W_over_Wmax >= r) { chosen = i; break; }
    else r -= regions[i].W_over_Wmax;
  }
  return chosen;
}

/* ---- spr_study.cpp:474-484 ---------------------------------------------------------------------- */
int32_t orc_spr_find_region(const orc_region* regions, int32_t n_regions, int32_t branch, double t) {
  for (int i = 0; i != n_regions; ++i)
    if (regions[i].branch == branch && regions[i].t_min < t && t <= regions[i].t_max) return i;
  return -1;
}

/* ---- spr_study.cpp:486-549 ---------------------------------------------------------------------- */
double orc_spr_log_alpha_in_region(const orc_emat* e, const orc_region* regions, int32_t n_regions,
                                   int32_t region_idx, double t,
                                   double lambda_X, double annealing_factor, double t_X, double t_max_tip,
                                   double sum_W_over_Wmax) {
  (void)n_regions;
  const orc_region* r = &regions[region_idx];
  double log_p_region = r->log_W_over_Wmax - log(sum_W_over_Wmax);
  if (!(r->t_min == -DBL_MAX)) {
    return log_p_region - log(r->t_max - r->t_min);
  } else {
    double f = annealing_factor;
    int m = r->min_muts;
    double t_S = e->t[r->branch];
    double s_min = fabs(t_X - t_S);
    double t_early = t_X < t_S ? t_X : t_S;
    double tree_span = t_max_tip - t_early;
    double s_max = s_min + 20.0 * tree_span;
    double x_min = lambda_X * f * s_min;
    double x_max = lambda_X * f * s_max;
    double s = t_X - t + t_S - t;
    if (s > s_max + 1e-6) return -INFINITY;
    if (x_max < 0.01) {
      double alpha = f * m + 1;
      return log_p_region + M_LN2 + log(alpha) + (alpha - 1) * log(s)
          + -alpha * log(s_max) + -log1p(-pow(s_min / s_max, alpha));
    } else {
      return log_p_region + M_LN2 + log(lambda_X * f) + f * m * log(lambda_X * f * s)
          + -lambda_X * f * s + -lgamma(f * m + 1) - orc_safe_log_gamma_integral(f * m + 1, x_min, x_max);
    }
  }
}

/* ---- convenience: a study of an ATTACHED X exactly as Subrun::spr1_move seeds it (core/subrun.cpp:540-553),
 * but without the preceding peel_graft (the builder only needs a self-consistent tree and the deltas from the
 * start region to X). */
int32_t orc_spr_study_from_attached(const orc_emat* e, const orc_sites* s, int32_t X,
                                    int32_t max_muts_from_start, int32_t can_change_root,
                                    double annealing_factor, double t_max_tip, const double* lambda_i,
                                    orc_region* out, int32_t cap, orc_study_summary* summary) {
  int L = s->num_sites;
  int P = e->parent[X];
  int S = (e->child0[P] == X) ? e->child1[P] : e->child0[P];
  int total_ivl = 0;
  for (int cur = X; cur != -1; cur = e->parent[cur]) total_ivl += e->miss_off[cur + 1] - e->miss_off[cur];
  int32_t* ms = (int32_t*)malloc(sizeof(int32_t) * (size_t)(total_ivl + 1));
  int32_t* me = (int32_t*)malloc(sizeof(int32_t) * (size_t)(total_ivl + 1));
  int n_missing = orc_missing_sites_at(e, X, ms, me, total_ivl + 1);
  int num_missing = 0;
  for (int i = 0; i != n_missing; ++i) num_missing += me[i] - ms[i];

  /* net deltas along branch P->X == calc_site_deltas_between(tree, P, X) (core/site_deltas.cpp:83-101) */
  int nm = nmuts(e, X);
  int32_t* isite = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nm + 1));
  uint8_t* ifrom = (uint8_t*)malloc((size_t)(nm + 1));
  uint8_t* ito = (uint8_t*)malloc((size_t)(nm + 1));
  int n_init = 0;
  for (int i = e->mut_off[X]; i != e->mut_off[X + 1]; ++i) {
    int l = e->mut_site[i];
    int j;
    for (j = 0; j != n_init; ++j) if (isite[j] == l) break;
    if (j == n_init) { isite[j] = l; ifrom[j] = e->mut_from[i]; ito[j] = e->mut_to[i]; ++n_init; }
    else {
      ito[j] = e->mut_to[i];
      if (ifrom[j] == ito[j]) { --n_init; isite[j] = isite[n_init]; ifrom[j] = ifrom[n_init]; ito[j] = ito[n_init]; }
    }
  }
  int n = orc_spr_study_build(e, L, X, e->t[X], ms, me, n_missing, S, 0, isite, ifrom, ito, n_init,
                              max_muts_from_start, can_change_root, out, cap);
  if (n >= 0 && summary) {
    if (n > 0) orc_spr_study_weights(e, L, num_missing, out, n, lambda_i[X], annealing_factor, e->t[X], t_max_tip, summary);
    else { memset(summary, 0, sizeof(*summary)); summary->num_missing_at_X = num_missing; }
  }
  free(ms); free(me); free(isite); free(ifrom); free(ito);
  return n;
}

/* ---- delphy.api.Tree: the FlatBuffers wire format (core/api.fbs:13-49) ------------------------------------------------------------- */
static uint32_t rd_u32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint16_t rd_u16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static int32_t rd_i32(const uint8_t* p) { return (int32_t)rd_u32(p); }
static float rd_f32(const uint8_t* p) { uint32_t u = rd_u32(p); float f; memcpy(&f, &u, 4); return f; }
static void wr_u32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
static void wr_u16(uint8_t* p, uint16_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }

/* one vector field of the table at `tpos`: vtable slot `slot` (4 = nodes, 6 = mutations, ...: core/api_generated.h:269-275) */
static int api_vector(const uint8_t* buf, int64_t len, int64_t tpos, int64_t vt, int vtsize, int slot, int64_t elem, const uint8_t** data, int64_t* count) {
  *data = NULL; *count = 0;
  if (slot + 2 > vtsize) return 0;                       /* field absent: an empty vector */
  int fo = rd_u16(buf + vt + slot);
  if (fo == 0) return 0;
  int64_t fpos = tpos + fo;
  if (fpos + 4 > len) return -1;
  int64_t vpos = fpos + (int64_t)rd_u32(buf + fpos);
  if (vpos + 4 > len) return -1;
  int64_t n = (int64_t)rd_u32(buf + vpos);
  if (vpos + 4 + n * elem > len) return -1;
  *data = buf + vpos + 4; *count = n;
  return 0;
}

int32_t orc_api_tree_parse(const uint8_t* buf, int64_t len, orc_api_tree_view* out) {
  memset(out, 0, sizeof(*out));
  if (len < 12) return -1;
  int64_t body = (int64_t)rd_u32(buf);                   /* FinishSizePrefixed (core/api.cpp:95) */
  if (body + 4 > len) return -1;
  len = body + 4;
  int64_t tpos = 4 + (int64_t)rd_u32(buf + 4);
  if (tpos + 4 > len || tpos < 8) return -1;
  int64_t vt = tpos - (int64_t)rd_i32(buf + tpos);
  if (vt < 4 || vt + 4 > len) return -1;
  int vtsize = rd_u16(buf + vt);
  if (vtsize < 4 || (vtsize & 1) || vt + vtsize > len) return -1;
  int tsize = rd_u16(buf + vt + 2);
  if (tpos + tsize > len) return -1;
  for (int slot = 4; slot + 2 <= vtsize; slot += 2) if (rd_u16(buf + vt + slot) + 4 > tsize && rd_u16(buf + vt + slot) != 0) return -1;
  int64_t n = 0, L = 0;
  if (api_vector(buf, len, tpos, vt, vtsize, 4, 16, &out->nodes, &n) != 0) return -1;
  if (api_vector(buf, len, tpos, vt, vtsize, 6, 16, &out->muts, &out->num_muts) != 0) return -1;
  if (api_vector(buf, len, tpos, vt, vtsize, 8, 12, &out->ivls, &out->num_ivls) != 0) return -1;
  if (api_vector(buf, len, tpos, vt, vtsize, 10, 1, &out->ref_seq, &L) != 0) return -1;
  if (n > INT32_MAX || L > INT32_MAX || out->num_muts > INT32_MAX || out->num_ivls > INT32_MAX) return -1;
  out->num_nodes = (int32_t)n; out->num_sites = (int32_t)L;
  out->root = (12 + 2 <= vtsize && rd_u16(buf + vt + 12) != 0) ? rd_i32(buf + tpos + rd_u16(buf + vt + 12)) : 0;   /* default 0 (:289) */
  return 0;
}

int32_t orc_api_tree_to_emat(const orc_api_tree_view* v, int32_t* counts, int32_t* parent, int32_t* child0, int32_t* child1, double* t,
                             int32_t* mut_off, int32_t* mut_site, uint8_t* mut_from, uint8_t* mut_to, double* mut_t,
                             int32_t* miss_off, int32_t* miss_start, int32_t* miss_end,
                             int32_t* fs_off, int32_t* fs_site, uint8_t* fs_from) {
  const int n = v->num_nodes, L = v->num_sites;
  const int64_t M = v->num_muts, I = v->num_ivls;
  if (n <= 0 || v->root < 0 || v->root >= n) return -1;
  for (int l = 0; l != L; ++l) if (v->ref_seq[l] > 3) return -1;                                  /* to_real_seq_letter throws (core/api.cpp:14-22) */
  /* the lists by branch (core/api.cpp:165-180 appends each record to its branch in file order; the writer emits them branch-major) */
  int32_t* moff = (int32_t*)calloc((size_t)n + 1, sizeof(int32_t));
  int32_t* ioff = (int32_t*)calloc((size_t)n + 1, sizeof(int32_t));
  int32_t* morder = (int32_t*)malloc(sizeof(int32_t) * (size_t)(M > 0 ? M : 1));
  int32_t* iorder = (int32_t*)malloc(sizeof(int32_t) * (size_t)(I > 0 ? I : 1));
  int32_t* foff = (int32_t*)calloc((size_t)n + 1, sizeof(int32_t));
  uint8_t* cur = (uint8_t*)malloc((size_t)(L > 0 ? L : 1));
  int32_t* nmiss = (int32_t*)calloc((size_t)(L > 0 ? L : 1), sizeof(int32_t));
  int32_t* stack = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)n);
  int32_t rc = 0;
  for (int64_t i = 0; i != M && rc == 0; ++i) {
    const uint8_t* r = v->muts + 16 * i;
    int b = rd_i32(r), l = rd_i32(r + 4);
    if (b < 0 || b >= n || l < 0 || l >= L || r[8] > 3 || r[9] > 3) rc = -1; else ++moff[b + 1];
  }
  for (int64_t i = 0; i != I && rc == 0; ++i) {
    const uint8_t* r = v->ivls + 12 * i;
    int b = rd_i32(r), s = rd_i32(r + 4), e = rd_i32(r + 8);
    if (b < 0 || b >= n || s < 0 || e > L || s >= e) rc = -1; else ++ioff[b + 1];
  }
  for (int x = 0; x != n && rc == 0; ++x) {
    const uint8_t* r = v->nodes + 16 * (size_t)x;
    int p = rd_i32(r), c0 = rd_i32(r + 4), c1 = rd_i32(r + 8);
    if (p < -1 || p >= n || c0 < -1 || c0 >= n || c1 < -1 || c1 >= n || ((c0 < 0) != (c1 < 0))) rc = -1;
  }
  if (rc == 0) {
    for (int x = 0; x != n; ++x) { moff[x + 1] += moff[x]; ioff[x + 1] += ioff[x]; }
    int32_t* fill = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    memcpy(fill, moff, sizeof(int32_t) * (size_t)n);
    for (int64_t i = 0; i != M; ++i) morder[fill[rd_i32(v->muts + 16 * i)]++] = (int32_t)i;
    memcpy(fill, ioff, sizeof(int32_t) * (size_t)n);
    for (int64_t i = 0; i != I; ++i) iorder[fill[rd_i32(v->ivls + 12 * i)]++] = (int32_t)i;
    free(fill);
    /* Interval_set::insert keeps a branch's intervals sorted and merged; a normal-form buffer already has them so */
    for (int x = 0; x != n && rc == 0; ++x)
      for (int k = ioff[x] + 1; k < ioff[x + 1]; ++k)
        if (rd_i32(v->ivls + 12 * (size_t)iorder[k] + 4) <= rd_i32(v->ivls + 12 * (size_t)iorder[k - 1] + 8)) rc = -2;
  }
  /* pass 1 of fix_up_missations (core/phylo_tree.cpp:383-398) must find nothing to bubble up: the children of an inner node share no missing site */
  for (int x = 0; x != n && rc == 0; ++x) {
    int c0 = rd_i32(v->nodes + 16 * (size_t)x + 4), c1 = rd_i32(v->nodes + 16 * (size_t)x + 8);
    if (c0 < 0) continue;
    int a = ioff[c0], b = ioff[c1];
    while (a < ioff[c0 + 1] && b < ioff[c1 + 1]) {
      int as = rd_i32(v->ivls + 12 * (size_t)iorder[a] + 4), ae = rd_i32(v->ivls + 12 * (size_t)iorder[a] + 8);
      int bs = rd_i32(v->ivls + 12 * (size_t)iorder[b] + 4), be = rd_i32(v->ivls + 12 * (size_t)iorder[b] + 8);
      if (as < be && bs < ae) { rc = -2; break; }
      if (ae <= be) ++a; else ++b;
    }
  }
  /* passes 2 and 3 (core/phylo_tree.cpp:400-478): a traversal carrying the current sequence and the set of missing sites */
  int F = 0;
  for (int pass = 0; pass != 2 && rc == 0; ++pass) {
    if (pass == 1 && !parent) break;
    if (L > 0) memcpy(cur, v->ref_seq, (size_t)L);
    int sp = 0, f = 0;
    stack[sp++] = v->root; stack[sp++] = 0;
    int visited = 0;
    while (sp > 0 && rc == 0) {
      int state = stack[--sp], x = stack[--sp];
      int c0 = rd_i32(v->nodes + 16 * (size_t)x + 4), c1 = rd_i32(v->nodes + 16 * (size_t)x + 8);
      if (state == 0) {
        if (++visited > n) { rc = -1; break; }
        if (x != v->root && rd_i32(v->nodes + 16 * (size_t)x) < 0) { rc = -1; break; }
        /* entering x: from_states = the sites of its missations whose current state is not the reference's (:451-458) */
        const int f_in = f;
        int w = foff[x];                                                               /* (pass 1) where this node's overrides go: CSR by node index */
        for (int k = ioff[x]; k != ioff[x + 1] && rc == 0; ++k) {
          const uint8_t* r = v->ivls + 12 * (size_t)iorder[k];
          for (int l = rd_i32(r + 4); l != rd_i32(r + 8); ++l) {
            if (nmiss[l]++ != 0) { rc = -2; break; }                                   /* pass 2 would rewrite this node's missations (:412-426) */
            if (cur[l] != v->ref_seq[l]) { if (pass == 1) { fs_site[w] = l; fs_from[w] = cur[l]; ++w; } ++f; }
          }
        }
        if (pass == 0) foff[x + 1] = f - f_in;
        for (int k = moff[x]; k != moff[x + 1] && rc == 0; ++k) {
          const uint8_t* r = v->muts + 16 * (size_t)morder[k];
          int l = rd_i32(r + 4);
          if (nmiss[l] != 0) { rc = -2; break; }                                      /* :461 would erase it */
          if (r[8] != cur[l]) { rc = -3; break; }                                     /* CHECK_EQ(m.from, cur_seq[m.site]) (:465) */
          cur[l] = r[9];
        }
        stack[sp++] = x; stack[sp++] = 1;
        if (c0 >= 0) { stack[sp++] = c1; stack[sp++] = 0; stack[sp++] = c0; stack[sp++] = 0; }
      } else {
        for (int k = moff[x + 1] - 1; k >= moff[x]; --k) { const uint8_t* r = v->muts + 16 * (size_t)morder[k]; cur[rd_i32(r + 4)] = r[8]; }
        for (int k = ioff[x]; k != ioff[x + 1]; ++k) {
          const uint8_t* r = v->ivls + 12 * (size_t)iorder[k];
          for (int l = rd_i32(r + 4); l != rd_i32(r + 8); ++l) --nmiss[l];
        }
      }
    }
    if (rc == 0 && visited != n) rc = -1;
    if (pass == 0) { F = f; for (int x = 0; x != n; ++x) foff[x + 1] += foff[x]; }
    else memcpy(fs_off, foff, sizeof(int32_t) * ((size_t)n + 1));
    /* an aborted traversal leaves marks behind; the arrays are only reused after a clean pass 0 */
  }
  counts[0] = n; counts[1] = v->root; counts[2] = (int32_t)M; counts[3] = (int32_t)I; counts[4] = F;
  if (rc == 0 && parent) {
    for (int x = 0; x != n; ++x) {
      const uint8_t* r = v->nodes + 16 * (size_t)x;
      parent[x] = rd_i32(r); child0[x] = rd_i32(r + 4); child1[x] = rd_i32(r + 8); t[x] = (double)rd_f32(r + 12);
    }
    memcpy(mut_off, moff, sizeof(int32_t) * ((size_t)n + 1));
    memcpy(miss_off, ioff, sizeof(int32_t) * ((size_t)n + 1));
    for (int64_t k = 0; k != M; ++k) {
      const uint8_t* r = v->muts + 16 * (size_t)morder[k];
      mut_site[k] = rd_i32(r + 4); mut_from[k] = r[8]; mut_to[k] = r[9]; mut_t[k] = (double)rd_f32(r + 12);
    }
    for (int64_t k = 0; k != I; ++k) {
      const uint8_t* r = v->ivls + 12 * (size_t)iorder[k];
      miss_start[k] = rd_i32(r + 4); miss_end[k] = rd_i32(r + 8);
    }
  }
  free(moff); free(ioff); free(morder); free(iorder); free(foff); free(cur); free(nmiss); free(stack);
  return rc;
}

int64_t orc_api_tree_write(const orc_emat* e, const uint8_t* ref, int32_t num_sites, uint8_t* out, int64_t cap) {
  const int n = e->num_nodes;
  const int64_t M = e->mut_off[n], I = e->miss_off[n];
  /* [size][root uoffset][vtable 14 B + 2][table 24 B][nodes][mutations][missation_intervals][ref_seq] */
  const int64_t vt = 8, tpos = 24, v_nodes = 48, v_muts = v_nodes + 4 + 16 * (int64_t)n, v_ivls = v_muts + 4 + 16 * M,
                v_ref = v_ivls + 4 + 12 * I, total = (v_ref + 4 + num_sites + 3) / 4 * 4;
  if (total > cap) return -1;
  memset(out, 0, (size_t)total);
  wr_u32(out, (uint32_t)(total - 4));
  wr_u32(out + 4, (uint32_t)(tpos - 4));
  wr_u16(out + vt, 14); wr_u16(out + vt + 2, 24);
  for (int i = 0; i != 5; ++i) wr_u16(out + vt + 4 + 2 * i, (uint16_t)(4 + 4 * i));
  wr_u32(out + tpos, (uint32_t)(tpos - vt));
  wr_u32(out + tpos + 4, (uint32_t)(v_nodes - (tpos + 4)));
  wr_u32(out + tpos + 8, (uint32_t)(v_muts - (tpos + 8)));
  wr_u32(out + tpos + 12, (uint32_t)(v_ivls - (tpos + 12)));
  wr_u32(out + tpos + 16, (uint32_t)(v_ref - (tpos + 16)));
  wr_u32(out + tpos + 20, (uint32_t)e->root);
  wr_u32(out + v_nodes, (uint32_t)n); wr_u32(out + v_muts, (uint32_t)M); wr_u32(out + v_ivls, (uint32_t)I); wr_u32(out + v_ref, (uint32_t)num_sites);
  for (int x = 0; x != n; ++x) {
    uint8_t* r = out + v_nodes + 4 + 16 * (size_t)x;
    float tf = (float)e->t[x]; uint32_t u; memcpy(&u, &tf, 4);
    wr_u32(r, (uint32_t)e->parent[x]); wr_u32(r + 4, (uint32_t)e->child0[x]); wr_u32(r + 8, (uint32_t)e->child1[x]); wr_u32(r + 12, u);   /* tips: -1, -1 (core/api.cpp:64-68) */
    for (int k = e->mut_off[x]; k != e->mut_off[x + 1]; ++k) {
      uint8_t* m = out + v_muts + 4 + 16 * (size_t)k;
      float mf = (float)e->mut_t[k]; memcpy(&u, &mf, 4);
      wr_u32(m, (uint32_t)x); wr_u32(m + 4, (uint32_t)e->mut_site[k]); m[8] = e->mut_from[k]; m[9] = e->mut_to[k]; wr_u32(m + 12, u);
    }
    for (int k = e->miss_off[x]; k != e->miss_off[x + 1]; ++k) {
      uint8_t* m = out + v_ivls + 4 + 12 * (size_t)k;
      wr_u32(m, (uint32_t)x); wr_u32(m + 4, (uint32_t)e->miss_start[k]); wr_u32(m + 8, (uint32_t)e->miss_end[k]);
    }
  }
  if (num_sites > 0) memcpy(out + v_ref + 4, ref, (size_t)num_sites);
  return total;
}

double orc_gamma_q_export(double a, double x) { return orc_gamma_q(a, x); }
