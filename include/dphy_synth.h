/* dphy_synth.h -- synthetic EMAT generator (libdphy_synth.so): INPUT GENERATOR for bench.py and the tests, not part of the
 * product library.  It lives in its own shared object so that the reference arm of bench.py (and anything else that only
 * needs inputs) loads no product code.  Shapes and parameters follow SURVEY.md section 8(d). */
#ifndef DPHY_SYNTH_H_
#define DPHY_SYNTH_H_

#include "delphy_b200.h"   /* dphy_emat_host / dphy_sites_host: the flat host layout the generator emits */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dphy_synth_params {
  int32_t num_tips;
  int32_t num_sites;
  uint64_t seed;
  double  muts_per_tip;          /* target M / n (SURVEY: ~1.5) */
  double  tip_date_span_years;   /* tips uniform over this span */
  double  growth_rate;           /* exponential-growth coalescent g (1/yr) */
  double  n0_years;              /* N(0) in years */
  double  kappa;                 /* HKY transition/transversion ratio */
  double  pi[4];                 /* stationary frequencies */
  int32_t site_rate_heterogeneity; /* 0: nu_l == 1;  1: nu_l ~ Gamma(alpha, alpha) */
  double  gamma_alpha;
  int32_t num_partitions;        /* 1, or 2 (mpox-hack-like split, core/run.cpp:359-435) */
  double  missing_mean_intervals_per_tip;   /* Geometric mean; 0 disables missing data */
  double  missing_len_min, missing_len_max; /* LogUniform interval length */
  int32_t end_gaps;              /* add 5'/3' end gaps */
  int32_t num_root_mutations;    /* root "mutations" at t=-DBL_MAX (as partition parts have, core/run.cpp:148-154) */
  int32_t caterpillar;           /* 1: ladder topology (the reference's random initial tree, core/phylo_tree.cpp) */
} dphy_synth_params;

typedef struct dphy_synth_emat {     /* owns its arrays; free with dphy_synth_free */
  dphy_emat_host emat;
  dphy_sites_host sites;
  double mu_used;
  double t_max_tip;
  int64_t num_mutations, num_intervals, num_from_states, num_missing_sites;
  int32_t max_depth;
  void* owner_;
} dphy_synth_emat;

void dphy_synth_default_params(dphy_synth_params* p, int32_t config /* 1..5 == BASELINE.json configs[0..4] */);
int  dphy_synth_generate(const dphy_synth_params* p, dphy_synth_emat** out);
void dphy_synth_free(dphy_synth_emat* s);

#ifdef __cplusplus
}
#endif
#endif /* DPHY_SYNTH_H_ */
