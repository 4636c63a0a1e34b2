// synth_emat.cpp -- synthetic EMAT generator (host, C++): the input generator for bench.py and the parity tests.
//
// Produces a valid Explicit Mutation-Annotated Tree in the flat host layout of include/delphy_b200.h, i.e. one
// that satisfies every invariant the reference asserts in core/phylo_tree.cpp:18-135 (mutations sorted by (t,site)
// inside [t_parent, t_node], `from` equal to the running state, no mutation on a site missing at-or-above it,
// missation intervals factored as far rootward as possible and never shared by both children, from_states only
// where the state differs from the reference sequence).  Shapes/distributions follow SURVEY.md section 8(d):
// coalescent tree under exponential growth with heterochronous tips, HKY mutations dropped as a Poisson process,
// per-tip gap intervals (LogUniform lengths + 5'/3' end gaps) factored upward with interval algebra.
#include "dphy_synth.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <numeric>
#include <utility>
#include <vector>

namespace {

struct Rng {   // xoshiro256** seeded by splitmix64
  uint64_t s[4];
  explicit Rng(uint64_t seed) {
    for (auto& x : s) {
      seed += 0x9E3779B97F4A7C15ULL;
      uint64_t z = seed;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
      x = z ^ (z >> 31);
    }
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }          // [0,1)
  double uniform_open() { double u; do { u = uniform(); } while (u == 0.0); return u; }     // (0,1)
  uint64_t below(uint64_t n) { return (uint64_t)(uniform() * (double)n) % n; }
  double exponential() { return -std::log(uniform_open()); }
  double normal() {
    double u1 = uniform_open(), u2 = uniform();
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  }
  double gamma(double a) {   // Marsaglia-Tsang, shape a, scale 1
    if (a < 1.0) { return gamma(a + 1.0) * std::pow(uniform_open(), 1.0 / a); }
    double d = a - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
    for (;;) {
      double x = normal(), v = 1.0 + c * x;
      if (v <= 0) continue;
      v = v * v * v;
      double u = uniform_open();
      if (std::log(u) < 0.5 * x * x + d - d * v + d * std::log(v)) return d * v;
    }
  }
  int poisson(double lam) {
    if (lam <= 0) return 0;
    if (lam < 30.0) {
      double L = std::exp(-lam), p = 1.0; int k = 0;
      do { ++k; p *= uniform_open(); } while (p > L);
      return k - 1;
    }
    // normal approximation with continuity correction is plenty for synthetic data
    double x = lam + std::sqrt(lam) * normal() + 0.5;
    return x < 0 ? 0 : (int)x;
  }
};

using Ivl = std::pair<int32_t, int32_t>;
using Ivls = std::vector<Ivl>;

Ivls canonicalize(Ivls v) {
  std::sort(v.begin(), v.end());
  Ivls out;
  for (auto& iv : v) {
    if (iv.first >= iv.second) continue;
    if (!out.empty() && iv.first <= out.back().second) out.back().second = std::max(out.back().second, iv.second);
    else out.push_back(iv);
  }
  return out;
}
Ivls intersect(const Ivls& a, const Ivls& b) {
  Ivls out; size_t i = 0, j = 0;
  while (i < a.size() && j < b.size()) {
    int32_t s = std::max(a[i].first, b[j].first), e = std::min(a[i].second, b[j].second);
    if (s < e) out.push_back({s, e});
    if (a[i].second < b[j].second) ++i; else ++j;
  }
  return out;
}
Ivls subtract(const Ivls& a, const Ivls& b) {   // a \ b
  Ivls out; size_t j = 0;
  for (auto iv : a) {
    int32_t cur = iv.first;
    while (j < b.size() && b[j].second <= cur) ++j;
    size_t k = j;
    while (k < b.size() && b[k].first < iv.second) {
      if (b[k].first > cur) out.push_back({cur, b[k].first});
      cur = std::max(cur, b[k].second);
      ++k;
    }
    if (cur < iv.second) out.push_back({cur, iv.second});
  }
  return out;
}
bool contains(const Ivls& a, int32_t l) {
  auto it = std::upper_bound(a.begin(), a.end(), l, [](int32_t x, const Ivl& iv) { return x < iv.first; });
  if (it == a.begin()) return false;
  --it;
  return l < it->second;
}

struct Owner {
  std::vector<int32_t> parent, child0, child1, mut_off, mut_site, miss_off, miss_start, miss_end, fs_off, fs_site, part;
  std::vector<uint8_t> mut_from, mut_to, fs_from, ref;
  std::vector<double> t, mut_t, nu, mu, pi, q;
};

struct Mut { double t; int32_t site; uint8_t from, to; };

}  // namespace

extern "C" void dphy_synth_default_params(dphy_synth_params* p, int32_t config) {
  std::memset(p, 0, sizeof(*p));
  p->seed = 20251017ULL + (uint64_t)config;
  p->muts_per_tip = 1.5;
  p->tip_date_span_years = 0.5;
  p->growth_rate = 10.0;
  p->n0_years = 3.0;
  p->kappa = 5.0;
  p->pi[0] = 0.30; p->pi[1] = 0.18; p->pi[2] = 0.20; p->pi[3] = 0.32;
  p->gamma_alpha = 0.5;
  p->num_partitions = 1;
  p->missing_mean_intervals_per_tip = 3.0;
  p->missing_len_min = 10.0; p->missing_len_max = 2000.0;
  p->end_gaps = 1;
  switch (config) {
    case 1: p->num_tips = 200; p->num_sites = 29903; break;
    case 2: p->num_tips = 1600; p->num_sites = 18959; p->site_rate_heterogeneity = 1;
            p->missing_mean_intervals_per_tip = 1.0; p->missing_len_max = 600.0; break;
    case 3: p->num_tips = 10000; p->num_sites = 29903; break;
    case 4: p->num_tips = 100000; p->num_sites = 29903; break;
    case 5: p->num_tips = 50000; p->num_sites = 197000; p->tip_date_span_years = 2.0;
            p->missing_mean_intervals_per_tip = 12.0; p->missing_len_min = 50.0; p->missing_len_max = 12000.0; break;
    default: p->num_tips = 64; p->num_sites = 2000; p->missing_len_max = 200.0; break;
  }
}

extern "C" int dphy_synth_generate(const dphy_synth_params* pp, dphy_synth_emat** out) {
  if (!pp || !out || pp->num_tips < 2 || pp->num_sites < 4 || pp->num_partitions < 1 || pp->num_partitions > 2) {
    return DPHY_ERR_INVALID_ARGUMENT;
  }
  const dphy_synth_params p = *pp;
  const int n = p.num_tips, N = 2 * n - 1, L = p.num_sites, P = p.num_partitions;
  Rng rng(p.seed);
  auto* ow = new Owner();
  auto* res = new dphy_synth_emat();
  std::memset(res, 0, sizeof(*res));
  res->owner_ = ow;

  // ---- reference sequence + evo model -------------------------------------------------------------------
  ow->ref.resize(L);
  for (int l = 0; l < L; ++l) {
    double u = rng.uniform(), c = 0; int a = 0;
    for (; a < 3; ++a) { c += p.pi[a]; if (u < c) break; }
    ow->ref[l] = (uint8_t)a;
  }
  ow->part.assign(L, 0);
  if (P == 2) { for (int l = 0; l < L; ++l) ow->part[l] = rng.uniform() < 0.1 ? 1 : 0; }
  ow->nu.assign(L, 1.0);
  if (p.site_rate_heterogeneity) {
    for (int l = 0; l < L; ++l) ow->nu[l] = std::max(1e-6, rng.gamma(p.gamma_alpha) / p.gamma_alpha);
  }
  // HKY: transitions A<->G (0<->2), C<->T (1<->3)
  ow->pi.resize(4 * P); ow->q.resize(16 * P); ow->mu.resize(P);
  for (int b = 0; b < P; ++b) {
    double q[4][4]; double tot = 0;
    for (int a = 0; a < 4; ++a) {
      double row = 0;
      for (int c = 0; c < 4; ++c) {
        if (a == c) continue;
        bool transition = ((a ^ c) == 2);
        q[a][c] = (transition ? p.kappa : 1.0) * p.pi[c];
        row += q[a][c];
      }
      q[a][a] = -row;
      tot += p.pi[a] * row;
    }
    for (int a = 0; a < 4; ++a) { ow->pi[b * 4 + a] = p.pi[a]; for (int c = 0; c < 4; ++c) ow->q[b * 16 + a * 4 + c] = q[a][c] / tot; }
  }

  // ---- tree: heterochronous coalescent under exponential growth (or a ladder) -------------------------------
  ow->parent.assign(N, -1); ow->child0.assign(N, -1); ow->child1.assign(N, -1); ow->t.assign(N, 0.0);
  std::vector<double> tip_t(n);
  for (int i = 0; i < n; ++i) tip_t[i] = -rng.uniform() * p.tip_date_span_years;   // latest tip near 0
  for (int i = 0; i < n; ++i) ow->t[i] = tip_t[i];
  std::vector<int> order(n); std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int a, int b) { return tip_t[a] > tip_t[b]; });   // latest first
  int next_inner = n;
  if (p.caterpillar) {
    // ((((t0,t1),t2),t3)...): the deepest pair is the two EARLIEST tips; each new inner node is older.
    std::vector<int> asc(order.rbegin(), order.rend());   // earliest first
    int cur = asc[0];
    double tt = tip_t[asc[0]];
    // build from the tips backwards in time: join latest first so that parents are always older than children
    cur = order[0];
    tt = tip_t[order[0]];
    for (int k = 1; k < n; ++k) {
      int tip = order[k];
      int v = next_inner++;
      double tv = std::min(tt, tip_t[tip]) - (0.002 + 0.02 * rng.uniform());
      ow->t[v] = tv;
      bool flip = rng.uniform() < 0.5;
      ow->child0[v] = flip ? tip : cur; ow->child1[v] = flip ? cur : tip;
      ow->parent[cur] = v; ow->parent[tip] = v;
      cur = v; tt = tv;
    }
  } else {
    std::vector<int> active;
    double g = p.growth_rate, n0 = p.n0_years;
    double tau = 0.0;                      // time before the latest tip
    int next_tip = 0;
    double t_latest = tip_t[order[0]];
    active.push_back(order[next_tip++]);
    while ((int)active.size() > 1 || next_tip < n) {
      double next_sample_tau = next_tip < n ? (t_latest - tip_t[order[next_tip]]) : INFINITY;
      size_t k = active.size();
      if (k < 2) { tau = next_sample_tau; active.push_back(order[next_tip++]); continue; }
      double rate_pairs = 0.5 * (double)k * (double)(k - 1);
      double E = rng.exponential() / rate_pairs;
      // Lambda(tau) = (exp(g tau) - 1) / (g n0); solve Lambda(tau') = Lambda(tau) + E
      double tau_new;
      if (g > 0) {
        double lam = std::expm1(g * tau) / (g * n0);
        tau_new = std::log1p(g * n0 * (lam + E)) / g;
      } else {
        tau_new = tau + E * n0;
      }
      if (tau_new >= next_sample_tau) { tau = next_sample_tau; active.push_back(order[next_tip++]); continue; }
      tau = tau_new;
      size_t i = rng.below(k), j = rng.below(k - 1); if (j >= i) ++j;
      int a = active[i], b = active[j];
      int v = next_inner++;
      double tv = t_latest - tau;
      double tmin_child = std::min(ow->t[a], ow->t[b]);
      if (!(tv < tmin_child - 1e-9)) tv = tmin_child - 1e-9 * (1.0 + rng.uniform());   // strictly positive branches
      ow->t[v] = tv;
      ow->child0[v] = a; ow->child1[v] = b; ow->parent[a] = v; ow->parent[b] = v;
      if (i > j) std::swap(i, j);
      active[j] = active.back(); active.pop_back();
      active[i] = v;
    }
  }
  const int root = N - 1;

  // pre-order (for bottom-up / top-down passes)
  std::vector<int> pre; pre.reserve(N);
  {
    std::vector<int> st{root};
    while (!st.empty()) {
      int v = st.back(); st.pop_back(); pre.push_back(v);
      if (ow->child0[v] >= 0) { st.push_back(ow->child1[v]); st.push_back(ow->child0[v]); }
    }
  }
  std::vector<int> depth(N, 0); int max_depth = 0;
  for (int v : pre) if (v != root) { depth[v] = depth[ow->parent[v]] + 1; max_depth = std::max(max_depth, depth[v]); }
  double T_total = 0;
  for (int v = 0; v < N; ++v) if (v != root) T_total += ow->t[v] - ow->t[ow->parent[v]];

  // ---- missing data: per-tip gaps, factored rootward ---------------------------------------------------------
  std::vector<Ivls> M(N);   // full missing set AT each node
  if (p.missing_mean_intervals_per_tip > 0 || p.end_gaps) {
    double pgeo = 1.0 / (1.0 + p.missing_mean_intervals_per_tip);
    double lmin = std::log(std::max(1.0, p.missing_len_min)), lmax = std::log(std::max(p.missing_len_min + 1, p.missing_len_max));
    for (int i = 0; i < n; ++i) {
      Ivls v;
      if (p.missing_mean_intervals_per_tip > 0) {
        int k = 0; while (rng.uniform() > pgeo && k < 200) ++k;
        for (int j = 0; j < k; ++j) {
          int len = (int)std::exp(lmin + rng.uniform() * (lmax - lmin));
          len = std::max(1, std::min(len, L / 4));
          int s = (int)rng.below((uint64_t)(L - len));
          v.push_back({s, s + len});
        }
      }
      if (p.end_gaps) {
        int g5 = 20 + (int)rng.below(60), g3 = 30 + (int)rng.below(100);
        g5 = std::min(g5, L / 8); g3 = std::min(g3, L / 8);
        v.push_back({0, g5}); v.push_back({L - g3, L});
      }
      M[i] = canonicalize(std::move(v));
    }
    for (int k = N - 1; k >= 0; --k) {   // reverse pre-order: children before parents
      int v = pre[k];
      if (ow->child0[v] >= 0) M[v] = intersect(M[ow->child0[v]], M[ow->child1[v]]);
    }
  }
  std::vector<Ivls> own(N);
  for (int v = 0; v < N; ++v) own[v] = (v == root) ? M[v] : subtract(M[v], M[ow->parent[v]]);

  // ---- mutations: Poisson process top-down with running state ------------------------------------------------
  // site sampling weights w_l = mu_rel(part) * nu_l ; mu chosen so that E[#mutations] ~= muts_per_tip * n
  std::vector<double> mu_rel(P, 1.0); if (P == 2) mu_rel[1] = 5.0;
  std::vector<double> cumw(L + 1, 0.0);
  for (int l = 0; l < L; ++l) cumw[l + 1] = cumw[l] + mu_rel[ow->part[l]] * ow->nu[l];
  double W = cumw[L];
  double mu_base = p.muts_per_tip * (double)n / (W * T_total);   // E[q_a] == 1 under pi
  for (int b = 0; b < P; ++b) ow->mu[b] = mu_base * mu_rel[b];
  double qmax = 0;
  for (int b = 0; b < P; ++b) for (int a = 0; a < 4; ++a) qmax = std::max(qmax, -ow->q[b * 16 + a * 5]);

  std::vector<uint8_t> state(ow->ref);                    // running state along the current root path
  std::vector<std::vector<Mut>> muts(N);
  std::vector<std::vector<std::pair<int32_t, uint8_t>>> fs(N);
  struct Undo { int32_t site; uint8_t prev; };
  std::vector<Undo> undo;                                   // global undo log
  std::vector<size_t> undo_mark(N, 0);
  // iterative DFS with explicit enter/exit
  std::vector<std::pair<int, int>> st; st.push_back({root, 0});
  std::vector<int32_t> touched;                             // scratch
  while (!st.empty()) {
    auto [v, phase] = st.back(); st.pop_back();
    if (phase == 1) {                                       // exit: undo this node's state changes
      while (undo.size() > undo_mark[v]) { state[undo.back().site] = undo.back().prev; undo.pop_back(); }
      continue;
    }
    undo_mark[v] = undo.size();
    // from_states of this node's own missations: sites overridden along the path (state != ref) at branch START
    if (!own[v].empty() && !undo.empty()) {
      touched.clear();
      for (auto& u : undo) touched.push_back(u.site);
      std::sort(touched.begin(), touched.end());
      touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
      for (int32_t l : touched) if (state[l] != ow->ref[l] && contains(own[v], l)) fs[v].push_back({l, state[l]});
    }
    if (v == root) {
      for (int k = 0; k < p.num_root_mutations; ++k) {
        int32_t l = (int32_t)rng.below((uint64_t)L);
        if (contains(M[v], l) || state[l] != ow->ref[l]) continue;
        uint8_t to = (uint8_t)((state[l] + 1 + rng.below(3)) & 3);
        muts[v].push_back({-DBL_MAX, l, state[l], to});
        undo.push_back({l, state[l]}); state[l] = to;
      }
      std::sort(muts[v].begin(), muts[v].end(), [](const Mut& a, const Mut& b) { return a.site < b.site; });
    } else {
      double tP = ow->t[ow->parent[v]], tv = ow->t[v], len = tv - tP;
      int K = rng.poisson(mu_base * W * qmax * len);
      if (K > 0) {
        std::vector<double> times(K);
        for (auto& x : times) { x = tP + rng.uniform_open() * len; if (x > tv) x = tv; if (x < tP) x = tP; }
        std::sort(times.begin(), times.end());
        for (double tm : times) {
          double u = rng.uniform() * W;
          int32_t l = (int32_t)(std::upper_bound(cumw.begin(), cumw.end(), u) - cumw.begin()) - 1;
          l = std::max(0, std::min(L - 1, l));
          if (contains(M[v], l)) continue;                  // missing at-or-above: no mutation allowed
          int b = ow->part[l]; uint8_t a = state[l];
          double qa = -ow->q[b * 16 + a * 5];
          if (rng.uniform() * qmax >= qa) continue;         // thinning
          double r = rng.uniform() * qa, c = 0; uint8_t to = a;
          for (int x = 0; x < 4; ++x) { if (x == a) continue; c += ow->q[b * 16 + a * 4 + x]; to = (uint8_t)x; if (r < c) break; }
          muts[v].push_back({tm, l, a, to});
          undo.push_back({l, a}); state[l] = to;
        }
        // ties in t are broken by site in the reference's ordering (core/mutations.h:43-45); times are a.s. distinct
        std::stable_sort(muts[v].begin(), muts[v].end(), [](const Mut& a, const Mut& b) {
          return a.t < b.t || (a.t == b.t && a.site < b.site); });
      }
    }
    st.push_back({v, 1});
    if (ow->child0[v] >= 0) { st.push_back({ow->child1[v], 0}); st.push_back({ow->child0[v], 0}); }
  }

  // ---- flatten ---------------------------------------------------------------------------------------------------
  ow->mut_off.assign(N + 1, 0); ow->miss_off.assign(N + 1, 0); ow->fs_off.assign(N + 1, 0);
  int64_t n_missing_sites = 0;
  for (int v = 0; v < N; ++v) {
    for (auto& m : muts[v]) { ow->mut_site.push_back(m.site); ow->mut_from.push_back(m.from); ow->mut_to.push_back(m.to); ow->mut_t.push_back(m.t); }
    ow->mut_off[v + 1] = (int32_t)ow->mut_site.size();
    for (auto& iv : own[v]) { ow->miss_start.push_back(iv.first); ow->miss_end.push_back(iv.second); }
    ow->miss_off[v + 1] = (int32_t)ow->miss_start.size();
    for (auto& f : fs[v]) { ow->fs_site.push_back(f.first); ow->fs_from.push_back(f.second); }
    ow->fs_off[v + 1] = (int32_t)ow->fs_site.size();
  }
  for (int i = 0; i < n; ++i) for (auto& iv : M[i]) n_missing_sites += iv.second - iv.first;
  // keep data() pointers valid for empty vectors
  auto nz = [](auto& v) { if (v.empty()) v.reserve(1); };
  nz(ow->mut_site); nz(ow->mut_from); nz(ow->mut_to); nz(ow->mut_t); nz(ow->miss_start); nz(ow->miss_end); nz(ow->fs_site); nz(ow->fs_from);

  auto& e = res->emat;
  e.num_nodes = N; e.root = root; e.includes_run_root = 1;
  e.parent = ow->parent.data(); e.child0 = ow->child0.data(); e.child1 = ow->child1.data(); e.t = ow->t.data();
  e.mut_off = ow->mut_off.data(); e.mut_site = ow->mut_site.data(); e.mut_from = ow->mut_from.data();
  e.mut_to = ow->mut_to.data(); e.mut_t = ow->mut_t.data();
  e.miss_off = ow->miss_off.data(); e.miss_start = ow->miss_start.data(); e.miss_end = ow->miss_end.data();
  e.fs_off = ow->fs_off.data(); e.fs_site = ow->fs_site.data(); e.fs_from = ow->fs_from.data();
  auto& s = res->sites;
  s.num_sites = L; s.num_partitions = P; s.ref = ow->ref.data(); s.partition_for_site = ow->part.data();
  s.nu_l = ow->nu.data(); s.mu = ow->mu.data(); s.pi_a = ow->pi.data(); s.q_ab = ow->q.data();
  res->mu_used = mu_base;
  res->t_max_tip = *std::max_element(tip_t.begin(), tip_t.end());
  res->num_mutations = (int64_t)ow->mut_site.size();
  res->num_intervals = (int64_t)ow->miss_start.size();
  res->num_from_states = (int64_t)ow->fs_site.size();
  res->num_missing_sites = n_missing_sites;
  res->max_depth = max_depth;
  *out = res;
  return DPHY_OK;
}

extern "C" void dphy_synth_free(dphy_synth_emat* s) {
  if (!s) return;
  delete static_cast<Owner*>(s->owner_);
  delete s;
}
