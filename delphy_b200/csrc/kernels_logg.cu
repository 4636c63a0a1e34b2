// kernels_logg.cu -- single-pass EMAT log-G evaluation for a whole forest (sm_100a).
//
// Replaces, for every tree of the forest in ONE launch:
//   calc_lambda_i                          core/phylo_tree_calc.cpp:420-436  (+ phylo_tree_calc.h:121-155)
//   calc_num_sites_missing_at_every_node   core/phylo_tree_calc.cpp:67-76
//   calc_log_root_prior                    core/phylo_tree_calc.cpp:467-504
//   calc_log_G_below_root                  core/phylo_tree_calc.cpp:515-543  (+ phylo_tree_calc.h:185-206)
//   calc_num_muts / calc_num_muts_ab / calc_T   core/phylo_tree_calc.cpp:577-597, :120-128
//
// Formulation.  The reference computes lambda_i with a pre-order walk (lambda_child = lambda_parent + delta of
// the branch) and then sums branch terms in node-index order.  Here nodes are stored in DFS pre-order, so the set
// of ancestors of position q is "everything opened at or before q and not yet closed".  With
//     diff[q] = delta[q] - sum_{a : subtree of a ends right before q} delta[a]
// lambda[q] = lambda_ref + inclusive_prefix_sum(diff)[q]: a plain scan.  The nodes closing at q are a contiguous
// slice of the tree's post-order list: post_node[c(q-1) .. c(q)) with c(q) = q - depth[q].  Each CTA owns a tile of
// kTile consecutive positions of one tree; it (1) computes delta / the mutation part of log G / missing-site counts
// for its own nodes straight from the CSR lists, (2) forms diff (re-deriving delta for the few nodes that opened in
// an earlier tile and close in this one), (3) block-scans and writes the tile-local lambda / nsmn (device order) plus
// per-tile partial sums.  No CTA ever waits for another one: a second, tiny kernel (one CTA per tree) scans the tile
// aggregates in tile order, folds log G = sum_tiles (A1 - prefix * A2) and evaluates the root prior, and a third
// streaming kernel adds each tile's prefix to lambda_i / nsmn in place.  Fixed reduction shapes everywhere =>
// bit-reproducible results.  Every input byte is read once (plus the re-derived closers).
#include "dphy_internal.h"
#include "device_utils.cuh"

#include <math_constants.h>
#include <cstdlib>

namespace dphy {

struct LogGParams {
  ForestDev f;
  double* lambda_out;    // [num_nodes] device order
  int32_t* nsmn_out;     // [num_nodes] device order
  double* tile_agg;      // [num_tiles] sum of diff over the tile; overwritten by the tile's exclusive prefix in pass 2
  int32_t* tile_iagg;
  double* tile_part;     // [num_tiles * 2]  (A1 = sum[-(lambda_ref+incl) len + g], A2 = sum len)
  int32_t* tile_ipart;   // [num_tiles * 17] (num_muts, num_muts_ab[16])
  double* tree_out;      // [num_trees * 4]: log_root_prior, log_G_below_root, T, lambda_root
  int32_t* tree_iout;    // [num_trees * 20]: num_muts, 0, num_muts_ab[16], 0, 0
  int32_t debug_mask;    // profiling only (DPHY_DEBUG_MASK): 1 skip mutations, 2 intervals, 4 from-states, 8 closers, 16 foreign closers
};

// Capacity (events per chunk) of the flat event buffers in shared memory.
constexpr int kEvCap = 1536;

// Sequential re-derivation of delta-lambda / missing-site count of ONE branch (used for the few "foreign closers":
// nodes that opened in an earlier tile and close inside this one).  phylo_tree_calc.h:121-155.
__device__ void branch_delta_seq(const ForestDev& f, const SitesDev& S, const double* __restrict__ s_dq,
                                 const double* __restrict__ smu, int p, double& delta, int& nmiss) {
  const bool uni = S.nu_uniform != 0;
  double dm = 0.0;
  for (int i = f.mut_off[p]; i < f.mut_off[p + 1]; ++i) {
    const int code = __ldg(f.mut_code + i);
    const double mn = uni ? smu[code >> 4] : __ldg(S.munu + __ldg(f.mut_site + i));
    dm += mn * s_dq[code];
  }
  double dmi = 0.0;
  int nm = 0;
  for (int i = f.miss_off[p]; i < f.miss_off[p + 1]; ++i) {
    const int2 se = __ldg(f.miss_se + i);
    dmi -= __ldg(S.cumQ + se.y) - __ldg(S.cumQ + se.x);
    nm += se.y - se.x;
  }
  for (int i = f.fs_off[p]; i < f.fs_off[p + 1]; ++i) {
    const int code = __ldg(f.fs_code + i);
    const double mn = uni ? smu[code >> 4] : __ldg(S.munu + __ldg(f.fs_site + i));
    dmi -= mn * s_dq[code];
  }
  delta = dm + dmi;
  nmiss = nm;
}

// Flat-over-events + segmented-sum-per-node helper.  All threads of the CTA call it with the same [r0, r1).
//   ev(i)          : thread-parallel over groups of kVec consecutive events starting at the kVec-aligned global index
//                    i (vector loads); writes the contributions of events i..i+kVec-1 to smem slots (i - c0)..
//                    Slots of events outside [r0, r1) hold garbage that is never read.
//   acc(k, slot)   : the thread owning node (tid + k*kLgThreads) folds the slots of that node's CSR slice in list
//                    order (deterministic)
template <int kVec, typename EventFn, typename AccFn>
__device__ __forceinline__ void flat_segmented(int r0, int r1, const int* __restrict__ s_off, int n_act, EventFn ev, AccFn acc) {
  const int tid = threadIdx.x;
  for (int c0 = r0 & ~(kVec - 1); c0 < r1; c0 += kEvCap) {
    const int c1 = min(c0 + kEvCap, r1);
    for (int i = c0 + kVec * tid; i < c1; i += kVec * kLgThreads) ev(i, i - c0);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kLgNPT; ++k) {
      const int q = tid + k * kLgThreads;
      if (q < n_act) {
        const int lo = max(s_off[q], c0), hi = min(s_off[q + 1], c1);
        for (int i = lo; i < hi; ++i) acc(k, i - c0);
      }
    }
    __syncthreads();
  }
}

// Tile kernel: kLgTile consecutive device positions of one tree per CTA, kLgNPT nodes per thread (node q = tid +
// k*kLgThreads, so every per-node global access is coalesced).
__global__ void __launch_bounds__(kLgThreads) emat_log_G_kernel(const LogGParams P) {
  __shared__ double s_dq[kMaxPartitions * 16];    // [part<<4 | x<<2 | y] = q_a(y) - q_a(x)
  __shared__ double s_lq[kMaxPartitions * 16];    // [part<<4 | from<<2 | to] = log(mu nu_const q_from,to)   (uniform nu)
  __shared__ double s_qft[kMaxPartitions * 16];   // q_ab
  __shared__ double s_md[kMaxPartitions * 16];    // mu nu_const * s_dq (uniform nu)
  __shared__ double s_mu[kMaxPartitions];
  __shared__ __align__(16) double s_bufA[kEvCap];
  __shared__ __align__(16) double s_bufB[kEvCap];
  __shared__ int s_off_m[kLgTile + 1];
  __shared__ int s_off_i[kLgTile + 1];
  __shared__ int s_off_f[kLgTile + 1];
  __shared__ double s_delta[kLgTile];             // delta per node, later diff / inclusive scan per node
  __shared__ int s_nmiss[kLgTile];
  __shared__ double s_wsd[kLgThreads / 32 * 3];
  __shared__ int s_wsi[kLgThreads / 32];
  __shared__ int s_ab[16];

  const ForestDev& f = P.f;
  const int tid = threadIdx.x;

  const int tile = blockIdx.x;
  const int tree = f.ctile_tree[tile];
  const TreeDev T = f.trees[tree];
  const SitesDev& S = f.sites[T.sites_id];
  const bool uni = S.nu_uniform != 0;
  if (tid < 16) s_ab[tid] = 0;
  if (tid < S.P) s_mu[tid] = S.mu[tid] * S.nu_const;
  if (tid < S.P * 16) {
    const int pt = tid >> 4, x = (tid >> 2) & 3, y = tid & 3;
    const double qxy = S.q[tid];
    s_qft[tid] = qxy;
    const double dq = (-S.q[pt * 16 + y * 5]) - (-S.q[pt * 16 + x * 5]);
    s_dq[tid] = dq;
    s_md[tid] = S.mu[pt] * S.nu_const * dq;
    s_lq[tid] = (x != y) ? log(S.mu[pt] * S.nu_const * qxy) : 0.0;
  } else if (tid < kMaxPartitions * 16) {
    s_dq[tid] = 0.0; s_md[tid] = 0.0; s_lq[tid] = 0.0; s_qft[tid] = 1.0;
  }

  const int tile_in_tree = tile - T.first_ctile;
  const int tile_start = T.node_base + tile_in_tree * kLgTile;              // global device position
  const int tile_end = min(tile_start + kLgTile, T.node_base + T.num_nodes);
  const int n_act = tile_end - tile_start;
  const bool has_root = tile_in_tree == 0;                                   // position 0 of a tree is its root

  // ---- (0) node records + CSR offsets of the tile ---------------------------------------------------------------------
  double tP[kLgNPT], tN[kLgNPT];
  int par[kLgNPT], dep[kLgNPT];
#pragma unroll
  for (int k = 0; k < kLgNPT; ++k) {
    const int q = tid + k * kLgThreads, p = tile_start + q;
    par[k] = -1; dep[k] = 0; tP[k] = 0.0; tN[k] = 0.0;
    if (q < n_act) {
      par[k] = f.parent_pos[p];
      dep[k] = f.depth[p];
      tN[k] = f.t[p];
      s_off_m[q] = f.mut_off[p]; s_off_i[q] = f.miss_off[p]; s_off_f[q] = f.fs_off[p];
      if (q == n_act - 1) { s_off_m[n_act] = f.mut_off[p + 1]; s_off_i[n_act] = f.miss_off[p + 1]; s_off_f[n_act] = f.fs_off[p + 1]; }
    }
  }
#pragma unroll
  for (int k = 0; k < kLgNPT; ++k) if (par[k] >= 0) tP[k] = f.t[par[k]];
  __syncthreads();

  // ---- (1) branch terms: flat over the tile's events, segmented sum per node --------------------------------------------
  double dm[kLgNPT], esum[kLgNPT], dmi[kLgNPT];
  int nmiss[kLgNPT];
#pragma unroll
  for (int k = 0; k < kLgNPT; ++k) { dm[k] = 0.0; esum[k] = 0.0; dmi[k] = 0.0; nmiss[k] = 0; }
  const int root_m1 = has_root ? s_off_m[1] : s_off_m[0];   // the root's list ("mutations" above the root) ends here
  // mutations: dm_i = mu nu (q_to - q_from);  e_i = dm_i * t_i + log(mu nu q_from,to)   [g_node = sum e_i - t_P * dm]
  {
    const int m_lo = max(root_m1, s_off_m[0]), m_hi = s_off_m[n_act];
    if (!(P.debug_mask & 1)) flat_segmented<4>(s_off_m[0], s_off_m[n_act], s_off_m, n_act,
      [&](int i, int slot) {
        const uchar4 c4 = __ldg(reinterpret_cast<const uchar4*>(f.mut_code + i));
        const double2 ta = __ldg(reinterpret_cast<const double2*>(f.mut_t + i));
        const double2 tb = __ldg(reinterpret_cast<const double2*>(f.mut_t + i + 2));
        const int code[4] = {c4.x & 63, c4.y & 63, c4.z & 63, c4.w & 63};
        const double tt[4] = {ta.x, ta.y, tb.x, tb.y};
        double d[4], e[4];
        if (uni) {
#pragma unroll
          for (int u = 0; u < 4; ++u) { d[u] = s_md[code[u]]; e[u] = d[u] * tt[u] + s_lq[code[u]]; }
        } else {
          const int4 l4 = __ldg(reinterpret_cast<const int4*>(f.mut_site + i));
          const int ll[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool ok = i + u >= s_off_m[0] && i + u < m_hi;
            const double mn = ok ? __ldg(S.munu + ll[u]) : 1.0;
            d[u] = mn * s_dq[code[u]];
            e[u] = d[u] * tt[u] + ((code[u] >> 2 & 3) != (code[u] & 3) ? log(mn * s_qft[code[u]]) : 0.0);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (i + u >= m_lo && i + u < m_hi) atomicAdd(&s_ab[code[u] & 15], 1); else e[u] = 0.0;
        }
        *reinterpret_cast<double2*>(s_bufA + slot) = make_double2(d[0], d[1]);
        *reinterpret_cast<double2*>(s_bufA + slot + 2) = make_double2(d[2], d[3]);
        *reinterpret_cast<double2*>(s_bufB + slot) = make_double2(e[0], e[1]);
        *reinterpret_cast<double2*>(s_bufB + slot + 2) = make_double2(e[2], e[3]);
      },
      [&](int k, int slot) { dm[k] += s_bufA[slot]; esum[k] += s_bufB[slot]; });
  }
  // missation intervals: -(cumQ[end] - cumQ[start]); count of sites going missing
  {
    const int i_lo = s_off_i[0], i_hi = s_off_i[n_act];
    if (!(P.debug_mask & 2)) flat_segmented<2>(i_lo, i_hi, s_off_i, n_act,
      [&](int i, int slot) {
        const int4 se = __ldg(reinterpret_cast<const int4*>(f.miss_se + i));
        const bool ok0 = i >= i_lo, ok1 = i + 1 < i_hi;
        const double a0 = ok0 ? __ldg(S.cumQ + se.y) - __ldg(S.cumQ + se.x) : 0.0;
        const double a1 = ok1 ? __ldg(S.cumQ + se.w) - __ldg(S.cumQ + se.z) : 0.0;
        *reinterpret_cast<double2*>(s_bufA + slot) = make_double2(a0, a1);
        *reinterpret_cast<int2*>(reinterpret_cast<int*>(s_bufB) + slot) = make_int2(se.y - se.x, se.w - se.z);
      },
      [&](int k, int slot) { dmi[k] -= s_bufA[slot]; nmiss[k] += reinterpret_cast<int*>(s_bufB)[slot]; });
  }
  // from-state overrides of missing sites
  {
    const int f_lo = s_off_f[0], f_hi = s_off_f[n_act];
    if (!(P.debug_mask & 4)) flat_segmented<4>(f_lo, f_hi, s_off_f, n_act,
      [&](int i, int slot) {
        const uchar4 c4 = __ldg(reinterpret_cast<const uchar4*>(f.fs_code + i));
        const int code[4] = {c4.x & 63, c4.y & 63, c4.z & 63, c4.w & 63};
        double d[4];
        if (uni) {
#pragma unroll
          for (int u = 0; u < 4; ++u) d[u] = s_md[code[u]];
        } else {
          const int4 l4 = __ldg(reinterpret_cast<const int4*>(f.fs_site + i));
          const int ll[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool ok = i + u >= f_lo && i + u < f_hi;
            d[u] = (ok ? __ldg(S.munu + ll[u]) : 1.0) * s_dq[code[u]];
          }
        }
        *reinterpret_cast<double2*>(s_bufA + slot) = make_double2(d[0], d[1]);
        *reinterpret_cast<double2*>(s_bufA + slot + 2) = make_double2(d[2], d[3]);
      },
      [&](int k, int slot) { dmi[k] -= s_bufA[slot]; });
  }
  double g[kLgNPT];
#pragma unroll
  for (int k = 0; k < kLgNPT; ++k) {
    const int q = tid + k * kLgThreads;
    g[k] = par[k] >= 0 ? esum[k] - tP[k] * dm[k] : 0.0;
    if (q < kLgTile) { s_delta[q] = q < n_act ? dm[k] + dmi[k] : 0.0; s_nmiss[q] = q < n_act ? nmiss[k] : 0; }
  }
  __syncthreads();

  // ---- (2) diff = own delta - deltas of the nodes whose subtree closes right before this position ---------------------------
  // The tile's closers are one contiguous slice of the tree's post-order list; their deltas are gathered in parallel
  // (in-tile from smem, foreign ones re-derived), then each position folds its own sub-slice in order.
  double diff[kLgNPT];
  int idiff[kLgNPT];
  {
    const int q_first = tile_start - T.node_base, q_last = tile_end - 1 - T.node_base;
    const int cl0 = q_first == 0 ? 0 : (q_first - 1) - f.depth[tile_start - 1];
    const int cl1 = (P.debug_mask & 8) ? cl0 : q_last - f.depth[tile_end - 1];
    int my0[kLgNPT], my1[kLgNPT];
#pragma unroll
    for (int k = 0; k < kLgNPT; ++k) {
      const int q = tid + k * kLgThreads, p = tile_start + q;
      my0[k] = 0; my1[k] = 0;
      diff[k] = q < n_act ? s_delta[q] : 0.0;
      idiff[k] = q < n_act ? s_nmiss[q] : 0;
      if (q < n_act) {
        const int tq = p - T.node_base;
        if (tq > 0) { my0[k] = (tq - 1) - f.depth[p - 1]; my1[k] = tq - dep[k]; }
      }
    }
    int* s_cn = reinterpret_cast<int*>(s_bufB);
    for (int c0 = cl0; c0 < cl1; c0 += kEvCap) {
      const int c1 = min(c0 + kEvCap, cl1);
      for (int j = c0 + tid; j < c1; j += kLgThreads) {
        const int a = f.post_node[T.node_base + j];
        double da; int na;
        if (a >= tile_start) { da = s_delta[a - tile_start]; na = s_nmiss[a - tile_start]; }
        else if (P.debug_mask & 16) { da = 0.0; na = 0; }
        else branch_delta_seq(f, S, s_dq, s_mu, a, da, na);
        s_bufA[j - c0] = da; s_cn[j - c0] = na;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kLgNPT; ++k) {
        const int lo = max(my0[k], c0), hi = min(my1[k], c1);
        for (int j = lo; j < hi; ++j) { diff[k] -= s_bufA[j - c0]; idiff[k] -= s_cn[j - c0]; }
      }
      __syncthreads();
    }
  }

  // ---- (3) inclusive scan of diff over the tile (position order): smem transpose -> blocked serial + block scan ------------
#pragma unroll
  for (int k = 0; k < kLgNPT; ++k) {
    const int q = tid + k * kLgThreads;
    s_delta[q] = diff[k]; s_nmiss[q] = idiff[k];
  }
  __syncthreads();
  double tot; int itot;
  {
    double loc[kLgNPT]; int iloc[kLgNPT];
    double run = 0.0; int irun = 0;
#pragma unroll
    for (int k = 0; k < kLgNPT; ++k) {
      run += s_delta[tid * kLgNPT + k]; irun += s_nmiss[tid * kLgNPT + k];
      loc[k] = run; iloc[k] = irun;
    }
    const double incl = block_scan_incl<double, kLgThreads>(run, s_wsd, &tot);
    const int iincl = block_scan_incl<int, kLgThreads>(irun, s_wsi, &itot);
    const double base = incl - run; const int ibase = iincl - irun;
#pragma unroll
    for (int k = 0; k < kLgNPT; ++k) { s_delta[tid * kLgNPT + k] = base + loc[k]; s_nmiss[tid * kLgNPT + k] = ibase + iloc[k]; }
  }
  __syncthreads();

  // ---- (4) tile-local lambda_i / nsmn (device order); per-tile partial sums ----------------------------------------------------
  const double lambda_ref = S.cumQ[S.L];
  double contrib = 0.0, tcontrib = 0.0;
  int nmut = 0;
#pragma unroll
  for (int k = 0; k < kLgNPT; ++k) {
    const int q = tid + k * kLgThreads, p = tile_start + q;
    if (q < n_act) {
      const double lam_local = lambda_ref + s_delta[q];     // + the tile's prefix, added in pass 3
      P.lambda_out[p] = lam_local;
      P.nsmn_out[p] = s_nmiss[q];
      if (par[k] >= 0) {
        const double len = tN[k] - tP[k];
        contrib += -lam_local * len + g[k];
        tcontrib += len;
        nmut += s_off_m[q + 1] - s_off_m[q];
      }
    }
  }
  // three sums with one barrier pair
  {
    const int lane = tid & 31, warp = tid >> 5;
    contrib = warp_sum(contrib); tcontrib = warp_sum(tcontrib); nmut = warp_sum(nmut);
    __syncthreads();
    if (lane == 0) { s_wsd[warp * 3 + 0] = contrib; s_wsd[warp * 3 + 1] = tcontrib; s_wsd[warp * 3 + 2] = (double)nmut; }
    __syncthreads();
    if (warp == 0) {
      double a = lane < kLgThreads / 32 ? s_wsd[lane * 3 + 0] : 0.0;
      double b = lane < kLgThreads / 32 ? s_wsd[lane * 3 + 1] : 0.0;
      double c = lane < kLgThreads / 32 ? s_wsd[lane * 3 + 2] : 0.0;
      a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
      if (lane == 0) {
        P.tile_agg[tile] = tot;
        P.tile_iagg[tile] = itot;
        P.tile_part[tile * 2 + 0] = a;
        P.tile_part[tile * 2 + 1] = b;
        P.tile_ipart[tile * 17 + 0] = (int)c;
      }
    }
  }
  if (tid < 16) P.tile_ipart[tile * 17 + 1 + tid] = s_ab[tid];
}

// ---- pass 2: one CTA per tree -- exclusive scan of the tile aggregates (tile order), log G fold, tallies, root prior -----
__global__ void __launch_bounds__(kTile) emat_log_G_tree_kernel(const LogGParams P) {
  __shared__ double s_wsd[kTile / 32];
  __shared__ int s_wsi[kTile / 32];
  __shared__ int s_cnt[kMaxPartitions * 4];
  __shared__ double s_carry;
  __shared__ int s_icarry;
  const ForestDev& f = P.f;
  const int tid = threadIdx.x;
  const int tree = blockIdx.x;
  const TreeDev T = f.trees[tree];
  const SitesDev& S = f.sites[T.sites_id];
  if (tid == 0) { s_carry = 0.0; s_icarry = 0; }
  __syncthreads();
  double a1 = 0.0, a2 = 0.0; int m = 0;
  int ab[16];
#pragma unroll
  for (int b = 0; b < 16; ++b) ab[b] = 0;
  for (int j0 = 0; j0 < T.num_ctiles; j0 += kTile) {
    const int j = T.first_ctile + j0 + tid;
    const bool ok = j0 + tid < T.num_ctiles;
    const double v = ok ? P.tile_agg[j] : 0.0;
    const int iv = ok ? P.tile_iagg[j] : 0;
    double tot; int itot;
    const double incl = block_scan_incl<double, kTile>(v, s_wsd, &tot);
    const int iincl = block_scan_incl<int, kTile>(iv, s_wsi, &itot);
    if (ok) {
      const double pre = s_carry + (incl - v);          // exclusive prefix of this tile
      const int ipre = s_icarry + (iincl - iv);
      P.tile_agg[j] = pre;
      P.tile_iagg[j] = ipre;
      const double A1 = P.tile_part[j * 2 + 0], A2 = P.tile_part[j * 2 + 1];
      a1 += A1 - pre * A2;                              // sum over the tile of -(lambda_local + pre) len + g
      a2 += A2;
      m += P.tile_ipart[j * 17];
#pragma unroll
      for (int b = 0; b < 16; ++b) ab[b] += P.tile_ipart[j * 17 + 1 + b];
    }
    __syncthreads();
    if (tid == 0) { s_carry += tot; s_icarry += itot; }
    __syncthreads();
  }
  a1 = block_sum<double, kTile>(a1, s_wsd);
  a2 = block_sum<double, kTile>(a2, s_wsd);
  m = block_sum<int, kTile>(m, s_wsi);
  if (tid == 0) {
    P.tree_out[tree * 4 + 1] = a1;
    P.tree_out[tree * 4 + 2] = a2;
    P.tree_out[tree * 4 + 3] = S.cumQ[S.L] + 0.0;   // (lambda at the root is lambda_out[node_base] after pass 3)
    P.tree_iout[tree * 20 + 0] = m;
    P.tree_iout[tree * 20 + 1] = 0;
  }
#pragma unroll
  for (int b = 0; b < 16; ++b) {
    const int c = block_sum<int, kTile>(ab[b], s_wsi);
    if (tid == 0) P.tree_iout[tree * 20 + 2 + b] = c;
  }
  // root prior (core/phylo_tree_calc.cpp:467-504): reference-sequence state counts per partition, adjusted by the
  // root's "mutations", missing sites and from-state overrides.
  __syncthreads();
  if (tid < kMaxPartitions * 4) s_cnt[tid] = tid < S.P * 4 ? S.ref_freq[tid] : 0;
  __syncthreads();
  {
    const int r = T.node_base;   // the root is the first position of its tree
    for (int i = f.mut_off[r] + tid; i < f.mut_off[r + 1]; i += kTile) {
      const int code = f.mut_code[i]; const int pt = code >> 4;
      atomicSub(&s_cnt[pt * 4 + ((code >> 2) & 3)], 1);
      atomicAdd(&s_cnt[pt * 4 + (code & 3)], 1);
    }
    for (int i = f.miss_off[r]; i < f.miss_off[r + 1]; ++i) {
      const int2 se = f.miss_se[i]; const int s = se.x, e = se.y;
      for (int l = s + tid; l < e; l += kTile) atomicSub(&s_cnt[S.part[l] * 4 + S.ref[l]], 1);
    }
    for (int i = f.fs_off[r] + tid; i < f.fs_off[r + 1]; i += kTile) {
      const int code = f.fs_code[i]; const int pt = code >> 4;
      atomicAdd(&s_cnt[pt * 4 + ((code >> 2) & 3)], 1);
      atomicSub(&s_cnt[pt * 4 + (code & 3)], 1);
    }
  }
  __syncthreads();
  if (tid == 0) {
    double lp = 0.0;
    bool impossible = false;
    for (int b = 0; b < S.P; ++b) {
      for (int a = 0; a < 4; ++a) {
        const double pi = S.pi[b * 4 + a];
        const int c = s_cnt[b * 4 + a];
        if (pi != 0.0) lp += c * log(pi);
        else if (c != 0) impossible = true;
      }
    }
    P.tree_out[tree * 4 + 0] = impossible ? -CUDART_INF : lp;
  }
}

// ---- pass 3: add each tile's prefix to lambda_i / nsmn in place (streaming) ------------------------------------------------------
__global__ void __launch_bounds__(kLgThreads) emat_log_G_finish_kernel(const LogGParams P) {
  const ForestDev& f = P.f;
  const int tile = blockIdx.x;
  const TreeDev T = f.trees[f.ctile_tree[tile]];
  const double pre = P.tile_agg[tile];
  const int ipre = P.tile_iagg[tile];
  const int p0 = T.node_base + (tile - T.first_ctile) * kLgTile;
#pragma unroll
  for (int k = 0; k < kLgNPT; ++k) {
    const int p = p0 + threadIdx.x + k * kLgThreads;
    if (p < T.node_base + T.num_nodes) {
      P.lambda_out[p] += pre;
      P.nsmn_out[p] += ipre;
    }
  }
}

// host-order gather for the getters: out[id] = src[pos_of_node[id]]
template <typename V>
__global__ void gather_host_order_kernel(const int32_t* __restrict__ pos_of_node, const V* __restrict__ src, V* __restrict__ dst,
                                         int node_base, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[node_base + pos_of_node[node_base + i]];
}

int gather_lambda_host_order(dphy_ctx* ctx, dphy_forest* fo, int tree, double* d_dst) {
  const TreeDev& T = fo->trees[tree];
  gather_host_order_kernel<double><<<(T.num_nodes + 255) / 256, 256, 0, ctx->stream>>>(fo->h.pos_of_node, fo->d_lambda, d_dst, T.node_base, T.num_nodes);
  ctx->launches += 1;
  return check_cuda(ctx, cudaGetLastError(), "gather_host_order_kernel<double>");
}
int gather_nsmn_host_order(dphy_ctx* ctx, dphy_forest* fo, int tree, int32_t* d_dst) {
  const TreeDev& T = fo->trees[tree];
  gather_host_order_kernel<int32_t><<<(T.num_nodes + 255) / 256, 256, 0, ctx->stream>>>(fo->h.pos_of_node, fo->d_nsmn, d_dst, T.node_base, T.num_nodes);
  ctx->launches += 1;
  return check_cuda(ctx, cudaGetLastError(), "gather_host_order_kernel<int>");
}

int launch_log_G(dphy_ctx* ctx, dphy_forest* fo) {
  if (fo->h.num_tiles == 0) return DPHY_OK;
  int st = refresh_sites(ctx, fo);
  if (st != DPHY_OK) return st;
  LogGParams P;
  P.f = fo->h;
  P.lambda_out = fo->d_lambda;
  P.nsmn_out = fo->d_nsmn;
  P.tile_agg = fo->d_tile_agg;
  P.tile_iagg = fo->d_tile_iagg;
  P.tile_part = fo->d_tile_part;
  P.tile_ipart = fo->d_tile_ipart;
  P.tree_out = fo->d_tree_out;
  P.tree_iout = fo->d_tree_iout;
  { const char* dm = getenv("DPHY_DEBUG_MASK"); P.debug_mask = dm ? atoi(dm) : 0; }
  emat_log_G_kernel<<<fo->h.num_ctiles, kLgThreads, 0, ctx->stream>>>(P);
  emat_log_G_tree_kernel<<<fo->h.num_trees, kTile, 0, ctx->stream>>>(P);
  emat_log_G_finish_kernel<<<fo->h.num_ctiles, kLgThreads, 0, ctx->stream>>>(P);
  ctx->launches += 3;
  return check_cuda(ctx, cudaGetLastError(), "emat_log_G kernels launch");
}

}  // namespace dphy
