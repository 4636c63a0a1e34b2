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
// an earlier tile and close in this one), (3) block-scans, publishes its tile aggregate and sums the aggregates of
// all earlier tiles of the same tree (fixed order => bit-reproducible), (4) writes lambda_i / nsmn and reduces
// -lambda*(t - t_parent) + mutation terms.  The last CTA to finish a tree folds the per-tile partials in tile
// order and evaluates the root prior.  Every input byte is read once (plus the re-derived closers).
#include "dphy_internal.h"
#include "device_utils.cuh"

#include <math_constants.h>

namespace dphy {

struct LogGParams {
  ForestDev f;
  double* lambda_out;
  int32_t* nsmn_out;
  double* tile_agg;
  int32_t* tile_iagg;
  uint32_t* tile_flag;
  double* tile_part;     // [num_tiles * 2]  (log G partial, T partial)
  int32_t* tile_ipart;   // [num_tiles * 17] (num_muts, num_muts_ab[16])
  uint32_t* tree_done;   // [num_trees]
  uint32_t* ticket;      // [2]: tile ticket, tiles done
  double* tree_out;      // [num_trees * 4]: log_root_prior, log_G_below_root, T, lambda_root
  int32_t* tree_iout;    // [num_trees * 20]: num_muts, 0, num_muts_ab[16], 0, 0
  uint32_t epoch;
};

// delta lambda across the branch ending at device position p (phylo_tree_calc.h:140-155), the number of sites that
// go missing on it, and (optionally) the mutation part of calc_branch_log_G (phylo_tree_calc.h:196-203).
template <bool kWantG>
__device__ __forceinline__ void branch_terms(const ForestDev& f, const SitesDev& S, const double* __restrict__ sq,
                                             int p, double tP, double& delta, int& nmiss, double& g,
                                             int* __restrict__ s_ab) {
  double dm = 0.0;
  g = 0.0;
  const int m0 = f.mut_off[p], m1 = f.mut_off[p + 1];
  for (int i = m0; i < m1; ++i) {
    const int l = __ldg(f.mut_site + i);
    const int ft = __ldg(f.mut_ft + i);
    const int from = ft >> 2, to = ft & 3;
    const int pt = __ldg(S.part + l);
    const double mn = __ldg(S.munu + l);
    const double qf = -sq[pt * 16 + from * 5], qt = -sq[pt * 16 + to * 5];
    dm += mn * (qt - qf);
    if (kWantG) {
      g -= mn * (qf - qt) * (__ldg(f.mut_t + i) - tP);
      g += log(mn * sq[pt * 16 + from * 4 + to]);
      atomicAdd(&s_ab[ft], 1);
    }
  }
  double dmi = 0.0;
  int nm = 0;
  const int i0 = f.miss_off[p], i1 = f.miss_off[p + 1];
  for (int i = i0; i < i1; ++i) {
    const int s = __ldg(f.miss_start + i), e = __ldg(f.miss_end + i);
    dmi -= __ldg(S.cumQ + e) - __ldg(S.cumQ + s);
    nm += e - s;
  }
  const int f0 = f.fs_off[p], f1 = f.fs_off[p + 1];
  for (int i = f0; i < f1; ++i) {
    const int l = __ldg(f.fs_site + i);
    const int from = __ldg(f.fs_from + i), rf = __ldg(S.ref + l);
    const int pt = __ldg(S.part + l);
    dmi -= __ldg(S.munu + l) * ((-sq[pt * 16 + from * 5]) - (-sq[pt * 16 + rf * 5]));
  }
  delta = dm + dmi;
  nmiss = nm;
}

__global__ void __launch_bounds__(kTile) emat_log_G_kernel(const LogGParams P) {
  __shared__ double s_q[kMaxPartitions * 16];
  __shared__ double s_delta[kTile];
  __shared__ int s_nmiss[kTile];
  __shared__ double s_wsd[kTile / 32];
  __shared__ int s_wsi[kTile / 32];
  __shared__ int s_ab[16];
  __shared__ int s_cnt[kMaxPartitions * 4];
  __shared__ double s_prefix;
  __shared__ int s_iprefix;
  __shared__ int s_tile;
  __shared__ int s_is_last;

  const ForestDev& f = P.f;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) s_tile = (int)atomicAdd(P.ticket, 1u);
  if (tid < 16) s_ab[tid] = 0;
  __syncthreads();
  const int tile = s_tile;
  const int tree = f.tile_tree[tile];
  const TreeDev T = f.trees[tree];
  const SitesDev& S = f.sites[T.sites_id];
  if (tid < S.P * 16) s_q[tid] = S.q[tid];
  __syncthreads();

  const int tile_in_tree = tile - T.first_tile;
  const int tile_start = T.node_base + tile_in_tree * kTile;              // global device position
  const int tile_end = min(tile_start + kTile, T.node_base + T.num_nodes);
  const int p = tile_start + tid;
  const bool active = p < tile_end;

  // ---- (1) own branch terms --------------------------------------------------------------------------------
  double delta = 0.0, g = 0.0, tP = 0.0, tN = 0.0;
  int nmiss = 0, par = -1, dep = 0;
  if (active) {
    par = f.parent_pos[p];
    dep = f.depth[p];
    tN = f.t[p];
    if (par >= 0) {
      tP = f.t[par];
      branch_terms<true>(f, S, s_q, p, tP, delta, nmiss, g, s_ab);
    } else {
      double gg; int dummy_ab[1];
      // the root's list holds ref->root-sequence "mutations" (t = -DBL_MAX): they shift lambda but are not counted
      branch_terms<false>(f, S, s_q, p, 0.0, delta, nmiss, gg, dummy_ab);
    }
  }
  s_delta[tid] = delta;
  s_nmiss[tid] = nmiss;
  __syncthreads();

  // ---- (2) diff = own delta - deltas of the nodes whose subtree closes right before this position ---------------
  double diff = delta;
  int idiff = nmiss;
  if (active) {
    const int q = p - T.node_base;
    if (q > 0) {
      const int dprev = f.depth[p - 1];
      const int c0 = (q - 1) - dprev, c1 = q - dep;
      double cs = 0.0; int ci = 0;
      for (int j = c0; j < c1; ++j) {
        const int a = f.post_node[T.node_base + j];
        if (a >= tile_start) {
          cs += s_delta[a - tile_start];
          ci += s_nmiss[a - tile_start];
        } else {   // opened in an earlier tile: re-derive its branch delta
          double da, ga; int na; int dummy_ab[1];
          branch_terms<false>(f, S, s_q, a, 0.0, da, na, ga, dummy_ab);
          cs += da; ci += na;
        }
      }
      diff -= cs;
      idiff -= ci;
    }
  }

  // ---- (3) block scan, publish tile aggregate, gather predecessors ---------------------------------------------------
  double tot; int itot;
  const double incl = block_scan_incl<double, kTile>(diff, s_wsd, &tot);
  const int iincl = block_scan_incl<int, kTile>(idiff, s_wsi, &itot);
  if (tid == 0) {
    P.tile_agg[tile] = tot;
    P.tile_iagg[tile] = itot;
    __threadfence();
    st_release_u32(P.tile_flag + tile, P.epoch);
  }
  if (warp == 0) {
    double acc = 0.0; int iacc = 0;
    for (int j0 = T.first_tile; j0 < tile; j0 += 32) {
      const int j = j0 + lane;
      if (j < tile) {
        while (ld_acquire_u32(P.tile_flag + j) != P.epoch) { __nanosleep(20); }
        acc += ld_cg_f64(P.tile_agg + j);
        iacc += ld_cg_i32(P.tile_iagg + j);
      }
    }
    acc = warp_sum(acc);
    iacc = warp_sum(iacc);
    if (lane == 0) { s_prefix = acc; s_iprefix = iacc; }
  }
  __syncthreads();

  // ---- (4) lambda_i, nsmn, log-G partials ----------------------------------------------------------------------------
  double contrib = 0.0, tcontrib = 0.0;
  int nmut = 0;
  if (active) {
    const double lambda_ref = S.cumQ[S.L];
    const double lam = lambda_ref + (s_prefix + incl);
    const int id = f.node_id[p];
    P.lambda_out[T.node_base + id] = lam;
    P.nsmn_out[T.node_base + id] = s_iprefix + iincl;
    if (par >= 0) {
      const double len = tN - tP;
      contrib = -lam * len + g;
      tcontrib = len;
      nmut = f.mut_off[p + 1] - f.mut_off[p];
    } else {
      P.tree_out[tree * 4 + 3] = lam;
    }
  }
  const double bsum = block_sum<double, kTile>(contrib, s_wsd);
  const double tsum = block_sum<double, kTile>(tcontrib, s_wsd);
  const int msum = block_sum<int, kTile>(nmut, s_wsi);
  __syncthreads();
  if (tid == 0) {
    P.tile_part[tile * 2 + 0] = bsum;
    P.tile_part[tile * 2 + 1] = tsum;
    P.tile_ipart[tile * 17 + 0] = msum;
  }
  if (tid < 16) P.tile_ipart[tile * 17 + 1 + tid] = s_ab[tid];
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const uint32_t done = atomicAdd(P.tree_done + tree, 1u);
    s_is_last = (done == (uint32_t)T.num_tiles - 1u);
    const uint32_t all = atomicAdd(P.ticket + 1, 1u);
    if (all == (uint32_t)f.num_tiles - 1u) { P.ticket[0] = 0u; P.ticket[1] = 0u; }   // re-arm for the next launch
  }
  __syncthreads();
  if (!s_is_last) return;

  // ---- (5) last CTA of this tree: fold per-tile partials in tile order + root prior --------------------------------------
  __threadfence();
  if (tid == 0) P.tree_done[tree] = 0u;
  if (warp == 0) {
    double a0 = 0.0, a1 = 0.0; int m = 0;
    for (int j = T.first_tile + lane; j < T.first_tile + T.num_tiles; j += 32) {
      a0 += ld_cg_f64(P.tile_part + j * 2 + 0);
      a1 += ld_cg_f64(P.tile_part + j * 2 + 1);
      m += ld_cg_i32(P.tile_ipart + j * 17);
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); m = warp_sum(m);
    if (lane == 0) {
      P.tree_out[tree * 4 + 1] = a0;
      P.tree_out[tree * 4 + 2] = a1;
      P.tree_iout[tree * 20 + 0] = m;
      P.tree_iout[tree * 20 + 1] = 0;
    }
  } else if (warp == 1) {
    for (int b = lane; b < 16; b += 32) {
      int c = 0;
      for (int j = T.first_tile; j < T.first_tile + T.num_tiles; ++j) c += ld_cg_i32(P.tile_ipart + j * 17 + 1 + b);
      P.tree_iout[tree * 20 + 2 + b] = c;
    }
  }
  // root prior (core/phylo_tree_calc.cpp:467-504): reference-sequence state counts per partition, adjusted by the
  // root's "mutations", missing sites and from-state overrides.
  if (tid < kMaxPartitions * 4) s_cnt[tid] = tid < S.P * 4 ? S.ref_freq[tid] : 0;
  __syncthreads();
  {
    const int r = T.node_base;   // the root is the first position of its tree
    for (int i = f.mut_off[r] + tid; i < f.mut_off[r + 1]; i += kTile) {
      const int l = f.mut_site[i]; const int ft = f.mut_ft[i]; const int pt = S.part[l];
      atomicSub(&s_cnt[pt * 4 + (ft >> 2)], 1);
      atomicAdd(&s_cnt[pt * 4 + (ft & 3)], 1);
    }
    for (int i = f.miss_off[r]; i < f.miss_off[r + 1]; ++i) {
      const int s = f.miss_start[i], e = f.miss_end[i];
      for (int l = s + tid; l < e; l += kTile) atomicSub(&s_cnt[S.part[l] * 4 + S.ref[l]], 1);
    }
    for (int i = f.fs_off[r] + tid; i < f.fs_off[r + 1]; i += kTile) {
      const int l = f.fs_site[i]; const int pt = S.part[l];
      atomicAdd(&s_cnt[pt * 4 + S.ref[l]], 1);
      atomicSub(&s_cnt[pt * 4 + f.fs_from[i]], 1);
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

int launch_log_G(dphy_ctx* ctx, dphy_forest* fo) {
  if (fo->h.num_tiles == 0) return DPHY_OK;
  LogGParams P;
  P.f = fo->h;
  P.lambda_out = fo->d_lambda;
  P.nsmn_out = fo->d_nsmn;
  P.tile_agg = fo->d_tile_agg;
  P.tile_iagg = fo->d_tile_iagg;
  P.tile_flag = fo->d_tile_flag;
  P.tile_part = fo->d_tile_part;
  P.tile_ipart = fo->d_tile_ipart;
  P.tree_done = fo->d_tree_done;
  P.ticket = fo->d_ticket;
  P.tree_out = fo->d_tree_out;
  P.tree_iout = fo->d_tree_iout;
  int st = refresh_sites(ctx, fo);
  if (st != DPHY_OK) return st;
  fo->epoch += 1;
  if (fo->epoch == 0) fo->epoch = 1;
  P.epoch = fo->epoch;
  emat_log_G_kernel<<<fo->h.num_tiles, kTile, 0, ctx->stream>>>(P);
  ctx->launches += 1;
  return check_cuda(ctx, cudaGetLastError(), "emat_log_G_kernel launch");
}

}  // namespace dphy
