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
};

// Capacity (events per chunk) of the flat event buffers in shared memory.
constexpr int kEvCap = 1536;
constexpr int kClCap = 1024;

// Sequential re-derivation of delta-lambda / missing-site count of ONE branch (used for the few "foreign closers":
// nodes that opened in an earlier tile and close inside this one).  phylo_tree_calc.h:121-155.
__device__ void branch_delta_seq(const ForestDev& f, const SitesDev& S, const double* __restrict__ sq,
                                 const double* __restrict__ smu, int p, double& delta, int& nmiss) {
  const bool uni = S.nu_uniform != 0;
  double dm = 0.0;
  for (int i = f.mut_off[p]; i < f.mut_off[p + 1]; ++i) {
    const int code = __ldg(f.mut_code + i);
    const int pt = code >> 4, from = (code >> 2) & 3, to = code & 3;
    const double mn = uni ? smu[pt] : __ldg(S.munu + __ldg(f.mut_site + i));
    dm += mn * ((-sq[pt * 16 + to * 5]) - (-sq[pt * 16 + from * 5]));
  }
  double dmi = 0.0;
  int nm = 0;
  for (int i = f.miss_off[p]; i < f.miss_off[p + 1]; ++i) {
    const int s = __ldg(f.miss_start + i), e = __ldg(f.miss_end + i);
    dmi -= __ldg(S.cumQ + e) - __ldg(S.cumQ + s);
    nm += e - s;
  }
  for (int i = f.fs_off[p]; i < f.fs_off[p + 1]; ++i) {
    const int code = __ldg(f.fs_code + i);
    const int pt = code >> 4, rf = (code >> 2) & 3, from = code & 3;
    const double mn = uni ? smu[pt] : __ldg(S.munu + __ldg(f.fs_site + i));
    dmi -= mn * ((-sq[pt * 16 + from * 5]) - (-sq[pt * 16 + rf * 5]));
  }
  delta = dm + dmi;
  nmiss = nm;
}

// Flat-over-events + segmented-sum-per-node helper.  All threads of the CTA call it with the same [r0, r1).
//   ev(i, slot)    : thread-parallel over events i of the tile, writes its contribution(s) to smem slot `slot`
//   acc(slot)      : the owning node's thread folds the slots of its CSR slice in list order (deterministic)
template <typename EventFn, typename AccFn>
__device__ __forceinline__ void flat_segmented(int r0, int r1, const int* __restrict__ s_off, int n_act, EventFn ev, AccFn acc) {
  const int tid = threadIdx.x;
  for (int c0 = r0; c0 < r1; c0 += kEvCap) {
    const int c1 = min(c0 + kEvCap, r1);
    for (int i = c0 + tid; i < c1; i += kTile) ev(i, i - c0);
    __syncthreads();
    if (tid < n_act) {
      const int lo = max(s_off[tid], c0), hi = min(s_off[tid + 1], c1);
      for (int i = lo; i < hi; ++i) acc(i - c0);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kTile) emat_log_G_kernel(const LogGParams P) {
  __shared__ double s_q[kMaxPartitions * 16];
  __shared__ double s_mu[kMaxPartitions];
  __shared__ double s_bufA[kEvCap];
  __shared__ double s_bufB[kEvCap];
  __shared__ int s_off_m[kTile + 1];
  __shared__ int s_off_i[kTile + 1];
  __shared__ int s_off_f[kTile + 1];
  __shared__ double s_delta[kTile];
  __shared__ int s_nmiss[kTile];
  __shared__ double s_wsd[kTile / 32];
  __shared__ int s_wsi[kTile / 32];
  __shared__ int s_ab[16];

  const ForestDev& f = P.f;
  const int tid = threadIdx.x;

  if (tid < 16) s_ab[tid] = 0;
  const int tile = blockIdx.x;
  const int tree = f.tile_tree[tile];
  const TreeDev T = f.trees[tree];
  const SitesDev& S = f.sites[T.sites_id];
  if (tid < S.P * 16) s_q[tid] = S.q[tid];
  if (tid < S.P) s_mu[tid] = S.mu[tid] * S.nu_const;
  const bool uni = S.nu_uniform != 0;

  const int tile_in_tree = tile - T.first_tile;
  const int tile_start = T.node_base + tile_in_tree * kTile;              // global device position
  const int tile_end = min(tile_start + kTile, T.node_base + T.num_nodes);
  const int n_act = tile_end - tile_start;
  const int p = tile_start + tid;
  const bool active = p < tile_end;
  const bool has_root = tile_in_tree == 0;                                 // position 0 of a tree is its root

  // ---- (0) node records + CSR offsets of the tile ---------------------------------------------------------------------
  double tP = 0.0, tN = 0.0;
  int par = -1, dep = 0;
  if (active) {
    par = f.parent_pos[p];
    dep = f.depth[p];
    tN = f.t[p];
    if (par >= 0) tP = f.t[par];
    s_off_m[tid] = f.mut_off[p]; s_off_i[tid] = f.miss_off[p]; s_off_f[tid] = f.fs_off[p];
    if (tid == n_act - 1) { s_off_m[n_act] = f.mut_off[p + 1]; s_off_i[n_act] = f.miss_off[p + 1]; s_off_f[n_act] = f.fs_off[p + 1]; }
  }
  __syncthreads();

  // ---- (1) branch terms: flat over the tile's events, segmented sum per node --------------------------------------------
  double dm = 0.0, esum = 0.0, dmi = 0.0;
  int nmiss = 0;
  const int root_m1 = has_root ? s_off_m[1] : s_off_m[0];   // the root's list ("mutations" above the root) ends here
  // mutations: dm_i = mu nu (q_to - q_from);  e_i = dm_i * t_i + log(mu nu q_from,to)   [g_node = sum e_i - t_P * dm]
  flat_segmented(s_off_m[0], s_off_m[n_act], s_off_m, n_act,
    [&](int i, int slot) {
      const int code = __ldg(f.mut_code + i);
      const int pt = code >> 4, from = (code >> 2) & 3, to = code & 3;
      const double mn = uni ? s_mu[pt] : __ldg(S.munu + __ldg(f.mut_site + i));
      const double d = mn * ((-s_q[pt * 16 + to * 5]) - (-s_q[pt * 16 + from * 5]));
      double e = 0.0;
      if (i >= root_m1) {
        e = d * __ldg(f.mut_t + i) + log(mn * s_q[pt * 16 + from * 4 + to]);
        atomicAdd(&s_ab[code & 15], 1);
      }
      s_bufA[slot] = d; s_bufB[slot] = e;
    },
    [&](int slot) { dm += s_bufA[slot]; esum += s_bufB[slot]; });
  // missation intervals: -(cumQ[end] - cumQ[start]); count of sites going missing
  flat_segmented(s_off_i[0], s_off_i[n_act], s_off_i, n_act,
    [&](int i, int slot) {
      const int s = __ldg(f.miss_start + i), e = __ldg(f.miss_end + i);
      s_bufA[slot] = __ldg(S.cumQ + e) - __ldg(S.cumQ + s);
      reinterpret_cast<int*>(s_bufB)[slot] = e - s;
    },
    [&](int slot) { dmi -= s_bufA[slot]; nmiss += reinterpret_cast<int*>(s_bufB)[slot]; });
  // from-state overrides of missing sites
  flat_segmented(s_off_f[0], s_off_f[n_act], s_off_f, n_act,
    [&](int i, int slot) {
      const int code = __ldg(f.fs_code + i);
      const int pt = code >> 4, rf = (code >> 2) & 3, from = code & 3;
      const double mn = uni ? s_mu[pt] : __ldg(S.munu + __ldg(f.fs_site + i));
      s_bufA[slot] = mn * ((-s_q[pt * 16 + from * 5]) - (-s_q[pt * 16 + rf * 5]));
    },
    [&](int slot) { dmi -= s_bufA[slot]; });
  const double delta = dm + dmi;
  const double g = (active && par >= 0) ? esum - tP * dm : 0.0;
  s_delta[tid] = active ? delta : 0.0;
  s_nmiss[tid] = active ? nmiss : 0;
  __syncthreads();

  // ---- (2) diff = own delta - deltas of the nodes whose subtree closes right before this position ---------------------------
  // The tile's closers are one contiguous slice of the tree's post-order list; their deltas are gathered in parallel
  // (in-tile from smem, foreign ones re-derived), then each position folds its own sub-slice in order.
  double diff = active ? delta : 0.0;
  int idiff = active ? nmiss : 0;
  {
    const int q_first = tile_start - T.node_base, q_last = tile_end - 1 - T.node_base;
    const int cl0 = q_first == 0 ? 0 : (q_first - 1) - f.depth[tile_start - 1];
    const int cl1 = q_last - f.depth[tile_end - 1];
    int my0 = 0, my1 = 0;
    if (active) {
      const int q = p - T.node_base;
      if (q > 0) { my0 = (q - 1) - f.depth[p - 1]; my1 = q - dep; }
    }
    int* s_cn = reinterpret_cast<int*>(s_bufB);
    for (int c0 = cl0; c0 < cl1; c0 += kClCap) {
      const int c1 = min(c0 + kClCap, cl1);
      for (int j = c0 + tid; j < c1; j += kTile) {
        const int a = f.post_node[T.node_base + j];
        double da; int na;
        if (a >= tile_start) { da = s_delta[a - tile_start]; na = s_nmiss[a - tile_start]; }
        else branch_delta_seq(f, S, s_q, s_mu, a, da, na);
        s_bufA[j - c0] = da; s_cn[j - c0] = na;
      }
      __syncthreads();
      const int lo = max(my0, c0), hi = min(my1, c1);
      for (int j = lo; j < hi; ++j) { diff -= s_bufA[j - c0]; idiff -= s_cn[j - c0]; }
      __syncthreads();
    }
  }

  // ---- (3) block scan; tile-local lambda_i / nsmn; per-tile partial sums ------------------------------------------------
  double tot; int itot;
  const double incl = block_scan_incl<double, kTile>(diff, s_wsd, &tot);
  const int iincl = block_scan_incl<int, kTile>(idiff, s_wsi, &itot);
  double contrib = 0.0, tcontrib = 0.0;
  int nmut = 0;
  if (active) {
    const double lam_local = S.cumQ[S.L] + incl;     // + the tile's prefix, added in pass 3
    P.lambda_out[p] = lam_local;
    P.nsmn_out[p] = iincl;
    if (par >= 0) {
      const double len = tN - tP;
      contrib = -lam_local * len + g;
      tcontrib = len;
      nmut = s_off_m[tid + 1] - s_off_m[tid];
    }
  }
  const double bsum = block_sum<double, kTile>(contrib, s_wsd);
  const double tsum = block_sum<double, kTile>(tcontrib, s_wsd);
  const int msum = block_sum<int, kTile>(nmut, s_wsi);
  __syncthreads();
  if (tid == 0) {
    P.tile_agg[tile] = tot;
    P.tile_iagg[tile] = itot;
    P.tile_part[tile * 2 + 0] = bsum;
    P.tile_part[tile * 2 + 1] = tsum;
    P.tile_ipart[tile * 17 + 0] = msum;
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
  for (int j0 = 0; j0 < T.num_tiles; j0 += kTile) {
    const int j = T.first_tile + j0 + tid;
    const bool ok = j0 + tid < T.num_tiles;
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
      const int s = f.miss_start[i], e = f.miss_end[i];
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
__global__ void __launch_bounds__(kTile) emat_log_G_finish_kernel(const LogGParams P) {
  const ForestDev& f = P.f;
  const int tile = blockIdx.x;
  const TreeDev T = f.trees[f.tile_tree[tile]];
  const int p = T.node_base + (tile - T.first_tile) * kTile + threadIdx.x;
  if (p < T.node_base + T.num_nodes) {
    P.lambda_out[p] += P.tile_agg[tile];
    P.nsmn_out[p] += P.tile_iagg[tile];
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
  emat_log_G_kernel<<<fo->h.num_tiles, kTile, 0, ctx->stream>>>(P);
  emat_log_G_tree_kernel<<<fo->h.num_trees, kTile, 0, ctx->stream>>>(P);
  emat_log_G_finish_kernel<<<fo->h.num_tiles, kTile, 0, ctx->stream>>>(P);
  ctx->launches += 3;
  return check_cuda(ctx, cudaGetLastError(), "emat_log_G kernels launch");
}

}  // namespace dphy
