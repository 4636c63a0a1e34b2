// kernels_sites.cu -- per-site derived tables of a (reference sequence, Global_evo_model) pair (sm_100a).
//
// Replaces calc_cum_Q_l_for_sequence (core/phylo_tree_calc.cpp:379-388) and
// calc_state_frequencies_per_partition_of (core/phylo_tree_calc.cpp:95-106), which the reference recomputes once per
// cycle per Subrun after set_evo (core/subrun.cpp:17-26) and which together were ~18 % of a 200-tip run
// (plans/2026-03-17-01-cache-ref-seq-derived-quantities.md:46-57).  Also builds the tables the tally kernels use:
// munu[l] = mu_{beta(l)} nu_l and the per-(partition, state) cumulative nu tables that turn the reference's
// per-missing-site loops (core/phylo_tree_calc.cpp:309-314) into two table look-ups per interval.
#include "dphy_internal.h"
#include "device_utils.cuh"

namespace dphy {

constexpr int kScanThreads = 1024;

// One CTA; each thread owns a contiguous chunk of sites (sequential left-to-right inside the chunk like the
// reference's loop), chunk totals are combined by a block scan.  L <= a few 1e5, so this is a ~10 us kernel.
__global__ void __launch_bounds__(kScanThreads) sites_derive_kernel(
    int L, int P, const uint8_t* __restrict__ ref, const uint8_t* __restrict__ part, const double* __restrict__ nu,
    const double* __restrict__ mu, const double* __restrict__ q, double* __restrict__ munu, double* __restrict__ cumQ,
    int32_t* __restrict__ ref_freq, double* __restrict__ cum_nu_ba) {
  __shared__ double s_q[kMaxPartitions * 16];
  __shared__ double s_mu[kMaxPartitions];
  __shared__ double s_ws[kScanThreads / 32];
  __shared__ int s_freq[kMaxPartitions * 4];
  const int tid = threadIdx.x;
  if (tid < P * 16) s_q[tid] = q[tid];
  if (tid < P) s_mu[tid] = mu[tid];
  if (tid < kMaxPartitions * 4) s_freq[tid] = 0;
  __syncthreads();
  const int chunk = (L + kScanThreads - 1) / kScanThreads;
  const int l0 = min(tid * chunk, L), l1 = min(l0 + chunk, L);

  // pass 1: chunk totals of Q_l = mu nu q_a(ref)
  double tot = 0.0;
  int cnt[kMaxPartitions * 4];
#pragma unroll
  for (int i = 0; i < kMaxPartitions * 4; ++i) cnt[i] = 0;
  for (int l = l0; l < l1; ++l) {
    const int pt = part[l], a = ref[l];
    const double mn = s_mu[pt] * nu[l];
    munu[l] = mn;
    tot += mn * (-s_q[pt * 16 + a * 5]);
#pragma unroll
    for (int i = 0; i < kMaxPartitions * 4; ++i) cnt[i] += (i == pt * 4 + a);
  }
  double btot;
  const double incl = block_scan_incl<double, kScanThreads>(tot, s_ws, &btot);
  double run = incl - tot;   // exclusive prefix of this chunk
  if (tid == 0) cumQ[0] = 0.0;
  for (int l = l0; l < l1; ++l) {
    const int pt = part[l], a = ref[l];
    run += s_mu[pt] * nu[l] * (-s_q[pt * 16 + a * 5]);
    cumQ[l + 1] = run;
  }
#pragma unroll
  for (int i = 0; i < kMaxPartitions * 4; ++i) if (cnt[i]) atomicAdd(&s_freq[i], cnt[i]);
  __syncthreads();
  if (tid < P * 4) ref_freq[tid] = s_freq[tid];

  // cumulative nu per (partition, state): cum_nu_ba[(b*4+a)*(L+1) + l] = sum_{l' < l, beta(l')=b, ref[l']=a} nu_l'
  if (cum_nu_ba != nullptr) {
    for (int k = 0; k < P * 4; ++k) {
      __syncthreads();
      double t = 0.0;
      for (int l = l0; l < l1; ++l) if (part[l] * 4 + ref[l] == k) t += nu[l];
      double bt;
      const double inc = block_scan_incl<double, kScanThreads>(t, s_ws, &bt);
      double r = inc - t;
      double* out = cum_nu_ba + (size_t)k * (L + 1);
      if (tid == 0) out[0] = 0.0;
      for (int l = l0; l < l1; ++l) {
        if (part[l] * 4 + ref[l] == k) r += nu[l];
        out[l + 1] = r;
      }
    }
  }
}

// Cumulative reference-state counts per (partition, state), interleaved by site:
// cref[l*(4P) + b*4+a] = #{l' < l : beta(l') == b, ref[l'] == a}  (one 16-byte look-up per partition and interval end).
// Structure only (reference sequence + partition map), so it is built once per sites table.  One CTA per (b, a) row.
// fold_branch_weights_kernel uses it to turn a missation interval into the 4P state counts of the sites it hides.
__global__ void __launch_bounds__(kScanThreads) sites_ref_counts_kernel(int L, const uint8_t* __restrict__ ref,
                                                                        const uint8_t* __restrict__ part, int32_t* __restrict__ cref) {
  __shared__ int s_ws[kScanThreads / 32];
  const int k = blockIdx.x, tid = threadIdx.x;
  const int chunk = (L + kScanThreads - 1) / kScanThreads;
  const int l0 = min(tid * chunk, L), l1 = min(l0 + chunk, L);
  int c = 0;
  for (int l = l0; l < l1; ++l) c += (part[l] * 4 + ref[l] == k);
  int tot;
  const int incl = block_scan_incl<int, kScanThreads>(c, s_ws, &tot);
  int run = incl - c;
  const size_t K = gridDim.x;
  if (tid == 0) cref[k] = 0;
  for (int l = l0; l < l1; ++l) {
    run += (part[l] * 4 + ref[l] == k);
    cref[(size_t)(l + 1) * K + k] = run;
  }
}

int launch_sites_ref_counts(dphy_ctx* ctx, dphy_sites* s) {
  sites_ref_counts_kernel<<<s->P * 4, kScanThreads, 0, ctx->stream>>>(s->L, s->d_ref, s->d_part, s->d_cref);
  ctx->launches += 1;
  return check_cuda(ctx, cudaGetLastError(), "sites_ref_counts_kernel launch");
}

int launch_sites_derive(dphy_ctx* ctx, dphy_sites* s) {
  // mu and q live in the SitesDev host mirror; stage them through the arena
  size_t mark = ctx->arena.mark();
  double* d_mu = (double*)ctx->arena.alloc(sizeof(double) * kMaxPartitions);
  double* d_q = (double*)ctx->arena.alloc(sizeof(double) * kMaxPartitions * 16);
  if (!d_mu || !d_q) return set_error(ctx, DPHY_ERR_OUT_OF_MEMORY, "arena exhausted in sites_derive");
  DPHY_CUDA(ctx, cudaMemcpyAsync(d_mu, s->h.mu, sizeof(double) * kMaxPartitions, cudaMemcpyHostToDevice, ctx->stream));
  DPHY_CUDA(ctx, cudaMemcpyAsync(d_q, s->h.q, sizeof(double) * kMaxPartitions * 16, cudaMemcpyHostToDevice, ctx->stream));
  sites_derive_kernel<<<1, kScanThreads, 0, ctx->stream>>>(s->L, s->P, s->d_ref, s->d_part, s->d_nu, d_mu, d_q, s->d_munu,
                                                           s->d_cumQ, s->d_ref_freq, s->d_cum_nu_ba);
  ctx->launches += 1;
  int st = check_cuda(ctx, cudaGetLastError(), "sites_derive_kernel launch");
  // the staging copies were enqueued before the kernel on the same stream; the arena slots may be reused by later
  // stream-ordered work only, so releasing the mark here is safe.
  ctx->arena.release(mark);
  return st;
}

}  // namespace dphy
