// kernels_spr_group.cuh -- full SPR studies of ONE tree, 32 at a time: what the grouped kernels share.
// (included by kernels_spr.cu inside namespace dphy; the kernels themselves are in kernels_spr_group2.cuh)
//
// In the per-study pipeline every study re-walks all N nodes and M mutations of its tree although the node records, the list
// offsets, the mutation sites / codes / times are the same for every study of that tree.  Full (unbounded) studies of the same
// tree are therefore processed 32 at a time: per study only what depends on X differs -- the state of X at a mutated site (one
// 32-byte row of the site-major table xT[l][32] per mutation), the cut-off t_X, the special nodes X / P / S, and where a region
// lands in that study's output.  Round 2 first walked the nodes with the lanes of a warp as studies (scan + staged emit, 1.50 ->
// 1.17 ms per 128 studies); the event-scan formulation of kernels_spr_group2.cuh replaced it (0.64 ms) and that code was removed.
//
// Nodes that need the general rules -- the root, P and S of a study (account_for_Xs_detachment), the nodes of its start->root
// path, O(depth) per study -- are emitted by spr_segments_kernel (one thread per path node) with eval_region and the segment
// arithmetic of the per-study kernels.  Region order, integers and doubles are those of the per-study pipeline (same parity tests).
constexpr int kGroup = 32;      // studies per group == lanes

struct SprGroupDev {
  int32_t tree, num, node_base, num_nodes;
  int32_t L, mut_base, pad0, pad;
  int32_t study[kGroup];        // indices into SprBatchDev::studies
  int64_t off_xT;               // u8  [L][32]      X's state (+ missing bit) per site, site-major
  // event-scan path (kernels_spr_group2.cuh)
  int32_t num_muts, num_templates, num_ev_chunks, num_t_chunks;   // M of the tree; N + M; chunks of 128 events / 32 templates
  int64_t off_S;                // i8  [num_ev_chunks * 128][32]  chunk-local prefix of the signed potentials before every event
  int64_t off_aggS;             // i32 [num_ev_chunks + 1][32]    chunk totals, then their exclusive prefixes
  int64_t off_mask;             // u32 [num_t_chunks + 1][32]     keep flags of the 32 templates of a chunk, per study
  int64_t off_aggK;             // i32 [num_t_chunks + 2][32]     kept regions per chunk, then their exclusive prefixes
  int64_t off_cbase;            // int2 [num_t_chunks + 1][32]    (output offset of the chunk's kept regions, 0) or (row in the study's hmix, 1)
  int64_t off_trec;             // G2Templ [num_templates]        study-independent template records (shared by the groups of a tree)
  int64_t off_consts, off_outs; // G2Const [32], G2Out [32]
  int64_t off_csort, off_cpm, off_cq;   // per chunk (shared like off_trec): double [..][32] sorted start times, u32 [..][32] running OR of their lane bits, int2 first / last position
  int32_t trec_owner, pad2;     // this group writes the tree's template records
};

__global__ void __launch_bounds__(256) spr_xT_kernel(SprBatchDev B, const SprGroupDev* __restrict__ groups) {
  const SprGroupDev& G = groups[blockIdx.y];
  const int l = blockIdx.x * 256 + threadIdx.x;
  if (l >= G.L) return;
  uint32_t w[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) w[k] = 0u;
  for (int s = 0; s < G.num; ++s) {
    const SprStudy& S = B.studies[G.study[s]];
    const uint32_t x = ((const uint8_t*)(B.slab + S.off_xtab))[l];      // coalesced across the threads of a warp for a fixed s
    w[s >> 2] |= x << ((s & 3) * 8);
  }
  uint4* dst = (uint4*)(B.slab + G.off_xT + (size_t)l * kGroup);
  dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
  dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

__device__ __forceinline__ int g_mut_dh(int xt, int code) {
  if (xt & 4) return 0;
  const int x = xt & 3, from = code >> 2, to = code & 3;
  return (int)(to != x) - (int)(from != x);
}

