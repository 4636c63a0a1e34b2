// kernels_spr_group.cuh -- full SPR studies of ONE tree, 32 at a time: the lanes of a warp are the studies.
// (included by kernels_spr.cu inside namespace dphy)
//
// In the per-study pipeline every study re-walks all N nodes and M mutations of its tree although the node records, the list
// offsets, the mutation sites / codes / times are the same for every study of that tree: with 64-128 studies per tree the integer
// work per region (list walks, index arithmetic, searches) is paid 64-128 times.  Here a warp owns a chunk of 32 consecutive
// device positions and walks it ONCE for 32 studies: everything that depends on the tree alone is loaded once per warp and
// broadcast; per lane (= study) only what depends on X differs -- the state of X at a mutated site (one 32-byte row of the
// site-major table xT[l][32] per mutation), the cut-off t_X, the special nodes X / P / S, and where the region lands in that
// study's output.  Per (node, 32 studies) the warp issues a few dozen instructions instead of a few dozen per study.
//
//   spr_xT_kernel      xT[l][lane] <- xtab_lane[l]                                     (site-major copy of the 32 state tables)
//   spr_gscan_kernel   per chunk: closers, per-mutation potentials (kept in dhT[m][lane]), H_end and the kept-region counts;
//                      chunk-local prefixes Hloc[q][lane], KBloc[q][lane] + chunk totals (same two-level scheme as spr_scan_kernel)
//   (spr_segments_kernel, unchanged, turns the totals into prefixes and lays out the DFS segments of every study)
//   spr_gemit_kernel   per chunk: regions of every node for the 32 studies; each lane writes its own study's 32-byte heads in that
//                      study's DFS order (the weights are the dense spr_weights_kernel pass over the heads)
//
// Nodes that need the general rules -- the root, P and S of a study (account_for_Xs_detachment), the nodes of its start->root
// path, O(depth) per study -- are emitted by spr_segments_kernel (one thread per path node) with eval_region and the segment
// arithmetic of the per-study kernels, so that no lane of a warp ever waits for another lane's special case.  Region order, integers and doubles are those of the per-study pipeline (same parity tests).
constexpr int kGroup = 32;      // studies per group == lanes
constexpr int kGChunk = 32;     // device positions per warp
constexpr int kGWarps = 4;      // warps per CTA
constexpr int kGRun = 4;        // records a lane parks before the warp writes the runs out (4 x 32 bytes = one line)

struct SprGroupDev {
  int32_t tree, num, node_base, num_nodes;
  int32_t L, mut_base, num_chunks, pad;
  int32_t study[kGroup];        // indices into SprBatchDev::studies
  int64_t off_xT;               // u8  [L][32]      X's state (+ missing bit) per site, site-major
  int64_t off_dhT;              // i8  [M_tree][32] potential of every mutation of the tree for every study
  int64_t off_dhP;              // u64 [M_tree / 32 + 4][32] the same, 2 bits each (dh + 1): 32 consecutive mutations per word
  int64_t off_H, off_KB;        // i32 [N][32], i32 [N + 1][32]
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

// ---- scan: H_end and kept-region counts of every node, 32 studies per warp --------------------------------------------------------------
__global__ void __launch_bounds__(kGWarps * 32) spr_gscan_kernel(ForestDev f, SprBatchDev B, const SprGroupDev* __restrict__ groups) {
  __shared__ int s_open[kGWarps][kGChunk][kGroup];       // dH of the nodes opened in this chunk (their closers may follow in it)
  const unsigned full = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const SprGroupDev& G = groups[blockIdx.y];
  const int chunk = blockIdx.x * kGWarps + warp;
  if (chunk >= G.num_chunks) return;
  const int nb = G.node_base, N = G.num_nodes;
  const int q0 = chunk * kGChunk, nq = min(kGChunk, N - q0);
  const bool active = lane < G.num;
  const int sidx = G.study[active ? lane : 0];
  SprStudy& S = B.studies[sidx];
  const bool ok = active && !S.error;
  const double tX = S.t_X;
  const int posX = S.posX, posS = S.posS, posP = S.posP;
  const int xe = posX >= 0 ? posX + f.subtree_size[posX] : -1;
  const uint8_t* __restrict__ xT = (const uint8_t*)(B.slab + G.off_xT);
  int8_t* __restrict__ dhT = (int8_t*)(B.slab + G.off_dhT);
  unsigned long long* __restrict__ dhP = (unsigned long long*)(B.slab + G.off_dhP);
  int32_t* __restrict__ Hloc = (int32_t*)(B.slab + G.off_H);
  int32_t* __restrict__ KBloc = (int32_t*)(B.slab + G.off_KB);
  int (*open)[kGroup] = s_open[warp];
  unsigned long long wacc = 0ULL;       // packed potentials of the word being filled (words at chunk borders are shared: atomicOr)
  int widx = -1;

  // node records of the chunk: lane i holds position q0 + i
  const int pm = nb + q0 + lane;
  const bool has = lane < nq;
  const int par_m = has ? f.parent_pos[pm] : -1;
  const int mo_m = has ? f.mut_off[pm] : 0, mo1_m = has ? f.mut_off[pm + 1] : 0;
  const int dep_m = has ? f.depth[pm] : 0;
  const double tn_m = has ? f.t[pm] : 0.0;
  const double tp_m = (has && par_m >= 0) ? f.t[par_m] : 0.0;

  int acc = 0, kacc = 0;
  int c_prev = q0 == 0 ? 0 : (q0 - 1) - f.depth[nb + q0 - 1];       // closers that precede position q0 are done
  for (int i = 0; i < nq; ++i) {
    const int p = nb + q0 + i;
    const int moi = __shfl_sync(full, mo_m, i), npi = __shfl_sync(full, mo1_m, i) - moi;
    const int pari = __shfl_sync(full, par_m, i);
    const double tni = __shfl_sync(full, tn_m, i), tpi = __shfl_sync(full, tp_m, i);
    // (1) the subtrees that close right before p: post_node[c(q-1) .. c(q))
    const int c_cur = (q0 + i) - __shfl_sync(full, dep_m, i);
    for (int j = c_prev; j < c_cur; ++j) {
      const int a = f.post_node[nb + j];
      int dha;
      if (a >= nb + q0) dha = open[a - nb - q0][lane];
      else {
        dha = 0;
        if (f.parent_pos[a] >= 0) {          // the root's own list is never crossed
          const int ma = f.mut_off[a], ma1 = f.mut_off[a + 1];
          for (int m = ma; m < ma1; ++m) dha += g_mut_dh(xT[(size_t)__ldg(f.mut_site + m) * kGroup + lane], __ldg(f.mut_code + m) & 15);
        }
      }
      acc -= dha;
    }
    c_prev = c_cur;
    // (2) open p: potentials of its own mutations (kept for the emit pass), mutations before t_X
    int dh = 0, below = 0;
    if (pari >= 0) {
      for (int k = 0; k < npi; ++k) {
        const int m = moi + k;
        const int d = g_mut_dh(xT[(size_t)__ldg(f.mut_site + m) * kGroup + lane], __ldg(f.mut_code + m) & 15);
        const int mr = m - G.mut_base;
        dhT[(size_t)mr * kGroup + lane] = (int8_t)d;
        if ((mr >> 5) != widx) {
          if (widx >= 0) atomicOr(dhP + (size_t)widx * kGroup + lane, wacc);
          widx = mr >> 5; wacc = 0ULL;
        }
        wacc |= (unsigned long long)(d + 1) << ((mr & 31) * 2);
        dh += d;
        below += (__ldg(f.mut_t + m) < tX) ? 1 : 0;
      }
    }
    open[i][lane] = dh;
    acc += dh;
    Hloc[(size_t)(q0 + i) * kGroup + lane] = acc;
    // (3) kept regions of branch p after account_for_Xs_detachment + remove_regions_in_Xs_future
    int kc = 0;
    if (ok) {
      if (p == posS || p == posP || pari < 0) {
        const SprView V = make_view(B, S, sidx);
        kc = node_kept_count(f, S, V, p, false, 0);
      } else if (posX >= 0 && p >= posX && p < xe) kc = 0;
      else if (tpi < tX) kc = 1 + (tni < tX ? npi : below);      // a region is kept iff it starts before t_X
    }
    KBloc[(size_t)(q0 + i) * kGroup + lane] = kacc;
    kacc += kc;
  }
  if (widx >= 0) atomicOr(dhP + (size_t)widx * kGroup + lane, wacc);
  if (ok && kacc > 65535) S.error = 5;      // the emit pass keeps the chunk-local counts in 16 bits
  if (active) {
    int32_t* agg = B.tile_agg + ((size_t)sidx * (B.max_tiles + 1) + chunk) * 3;
    agg[0] = acc; agg[1] = 0; agg[2] = kacc;
    if (q0 + nq == N) KBloc[(size_t)N * kGroup + lane] = (N % kGChunk) ? kacc : 0;     // KB(N): same chunk unless N is a chunk multiple
  }
}

// ---- emit --------------------------------------------------------------------------------------------------------------------------------------
// (GLane, g_store_region and the general per-node rules live in kernels_spr.cu: spr_segments_kernel emits the nodes that need them)
__global__ void __launch_bounds__(kGWarps * 32, 4) spr_gemit_kernel(ForestDev f, SprBatchDev B, const SprGroupDev* __restrict__ groups) {
  __shared__ int s_h[kGWarps][kGChunk][kGroup];     // H at the top of every branch of the chunk, per study
  __shared__ unsigned short s_k[kGWarps][kGChunk][kGroup];   // kept regions before every node of the chunk, per study (chunk-local;
                                                             // spr_gscan_kernel flags a chunk with more than 65535 of them)
  // Output staging.  A lane's regions of a chunk are consecutive records of ITS study's array, so written directly every store
  // instruction of the warp touches 32 different lines with 32 bytes each, and L2 / DRAM see 20 M scattered sector writes (measured:
  // half of the kernel's time).  Instead each lane parks up to kGRun records (128 bytes) here and the warp writes all parked runs
  // together, 8 threads x 16 bytes per run: four whole-line runs per store instruction.
  __shared__ uint4 s_stage[kGWarps][kGroup][kGRun * 2 + 1];      // + 1: 144-byte pitch, conflict-free 16-byte shared stores
  const unsigned full = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const SprGroupDev& G = groups[blockIdx.y];
  const int chunk = blockIdx.x * kGWarps + warp;
  if (chunk >= G.num_chunks) return;
  const int nb = G.node_base, N = G.num_nodes;
  const int q0 = chunk * kGChunk, nq = min(kGChunk, N - q0);
  const bool active = lane < G.num;
  const int sidx = G.study[active ? lane : 0];
  const SprStudy& S = B.studies[sidx];
  const SprView V = make_view(B, S, sidx);
  uint4 (*stage)[kGRun * 2 + 1] = s_stage[warp];
  int st_cnt = 0, st_base = 0;                       // parked records of this lane and the output index of the first one
  const long long out_off = S.off_regions;           // this lane's output array, as a slab offset (shuffled to the writers)
  auto flush = [&]() {
    __syncwarp();
#pragma unroll
    for (int it = 0; it < kGRun * 2 * kGroup / 32; ++it) {
      const int item = it * 32 + lane, run = item / (kGRun * 2), piece = item % (kGRun * 2);
      const int c = __shfl_sync(full, st_cnt, run), b = __shfl_sync(full, st_base, run);
      const long long o = __shfl_sync(full, out_off, run);
      if (piece < 2 * c) reinterpret_cast<uint4*>(B.slab + o + (size_t)b * sizeof(RegionHead))[piece] = stage[run][piece];
    }
    __syncwarp();
    st_cnt = 0;
  };
  GLane L;
  L.out = (RegionHead*)(B.slab + S.off_regions);
  L.pae = V.pae; L.seg = V.seg; L.agg = V.agg;
  L.region_cap = S.region_cap; L.path_len = S.path_len; L.H0 = S.H0; L.init_min_muts = S.init_min_muts;
  L.tX = S.t_X;
  // lanes with nothing kept in this chunk only take part in the shuffles
  const int kb_chunk = L.agg[chunk * 3 + 2];
  const bool live = active && !S.error && (L.agg[(chunk + 1) * 3 + 2] != kb_chunk);
  if (!__any_sync(full, live)) return;
  const int posX = S.posX, posS = S.posS, posP = S.posP;
  const int xe = posX >= 0 ? posX + f.subtree_size[posX] : -1;
  const int8_t* __restrict__ dhT = (const int8_t*)(B.slab + G.off_dhT);
  const unsigned long long* __restrict__ dhP = (const unsigned long long*)(B.slab + G.off_dhP);
  const int32_t* __restrict__ Hloc = (const int32_t*)(B.slab + G.off_H);
  const int32_t* __restrict__ KBloc = (const int32_t*)(B.slab + G.off_KB);
  const int agg_h_chunk = L.agg[chunk * 3 + 0];

  // ---- everything the walk needs, loaded up front (independent loads, all in flight together) ----
  const int pm = nb + q0 + lane;
  const bool has = lane < nq;
  const int par_m = has ? f.parent_pos[pm] : -1;
  const int mo_m = has ? f.mut_off[pm] : 0, mo1_m = has ? f.mut_off[pm + 1] : 0;
  const int id_m = has ? f.node_id[pm] : 0;
  const double tn_m = has ? f.t[pm] : 0.0;
  const double tp_m = (has && par_m >= 0) ? f.t[par_m] : 0.0;
  const int M0 = __shfl_sync(full, mo_m, 0), M1 = __shfl_sync(full, mo1_m, nq - 1);     // the chunk's mutations: one CSR range
  const double mtA = (M0 + lane < M1) ? f.mut_t[M0 + lane] : 0.0;                       // lane i holds the time of mutation M0 + i
  const double mtB = (M0 + 32 + lane < M1) ? f.mut_t[M0 + 32 + lane] : 0.0;
  const int wbase = (M0 - G.mut_base) >> 5;
  const unsigned long long wA = dhP[(size_t)wbase * kGroup + lane], wB = dhP[(size_t)(wbase + 1) * kGroup + lane],
                           wC = dhP[(size_t)(wbase + 2) * kGroup + lane];
  int (*sh)[kGroup] = s_h[warp];
  unsigned short (*sk)[kGroup] = s_k[warp];
  // two batches of 16 nodes: all 48 row loads of a batch are issued before the first one is consumed
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    int hv[16], av[16], kv[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int i = half * 16 + u;
      const int pari = __shfl_sync(full, par_m, i);
      const int qp = pari - nb;
      const bool v = i < nq && pari >= 0;
      hv[u] = v ? Hloc[(size_t)qp * kGroup + lane] : 0;
      av[u] = v ? ((qp >> 5) == chunk ? agg_h_chunk : L.agg[(qp >> 5) * 3 + 0]) : 0;
      kv[u] = i < nq ? KBloc[(size_t)(q0 + i) * kGroup + lane] : 0;
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) { sh[half * 16 + u][lane] = hv[u] + av[u]; sk[half * 16 + u][lane] = (unsigned short)kv[u]; }
  }
  __syncwarp();

  // deepest node of the study's start->root path that contains the current position: monotone along the positions on either
  // side of the start node, so it is searched once per chunk and then stepped
  int j = 0, cur_x = 0, cur_y = 0, dn_x = INT_MAX, dn_y = 0, hang = 0;
  auto load_j = [&]() {
    const int2 ae = __ldg(L.pae + j);
    cur_x = ae.x; cur_y = ae.y;
    if (j > 0) { const int2 d = __ldg(L.pae + j - 1); dn_x = d.x; dn_y = d.y; } else { dn_x = INT_MAX; dn_y = 0; }
    const int32_t* sg = L.seg + (size_t)j * kSegStride;
    hang = sg[2] - V.KB(sg[5] - nb);
  };
  if (live) { j = classify(V, nb + q0); load_j(); }

  for (int i = 0; i < nq; ++i) {
    const int p = nb + q0 + i;
    const int moi = __shfl_sync(full, mo_m, i), npi = __shfl_sync(full, mo1_m, i) - moi;
    const int pari = __shfl_sync(full, par_m, i);
    const int idi = __shfl_sync(full, id_m, i);
    const double tni = __shfl_sync(full, tn_m, i), tpi = __shfl_sync(full, tp_m, i);
    bool fast = false;
    int Hk = 0, kb = 0;
    if (live) {
      // step the path index: up when p leaves the current path node's subtree, down when it enters the next path node's
      if (p >= cur_y) { do { ++j; } while (j < L.path_len - 1 && p >= __ldg(L.pae + j).y); load_j(); }
      while (p >= dn_x && p < dn_y) { --j; load_j(); }
      const bool on_path = p == cur_x;
      const bool in_X = posX >= 0 && p >= posX && p < xe;
      if (!in_X) {
        // the root, P, S and the nodes of the start->root path follow the general rules: spr_segments_kernel has emitted them
        if (!(on_path || p == posS || p == posP || pari < 0) && tpi < L.tX) { fast = true; Hk = sh[i][lane]; kb = hang + kb_chunk + sk[i][lane]; }
      }
    }
    if (__any_sync(full, fast)) {
      // ordinary branch: region k spans (t_min, t_max] with t_min = parent time / previous mutation, kept iff it starts before t_X
      double t_min = tpi;
      for (int k = 0; k <= npi; ++k) {
        const int ci = moi + k - M0;                    // index of mutation k of this branch inside the chunk
        double t_max = tni;
        if (k < npi) {
          const double a = __shfl_sync(full, mtA, ci & 31), b = __shfl_sync(full, mtB, ci & 31);
          t_max = ci < 32 ? a : (ci < 64 ? b : __ldg(f.mut_t + moi + k));
        }
        const bool emit = fast && t_min < L.tX && kb >= 0 && kb < L.region_cap;
        // park the record; the warp writes the parked runs out when a lane's run is full or stops being contiguous
        if (__any_sync(full, emit && (st_cnt == kGRun || (st_cnt > 0 && kb != st_base + st_cnt)))) flush();
        if (emit) {
          const double tmx = t_max > L.tX ? L.tX : t_max;
          const int m = L.init_min_muts + (Hk - L.H0);
          if (st_cnt == 0) st_base = kb;
          stage[lane][st_cnt * 2] = make_uint4((unsigned)idi, (unsigned)k, (unsigned)__double2loint(t_min), (unsigned)__double2hiint(t_min));
          stage[lane][st_cnt * 2 + 1] = make_uint4((unsigned)__double2loint(tmx), (unsigned)__double2hiint(tmx), (unsigned)m, 0u);
          ++st_cnt;
        }
        if (fast && t_min < L.tX) ++kb;
        if (k < npi) {
          const int rel = moi + k - G.mut_base - (wbase << 5);     // position inside the three preloaded words
          int d;
          if (rel < 32) d = (int)((wA >> (rel * 2)) & 3ULL) - 1;
          else if (rel < 64) d = (int)((wB >> ((rel - 32) * 2)) & 3ULL) - 1;
          else if (rel < 96) d = (int)((wC >> ((rel - 64) * 2)) & 3ULL) - 1;
          else d = dhT[(size_t)(moi + k - G.mut_base) * kGroup + lane];
          Hk += d; t_min = t_max;
        }
      }
    }
  }
  __syncwarp();
  if (__any_sync(full, st_cnt > 0)) flush();
}
