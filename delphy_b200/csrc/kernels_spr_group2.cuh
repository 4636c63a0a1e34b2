// kernels_spr_group2.cuh -- full SPR studies of ONE tree, 32 at a time, as a flat scan over mutation EVENTS and a flat emit over
// region TEMPLATES.  (included by kernels_spr.cu inside namespace dphy; replaces spr_gscan / spr_gemit of kernels_spr_group.cuh)
//
// What depends on the tree alone is tabulated once per forest (spr_tab_* kernels, cached in dphy_forest::d_spr_*):
//   * region templates: the regions (q, k), k = 0 .. np(q), of every node q in device (DFS) order -- N + M per tree; template index
//     of (q, k) = q + (mut_off[q] - mut_base) + k; tnode[] maps a template back to its node.
//   * events: walking the tree in DFS order, "enter" events (one per mutation of q, when q opens) and "exit" events (the same
//     mutations again when q's subtree closes) -- 2 M per tree.  The nodes closing right before position q are a contiguous slice of
//     the post-order list, so with PM = prefix of np over the post-order list the event index at which q opens is
//     eopen(q) = (mut_off[q] - mut_base) + PM[q - depth(q)], and the exits of a node with post-order index j, closing right before
//     position c, start at (mut_off[c] - mut_base) + PM[j].
// With d(m, s) the potential of mutation m for study s (kernels_spr.cu header) and S[e] = sum of +-d over the events before e, the
// Hamming potential at region (q, k) is simply S[eopen(q) + k]: the tree prefix sum with its closer correction becomes ONE plain
// prefix sum over events, the same for every node shape.
//
//   spr_g2_scan_kernel    lanes = studies.  A warp takes 128 consecutive events: event records loaded 32 at a time (coalesced) and
//                         broadcast, one 32-byte row of xT[site][32] gathered per event, chunk-local prefix stored as int8 rows
//                         S[e][32] (|prefix| <= 127 inside a chunk) + the chunk total.
//   spr_g2_emit_kernel<1> lanes = templates.  A warp takes 32 consecutive templates and loops over the 32 studies: keep flag of every
//                         (template, study) -> one ballot = the 32-bit keep mask of (chunk, study), stored with its popcount.
//   spr_g2_prefix_kernel  exclusive prefixes of the chunk totals (events) and of the kept counts (templates), per study.
//   (spr_segments_kernel lays out the DFS segments of every study from these and emits the special nodes: root, P, S, path nodes)
//   spr_g2_emit_kernel<0> the same loop again: rank inside the mask -> output index; the kept lanes of a warp write CONSECUTIVE
//                         32-byte heads of one study's array, so every store instruction covers whole lines and nothing is staged.
constexpr int kEvChunk = 128;      // events per warp of the scan (int8 chunk-local prefixes)
constexpr int kG2Warps = 4;

// ---- study-independent tables -----------------------------------------------------------------------------------------------------------
// PM[j] = number of mutations of the first j nodes of the tree's post-order list: one CTA per tree, 8 elements per thread per round
__global__ void __launch_bounds__(1024) spr_tab_pm_kernel(ForestDev f, int32_t* __restrict__ PM) {
  __shared__ int s_ws[32];
  __shared__ int s_carry;
  const TreeDev T = f.trees[blockIdx.x];
  const int nb = T.node_base, N = T.num_nodes;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int j0 = 0; j0 < N; j0 += 1024 * 8) {
    int v[8], mine = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = j0 + (int)threadIdx.x * 8 + u;
      v[u] = 0;
      if (j < N) { const int a = f.post_node[nb + j]; v[u] = f.mut_off[a + 1] - f.mut_off[a]; }
      mine += v[u];
    }
    int tot;
    const int incl = block_scan_incl<int, 1024>(mine, s_ws, &tot);
    int run = s_carry + incl - mine;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = j0 + (int)threadIdx.x * 8 + u;
      if (j < N) PM[nb + j] = run;
      run += v[u];
    }
    __syncthreads();
    if (threadIdx.x == 0) s_carry += tot;
    __syncthreads();
  }
}

// eopen, tnode and the event list: one thread per node (grid.y = tree)
__global__ void __launch_bounds__(256) spr_tab_fill_kernel(ForestDev f, const int32_t* __restrict__ PM, int32_t* __restrict__ eopen,
                                                           int32_t* __restrict__ tnode, int32_t* __restrict__ ev) {
  const TreeDev T = f.trees[blockIdx.y];
  const int nb = T.node_base, N = T.num_nodes;
  const int q = blockIdx.x * 256 + threadIdx.x;
  if (q >= N) return;
  const int p = nb + q;
  const int mb = f.mut_off[nb];
  const int moff = f.mut_off[p] - mb, np = f.mut_off[p + 1] - f.mut_off[p];
  const int dep = f.depth[p], size = f.subtree_size[p];
  const int eo = moff + PM[nb + q - dep];
  eopen[p] = eo;
  int32_t* tn = tnode + (size_t)nb + mb + q + moff;          // global template base of the tree = nb + mb
  for (int k = 0; k <= np; ++k) tn[k] = p;
  if (np > 0) {
    int32_t* evt = ev + 2 * (size_t)mb;
    const int xo = (f.mut_off[p + size] - mb) + PM[nb + q + size - 1 - dep];   // post-order index of q = q + size - 1 - depth
    for (int k = 0; k < np; ++k) { evt[eo + k] = moff + k; evt[xo + k] = (moff + k) | (int)0x80000000; }
  }
}

// ---- scan over events -------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kG2Warps * 32) spr_g2_scan_kernel(ForestDev f, SprBatchDev B, const SprGroupDev* __restrict__ groups) {
  const unsigned full = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const SprGroupDev& G = groups[blockIdx.y];
  const int chunk = blockIdx.x * kG2Warps + warp;
  if (chunk >= G.num_ev_chunks) return;
  const int nb = G.node_base, mb = G.mut_base, M2 = 2 * G.num_muts;
  const int np_root = f.mut_off[nb + 1] - f.mut_off[nb];       // the root's own list is never crossed: its events carry no potential
  const uint8_t* __restrict__ xT = (const uint8_t*)(B.slab + G.off_xT);
  int8_t* __restrict__ Srow = (int8_t*)(B.slab + G.off_S);
  const int32_t* __restrict__ ev = B.g2_ev + 2 * (size_t)mb;
  const int e0 = chunk * kEvChunk;
  int acc = 0;
#pragma unroll 1
  for (int b = 0; b < kEvChunk / 32; ++b) {
    // lane i holds event e0 + 32 b + i: (site, from|to, sign) -- pk bit 5 = carries a potential, bit 4 = exit
    const int e = e0 + b * 32 + lane;
    int site = 0, pk = 0;
    if (e < M2) {
      const int evv = __ldg(ev + e);
      const int mr = evv & 0x7fffffff;
      if (mr >= np_root) { site = __ldg(f.mut_site + mb + mr); pk = (__ldg(f.mut_code + mb + mr) & 15) | (evv < 0 ? 16 : 0) | 32; }
    }
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      const int si = __shfl_sync(full, site, i), pi = __shfl_sync(full, pk, i);
      Srow[(size_t)(e0 + b * 32 + i) * kGroup + lane] = (int8_t)acc;        // one 32-byte sector per event
      if (pi & 32) {
        const int d = g_mut_dh(xT[(size_t)si * kGroup + lane], pi & 15);
        acc += (pi & 16) ? -d : d;
      }
    }
  }
  ((int32_t*)(B.slab + G.off_aggS))[(size_t)chunk * kGroup + lane] = acc;
}

// ---- exclusive prefixes over the chunk rows of a group: [rows][32] -> in place, totals in row [rows] ---------------------------------------
// A thread-block cluster of 8 CTAs per (group, table): CTA r owns a contiguous eighth of the rows (warp w a 32nd of that), sums it,
// publishes its column totals into the shared memory of every CTA of the cluster (distributed shared memory) and, after one cluster
// barrier, rewrites its rows as exclusive prefixes.  grid = (8, groups, 2).
constexpr int kPfxCtas = 8;
__global__ void __cluster_dims__(kPfxCtas, 1, 1) __launch_bounds__(1024) spr_g2_prefix_kernel(SprBatchDev B, const SprGroupDev* __restrict__ groups) {
  __shared__ int s_w[32][33];                 // per-warp column sums of this CTA
  __shared__ int s_cta[kPfxCtas][32];         // column totals of every CTA of the cluster
  namespace cgx = cooperative_groups;
  cgx::cluster_group cluster = cgx::this_cluster();
  const int rank = (int)cluster.block_rank();
  const SprGroupDev& G = groups[blockIdx.y];
  const int rows = blockIdx.z == 0 ? G.num_ev_chunks : G.num_t_chunks + 1;
  int32_t* a = (int32_t*)(B.slab + (blockIdx.z == 0 ? G.off_aggS : G.off_aggK));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = (rows + kPfxCtas * 32 - 1) / (kPfxCtas * 32);
  const int r0 = min((rank * 32 + warp) * per, rows), r1 = min(r0 + per, rows);
  int sum = 0;
#pragma unroll 8
  for (int r = r0; r < r1; ++r) sum += a[(size_t)r * kGroup + lane];
  s_w[warp][lane] = sum;
  __syncthreads();
  if (warp < kPfxCtas) {
    // warp w of every CTA posts this CTA's column totals into CTA w
    int tot = 0;
#pragma unroll
    for (int w = 0; w < 32; ++w) tot += s_w[w][lane];
    *cluster.map_shared_rank(&s_cta[rank][lane], warp) = tot;
  }
  cluster.sync();
  int run = 0;
  for (int c = 0; c < rank; ++c) run += s_cta[c][lane];
  for (int w = 0; w < warp; ++w) run += s_w[w][lane];
#pragma unroll 4
  for (int r = r0; r < r1; ++r) {
    const int v = a[(size_t)r * kGroup + lane];
    a[(size_t)r * kGroup + lane] = run;
    run += v;
  }
  if (rank == kPfxCtas - 1 && warp == 31) a[(size_t)rows * kGroup + lane] = run;     // its (possibly empty) range ends at `rows`
}

// ---- keep masks (kCount) / emit over templates ---------------------------------------------------------------------------------------------
struct G2Study {            // what the template loop needs of a study: 3 broadcast 16-byte shared loads per (chunk, study)
  double tX; int qX, xe;    // X's subtree = [qX, xe) in tree-local positions (empty when X is detached)
  int qS, qP, q0, cap;      // S, P, the start node; region capacity
  int base, mH; unsigned out_lo, out_hi;   // hang + kept regions before the chunk; init_min_muts - H0; the study's head array
};

// ---- per (chunk, study): where the chunk's kept regions go -------------------------------------------------------------------------------
// The regions of an off-path node go to  hang[j] + KB(node) + rank,  j = the deepest node of the study's start->root path that
// contains the node.  j is constant over long runs of positions, so it is found ONCE per (template chunk, study) here -- one thread
// each, the binary search over the study's nested path ranges running out of L1 (every thread of a CTA searches the same study) --
// instead of in the emit kernel's prologue, where ten dependent loads per warp were its critical path.  A chunk whose positions
// straddle a segment boundary of the study is flagged; its lanes step from j on their own (g2_lane_hang).
__global__ void __launch_bounds__(256) spr_g2_bases_kernel(SprBatchDev B, const SprGroupDev* __restrict__ groups) {
  const SprGroupDev& G = groups[blockIdx.z];
  const int s = blockIdx.y;
  const int tc = blockIdx.x * 256 + threadIdx.x;
  if (s >= G.num || tc > G.num_t_chunks) return;
  const SprStudy& S = B.studies[G.study[s]];
  int2* out = (int2*)(B.slab + G.off_cbase) + (size_t)tc * kGroup + s;
  const uint32_t mask = ((const uint32_t*)(B.slab + G.off_mask))[(size_t)tc * kGroup + s];
  if (S.error || mask == 0u) { *out = make_int2(0, 0); return; }
  const int nb = G.node_base, NT = G.num_templates;
  const int32_t* tn = B.g2_tnode + (size_t)nb + G.mut_base;
  const int pfirst = __ldg(tn + tc * 32), plast = __ldg(tn + min(tc * 32 + 31, NT - 1));
  const int2* __restrict__ pae = (const int2*)(B.slab + S.off_pae);
  int lo = 0, hi = S.path_len - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int2 ae = __ldg(pae + mid);
    if (pfirst >= ae.x && pfirst < ae.y) hi = mid; else lo = mid + 1;
  }
  const int j = lo;
  const int2 cur = __ldg(pae + j);
  const int dnx = j > 0 ? __ldg(pae + j - 1).x : INT_MAX;
  const bool mixed = (plast >= cur.y) || (dnx > pfirst && dnx <= plast);
  const int base = ((const int32_t*)(B.slab + S.off_hang))[j] + ((const int32_t*)(B.slab + G.off_aggK))[(size_t)tc * kGroup + s];
  *out = make_int2(base, j | (mixed ? (1 << 30) : 0));
}

// slow path of a (chunk, study) pair that straddles a segment boundary of the study: the lane's own segment, stepped from the
// chunk's first node (the deepest path node containing p is monotone in p on either side of the start node)
__device__ __noinline__ int g2_lane_hang(const SprBatchDev& B, int sidx, int j, int p) {
  const SprStudy& S = B.studies[sidx];
  const int2* __restrict__ pae = (const int2*)(B.slab + S.off_pae);
  const int last = S.path_len - 1;
  while (j < last && p >= __ldg(pae + j).y) ++j;
  while (j > 0) { const int2 d = __ldg(pae + j - 1); if (p >= d.x && p < d.y) --j; else break; }
  return ((const int32_t*)(B.slab + S.off_hang))[j];
}

__device__ __noinline__ bool g2_special_keep(const ForestDev& f, const SprBatchDev& B, int sidx, int p, int k, int np, int moff, double tp, double tn) {
  return eval_region(f, B.studies[sidx], p, k, np, moff, tp, tn).keep;
}

template <bool kCount>
__global__ void __launch_bounds__(kG2Warps * 32) spr_g2_emit_kernel(ForestDev f, SprBatchDev B, const SprGroupDev* __restrict__ groups) {
  __shared__ G2Study s_st[kG2Warps][kGroup];
  const unsigned full = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const SprGroupDev& G = groups[blockIdx.y];
  const int tc = blockIdx.x * kG2Warps + warp;
  if (tc > G.num_t_chunks) return;                    // chunk num_t_chunks exists (empty): KB(N) reads its row
  const int nb = G.node_base, mb = G.mut_base, NT = G.num_templates;
  uint32_t* maskp = (uint32_t*)(B.slab + G.off_mask) + (size_t)tc * kGroup;
  int32_t* aggKp = (int32_t*)(B.slab + G.off_aggK) + (size_t)tc * kGroup;

  // ---- the lane's template -----------------------------------------------------------------------------------------------------------------
  const int r = tc * 32 + lane;
  const bool valid = r < NT;
  const int p = valid ? __ldg(B.g2_tnode + (size_t)nb + mb + r) : nb;
  const int q = p - nb;
  const int moff = f.mut_off[p], np = f.mut_off[p + 1] - moff;
  const int k = r - (q + (moff - mb));
  const int par = f.parent_pos[p];
  const bool nonroot = par >= 0;
  const double tn = f.t[p], tp = nonroot ? f.t[par] : 0.0;
  const double t_min = (k == 0 || !valid) ? tp : f.mut_t[moff + k - 1];
  const double t_max = (k >= np || !valid) ? tn : f.mut_t[moff + k];

  // ---- the lane's study ---------------------------------------------------------------------------------------------------------------------
  const bool active = lane < G.num;
  const int sidx = G.study[active ? lane : 0];
  const SprStudy& S = B.studies[sidx];
  const bool ok = active && !S.error;
  uint32_t mymask = 0u;
  int jst = 0;                  // classify() of the chunk's first node
  bool mixed = false;
  {
    G2Study st;
    st.tX = ok ? S.t_X : -DBL_MAX;             // nothing starts before -DBL_MAX: an inactive lane keeps nothing
    const int posX = S.posX;
    st.qX = posX >= 0 ? posX - nb : -1; st.xe = posX >= 0 ? posX - nb + f.subtree_size[posX] : -1;
    st.qS = S.posS >= 0 ? S.posS - nb : -1; st.qP = S.posP >= 0 ? S.posP - nb : -1; st.q0 = S.pos0 - nb; st.cap = S.region_cap;
    st.base = 0; st.mH = 0; st.out_lo = 0u; st.out_hi = 0u;
    if (!kCount) {
      mymask = maskp[lane];
      if (ok && mymask != 0u) {
        const int2 cb = ((const int2*)(B.slab + G.off_cbase))[(size_t)tc * kGroup + lane];
        st.base = cb.x; jst = cb.y & 0x3fffffff; mixed = (cb.y >> 30) & 1;
        st.mH = S.init_min_muts - S.H0;
        const unsigned long long o = (unsigned long long)(B.slab + S.off_regions);
        st.out_lo = (unsigned)o; st.out_hi = (unsigned)(o >> 32);
      } else mymask = 0u;
    }
    s_st[warp][lane] = st;
  }
  __syncwarp();
  if (!kCount && !__any_sync(full, mymask != 0u)) return;

  int cnt_mine = 0;
  const int qend = q + f.subtree_size[p];
  const int idn = f.node_id[p];
  // chunk-local potentials of the lane's template for the 32 studies: one 32-byte row
  uint4 row0 = make_uint4(0u, 0u, 0u, 0u), row1 = row0;
  const int32_t* aggrow = nullptr;
  if (!kCount) {
    const int e = valid ? B.g2_eopen[p] + k : 0;
    const uint4* rp = reinterpret_cast<const uint4*>(B.slab + G.off_S + (size_t)e * kGroup);
    row0 = __ldg(rp); row1 = __ldg(rp + 1);
    aggrow = (const int32_t*)(B.slab + G.off_aggS) + (size_t)(e >> 7) * kGroup;
  }
  const unsigned lt = (1u << lane) - 1u;

#pragma unroll
  for (int s = 0; s < kGroup; ++s) {
    const G2Study& st = s_st[warp][s];
    if (kCount) {
      const double tX = st.tX;
      bool keep = valid && nonroot && t_min < tX && !(q >= st.qX && q < st.xe);
      if (valid && (q == st.qS || q == st.qP || !nonroot)) {
        // S, P (relabelled by account_for_Xs_detachment) and the root follow the general rules, as in node_kept_count
        keep = false;
        if (tX != -DBL_MAX && (nonroot || k == np)) keep = g2_special_keep(f, B, G.study[s], p, k, np, moff, tp, tn);
      }
      const unsigned bal = __ballot_sync(full, keep);
      if (lane == s) { mymask = bal; cnt_mine = __popc(bal); }
    } else {
      const unsigned bal = __shfl_sync(full, mymask, s);
      if (bal == 0u) continue;
      const bool special = q == st.qS || q == st.qP || !nonroot || (q <= st.q0 && st.q0 < qend);
      const bool emit = ((bal >> lane) & 1u) && !special;
      int base = st.base;
      if (__shfl_sync(full, (int)mixed, s)) {
        const int js = __shfl_sync(full, jst, s);
        if (emit) base = g2_lane_hang(B, G.study[s], js, p) + aggKp[s];
      }
      const int idx = base + __popc(bal & lt);
      if (emit && idx >= 0 && idx < st.cap) {
        const unsigned w = s < 16 ? (s < 8 ? (s < 4 ? row0.x : row0.y) : (s < 12 ? row0.z : row0.w))
                                  : (s < 24 ? (s < 20 ? row1.x : row1.y) : (s < 28 ? row1.z : row1.w));
        const int hloc = (int)(w << (24 - 8 * (s & 3))) >> 24;                // sign-extended byte s of the row
        const int m = st.mH + hloc + __ldg(aggrow + s);
        const double tmx = t_max > st.tX ? st.tX : t_max;
        char* o = (char*)(((unsigned long long)st.out_hi << 32) | st.out_lo) + (size_t)idx * sizeof(RegionHead);
        const unsigned long long w0 = (unsigned long long)(unsigned)idn | ((unsigned long long)(unsigned)k << 32);
        asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(o), "l"(w0), "l"(__double_as_longlong(t_min)),
                     "l"(__double_as_longlong(tmx)), "l"((unsigned long long)(unsigned)m) : "memory");
      }
    }
  }
  if (kCount) { maskp[lane] = mymask; aggKp[lane] = cnt_mine; }
}
