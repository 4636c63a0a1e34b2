// partition.cpp -- cut an EMAT into independent sub-tree parts (host, C++), one part (or several) per GPU.
//
// Mirrors the reference's partitioning of a Run into Subruns:
//   generate_random_partition_stencil   core/tree_partitioning.h:139-194  (cut points: subtrees of >= max(10,
//                                        branches_left/(parts_left+1)) branches, found in a post-order sweep)
//   partition_tree / make_partition_part core/tree_partitioning.h:88-137,196-239 (a part runs from its cut point down to
//                                        tips or to other cut points, which become frozen "tips" of the part)
//   Run::repartition                     core/run.cpp:110-193 (each part becomes a standalone Phylo_tree: the part root
//                                        carries the sequence at the cut point as "mutations" from the reference sequence
//                                        at t = -DBL_MAX and the sites missing there as its missations, with no
//                                        from_states; includes_run_root only for the part that holds the tree's root)
// The additive tallies (log G, num_muts, num_muts_ab, T, Ttwiddle) of the parts sum to those of the whole tree
// (Run::check_global_and_local_totals_match, core/run.cpp:340-357), which is what lets every part be evaluated on a
// different GPU with one small all-reduce per cycle.
#include "delphy_b200.h"

#include <algorithm>
#include <cfloat>
#include <cstring>
#include <map>
#include <vector>

namespace {

struct PartOwner {
  std::vector<int32_t> parent, child0, child1, mut_off, mut_site, miss_off, miss_start, miss_end, fs_off, fs_site, orig;
  std::vector<uint8_t> mut_from, mut_to, fs_from;
  std::vector<double> t, mut_t;
  dphy_emat_host view{};
};

}  // namespace

struct dphy_partition {
  std::vector<PartOwner*> parts;
  std::vector<int32_t> part_of_node;   // original node -> part index (cut points belong to the part they root)
  PartOwner* merged = nullptr;         // result of the last dphy_partition_reassemble
  ~dphy_partition() { for (auto* p : parts) delete p; delete merged; }
};

// The reference draws its "bit of randomness" as std::bernoulli_distribution{0.5}(bitgen) with bitgen an absl::BitGenRef over the
// run's std::mt19937 (core/run.h:20, core/tree.h:338, core/tree_partitioning.h:172).  Restated from the published pieces so that
// the same seed yields the same cut points as the reference: MT19937 (Matsumoto & Nishimura 1998; 32-bit outputs); BitGenRef
// composes a 64-bit value from two outputs, first one high (absl FastUniformBits, power-of-two range); libstdc++'s bernoulli
// compares generate_canonical<double, 53> = double(u64) / 2^64 (conversion rounds to nearest) with p.
namespace {
struct Mt19937 {
  uint32_t mt[624];
  int idx = 624;
  explicit Mt19937(uint32_t seed) {
    mt[0] = seed;
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
  }
  uint32_t next() {
    if (idx >= 624) {
      for (int i = 0; i < 624; ++i) {
        const uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
        mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    uint32_t y = mt[idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
  }
  bool bernoulli_half() {
    const uint64_t hi = next(), lo = next();
    const uint64_t u = (hi << 32) + lo;
    return static_cast<double>(u) / 18446744073709551616.0 < 0.5;
  }
};

// randomized_post_order_traversal (core/tree.h:320-365) as a pull iterator: the reference's generator is lazy, so its coin flips
// interleave with the ones the stencil loop makes between two pulls
struct RandomizedPostOrder {
  const dphy_emat_host* e;
  Mt19937* rng;
  std::vector<std::pair<int32_t, int32_t>> stack;   // (node, children_so_far), -1 == not expanded yet
  RandomizedPostOrder(const dphy_emat_host* e_, Mt19937* r) : e(e_), rng(r) { stack.push_back({e->root, -1}); }
  // next node for which children_so_far == number of children, or -1 when the traversal is over
  int32_t next() {
    while (!stack.empty()) {
      auto [node, so_far] = stack.back();
      stack.pop_back();
      const bool tip = e->child0[node] < 0;
      const int nchild = tip ? 0 : 2;
      if (so_far != -1) {
        if (so_far == nchild) return node;
        continue;
      }
      stack.push_back({node, nchild});
      if (!tip) {
        const int32_t c0 = e->child0[node], c1 = e->child1[node];
        if (rng->bernoulli_half()) { stack.push_back({c0, -1}); stack.push_back({node, 1}); stack.push_back({c1, -1}); stack.push_back({node, 0}); }
        else { stack.push_back({c1, -1}); stack.push_back({node, 1}); stack.push_back({c0, -1}); stack.push_back({node, 0}); }
      }
    }
    return -1;
  }
};
}  // namespace

extern "C" int dphy_partition_generate_stencil(const dphy_emat_host* e, int32_t num_parts, uint64_t seed, int32_t* cut_points,
                                               int32_t* num_cut_points) {
  if (!e || !cut_points || !num_cut_points || num_parts < 1) return DPHY_ERR_INVALID_ARGUMENT;
  const int n = e->num_nodes;
  *num_cut_points = 0;
  Mt19937 rng((uint32_t)seed);
  RandomizedPostOrder order(e, &rng);
  std::vector<int32_t> desc(n, 0);
  long branches_left = n, parts_left = num_parts;
  int ncut = 0;
  for (int32_t v = order.next(); v >= 0; v = order.next()) {
    if (v == e->root) break;                  // the root never goes explicitly into the stencil
    if (ncut == num_parts - 1) break;         // the last part implicitly starts at the root
    desc[v] = 1;
    if (e->child0[v] >= 0) desc[v] += desc[e->child0[v]] + desc[e->child1[v]];
    const long min_size = std::max(10L, branches_left / (parts_left + 1));
    if (desc[v] >= min_size) {
      bool allowed = true;
      if (branches_left - (desc[v] - 1) < min_size) allowed = false;       // the remaining stump would be too small
      if (allowed && rng.bernoulli_half()) allowed = false;                  // "a bit of randomness"
      if (allowed) {
        branches_left -= desc[v] - 1;
        cut_points[ncut++] = v;
        desc[v] = 1;
        --parts_left;
      }
    }
  }
  *num_cut_points = ncut;
  return DPHY_OK;
}

extern "C" {

int dphy_partition_split(const dphy_emat_host* e, const dphy_sites_host* s, int32_t num_cut_points, const int32_t* cut_points,
                         dphy_partition** out) {
  if (!e || !s || !out || num_cut_points < 0 || (num_cut_points > 0 && !cut_points)) return DPHY_ERR_INVALID_ARGUMENT;
  const int n = e->num_nodes, L = s->num_sites;
  std::vector<char> is_cut(n, 0);
  std::vector<int32_t> roots;
  for (int i = 0; i < num_cut_points; ++i) {
    const int c = cut_points[i];
    if (c < 0 || c >= n) return DPHY_ERR_OUT_OF_RANGE;
    if (c == e->root || is_cut[c]) continue;
    is_cut[c] = 1; roots.push_back(c);
  }
  roots.push_back(e->root);   // the implicit final part starts at the root (core/tree_partitioning.h:196-222)
  auto* P = new dphy_partition();
  P->part_of_node.assign(n, -1);
  for (size_t pi = 0; pi < roots.size(); ++pi) {
    const int subroot = roots[pi];
    auto* po = new PartOwner();
    P->parts.push_back(po);
    // ---- topology: DFS from the cut point, stopping at tips and at other cut points ------------------------------------
    // node numbering exactly as details::make_partition_part (core/tree_partitioning.h:88-135): the part root is 0; a node's two
    // children get consecutive indices when the node is popped from the work stack (left pushed first, so right is popped first)
    auto add_node = [&](int32_t pd) {
      po->orig.push_back(-1); po->parent.push_back(pd); po->child0.push_back(-1); po->child1.push_back(-1); po->t.push_back(0.0);
      return (int32_t)po->orig.size() - 1;
    };
    std::vector<std::pair<int32_t, int32_t>> st{{subroot, add_node(-1)}};   // (orig node, dst node)
    while (!st.empty()) {
      auto [v, d] = st.back(); st.pop_back();
      po->orig[d] = v; po->t[d] = e->t[v];
      P->part_of_node[v] = (int)pi;
      const bool frozen_tip = is_cut[v] && v != subroot;
      if (e->child0[v] >= 0 && !frozen_tip) {
        const int32_t dl = add_node(d), dr = add_node(d);
        po->child0[d] = dl; po->child1[d] = dr;
        st.push_back({e->child0[v], dl});
        st.push_back({e->child1[v], dr});
      }
    }
    const int pn = (int)po->orig.size();
    // ---- lists ---------------------------------------------------------------------------------------------------------
    po->mut_off.assign(pn + 1, 0); po->miss_off.assign(pn + 1, 0); po->fs_off.assign(pn + 1, 0);
    for (int d = 0; d < pn; ++d) {
      const int v = po->orig[d];
      if (d == 0 && subroot != e->root) {
        // missing at the cut point: union of the intervals on the path to the root (touching intervals coalesce)
        std::vector<std::pair<int32_t, int32_t>> iv;
        std::vector<int32_t> path;
        for (int a = v; a >= 0; a = e->parent[a]) {
          path.push_back(a);
          for (int i = e->miss_off[a]; i < e->miss_off[a + 1]; ++i) iv.push_back({e->miss_start[i], e->miss_end[i]});
        }
        std::sort(iv.begin(), iv.end());
        std::vector<std::pair<int32_t, int32_t>> merged;
        for (auto& x : iv) {
          if (!merged.empty() && x.first <= merged.back().second) merged.back().second = std::max(merged.back().second, x.second);
          else merged.push_back(x);
        }
        auto missing = [&](int l) {
          auto it = std::upper_bound(merged.begin(), merged.end(), l, [](int x, const std::pair<int32_t, int32_t>& p) { return x < p.first; });
          if (it == merged.begin()) return false;
          --it; return l < it->second;
        };
        // sequence at the cut point: reference overlaid with the mutations on the root->cut-point path, in order
        std::map<int32_t, uint8_t> state;
        for (auto it = path.rbegin(); it != path.rend(); ++it)
          for (int i = e->mut_off[*it]; i < e->mut_off[*it + 1]; ++i) state[e->mut_site[i]] = e->mut_to[i];
        for (auto& [l, b] : state) {
          if (l < 0 || l >= L) { delete P; return DPHY_ERR_OUT_OF_RANGE; }
          if (b != s->ref[l] && !missing(l)) {
            po->mut_site.push_back(l); po->mut_from.push_back(s->ref[l]); po->mut_to.push_back(b); po->mut_t.push_back(-DBL_MAX);
          }
        }
        for (auto& x : merged) { po->miss_start.push_back(x.first); po->miss_end.push_back(x.second); }
      } else {
        for (int i = e->mut_off[v]; i < e->mut_off[v + 1]; ++i) {
          po->mut_site.push_back(e->mut_site[i]); po->mut_from.push_back(e->mut_from[i]); po->mut_to.push_back(e->mut_to[i]);
          po->mut_t.push_back(e->mut_t[i]);
        }
        for (int i = e->miss_off[v]; i < e->miss_off[v + 1]; ++i) { po->miss_start.push_back(e->miss_start[i]); po->miss_end.push_back(e->miss_end[i]); }
        for (int i = e->fs_off[v]; i < e->fs_off[v + 1]; ++i) { po->fs_site.push_back(e->fs_site[i]); po->fs_from.push_back(e->fs_from[i]); }
      }
      po->mut_off[d + 1] = (int32_t)po->mut_site.size();
      po->miss_off[d + 1] = (int32_t)po->miss_start.size();
      po->fs_off[d + 1] = (int32_t)po->fs_site.size();
    }
    auto nz = [](auto& v) { if (v.empty()) v.reserve(1); };
    nz(po->mut_site); nz(po->mut_from); nz(po->mut_to); nz(po->mut_t); nz(po->miss_start); nz(po->miss_end); nz(po->fs_site); nz(po->fs_from);
    auto& w = po->view;
    w.num_nodes = pn; w.root = 0; w.includes_run_root = subroot == e->root ? e->includes_run_root : 0; w.reserved = 0;
    w.parent = po->parent.data(); w.child0 = po->child0.data(); w.child1 = po->child1.data(); w.t = po->t.data();
    w.mut_off = po->mut_off.data(); w.mut_site = po->mut_site.data(); w.mut_from = po->mut_from.data(); w.mut_to = po->mut_to.data();
    w.mut_t = po->mut_t.data(); w.miss_off = po->miss_off.data(); w.miss_start = po->miss_start.data(); w.miss_end = po->miss_end.data();
    w.fs_off = po->fs_off.data(); w.fs_site = po->fs_site.data(); w.fs_from = po->fs_from.data();
  }
  *out = P;
  return DPHY_OK;
}

int32_t dphy_partition_num_parts(const dphy_partition* p) { return p ? (int32_t)p->parts.size() : 0; }
const dphy_emat_host* dphy_partition_part(const dphy_partition* p, int32_t i) {
  return (p && i >= 0 && i < (int32_t)p->parts.size()) ? &p->parts[i]->view : nullptr;
}
const int32_t* dphy_partition_orig_index(const dphy_partition* p, int32_t i) {
  return (p && i >= 0 && i < (int32_t)p->parts.size()) ? p->parts[i]->orig.data() : nullptr;
}
void dphy_partition_free(dphy_partition* p) { delete p; }

// reassemble_tree / Run::reassemble (core/tree_partitioning.cpp:55-83, core/run.cpp:195-256): transpose every part's node times,
// lists and topology back onto the whole tree through orig_tree_index.  `parts` are the (possibly edited) parts in the order of
// dphy_partition_part(p, i) -- same node counts as the split produced, any topology / lists / times.  The part that holds the
// run root also sets the tree's root and the root's lists.  The result is a freshly built EMAT owned by `p`.
const dphy_emat_host* dphy_partition_reassemble(dphy_partition* P, const dphy_emat_host* whole, int32_t num_parts, const dphy_emat_host* parts) {
  if (!P || !whole || !parts || num_parts != (int32_t)P->parts.size()) return nullptr;
  const int n = whole->num_nodes;
  struct NodeLists { const dphy_emat_host* src; int32_t v; };
  std::vector<int32_t> parent(whole->parent, whole->parent + n), child0(whole->child0, whole->child0 + n), child1(whole->child1, whole->child1 + n);
  std::vector<double> t(whole->t, whole->t + n);
  std::vector<NodeLists> lists(n);
  for (int v = 0; v < n; ++v) lists[v] = {whole, v};
  int32_t root = whole->root;
  for (int32_t i = 0; i < num_parts; ++i) {
    const dphy_emat_host& sub = parts[i];
    const std::vector<int32_t>& orig = P->parts[i]->orig;
    if (sub.num_nodes != (int32_t)orig.size()) return nullptr;
    for (int32_t sv = 0; sv < sub.num_nodes; ++sv) {
      const int32_t v = orig[sv];
      t[v] = sub.t[sv];
      if (sv != sub.root) lists[v] = {&sub, sv};
      if (sub.child0[sv] >= 0) {     // inner node OF THE PART (cut points that are sub-tips keep their own children)
        const int32_t l = orig[sub.child0[sv]], r = orig[sub.child1[sv]];
        child0[v] = l; child1[v] = r; parent[l] = v; parent[r] = v;
      }
    }
    if (sub.includes_run_root) {
      root = orig[sub.root];
      parent[root] = -1;
      lists[root] = {&sub, sub.root};
    }
  }
  auto* po = new PartOwner();
  po->parent = std::move(parent); po->child0 = std::move(child0); po->child1 = std::move(child1); po->t = std::move(t);
  po->mut_off.assign(n + 1, 0); po->miss_off.assign(n + 1, 0); po->fs_off.assign(n + 1, 0);
  for (int v = 0; v < n; ++v) {
    const dphy_emat_host& e = *lists[v].src; const int32_t u = lists[v].v;
    for (int k = e.mut_off[u]; k < e.mut_off[u + 1]; ++k) {
      po->mut_site.push_back(e.mut_site[k]); po->mut_from.push_back(e.mut_from[k]); po->mut_to.push_back(e.mut_to[k]); po->mut_t.push_back(e.mut_t[k]);
    }
    for (int k = e.miss_off[u]; k < e.miss_off[u + 1]; ++k) { po->miss_start.push_back(e.miss_start[k]); po->miss_end.push_back(e.miss_end[k]); }
    for (int k = e.fs_off[u]; k < e.fs_off[u + 1]; ++k) { po->fs_site.push_back(e.fs_site[k]); po->fs_from.push_back(e.fs_from[k]); }
    po->mut_off[v + 1] = (int32_t)po->mut_site.size(); po->miss_off[v + 1] = (int32_t)po->miss_start.size(); po->fs_off[v + 1] = (int32_t)po->fs_site.size();
  }
  auto nz = [](auto& x) { if (x.empty()) x.reserve(1); };
  nz(po->mut_site); nz(po->mut_from); nz(po->mut_to); nz(po->mut_t); nz(po->miss_start); nz(po->miss_end); nz(po->fs_site); nz(po->fs_from);
  auto& w = po->view;
  w.num_nodes = n; w.root = root; w.includes_run_root = whole->includes_run_root; w.reserved = 0;
  w.parent = po->parent.data(); w.child0 = po->child0.data(); w.child1 = po->child1.data(); w.t = po->t.data();
  w.mut_off = po->mut_off.data(); w.mut_site = po->mut_site.data(); w.mut_from = po->mut_from.data(); w.mut_to = po->mut_to.data();
  w.mut_t = po->mut_t.data(); w.miss_off = po->miss_off.data(); w.miss_start = po->miss_start.data(); w.miss_end = po->miss_end.data();
  w.fs_off = po->fs_off.data(); w.fs_site = po->fs_site.data(); w.fs_from = po->fs_from.data();
  delete P->merged;
  P->merged = po;
  return &po->view;
}

}  // extern "C"
