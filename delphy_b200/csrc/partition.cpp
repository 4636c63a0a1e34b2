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
  ~dphy_partition() { for (auto* p : parts) delete p; }
};

extern "C" {

int dphy_partition_generate_stencil(const dphy_emat_host* e, int32_t num_parts, uint64_t seed, int32_t* cut_points,
                                    int32_t* num_cut_points) {
  if (!e || !cut_points || !num_cut_points || num_parts < 1) return DPHY_ERR_INVALID_ARGUMENT;
  const int n = e->num_nodes;
  *num_cut_points = 0;
  if (num_parts == 1 || n < 3) return DPHY_OK;
  // splitmix64 stream for the "bit of randomness" of the reference (child visiting order + 50% veto)
  auto next = [&seed]() {
    seed += 0x9E3779B97F4A7C15ULL;
    uint64_t z = seed;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
  };
  // randomized post-order (core/tree.h:320-357)
  std::vector<int32_t> post; post.reserve(n);
  {
    std::vector<int32_t> st{e->root};
    std::vector<int32_t> pre; pre.reserve(n);
    while (!st.empty()) {
      int v = st.back(); st.pop_back(); pre.push_back(v);
      if (e->child0[v] >= 0) {
        if (next() & 1) { st.push_back(e->child0[v]); st.push_back(e->child1[v]); }
        else { st.push_back(e->child1[v]); st.push_back(e->child0[v]); }
      }
    }
    post.assign(pre.rbegin(), pre.rend());   // reverse pre-order visits children before parents
  }
  std::vector<int32_t> desc(n, 0);
  long branches_left = n, parts_left = num_parts;
  int ncut = 0;
  for (int v : post) {
    if (v == e->root) break;
    if (ncut == num_parts - 1) break;
    desc[v] = 1;
    if (e->child0[v] >= 0) desc[v] += desc[e->child0[v]] + desc[e->child1[v]];
    const long min_size = std::max(10L, branches_left / (parts_left + 1));
    if (desc[v] >= min_size) {
      bool allowed = true;
      if (branches_left - (desc[v] - 1) < min_size) allowed = false;
      if (allowed && (next() & 1)) allowed = false;
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
    std::vector<std::pair<int32_t, int32_t>> st{{subroot, -1}};   // (orig node, parent dst)
    while (!st.empty()) {
      auto [v, pd] = st.back(); st.pop_back();
      const int d = (int)po->orig.size();
      po->orig.push_back(v); po->parent.push_back(pd); po->child0.push_back(-1); po->child1.push_back(-1);
      po->t.push_back(e->t[v]);
      if (pd >= 0) { if (po->child0[pd] < 0) po->child0[pd] = d; else po->child1[pd] = d; }
      P->part_of_node[v] = (int)pi;
      const bool frozen_tip = is_cut[v] && v != subroot;
      if (e->child0[v] >= 0 && !frozen_tip) {
        st.push_back({e->child1[v], d});   // pushed first => visited second => becomes children[1]
        st.push_back({e->child0[v], d});
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

}  // extern "C"
