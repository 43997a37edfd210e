// host_bvh.cpp -- builds the bounding-volume hierarchy of bvh_types.h over the bounded primitives of a scene.
//
// The scene is static over the 1000 passes of a job, so the hierarchy is built once per ppm_scene_set, on the host,
// top-down with a binned surface-area heuristic (16 bins per axis): traversal quality matters, build time does not
// (~1 s per million primitives).  Deep in the tree (PPM_BVH_SAH_DEPTH) the split falls back to the object median,
// which bounds the depth by PPM_BVH_SAH_DEPTH + log2(N) < PPM_BVH_STACK whatever the geometry.
//
// Conservativeness (the traversal must never cull a primitive whose hit test would succeed): every box is the exact
// f64 box of its primitives' corners (sphere: centre +- radius) grown by PAD = 1e-6 x the largest coordinate
// magnitude of the scene (at least 1e-3).  The hit tests of dev_core.cuh accept points whose distance from the
// primitive is a few roundings of the ray arithmetic (~1e-15 x scale, larger only for rays within ~1e-9 rad of the
// primitive's plane), and the slab test itself is good to ~1e-15 x scale: both are far inside the pad.
#include "host_common.h"
#include "bvh_types.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace ppmhost {

namespace {
struct Box {
  double lo[3], hi[3];
  void clear() { for (int a = 0; a < 3; ++a) { lo[a] = INFINITY; hi[a] = -INFINITY; } }
  void add(const double p[3]) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
  void add(const Box& b) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
  double area() const {
    const double x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
    return (x < 0.0 || y < 0.0 || z < 0.0) ? 0.0 : 2.0 * (x * y + y * z + z * x);
  }
};
struct Item { Box b; double c[3]; int32_t prim; };

struct Builder {
  std::vector<Item> items;
  std::vector<BvhNode> nodes;
  std::vector<uint32_t> leaf_order;     // item indices in leaf order
  double pad = 0.0;
  int max_depth = 0;

  uint32_t make_leaf(size_t b, size_t e) {
    const uint32_t first = (uint32_t)leaf_order.size();
    for (size_t i = b; i < e; ++i) leaf_order.push_back((uint32_t)i);
    return PPM_BVH_LEAF | ((uint32_t)(e - b - 1) << 28) | first;
  }
  // builds the subtree over items [b, e); returns its reference and its (unpadded) box
  uint32_t build(size_t b, size_t e, int depth, Box& box) {
    max_depth = std::max(max_depth, depth);
    box.clear();
    Box cb; cb.clear();
    for (size_t i = b; i < e; ++i) { box.add(items[i].b); cb.add(items[i].c); }
    if (e - b <= PPM_BVH_LEAF_MAX) return make_leaf(b, e);
    int axis = 0;
    for (int a = 1; a < 3; ++a) if (cb.hi[a] - cb.lo[a] > cb.hi[axis] - cb.lo[axis]) axis = a;
    size_t mid = b;
    if (depth < PPM_BVH_SAH_DEPTH && cb.hi[axis] > cb.lo[axis]) {
      // binned SAH over the three axes
      constexpr int NB = 16;
      double best = INFINITY; int best_axis = -1, best_split = 0;
      for (int a = 0; a < 3; ++a) {
        const double ext = cb.hi[a] - cb.lo[a];
        if (!(ext > 0.0)) continue;
        Box bb[NB]; size_t cnt[NB];
        for (int k = 0; k < NB; ++k) { bb[k].clear(); cnt[k] = 0; }
        const double scale = NB / ext;
        for (size_t i = b; i < e; ++i) {
          int k = (int)((items[i].c[a] - cb.lo[a]) * scale);
          k = std::min(std::max(k, 0), NB - 1);
          bb[k].add(items[i].b); ++cnt[k];
        }
        double right_area[NB]; size_t right_cnt[NB];
        Box acc; acc.clear(); size_t n = 0;
        for (int k = NB - 1; k > 0; --k) { acc.add(bb[k]); n += cnt[k]; right_area[k] = acc.area(); right_cnt[k] = n; }
        acc.clear(); n = 0;
        for (int k = 0; k + 1 < NB; ++k) {
          acc.add(bb[k]); n += cnt[k];
          if (n == 0 || right_cnt[k + 1] == 0) continue;
          const double cost = acc.area() * (double)n + right_area[k + 1] * (double)right_cnt[k + 1];
          if (cost < best) { best = cost; best_axis = a; best_split = k; }
        }
      }
      if (best_axis >= 0) {
        const double ext = cb.hi[best_axis] - cb.lo[best_axis], scale = NB / ext, lo = cb.lo[best_axis];
        const int a = best_axis, ks = best_split;
        auto it = std::partition(items.begin() + (ptrdiff_t)b, items.begin() + (ptrdiff_t)e, [&](const Item& it_) {
          int k = (int)((it_.c[a] - lo) * scale);
          k = std::min(std::max(k, 0), NB - 1);
          return k <= ks;
        });
        mid = (size_t)(it - items.begin());
      }
    }
    if (mid == b || mid == e) {
      // object median along the widest centroid axis (also the fallback when all centroids coincide)
      mid = b + (e - b) / 2;
      std::nth_element(items.begin() + (ptrdiff_t)b, items.begin() + (ptrdiff_t)mid, items.begin() + (ptrdiff_t)e,
                       [axis](const Item& x, const Item& y) { return x.c[axis] < y.c[axis] || (x.c[axis] == y.c[axis] && x.prim < y.prim); });
    }
    const uint32_t me = (uint32_t)nodes.size();
    nodes.emplace_back();
    Box b0, b1;
    const uint32_t c0 = build(b, mid, depth + 1, b0);
    const uint32_t c1 = build(mid, e, depth + 1, b1);
    BvhNode& nd = nodes[me];
    std::memset(&nd, 0, sizeof nd);
    nd.child[0] = c0; nd.child[1] = c1;
    for (int a = 0; a < 3; ++a) {
      nd.box[0][a] = b0.lo[a] - pad; nd.box[0][3 + a] = b0.hi[a] + pad;
      nd.box[1][a] = b1.lo[a] - pad; nd.box[1][3 + a] = b1.hi[a] + pad;
    }
    return me;
  }
};
}  // namespace

// out_nodes: node 0 is the root (empty when the scene has no bounded primitive); out_prims: leaf order.
// Returns false (message in err) when the scene cannot be indexed.
bool bvh_build(const ppm_prim* prims, int64_t n, std::vector<BvhNode>& out_nodes, std::vector<BvhPrim>& out_prims, int* depth,
               std::string& err) {
  out_nodes.clear(); out_prims.clear();
  if (depth) *depth = 0;
  Builder B;
  double scale = 1e-3;
  for (int64_t i = 0; i < n; ++i) {
    const ppm_prim& s = prims[i];
    Item it; it.b.clear(); it.prim = (int32_t)i;
    if (s.type == PPM_SHAPE_SPHERE) {
      const double r = std::fabs(s.scalar);
      double lo[3], hi[3];
      for (int a = 0; a < 3; ++a) { lo[a] = s.position[a] - r; hi[a] = s.position[a] + r; }
      it.b.add(lo); it.b.add(hi);
    } else if (s.type == PPM_SHAPE_POLYGON || s.type == PPM_SHAPE_PARALLELOGRAM) {
      double p1[3], p2[3], p3[3];
      for (int a = 0; a < 3; ++a) { p1[a] = s.position[a] + s.dir1[a]; p2[a] = s.position[a] + s.dir2[a]; p3[a] = (s.position[a] + s.dir1[a]) + s.dir2[a]; }
      it.b.add(s.position); it.b.add(p1); it.b.add(p2);
      if (s.type == PPM_SHAPE_PARALLELOGRAM) it.b.add(p3);
    } else {
      continue;                                            // planes stay in the constant list; points are never hit
    }
    for (int a = 0; a < 3; ++a) {
      if (!std::isfinite(it.b.lo[a]) || !std::isfinite(it.b.hi[a])) { err = "primitive " + std::to_string(i) + " has a non-finite extent"; return false; }
      it.c[a] = 0.5 * (it.b.lo[a] + it.b.hi[a]);
      scale = std::max(scale, std::max(std::fabs(it.b.lo[a]), std::fabs(it.b.hi[a])));
    }
    B.items.push_back(it);
  }
  if (B.items.size() > PPM_BVH_MAX_PRIMS) { err = "more than 2^26 bounded primitives"; return false; }
  if (B.items.empty()) return true;
  B.pad = 1e-6 * scale;
  B.nodes.reserve(B.items.size());
  B.leaf_order.reserve(B.items.size());
  Box root;
  const uint32_t r = B.build(0, B.items.size(), 0, root);
  if (r & PPM_BVH_LEAF) {
    // a single leaf: give it a root node whose second child does not exist
    BvhNode nd;
    std::memset(&nd, 0, sizeof nd);
    nd.child[0] = r; nd.child[1] = PPM_BVH_NONE;
    for (int a = 0; a < 3; ++a) { nd.box[0][a] = root.lo[a] - B.pad; nd.box[0][3 + a] = root.hi[a] + B.pad; nd.box[1][a] = 1.0; nd.box[1][3 + a] = -1.0; }
    B.nodes.push_back(nd);
  }
  if (B.max_depth + 1 >= PPM_BVH_STACK) { err = "hierarchy deeper than the traversal stack"; return false; }
  out_nodes.swap(B.nodes);
  out_prims.resize(B.leaf_order.size());
  for (size_t k = 0; k < B.leaf_order.size(); ++k) {
    const ppm_prim& s = prims[B.items[B.leaf_order[k]].prim];
    BvhPrim& q = out_prims[k];
    std::memset(&q, 0, sizeof q);
    for (int a = 0; a < 3; ++a) { q.p0[a] = s.position[a]; q.d1[a] = s.dir1[a]; q.d2[a] = s.dir2[a]; }
    if (s.type == PPM_SHAPE_SPHERE) { q.d1[0] = s.scalar; q.d1[1] = q.d1[2] = 0.0; q.d2[0] = q.d2[1] = q.d2[2] = 0.0; }
    q.obj = B.items[B.leaf_order[k]].prim;
    q.type = s.type;
  }
  if (depth) *depth = B.max_depth + 1;
  return true;
}

}  // namespace ppmhost

// ---- C ABI: inspection of the hierarchy (host only; include/ppm.h) ---------------------------------------------
extern "C" int ppm_bvh_inspect(const ppm_prim* prims, int32_t nprims, int64_t* n_nodes, int64_t* n_leaf_prims, int32_t* depth,
                               double* sah_cost) {
  if (!prims || nprims <= 0) return PPM_ERR_ARG;
  std::vector<BvhNode> nodes;
  std::vector<BvhPrim> bp;
  int d = 0;
  std::string err;
  if (!ppmhost::bvh_build(prims, nprims, nodes, bp, &d, err)) return PPM_ERR_CAPACITY;
  int64_t bounded = 0;
  for (int32_t i = 0; i < nprims; ++i)
    if (prims[i].type == PPM_SHAPE_SPHERE || prims[i].type == PPM_SHAPE_POLYGON || prims[i].type == PPM_SHAPE_PARALLELOGRAM) ++bounded;
  if ((int64_t)bp.size() != bounded) return PPM_ERR_STATE;
  // self-check: walk the tree; every leaf primitive's corners lie inside every box on its path; every primitive once
  std::vector<uint8_t> seen((size_t)nprims, 0);
  double cost = 0.0, root_area = 0.0;
  struct Frame { uint32_t ref; double lo[3], hi[3]; int depth; };
  std::vector<Frame> st;
  auto area = [](const double* b) { const double x = b[3] - b[0], y = b[4] - b[1], z = b[5] - b[2]; return 2.0 * (x * y + y * z + z * x); };
  if (!nodes.empty()) {
    Frame f; f.ref = 0; f.depth = 0;
    for (int a = 0; a < 3; ++a) { f.lo[a] = -INFINITY; f.hi[a] = INFINITY; }
    st.push_back(f);
    double rb[6];
    for (int a = 0; a < 3; ++a) { rb[a] = std::min(nodes[0].box[0][a], nodes[0].child[1] == PPM_BVH_NONE ? INFINITY : nodes[0].box[1][a]);
                                  rb[3 + a] = std::max(nodes[0].box[0][3 + a], nodes[0].child[1] == PPM_BVH_NONE ? -INFINITY : nodes[0].box[1][3 + a]); }
    root_area = area(rb);
  }
  int maxd = 0;
  while (!st.empty()) {
    Frame f = st.back(); st.pop_back();
    maxd = std::max(maxd, f.depth);
    if (f.ref & PPM_BVH_LEAF) {
      const uint32_t first = f.ref & 0x0FFFFFFFu, cnt = ((f.ref >> 28) & 7u) + 1u;
      if (cnt > PPM_BVH_LEAF_MAX || first + cnt > bp.size()) return PPM_ERR_STATE;
      for (uint32_t k = 0; k < cnt; ++k) {
        const BvhPrim& q = bp[first + k];
        if (q.obj < 0 || q.obj >= nprims || seen[(size_t)q.obj]) return PPM_ERR_STATE;
        seen[(size_t)q.obj] = 1;
        const ppm_prim& s = prims[q.obj];
        if ((q.type & 0xff) != s.type) return PPM_ERR_STATE;
        double c[4][3]; int nc = 0;
        if (s.type == PPM_SHAPE_SPHERE) {
          for (int a = 0; a < 3; ++a) { c[0][a] = s.position[a] - std::fabs(s.scalar); c[1][a] = s.position[a] + std::fabs(s.scalar); }
          nc = 2;
        } else {
          for (int a = 0; a < 3; ++a) { c[0][a] = s.position[a]; c[1][a] = s.position[a] + s.dir1[a]; c[2][a] = s.position[a] + s.dir2[a];
                                        c[3][a] = (s.position[a] + s.dir1[a]) + s.dir2[a]; }
          nc = s.type == PPM_SHAPE_PARALLELOGRAM ? 4 : 3;
        }
        for (int j = 0; j < nc; ++j)
          for (int a = 0; a < 3; ++a)
            if (!(c[j][a] > f.lo[a] && c[j][a] < f.hi[a])) return PPM_ERR_STATE;      // strictly inside: the pad
      }
      continue;
    }
    if (f.ref >= nodes.size() || f.depth >= PPM_BVH_STACK) return PPM_ERR_STATE;
    const BvhNode& nd = nodes[f.ref];
    for (int k = 0; k < 2; ++k) {
      if (nd.child[k] == PPM_BVH_NONE) continue;
      Frame g; g.ref = nd.child[k]; g.depth = f.depth + 1;
      for (int a = 0; a < 3; ++a) { g.lo[a] = std::max(f.lo[a], nd.box[k][a]); g.hi[a] = std::min(f.hi[a], nd.box[k][3 + a]); }
      cost += area(nd.box[k]) * ((nd.child[k] & PPM_BVH_LEAF) ? (double)(((nd.child[k] >> 28) & 7u) + 1u) : 1.0);
      st.push_back(g);
    }
  }
  for (int32_t i = 0; i < nprims; ++i) {
    const bool b = prims[i].type == PPM_SHAPE_SPHERE || prims[i].type == PPM_SHAPE_POLYGON || prims[i].type == PPM_SHAPE_PARALLELOGRAM;
    if (b != (seen[(size_t)i] != 0)) return PPM_ERR_STATE;
  }
  if (n_nodes) *n_nodes = (int64_t)nodes.size();
  if (n_leaf_prims) *n_leaf_prims = (int64_t)bp.size();
  if (depth) *depth = maxd;
  if (sah_cost) *sah_cost = root_area > 0.0 ? cost / root_area : 0.0;
  return PPM_OK;
}
