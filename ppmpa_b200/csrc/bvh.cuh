// bvh.cuh -- traversal of the hierarchy of bvh_types.h (included by dev_core.cuh after the candidate tests).
//
// One thread walks one ray: depth-first, near child first, a 64-entry stack in local memory.  A node fetch is
// 8 x 16 bytes and decides both children with two f64 slab tests (12 DFMA + min/max); the boxes are padded
// (host_bvh.cpp) far beyond the rounding of the slab arithmetic, so FMA is allowed HERE (and only here): the slab test
// culls, it never produces a value.  Candidates go through consider_polygon / consider_sphere with their object index,
// i.e. the nearest hit and the tie rule are calc_intersection's (tracer.rs:306-350, :335-336).
//
// Subtrees whose entry distance is beyond the best candidate so far are skipped (also when popped: the stack keeps the
// entry distance), which is what makes the walk sub-linear; `<=` keeps equal distances, since a tie is decided by the
// object index.
#ifndef PPM_BVH_CUH_
#define PPM_BVH_CUH_

// true when the ray (o + t d, t >= 0) meets the box before `limit`; tn = entry distance.
// oi = o * inv.  min/max ignore a NaN operand (0 * inf when the origin lies in the plane of a face and the ray is
// parallel to it), which leaves the other slabs to decide.
__device__ __forceinline__ bool bvh_slab(const double* __restrict__ b, D3 inv, D3 oi, double limit, double& tn) {
  const double x1 = __fma_rn(b[0], inv.x, -oi.x), x2 = __fma_rn(b[3], inv.x, -oi.x);
  const double y1 = __fma_rn(b[1], inv.y, -oi.y), y2 = __fma_rn(b[4], inv.y, -oi.y);
  const double z1 = __fma_rn(b[2], inv.z, -oi.z), z2 = __fma_rn(b[5], inv.z, -oi.z);
  const double n = fmax(fmax(fmin(x1, x2), fmin(y1, y2)), fmin(z1, z2));
  const double f = fmin(fmin(fmax(x1, x2), fmax(y1, y2)), fmax(z1, z2));
  tn = n;
  return n <= f && f >= 0.0 && n <= limit;
}

__device__ __forceinline__ double bvh_inv(double d) { return fabs(d) > 1e-300 ? 1.0 / d : copysign(1e300, d); }

__device__ __forceinline__ void bvh_leaf(const DevScene& sc, uint32_t ref, D3 pos, D3 dir, double& best_t, int& best_o) {
  const uint32_t first = ref & 0x0FFFFFFFu, cnt = ((ref >> 28) & 7u) + 1u;
  for (uint32_t k = 0; k < cnt; ++k) {
    const double2* q = reinterpret_cast<const double2*>(sc.bprims + first + k);
    const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3), e = __ldg(q + 4);
    const long long meta = __double_as_longlong(e.y);
    const int obj = (int)(meta & 0xffffffffll), type = (int)(meta >> 32) & 0xff;
    const D3 p0 = mk3(a.x, a.y, b.x);
    if (type == PPM_SHAPE_SPHERE) consider_sphere(p0, b.y, pos, dir, obj, best_t, best_o);
    else consider_polygon(type == PPM_SHAPE_PARALLELOGRAM ? 2.0 : 1.0, p0, mk3(b.y, c.x, c.y), mk3(d.x, d.y, e.x), pos, dir, obj, best_t, best_o);
  }
}

// Tests the bounded primitives the ray can reach and folds them into (best_t, best_o) -- best_o < 0: no candidate yet.
// any_t2 >= 0 (shadow rays): the walk may stop at the first candidate with t^2 < any_t2 -- the caller only asks whether
// the nearest hit is nearer than that, and the nearest hit is at most as far as any candidate.
__device__ __forceinline__ void bvh_traverse(const DevScene& sc, D3 pos, D3 dir, double& best_t, int& best_o, double any_t2 = -1.0) {
  if (!sc.bvh) return;
  if (best_o >= 0 && best_t * best_t < any_t2) return;
  // 1/d, with |1/d| capped at 1e300 so that a ray parallel to an axis keeps a finite o * (1/d): its slab then
  // evaluates to -+huge on the two sides of the origin, as it should, instead of inf - inf
  const D3 inv = mk3(bvh_inv(dir.x), bvh_inv(dir.y), bvh_inv(dir.z));
  const D3 oi = mk3(pos.x * inv.x, pos.y * inv.y, pos.z * inv.z);
  uint32_t st_ref[PPM_BVH_STACK];
  float st_tn[PPM_BVH_STACK];
  int sp = 0;
  uint32_t cur = 0u;                                        // the root is an inner node
  for (;;) {
    if (cur & PPM_BVH_LEAF) {
      bvh_leaf(sc, cur, pos, dir, best_t, best_o);
      if (best_o >= 0 && best_t * best_t < any_t2) return;
    } else {
      const double2* q = reinterpret_cast<const double2*>(sc.bvh + cur);
      double box[12];
#pragma unroll
      for (int k = 0; k < 6; ++k) { const double2 v = __ldg(q + k); box[2 * k] = v.x; box[2 * k + 1] = v.y; }
      const uint2 ch = __ldg(reinterpret_cast<const uint2*>(q + 6));
      const double limit = best_o >= 0 ? best_t : 1.7976931348623157e308;
      double t0, t1;
      const bool h0 = bvh_slab(box, inv, oi, limit, t0);
      const bool h1 = ch.y != PPM_BVH_NONE && bvh_slab(box + 6, inv, oi, limit, t1);
      if (h0 && h1) {
        const bool first0 = t0 <= t1;
        st_ref[sp] = first0 ? ch.y : ch.x;
        st_tn[sp] = __double2float_rd(first0 ? t1 : t0);    // rounded down: the re-test when popped stays conservative
        ++sp;
        cur = first0 ? ch.x : ch.y;
        continue;
      }
      if (h0) { cur = ch.x; continue; }
      if (h1) { cur = ch.y; continue; }
    }
    // next subtree that can still hold a nearer (or equally near) candidate
    for (;;) {
      if (sp == 0) return;
      --sp;
      if (best_o < 0 || !((double)st_tn[sp] > best_t)) break;
    }
    cur = st_ref[sp];
  }
}

#endif
