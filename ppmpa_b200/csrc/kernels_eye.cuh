// kernels_eye.cuh -- camera rays, eye-path expansion, classic direct light, combine + accumulate
// Part of the single translation unit engine.cu (compiled -fmad=false, sm_100a); see DESIGN.md section 6.
#ifndef PPM_KERNELS_EYE_CUH_
#define PPM_KERNELS_EYE_CUH_

#include "dev_core.cuh"
#include "pass_state.cuh"

// ---- camera ---------------------------------------------------------------------
__device__ __forceinline__ void camera_ray(const ppm_camera& cam, int64_t pix, uint64_t seed, uint32_t pass, D3& pos, D3& dir) {
  Philox rng(seed, pass, PPM_DOMAIN_EYE, (uint64_t)pix, 0);
  // (32-bit division when the pixel index allows it: the 64-bit one is a ~100-instruction subroutine per pixel)
  int64_t py, px;
  if (pix >= 0 && pix < (int64_t)0x7fffffff) { const uint32_t q = (uint32_t)pix / (uint32_t)cam.xreso; py = q; px = (int64_t)((uint32_t)pix - q * (uint32_t)cam.xreso); }
  else { py = pix / cam.xreso; px = pix % cam.xreso; }
  double y = (double)py, x = (double)px;
  D3 blur = mk3(0.0, 0.0, 0.0);
  if (cam.blur) {
    double r1 = rng.range(-0.5, 0.5);
    double r2 = rng.range(-0.5, 0.5);
    blur = r1 * ld3(cam.eex) + r2 * ld3(cam.eey);
  }
  double r3 = 0.0, r4 = 0.0;
  if (cam.progressive && cam.antialias) { r3 = rng.range(-0.5, 0.5); r4 = rng.range(-0.5, 0.5); }
  pos = ld3(cam.eye_pos) + blur;
  D3 ed = ((ld3(cam.origin) + (x + r3) * ld3(cam.esx)) + (y + r4) * ld3(cam.esy)) - blur;
  dir = mk3(1.0, 0.0, 0.0);
  normalize(ed, dir);
}
__global__ void k_gen_rays(const __grid_constant__ ppm_camera cam, uint64_t seed, uint32_t pass, int64_t n,
                           double* __restrict__ rays6) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  D3 p, d;
  camera_ray(cam, i, seed, pass, p, d);
  st3(rays6 + i * 6, p); st3(rays6 + i * 6 + 3, d);
}

// ---- eye path expansion ------------------------------------------------------------
// The binary recursion of trace_ray becomes a per-pixel depth-first walk with an explicit
// stack and a top-down RGB throughput W.  Each visited node that has a non-zero diffuse
// coefficient becomes one "gather node" (hit point, normal, W (.) kd) in a global pool.
// Single pass: slots are claimed with one atomic per warp (opportunistic warp aggregation);
// every node stores the slot of the previous node of its pixel, and the pixel stores the last
// one, so k_combine can walk a pixel's nodes in a fixed order (reverse creation order) --
// the image does not depend on where the atomics placed the nodes.
struct EyeNodes {
  double* pos3;     // [cap][3] hit position     (gather / direct-light query)
  double* nrm3;     // [cap][3] facing normal
  double* w3;       // [cap][3] W (.) kd
  uint32_t* prev;   // [cap]    previous node of the same pixel, EYE_NONE = first
};
#define EYE_NONE 0xFFFFFFFFu
struct EyeStack {
  D3 pos, dir, W;
  int medium, depth;
  uint32_t node;
};
#ifndef PPM_EYE_MINB
#define PPM_EYE_MINB 6
#endif
template <bool BVH>
__global__ void __launch_bounds__(128, PPM_EYE_MINB)
k_eye_expand(const __grid_constant__ DevScene sc, const __grid_constant__ ppm_camera cam, const double* __restrict__ rays6,
             int64_t n, int64_t first_pixel, PassDev* ps, EyeNodes nodes, uint32_t cap,
             uint32_t* __restrict__ head, double* __restrict__ emit3, int classic, int stamp_slot) {
  stamp(ps, stamp_slot);
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t seed = ps->seed;                       // RNG keys and the node counter live in the pass state
  const uint32_t pass = ps->pass;
  unsigned long long* const pool_counter = &ps->n_nodes;
  unsigned long long* const n_visited = &ps->n_visited;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int64_t pix = first_pixel + i;
  EyeStack st[PPM_MAX_TRACE + 2];
  int sp = 0;
  if (rays6) { st[0].pos = ld3(rays6 + i * 6); st[0].dir = ld3(rays6 + i * 6 + 3); }
  else camera_ray(cam, pix, seed, pass, st[0].pos, st[0].dir);
  st[0].W = mk3(1.0, 1.0, 1.0); st[0].medium = -1; st[0].depth = 0; st[0].node = 1;
  sp = 1;
  D3 emit = mk3(0.0, 0.0, 0.0);
  uint32_t last = EYE_NONE, visited = 0;
  const double SR_HALF = 1.0 / (2.0 * PPM_PI);
  while (sp > 0) {
    EyeStack e = st[--sp];
    if (e.depth >= PPM_MAX_TRACE) continue;
    Isect is;
    if (!nearest_hit<BVH>(sc, e.pos, e.dir, is)) continue;
    ++visited;
    Philox rng(seed, pass, PPM_DOMAIN_EYE, (uint64_t)pix, e.node);
    EyeNode nd;
    eye_node(sc, is, e.dir, e.medium, rng, nd, classic != 0);
    const ppm_material& m = sc.mats[is.mat];
    emit = emit + cmul(e.W, ld3(m.emittance) * SR_HALF);
    const D3 wd = cmul(e.W, nd.kd);
    {
      const bool create = any_nz(wd);
      const unsigned conv = __activemask();                 // lanes that reached this point together
      const unsigned cm = __ballot_sync(conv, create);
      if (cm) {
        unsigned long long base = 0;
        const int leader = __ffs(cm) - 1;
        if ((int)lane == leader) base = atomicAdd(pool_counter, (unsigned long long)__popc(cm));
        base = __shfl_sync(conv, base, leader);
        if (create) {
          const unsigned long long s = base + __popc(cm & lt_mask);
          if (s < cap) {
            st3(nodes.pos3 + s * 3, is.pos); st3(nodes.nrm3 + s * 3, is.nvec); st3(nodes.w3 + s * 3, wd);
            nodes.prev[s] = last;
            last = (uint32_t)s;
          }                                                 // else: pool overflow, the host grows it and re-runs
        }
      }
    }
    // push the refract child first so that the reflect subtree is walked first
    // (reference order: si is evaluated before ti, tracer.rs:152-171)
    if (nd.refract) {
      D3 wt = cmul(e.W, nd.kt);
      if (any_nz(wt) && sp < PPM_MAX_TRACE + 2) {
        EyeStack& c = st[sp++];
        c.pos = is.pos; c.dir = nd.tdir; c.W = wt; c.medium = nd.t_medium; c.depth = e.depth + 1; c.node = e.node * 2 + 1;
      }
    }
    if (nd.reflect) {
      D3 ws = cmul(e.W, nd.ks);
      if (any_nz(ws) && sp < PPM_MAX_TRACE + 2) {
        EyeStack& c = st[sp++];
        c.pos = is.pos; c.dir = nd.rdir; c.W = ws; c.medium = e.medium; c.depth = e.depth + 1; c.node = e.node * 2;
      }
    }
  }
  head[i] = last;
  st3(emit3 + i * 3, emit);
  {
    const unsigned conv = __activemask();
    const unsigned v = __reduce_add_sync(conv, visited);
    if (lane == (unsigned)(__ffs(conv) - 1) && v) atomicAdd(n_visited, (unsigned long long)v);
  }
}

// ---- direct light: one thread per gather node, samples walked sequentially ----------
// get_radiance_from_light (tracer.rs:263-270) pairs [0, L(d0), L(d1), ...] with
// [c0, c1, c2, ...] (the RADIANCE0 seed of light.rs:132): the i-th surviving sample is
// weighted with the radiance of the (i-1)-th.  Walking the 25 samples in order inside one
// thread turns that pairing into a register recurrence and reproduces the reference's
// summation order.  Point and sun lights have a single sample, which is paired with the
// zero -> they contribute nothing and are skipped.
//
// Conservative per-node culling (cull_classify).  All 25 shadow rays of a node start at the node
// and end inside the light's quad, so most primitives can be ruled out ONCE per node instead of
// being tested 25 times.  `illuminated` (tracer.rs:272-290) only uses the nearest hit through
// "is there one" and "sq_ldist - |hit - p|^2 > 0.002", so a primitive may be skipped for a node
// when, for EVERY ray from the node to a point of the light quad, one of these is certain:
//  (a) it yields no candidate at all:
//      - bounded primitive whose bounding sphere (radius inflated by 1e-6) lies outside the cone
//        apex = node, through the light's bounding sphere, with a 2e-7 margin on the cosine -- ten
//        orders above the rounding of the tests;
//      - polygon / parallelogram that passes the sphere test but lies entirely outside one side plane of the
//        pyramid apex = node over the light quad (all four corners beyond the plane through the node and one
//        light edge, by an angular margin of 1e-6; only evaluated when the pyramid is well conditioned);
//      - the plane the node itself lies on: with h = dist + n.p (the very value calc_distance
//        computes, identical for all 25 rays) and h_g the same at the sample, den = (h - h_g)/ldist
//        +- 1e-15, so |t| <= |h| ldist / (|h_g| - |h|) < NEARLY0 / 2 when the light is well off
//        the plane: the root is always discarded by `t < NEARLY0`;
//  (b) its root can never be an occluder, i.e. never makes sq_ldist - |po|^2 exceed 0.002:
//      - a plane with the node and all four corners of the light quad on the same side by more than
//        D = 1e-6: the computed t is < 0, or > ldist (1 + 1e-8), or > 1e8;
//      - a polygon / parallelogram lying IN the light's plane (the emitter's own geometry): its root
//        is t = ldist (1 +- 1e-7/(1+L)) as long as the node is off that plane by 1e-6 (1+L).
// Primitives of class (b) can still be what makes calc_intersection return Some, so they are only
// skipped when some plane is a CERTIFICATE: node and light strictly on one side, the light closer by
// more than `gap` => every ray has a valid root (t > ldist >= NEARLY0) on it.  Then "no candidate
// among the tested primitives" means "nearest hit is at or beyond the light" = lit.  Without a
// certificate class (b) is tested like everything else.  Masks are OR-ed across the warp (testing
// extra primitives never changes a result), so the primitive loops stay warp-uniform.
// tests/test_gpu_parity.py::test_direct_light_cull_is_exact compares culled and unculled bit for bit.
struct CullPrim {
  double c[3];      // bounded: bounding-sphere centre | plane: c[0] = D, c[1] = gap
  double r;         // bounded: inflated radius
  int32_t kind;     // 0 = can never be hit (Point), 1 = plane, 2 = bounded (sphere, polygon, parallelogram), 3 = always tested
  int32_t nvtx;     // 4 for polygons / parallelograms (vtx = the parallelogram's corners, a superset of the triangle), else 0
  double vtx[4][3];
};
struct CullLight {
  double c[3], r;                                   // bounding sphere of the light quad (inflated)
  double nl[3];                                     // unit normal of the light's plane (through c)
  double corner[4][3];                              // the quad's corners in cyclic order: pos, +dir1, +dir1+dir2, +dir2
  unsigned long long coplanar;                      // polygons / parallelograms lying in that plane
  double hmin[PPM_MAX_PRIMS], hmax[PPM_MAX_PRIMS];  // planes: min / max of dist + n.corner over the quad's corners
  double side_e;                                    // |nvec.dir1| + |nvec.dir2|: spread of nvec.(sample - p) over the 25 samples
  double ext, ln1;                                  // |pos|_1 + |dir1|_1 + |dir2|_1 and |nvec|_1 (rounding scale of that test)
};
struct DevCull {
  CullPrim prim[PPM_MAX_PRIMS];
  CullLight light[PPM_MAX_LIGHTS];
};
__device__ __forceinline__ unsigned long long cull_classify(const DevScene& sc, const DevCull* __restrict__ cull, int li, D3 p,
                                                            unsigned long long all, bool& cert) {
  cert = false;
  const CullLight& cl = cull->light[li];
  const D3 u = ld3(cl.c) - p;
  const double uu = dot(u, u);
  if (!(uu < 1e6)) return all;                      // absurd scale or NaN: no culling
  const double rl = cl.r;
  const double ucone = uu - rl * rl;
  const bool cone = ucone > 0.0;                    // the node is outside the light's bounding sphere
  const double L = sqrt(uu) + rl;                   // every ldist is below this
  const bool off_light_plane = fabs(dot(ld3(cl.nl), u)) > 1e-6 * (1.0 + L);
  unsigned long long mask = 0, harmless = 0;
  int pyr_state = 0;                                // 0 = side planes not built yet, 1 = usable, 2 = ill conditioned
  D3 pn[4];
  double pnn[4];
  const int np = sc.nprims;
  for (int o = 0; o < np; ++o) {
    const CullPrim& cp = cull->prim[o];
    const unsigned long long bit = 1ull << o;
    if (cp.kind == 1) {
      const ppm_prim& s = sc.prims[o];
      const double num = s.scalar + dot(ld3(s.nvec), p);
      const double D = cp.c[0], gap = cp.c[1];
      const double hmin = cl.hmin[o], hmax = cl.hmax[o];
      if (num > D && hmin > D) { harmless |= bit; if (hmax < num - gap) cert = true; }
      else if (num < -D && hmax < -D) { harmless |= bit; if (hmin > num + gap) cert = true; }
      else {
        const double habs = hmin > D ? hmin : (hmax < -D ? -hmax : 0.0);   // light quad's distance from the plane
        const double an = fabs(num);
        if (!(habs > 0.0 && an * L < 0.4999e-4 * (habs - an))) mask |= bit;  // else: |t| < NEARLY0/2 for every ray
      }
    } else if (cp.kind == 2) {
      if (off_light_plane && ((cl.coplanar >> o) & 1ull)) { harmless |= bit; continue; }
      bool keep = true;
      if (cone) {
        const D3 v = ld3(cp.c) - p;
        const double vv = dot(v, v);
        const double vcone = vv - cp.r * cp.r;
        if (vcone > 0.0 && vv < 1e12) {             // the node is outside the primitive's bounding sphere
          // angle(u, v) > asin(rl/|u|) + asin(r/|v|)  <=>  u.v < sqrt((uu - rl^2)(vv - r^2)) - rl r
          // (without the square root: x < sqrt(a), a > 0  <=>  x < 0 or x^2 < a; the 1e-7 margin dwarfs the rounding)
          const double x = (dot(u, v) + rl * cp.r) + 1e-7 * (uu + vv);
          keep = !(x < 0.0 || x * x < ucone * vcone);
        }
      }
      if (keep && cp.nvtx == 4 && off_light_plane) {
        if (pyr_state == 0) {                           // side planes of the pyramid, built on first use
          pyr_state = 2;
          D3 a[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) a[j] = ld3(cl.corner[j]) - p;
          bool ok = true;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const D3 n = cross(a[j], a[(j + 1) & 3]);
            const double nn = dot(n, n);
            const double s = dot(n, a[(j + 2) & 3]);    // the opposite corner is inside
            // well conditioned: the two edge directions are not (nearly) collinear and the opposite corner is
            // clearly off the side plane, so the plane's orientation is known to ~1e-10 rad
            ok = ok && nn > 1e-12 * (dot(a[j], a[j]) * dot(a[(j + 1) & 3], a[(j + 1) & 3])) &&
                 s * s > 1e-12 * (nn * dot(a[(j + 2) & 3], a[(j + 2) & 3]));
            pn[j] = s < 0.0 ? -n : n;                   // oriented inward
            pnn[j] = nn;
          }
          if (ok) pyr_state = 1;
        }
        if (pyr_state == 1) {
          D3 w[4];
          double ww[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) { w[k] = ld3(cp.vtx[k]) - p; ww[k] = dot(w[k], w[k]); }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (keep) {                                  // (static indices: pn / pnn stay in registers)
              bool all_out = true;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const double d = dot(pn[j], w[k]);
                all_out = all_out && d < 0.0 && d * d > 1e-12 * (pnn[j] * ww[k]);   // outside by > 1e-6 rad
              }
              if (all_out) keep = false;
            }
          }
        }
      }
      if (keep) mask |= bit;
    } else if (cp.kind == 3) {
      mask |= bit;
    }
  }
  if (!cert) mask |= harmless;
  return mask;
}

// ---- BVH mode: the bounded primitives against the SHAFT of a node -------------------------------------------------
// All 25 shadow rays of a node run from the node p to points of the light's quad, i.e. inside the pyramid with apex p
// over the quad.  Whatever lies outside one of its four side planes (by the same > 1e-6 rad margin as the pyramid test
// of cull_classify) cannot be met by any of them.  One walk of the hierarchy with the pyramid instead of 25 walks with
// rays: a subtree is dropped when its (padded) box is outside a side plane; at a leaf the primitives' own corners
// (polygon: the parallelogram's four, a superset of the triangle) or box (sphere) are tested the same way.
// Returns 2 = some primitive may be met (test the hierarchy for this node), 1 = only primitives lying in the light's
// plane (class (b): harmless given a certificate), 0 = none.
__device__ __forceinline__ bool shaft_outside(const D3 pn[4], const double pnn[4], D3 lo, D3 hi) {   // lo, hi relative to p
  const double ww = (fmax(lo.x * lo.x, hi.x * hi.x) + fmax(lo.y * lo.y, hi.y * hi.y)) + fmax(lo.z * lo.z, hi.z * hi.z);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    // the largest n . w over the box's corners; all corners are outside when even that one is
    const double d = (fmax(pn[j].x * lo.x, pn[j].x * hi.x) + fmax(pn[j].y * lo.y, pn[j].y * hi.y)) + fmax(pn[j].z * lo.z, pn[j].z * hi.z);
    if (d < 0.0 && d * d > 1e-12 * (pnn[j] * ww)) return true;
  }
  return false;
}
// The walk visits at most PPM_SHAFT_BUDGET nodes and leaves: a shaft that grazes a finely tessellated surface would
// otherwise visit hundreds of boxes to prove what 25 ray walks find directly.  Measured on a 1 M-triangle glass sphere
// at 1080p (profiles/r2bvh_budget.txt): pass 32.5 ms without the classification, 38.2 ms with an unbounded walk (it costs
// 10 ms and saves 5), 35.2 / 33.6 / 30.6 / 28.8 ms with budgets 200 / 64 / 24 / 8.  Running out of budget answers
// "test", which is always safe.
#ifndef PPM_SHAFT_BUDGET
#define PPM_SHAFT_BUDGET 8
#endif
// squared distance from the origin to the box [lo, hi] (0 inside)
__device__ __forceinline__ double box_dist2(D3 lo, D3 hi) {
  const double x = fmax(fmax(lo.x, -hi.x), 0.0), y = fmax(fmax(lo.y, -hi.y), 0.0), z = fmax(fmax(lo.z, -hi.z), 0.0);
  return (x * x + y * y) + z * z;
}
__device__ __forceinline__ int bvh_shaft_classify(const DevScene& sc, const CullLight& cl, int li, D3 p) {
  if (!sc.bvh) return 0;
  const D3 u = ld3(cl.c) - p;
  const double uu = dot(u, u);
  if (!(uu < 1e6)) return 2;
  const double L = sqrt(uu) + cl.r;
  if (!(fabs(dot(ld3(cl.nl), u)) > 1e-6 * (1.0 + L))) return 2;      // the node lies (nearly) in the light's plane
  D3 pn[4];
  double pnn[4];
  {
    D3 a[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = ld3(cl.corner[j]) - p;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 4; ++j) {                          // as in cull_classify
      const D3 n = cross(a[j], a[(j + 1) & 3]);
      const double nn = dot(n, n);
      const double s = dot(n, a[(j + 2) & 3]);
      ok = ok && nn > 1e-12 * (dot(a[j], a[j]) * dot(a[(j + 1) & 3], a[(j + 1) & 3])) &&
           s * s > 1e-12 * (nn * dot(a[(j + 2) & 3], a[(j + 2) & 3]));
      pn[j] = s < 0.0 ? -n : n;
      pnn[j] = nn;
    }
    if (!ok) return 2;
  }
  int res = 0;
  uint32_t stack[PPM_BVH_STACK];
  int sp = 0, budget = PPM_SHAFT_BUDGET;
  uint32_t cur = 0u;
  for (;;) {
    if (--budget < 0) return 2;                            // not settled within the budget: let the rays decide
    if (cur & PPM_BVH_LEAF) {
      const uint32_t first = cur & 0x0FFFFFFFu, cnt = ((cur >> 28) & 7u) + 1u;
      for (uint32_t k = 0; k < cnt; ++k) {
        const double2* q = reinterpret_cast<const double2*>(sc.bprims + first + k);
        const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3), e = __ldg(q + 4);
        const int type = (int)(__double_as_longlong(e.y) >> 32);
        const D3 p0 = mk3(a.x, a.y, b.x) - p;
        bool out;
        if ((type & 0xff) == PPM_SHAPE_SPHERE) {
          const double r = fabs(b.y) * (1.0 + 1e-6) + 1e-6 * (1.0 + (fabs(a.x) + fabs(a.y) + fabs(b.x)));
          out = shaft_outside(pn, pnn, p0 - mk3(r, r, r), p0 + mk3(r, r, r));
        } else {
          const D3 d1 = mk3(b.y, c.x, c.y), d2 = mk3(d.x, d.y, e.x);
          D3 w[4];
          w[0] = p0; w[1] = p0 + d1; w[2] = (p0 + d1) + d2; w[3] = p0 + d2;
          out = false;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            bool all_out = true;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const double dd = dot(pn[j], w[v]);
              all_out = all_out && dd < 0.0 && dd * dd > 1e-12 * (pnn[j] * dot(w[v], w[v]));
            }
            out = out || all_out;
          }
        }
        if (out) continue;
        if ((type >> (8 + li)) & 1) res = 1;               // lies in the light's plane
        else return 2;
      }
    } else {
      const double2* q = reinterpret_cast<const double2*>(sc.bvh + cur);
      double box[12];
#pragma unroll
      for (int k = 0; k < 6; ++k) { const double2 v = __ldg(q + k); box[2 * k] = v.x; box[2 * k + 1] = v.y; }
      const uint2 ch = __ldg(reinterpret_cast<const uint2*>(q + 6));
      const D3 lo0 = mk3(box[0], box[1], box[2]) - p, hi0 = mk3(box[3], box[4], box[5]) - p;
      const D3 lo1 = mk3(box[6], box[7], box[8]) - p, hi1 = mk3(box[9], box[10], box[11]) - p;
      const bool h0 = !shaft_outside(pn, pnn, lo0, hi0);
      const bool h1 = ch.y != PPM_BVH_NONE && !shaft_outside(pn, pnn, lo1, hi1);
      if (h0 && h1) {
        // the child nearer to the apex first: a node ON the indexed surface finds its own neighbourhood at once, and the
        // walk ends with the first primitive that has to be tested
        const double e0 = box_dist2(lo0, hi0), e1 = box_dist2(lo1, hi1);
        const bool first0 = e0 <= e1;
        stack[sp++] = first0 ? ch.y : ch.x;
        cur = first0 ? ch.x : ch.y;
        continue;
      }
      if (h0) { cur = ch.x; continue; }
      if (h1) { cur = ch.y; continue; }
    }
    if (sp == 0) return res;
    cur = stack[--sp];
  }
}

#define PPM_CULL_CERT (1ull << 63)
#define PPM_CULL_BVH (1ull << 62)      // BVH mode: the node's shadow rays must walk the hierarchy
// Conservative "no sample of this light can ever reach the hit test" for a node: every sample fails
// `dot(lnv, d) < 0` (light.rs:112) or every sample has cos0 < 0 (tracer.rs:277-278).  The bands are > 1000 x the
// rounding of the reference's own tests, so this only fires when each of the 25 reference decisions is certain.
__device__ __forceinline__ bool light_never_tested(const ppm_light& l, const CullLight& cl, D3 p, D3 nv) {
  const D3 lp = ld3(l.pos) - p;
  const double p1 = fabs(p.x) + fabs(p.y) + fabs(p.z);
  const double a0 = dot(ld3(l.nvec), lp);
  if (a0 > 4.0 * (cl.side_e + 1e-12 * (cl.ln1 * (1.0 + cl.ext + p1)))) return true;
  const double x = dot(nv, ld3(l.dir1)), y = dot(nv, ld3(l.dir2));
  const double bmax = (dot(nv, lp) + fmax(0.1 * x, 0.9 * x)) + fmax(0.1 * y, 0.9 * y);
  const double n1 = fabs(nv.x) + fabs(nv.y) + fabs(nv.z);
  return bmax < -1e-11 * (n1 * (1.0 + cl.ext + p1));
}
// One thread per node: the conservative classification above, for every area light.  masks[li * cap + node] = the
// primitives the node's shadow rays towards light li must test, bit 63 = the node has a certificate.  A separate
// kernel so that it has its own register budget (k_direct_light is compiled for 64 registers) and can run right after
// the eye-path expansion, concurrently with the photon branch.  The node count comes from the pass state.
template <bool BVH>
#ifdef PPM_CLS_MINB                                   // tuning builds (tools/build_variants.sh): 5 / 6 / 8 CTAs per SM are all slower
__global__ void __launch_bounds__(128, PPM_CLS_MINB)
#else
__global__ void __launch_bounds__(128)
#endif
k_dl_classify(const __grid_constant__ DevScene sc, const DevCull* __restrict__ cull, PassDev* ps, uint32_t cap,
              const double* __restrict__ pos3, const double* __restrict__ nrm3, unsigned long long* __restrict__ masks, int stamp_slot) {
  stamp(ps, stamp_slot);
  const unsigned long long made = ps->n_nodes;
  const int64_t n = made > (unsigned long long)cap ? (int64_t)cap : (int64_t)made;
  const int64_t node = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= n) return;
  const D3 p = ld3(pos3 + node * 3);
  const D3 nv = ld3(nrm3 + node * 3);
  const unsigned long long all = (1ull << sc.nprims) - 1ull;       // nprims <= 62 here (bits 62 / 63: hierarchy / certificate)
  for (int li = 0; li < sc.nlights; ++li) {
    unsigned long long m = 0;
    if (sc.lights[li].type == PPM_LIGHT_PARALLELOGRAM && !light_never_tested(sc.lights[li], cull->light[li], p, nv)) {
      bool cert;
      m = cull_classify(sc, cull, li, p, all, cert);      // BVH mode: the scene's planes
      if (BVH) {
        const int sh = bvh_shaft_classify(sc, cull->light[li], li, p);
        if (sh == 2 || (sh == 1 && !cert)) m |= PPM_CULL_BVH;
      }
      if (cert) m |= PPM_CULL_CERT;
    }
    masks[(int64_t)li * cap + node] = m;
  }
}

__device__ __forceinline__ double ts5(unsigned i) {   // the literals 0.1, 0.3, 0.5, 0.7, 0.9 (light.rs:164-170)
  return i == 0 ? 0.1 : (i == 1 ? 0.3 : (i == 2 ? 0.5 : (i == 3 ? 0.7 : 0.9)));
}
// 64 registers (8 CTAs per SM): the kernel is latency bound, so occupancy beats the ~100 bytes of spills
// (96 registers / 5 CTAs: 2.1 ms, 64 / 8: 1.35 ms on config 2).
//
// Arithmetic (tolerance tier, 1e-12 held by the tests).  The reference evaluates, per surviving sample i,
//   rad += (color * (lnum / (4 pi |d_{i-1}|^2))) * cos0_i^2,   cos0_i = nvec . normalize(d_i)
// which costs a square root, a reciprocal, a division and nine multiplications per sample.  Here
//   rad = (color * lnum / (4 pi)) * sum_i (1 / |d_{i-1}|^2) * cos0_i^2,   cos0_i^2 = (nvec . d_i)^2 / |d_i|^2
// with ONE reciprocal per sample: no square root and no normalised direction unless the node has primitives to test
// (then the shadow ray needs the reference's exact direction, and cos0 is taken from it as the reference does).  The
// DECISIONS stay the reference's: `dot(lnv, d) < 0` is evaluated per sample unless the node is far enough from the
// light's plane that all 25 outcomes are certain (band > 1000 x the rounding), and the sign of nvec . d stands for
// the sign of cos0 only when |nvec . d| is 10^7 roundings away from zero; anything closer takes the exact path.
#ifndef PPM_DL_MINB
#define PPM_DL_MINB 8
#endif
#ifndef PPM_DL_UNROLL
#define PPM_DL_UNROLL 5      // samples per trip of the straight-line loop: 1 / 3 / 5 / 25 -> 930 / 893 / 867 / 950 us (profiles/r2_dl_straight_unroll.txt)
#endif
#define PPM_PRAGMA_(x) _Pragma(#x)
#define PPM_UNROLL(n) PPM_PRAGMA_(unroll n)
#ifndef PPM_DL_STRAIGHT
#define PPM_DL_STRAIGHT 1
#endif
template <bool BVH>
__global__ void __launch_bounds__(128, PPM_DL_MINB)
k_direct_light(const __grid_constant__ DevScene sc, PassDev* ps, uint32_t cap, const unsigned long long* __restrict__ masks,
               const uint32_t* __restrict__ order, const double* __restrict__ pos3, const double* __restrict__ nrm3,
               double* __restrict__ out3, unsigned long long* __restrict__ dbg, int stamp_slot) {
  __shared__ double s_gp[25][3];
  __shared__ double s_lc[3];
  stamp(ps, stamp_slot);
  const unsigned long long made = ps->n_nodes;
  const int64_t n = made > (unsigned long long)cap ? (int64_t)cap : (int64_t)made;
  if ((int64_t)blockIdx.x * blockDim.x >= n) return;     // block-uniform: grids are sized for the capacity
  const int64_t node0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = node0 < n;
  // `order` (optional): visit the nodes in this order -- a pass hands in the cell-sorted query order of the gather, so
  // the 32 nodes of a warp lie in the same or adjacent grid cells and have (almost) the same culling mask: the
  // warp-wide OR then costs nothing (2.35 -> 0.87 tested primitives per node on config 2).
  const int64_t slot = live ? node0 : n - 1;           // idle lanes of the last block shadow the last node (no store)
  const int64_t node = order ? (int64_t)order[slot] : slot;
  const D3 p = ld3(pos3 + node * 3), nv = ld3(nrm3 + node * 3);
  const double p1 = fabs(p.x) + fabs(p.y) + fabs(p.z);
  const double nn_thr = 1e-28 * dot(nv, nv);
  const unsigned long long all = sc.nprims >= 64 ? ~0ull : ((1ull << sc.nprims) - 1ull);
  const PrimMasks tmask = sc.types;
  D3 total = mk3(0.0, 0.0, 0.0);
  for (int li = 0; li < sc.nlights; ++li) {
    const ppm_light& l = sc.lights[li];
    if (l.type != PPM_LIGHT_PARALLELOGRAM) continue;
    const D3 lnv = ld3(l.nvec);
    __syncthreads();
    if (threadIdx.x < 25) {
      const unsigned s = threadIdx.x;
      const D3 gp = (ld3(l.pos) + ts5(s / 5) * ld3(l.dir1)) + ts5(s % 5) * ld3(l.dir2);   // gen_pos, light.rs:152-154
      s_gp[s][0] = gp.x; s_gp[s][1] = gp.y; s_gp[s][2] = gp.z;
    } else if (threadIdx.x == 32) {
      const D3 d1 = ld3(l.dir1), d2 = ld3(l.dir2);
      s_lc[0] = fabs(dot(lnv, d1)) + fabs(dot(lnv, d2));
      s_lc[1] = (fabs(l.pos[0]) + fabs(l.pos[1]) + fabs(l.pos[2])) + (fabs(d1.x) + fabs(d1.y) + fabs(d1.z)) + (fabs(d2.x) + fabs(d2.y) + fabs(d2.z));
      s_lc[2] = fabs(lnv.x) + fabs(lnv.y) + fabs(lnv.z);
    }
    __syncthreads();
    bool cert = false, own_none = false, walk = BVH;         // walk (BVH mode, warp-uniform): shadow rays go through the hierarchy
    unsigned long long mask = all;
    if (masks) {
      const unsigned long long own = masks[(int64_t)li * cap + node];   // k_dl_classify
      cert = (own & PPM_CULL_CERT) != 0ull;
      mask = own & ~(PPM_CULL_CERT | PPM_CULL_BVH);
      own_none = (own & ~PPM_CULL_CERT) == 0ull;              // a property of the node alone (the OR below is not)
      if (BVH) walk = __any_sync(0xffffffffu, (own & PPM_CULL_BVH) != 0ull);
      const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)mask);
      const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(mask >> 32));
      mask = ((unsigned long long)hi << 32) | lo;
      if (dbg && live) {                                      // diagnostic ("dl_stats"): primitives tested per node
        const unsigned long long o1 = own & ~PPM_CULL_CERT;
        atomicAdd(dbg, 1ull); atomicAdd(dbg + 1, (unsigned long long)__popcll(o1));
        atomicAdd(dbg + 2, (unsigned long long)__popcll(mask)); atomicAdd(dbg + 3, cert ? 1ull : 0ull);
        atomicAdd(dbg + 4 + min(__popcll(o1), 7), 1ull); atomicAdd(dbg + 12 + min(__popcll(mask), 7), 1ull);
      }
    }
    // which side of the light's plane: +1 = every sample passes `dot(lnv, d) < 0` (light.rs:112), -1 = none does,
    // 0 = too close to the plane to say, evaluate per sample
    const double a0 = dot(lnv, ld3(l.pos) - p);
    const double band = s_lc[0] + 1e-12 * (s_lc[2] * (1.0 + s_lc[1] + p1));
    const int side = a0 < -band ? 1 : (a0 > band ? -1 : 0);
    if (side >= 0) {
      PrimMasks pm;
      pm.plain = mask & tmask.plain; pm.sphere = mask & tmask.sphere; pm.poly = mask & tmask.poly; pm.para = mask & tmask.para;
      pm.nwords = tmask.nwords; pm._pad = 0;
      const bool need_ld = walk || mask != 0ull;              // warp-uniform (masks are OR-ed across the warp)
      const double C = (2.0 * l.flux * 0.2 * 0.2) / (PPM_PI * 4.0);   // 2 * flux * PARA_DIV^2 / (4 pi), light.rs:142
      double acc = 0.0, inv_prev = 0.0;
      bool have_prev = false;
#if PPM_DL_STRAIGHT
      if (!need_ld) {
        // No node of the warp has a primitive to test (two thirds of the warps in cell-sorted order): the same samples,
        // decisions and operations as the loop below, but as straight-line selects, so that the reciprocal chains of
        // consecutive samples overlap instead of waiting behind the `continue` branches.
PPM_UNROLL(PPM_DL_UNROLL)
        for (unsigned s = 0; s < 25; ++s) {
          const D3 d = mk3(s_gp[s][0], s_gp[s][1], s_gp[s][2]) - p;
          const double dd = dot(d, d);
          const double b = dot(nv, d);
          bool ok = cert && dd != 0.0 && (side != 0 || dot(lnv, d) < 0.0);
          const double inv = 1.0 / dd;
          double cc = (b * b) * inv;
          if (b * b > nn_thr * dd) {
            ok = ok && !(b < 0.0);
          } else if (ok) {                                    // grazing: the reference's own cos0 (rare)
            D3 ld = d;
            normalize(d, ld);
            const double cos0 = dot(nv, ld);
            ok = !(cos0 < 0.0);
            cc = cos0 * cos0;
          }
          if (ok && have_prev) acc = acc + inv_prev * cc;
          if (ok) { have_prev = true; inv_prev = inv; }
        }
      } else
#endif
      for (unsigned s = 0; s < 25; ++s) {
        const D3 gp = mk3(s_gp[s][0], s_gp[s][1], s_gp[s][2]);
        const D3 d = gp - p;
        if (side == 0 && !(dot(lnv, d) < 0.0)) continue;      // light.rs:112
        const double dd = dot(d, d);                          // sq_ldist
        if (dd == 0.0) continue;                              // normalize(d) is None exactly when |d|^2 == 0, tracer.rs:275-276
        const double inv = 1.0 / dd;
        const double b = dot(nv, d);
        double cc;
        D3 ld = d;
        // Which arithmetic a node gets depends on the node alone (own_none), never on the warp it sits in, so the
        // result is reproducible whatever order the nodes come in; the warp-uniform need_ld only decides whether the
        // normalised direction has to exist for the hit test.
        const bool lean = own_none && b * b > nn_thr * dd;
        if (need_ld || !lean) normalize(d, ld);
        if (lean) {
          if (b < 0.0) continue;
          cc = (b * b) * inv;
        } else {
          const double cos0 = dot(nv, ld);
          if (cos0 < 0.0) continue;
          cc = cos0 * cos0;
        }
        if (need_ld) {
          D3 hp;
          const int hit = BVH ? nearest_hit_masked_bvh(sc, p, ld, pm, walk, (dd - 0.002) * (1.0 - 1e-6), hp) : nearest_hit_masked(sc, p, ld, pm, hp);
          if (hit == 1) {
            const D3 po = hp - p;
            if (dd - dot(po, po) > 0.002) continue;
          } else if (hit == 2 || !cert) {
            continue;                                         // no hit counts as occluded, tracer.rs:282
          }                                                   // hit == 0 with a certificate: nearest hit beyond the light
        } else if (!cert) {
          continue;
        }
        if (have_prev) acc = acc + inv_prev * cc;             // the off-by-one pairing [0, L(d0), ...] . [c0, c1, ...]
        have_prev = true;
        inv_prev = inv;
      }
      total = total + mk3((l.color[0] * C) * acc, (l.color[1] * C) * acc, (l.color[2] * C) * acc);
    }
  }
  if (live) st3(out3 + node * 3, total);
}

// ---- combine + accumulate ------------------------------------------------------------
// pixel = sum_nodes W(.)kd (.) (direct + photon estimate) + sum emittance terms;
// then the pass image is added to the running sum (util/averager2.rb:49-62).
__global__ void k_combine(PassDev* ps, const uint32_t* __restrict__ head, const uint32_t* __restrict__ prev, const double* __restrict__ w3,
                          const double* __restrict__ direct3, const double* __restrict__ photon3,
                          const double* __restrict__ emit3, int64_t n, double* __restrict__ out3,
                          double* __restrict__ accum3, int64_t accum_first, D3 ambient, int stamp_slot) {
  stamp(ps, stamp_slot);
  if (accum3 && ps->status != 0u) accum3 = nullptr;      // a pass that overflowed a buffer does not count: the host re-renders it
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  D3 rad = ld3(emit3 + i * 3);
  for (uint32_t s = head[i]; s != EYE_NONE; s = prev[s]) {
    D3 di;
    if (photon3) {
      di = ld3(photon3 + (uint64_t)s * 3);
      if (direct3) di = ld3(direct3 + (uint64_t)s * 3) + di;   // di = direct + estimate, tracer.rs:136-145
    } else {
      di = ld3(direct3 + (uint64_t)s * 3) + ambient;           // classic: di = direct + cam.ambient, tracer.rs:234-238
    }
    rad = rad + cmul(ld3(w3 + (uint64_t)s * 3), di);
  }
  st3(out3 + i * 3, rad);
  if (accum3) {
    double* a = accum3 + (accum_first + i) * 3;
    a[0] += rad.x; a[1] += rad.y; a[2] += rad.z;
  }
}
// acc += other; other = 0   (merging the twin lane's accumulator, incl. the pass counter)
__global__ void k_accum_merge(double* __restrict__ acc, double* __restrict__ other, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { acc[i] += other[i]; other[i] = 0.0; }
}
// acc[0..n) += src; the pass counter acc[n] += npass   (ppm_accum_add: a saved sum image put back)
__global__ void k_accum_add(double* __restrict__ acc, const double* __restrict__ src, int64_t n, double npass) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) acc[i] += src[i];
  if (i == 0) acc[n] += npass;
}
__global__ void k_scale(const double* __restrict__ in, const double* __restrict__ npass, int64_t n, double* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] / npass[0];
}

#endif
