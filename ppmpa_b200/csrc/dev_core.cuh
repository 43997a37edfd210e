// dev_core.cuh -- device-side building blocks of the photon-mapping kernels.
//
// All arithmetic is IEEE binary64 in the reference's operation order
// (SURVEY.md Appendix C).  This translation unit is compiled with -fmad=false,
// so a*b+c is never contracted into DFMA (rustc/LLVM never contracts); double
// division and sqrt are IEEE-correct in CUDA.  exp/pow/sin/cos come from
// libdevice and may differ from glibc in the last bits: they only feed the
// tolerance tier (filters, Fresnel, sampled directions).
#ifndef PPM_DEV_CORE_CUH_
#define PPM_DEV_CORE_CUH_

#include "../../include/ppm.h"
#include "bvh_types.h"

#include <cuda_runtime.h>
#include <stdint.h>

#define PPM_MAX_PRIMS 64
#define PPM_MAX_MATS 48
#define PPM_MAX_LIGHTS 8
#define PPM_MAX_TRACE 10          // tracer.rs:27
#define PPM_NEARLY0 0.0001        // ray/mod.rs:15
#define PPM_PI 3.14159265358979323846264338327950288

// The whole scene travels as a __grid_constant__ kernel parameter (<32 KB):
// primitive loops are warp-uniform, so every read is a constant-cache broadcast.
// primitives by shape (bit o = primitive o); nwords = 1 if the scene has <= 32 primitives
struct PrimMasks { unsigned long long plain, sphere, poly, para; int nwords, _pad; };
//
// BVH mode (bvh_on; scenes beyond PPM_MAX_PRIMS primitives, or the "bvh" option): `prims` holds only the scene's
// infinite planes (nprims of them, unb_obj[] = their object indices), the bounded primitives are reached through the
// hierarchy in device memory (bvh.cuh) and `gprims` is the whole primitive array in object order.
struct DevScene {
  int32_t nprims, nmats, nlights, bvh_on;
  ppm_prim prims[PPM_MAX_PRIMS];
  ppm_material mats[PPM_MAX_MATS];
  ppm_light lights[PPM_MAX_LIGHTS];
  PrimMasks types;                 // filled by ppm_scene_set
  const BvhNode* bvh;              // node 0 = root; nullptr when the scene has no bounded primitive
  const BvhPrim* bprims;           // leaf order
  const ppm_prim* gprims;          // [nprims_total], object order
  int32_t nprims_total, _pad;
  int32_t unb_obj[PPM_MAX_PRIMS];
};
static_assert(sizeof(DevScene) < 32000, "scene must fit the kernel parameter space");

struct D3 {
  double x, y, z;
};
__device__ __forceinline__ D3 mk3(double x, double y, double z) { D3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ D3 ld3(const double* p) { return mk3(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(double* p, D3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ D3 operator-(D3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ D3 operator*(D3 a, double s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ D3 operator*(double s, D3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ D3 cmul(D3 a, D3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
// algebra.rs:146-148: (x*x' + y*y') + z*z'
__device__ __forceinline__ double dot(D3 a, D3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
// algebra.rs:174-180
__device__ __forceinline__ D3 cross(D3 a, D3 b) {
  return mk3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
// algebra.rs:151-158: v * (1/|v|); false for the zero vector (Option::None)
__device__ __forceinline__ bool normalize(D3 a, D3& out) {
  double n = sqrt(dot(a, a));
  if (n == 0.0) return false;
  out = a * (1.0 / n);
  return true;
}
__device__ __forceinline__ double chan(D3 a, int wl) { return wl == 0 ? a.x : (wl == 1 ? a.y : a.z); }
__device__ __forceinline__ bool any_nz(D3 a) { return a.x != 0.0 || a.y != 0.0 || a.z != 0.0; }

// ---------------------------------------------------------------------------
// Philox4x32-10 counter RNG replacing rand::thread_rng() (light.rs:74,
// physics.rs:241,324, algebra.rs:212,226, camera.rs:59).  One stream per
// photon path / eye ray / eye-path node:
//   key = (seed_lo, seed_hi ^ pass), ctr = (index_lo, index_hi, sub, domain<<24 | block)
// draw k uses block k>>1 and words (0,1) / (2,3);  u = (64 bits >> 11) * 2^-53.
// ---------------------------------------------------------------------------
#define PPM_DOMAIN_PHOTON 1u
#define PPM_DOMAIN_EYE 2u

struct Philox {
  uint32_t k0, k1, c0, c1, c2, dom, k;
  uint32_t w[4];
  __device__ __forceinline__ Philox(uint64_t seed, uint32_t pass, uint32_t domain, uint64_t index, uint32_t sub)
      : k0((uint32_t)seed), k1((uint32_t)(seed >> 32) ^ pass), c0((uint32_t)index), c1((uint32_t)(index >> 32)),
        c2(sub), dom(domain), k(0) {}
  __device__ __forceinline__ void block(uint32_t b) {
    uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = (dom << 24) | b;
    uint32_t ka = k0, kb = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
      uint32_t n0 = hi1 ^ x1 ^ ka, n2 = hi0 ^ x3 ^ kb;
      x0 = n0; x1 = lo1; x2 = n2; x3 = lo0;
      ka += 0x9E3779B9u; kb += 0xBB67AE85u;
    }
    w[0] = x0; w[1] = x1; w[2] = x2; w[3] = x3;
  }
  __device__ __forceinline__ double next01() {
    if ((k & 1u) == 0) block(k >> 1);
    uint32_t lo = (k & 1u) ? w[2] : w[0], hi = (k & 1u) ? w[3] : w[1];
    ++k;
    uint64_t bits = (((uint64_t)hi << 32) | lo) >> 11;
    return (double)bits * (1.0 / 9007199254740992.0);
  }
  // Rng::gen_range(lo, hi): half-open, defined as lo + (hi - lo) * u
  __device__ __forceinline__ double range(double lo, double hi) { return lo + (hi - lo) * next01(); }
};

// ---------------------------------------------------------------------------
// Nearest hit: calc_intersection, tracer.rs:306-350.
// Every object is tested; roots with t >= NEARLY0 are kept in (object, root)
// order and the reference stable-sorts by t and takes the first (:335-336), i.e.
// the minimum t with ties going to the earliest candidate.  The scan below runs
// shape by shape (bit loops over the scene's per-shape masks, no per-primitive
// dispatch) instead of in object order, so consider() keeps the tie rule
// explicitly: equal t -> lower object index (within one object the roots come
// in order, and the later equal root does not replace the earlier one).
// ---------------------------------------------------------------------------
struct Isect {
  D3 pos, nvec;
  double t;
  int obj, mat, io;   // io: 0 = In, 1 = Out (normal was flipped)
};

__device__ __forceinline__ void consider(double t, int o, double& best_t, int& best_o) {
  if (t < PPM_NEARLY0) return;          // `if i.0 < NEARLY0 { continue; }`
  if (best_o < 0 || t < best_t || (t == best_t && o < best_o)) { best_t = t; best_o = o; }
}

// consider(num / den) without the IEEE division whenever the outcome is certain.
// A candidate only matters if NEARLY0 <= t and (no best yet or t < best_t).  The exact
// quotient q = num/den satisfies fl(q) = q(1+e), |e| <= 2^-53, so with the 1e-9 guard
// bands below every skipped case is provably one that consider() would also discard; all
// other cases (and NaN/inf) take the exact division.  Results are bit-identical to the
// plain `consider(num / den, ...)`.
#define PPM_GUARD 1e-9
__device__ __forceinline__ void consider_ratio(double num, double den, int o, double& best_t, int& best_o) {
  const double an = fabs(num), ad = fabs(den);
  if (num == 0.0 && ad > 0.0 && ad < 1e300) return;                                          // t = +-0 < NEARLY0
  if (an > 0.0 && ad > 0.0 && an < 1e300 && ad < 1e300) {
    if ((num < 0.0) != (den < 0.0)) return;                                                  // t < 0 (or -0)
    if (an < ad * (PPM_NEARLY0 * (1.0 - PPM_GUARD))) return;                                 // t < NEARLY0
    if (best_o >= 0 && best_t < 1e300 && an > (ad * best_t) * (1.0 + PPM_GUARD)) return;     // t > best_t
  }
  consider(num / den, o, best_t, best_o);
}

// Polygon / parallelogram candidate for the nearest-hit scan (method_moller, geometry.rs:149-165: u, v, t
// are all computed before the rejection test): same decisions and the same t as the reference, but the three divisions are only executed when a
// guard-banded sign/magnitude test cannot settle u, v, u+v or t (see consider_ratio).
// Guard-banded classification of fl(x / det) against [0, 1]:
//   -1 = certainly rejected (u < 0 or u > 1), +1 = certainly inside, 0 = undecided (divide).
// Requires 1e-300 < |det| < 1e300; ad_hi = |det|(1+G), ad_lo = |det|(1-G) are shared by the u
// and v tests.  A negative quotient is only "certain" when it cannot underflow to -0 (which
// `u < 0.0` would not reject).
__device__ __forceinline__ int classify01(double x, double det, double ad, double ad_hi, double ad_lo) {
  const double ax = fabs(x);
  if (!(ax < 1e300)) return 0;                               // inf / NaN -> exact path
  if ((x < 0.0) != (det < 0.0)) return ax > ad * 1e-300 ? -1 : 0;
  if (ax > ad_hi) return -1;
  return ax < ad_lo ? 1 : 0;
}
__device__ __forceinline__ void consider_polygon(double l, D3 p0, D3 d1, D3 d2, D3 p, D3 d, int o, double& best_t, int& best_o) {
  const D3 re2 = cross(d, d2);
  const double det = dot(re2, d1);
  if (det == 0.0) return;                                    // `det_a == 0.0 ||` -> None
  const D3 pp = p - p0;
  const double a = dot(re2, pp);                             // u = a / det
  const double ad = fabs(det);
  const double ad_hi = ad * (1.0 + PPM_GUARD), ad_lo = ad * (1.0 - PPM_GUARD);
  bool exact = !(ad < 1e300 && ad > 1e-300);
  if (!exact) {
    const int cu = classify01(a, det, ad, ad_hi, ad_lo);
    if (cu < 0) return;
    exact = cu == 0;
  }
  const D3 te1 = cross(pp, d1);
  const double b = dot(te1, d);                              // v = b / det
  if (!exact) {
    const int cv = classify01(b, det, ad, ad_hi, ad_lo);
    if (cv < 0) return;
    exact = cv == 0;
    if (!exact) {
      const double sum = fabs(a) + fabs(b);                  // u + v against l (both quotients are >= 0 here)
      if (sum > ad_hi * l) return;
      exact = !(sum < ad_lo * l);
    }
  }
  const double c = dot(te1, d2);                             // t = c / det
  if (exact) {
    const double u = a / det, v = b / det, t = c / det;
    if (u < 0.0 || u > 1.0 || v < 0.0 || v > 1.0 || u + v > l) return;
    consider(t, o, best_t, best_o);
    return;
  }
  consider_ratio(c, det, o, best_t, best_o);
}

// geometry.rs:179-193
__device__ __forceinline__ void consider_sphere(D3 center, double rad, D3 pos, D3 dir, int o, double& best_t, int& best_o) {
  D3 oc = center - pos;
  double t0 = dot(oc, dir);
  double t1 = rad * rad - (dot(oc, oc) - (t0 * t0));
  if (t1 > 0.0) {
    double t2 = sqrt(t1);
    if (t2 == 0.0) consider(t0, o, best_t, best_o);
    else { consider(t0 - t2, o, best_t, best_o); consider(t0 + t2, o, best_t, best_o); }
  }
}

#define PPM_FOR_EACH_BIT(mask64, o)                                                        \
  if (mask64) _Pragma("unroll 1") for (int w__ = 0; w__ < pm.nwords; ++w__)                \
    for (unsigned m__ = (unsigned)((mask64) >> (32 * w__)), o = 0; m__ && ((o = __ffs(m__) - 1 + 32 * w__), true); m__ &= m__ - 1)
__device__ __forceinline__ void scan_prims(const DevScene& sc, const PrimMasks& pm, D3 pos, D3 dir, double& best_t, int& best_o) {
  PPM_FOR_EACH_BIT(pm.plain, o) {
    // geometry.rs:170-177: t = (dist + n.pos) / -cos0
    const ppm_prim& s = sc.prims[o];
    D3 n = ld3(s.nvec);
    double cos0 = dot(n, dir);
    if (cos0 != 0.0) consider_ratio(s.scalar + dot(n, pos), -cos0, (int)o, best_t, best_o);
  }
  PPM_FOR_EACH_BIT(pm.para, o) {
    // geometry.rs:195-202, l = 2 (parallelogram), :141-143
    const ppm_prim& s = sc.prims[o];
    consider_polygon(2.0, ld3(s.position), ld3(s.dir1), ld3(s.dir2), pos, dir, (int)o, best_t, best_o);
  }
  PPM_FOR_EACH_BIT(pm.poly, o) {
    const ppm_prim& s = sc.prims[o];
    consider_polygon(1.0, ld3(s.position), ld3(s.dir1), ld3(s.dir2), pos, dir, (int)o, best_t, best_o);
  }
  PPM_FOR_EACH_BIT(pm.sphere, o) {
    // geometry.rs:179-193
    const ppm_prim& s = sc.prims[o];
    consider_sphere(ld3(s.position), s.scalar, pos, dir, (int)o, best_t, best_o);
  }
}

#include "bvh.cuh"

// Nearest hit: calc_intersection, tracer.rs:306-350.  Every object is tested; roots with t >= NEARLY0
// are kept and the reference stable-sorts by t and takes the first.
// BVH = true (a scene in BVH mode): the planes from the constant list, then the hierarchy; same candidates' arithmetic,
// same tie rule on the object index, so the result is the brute-force scan's.
template <bool BVH = false>
__device__ __forceinline__ bool nearest_hit(const DevScene& sc, D3 pos, D3 dir, Isect& is) {
  double best_t = 0.0;
  int best_o = -1;
  scan_prims(sc, sc.types, pos, dir, best_t, best_o);
  if (BVH) {
    if (best_o >= 0) best_o = sc.unb_obj[best_o];          // plane order = object order, so ties among planes were right
    bvh_traverse(sc, pos, dir, best_t, best_o);
  }
  if (best_o < 0) return false;
  const ppm_prim& s = BVH ? sc.gprims[best_o] : sc.prims[best_o];
  D3 p = pos + dir * best_t;            // Ray::target, geometry.rs:56-58
  D3 n;
  if (s.type == PPM_SHAPE_SPHERE) {
    if (!normalize(p - ld3(s.position), n)) return false;   // get_normal -> None
  } else {
    n = ld3(s.nvec);
  }
  is.pos = p; is.t = best_t; is.obj = best_o; is.mat = s.material;
  if (dot(n, dir) > 0.0) { is.nvec = -n; is.io = 1; }
  else { is.nvec = n; is.io = 0; }
  return true;
}
// Shadow ray of a scene in BVH mode: what `illuminated` (tracer.rs:272-290) consumes of calc_intersection, over the planes
// in `pm` (indices into the constant list) and, when `walk`, the hierarchy.
// Returns 0 = no candidate, 1 = hit (hit_pos set), 2 = calc_intersection is None because get_normal failed.
// occl_t2: `illuminated` only asks whether sq_ldist - |hit - p|^2 > 0.002 for the NEAREST hit (dir is a unit vector,
// so |hit - p| = t).  A candidate with t^2 < occl_t2 = (sq_ldist - 0.002)(1 - 1e-6) already says yes, for itself and
// for whatever is nearer (the 1e-6 margin is ten orders above the rounding of |hit - p|^2): the walk stops there and the
// caller reaches the same verdict with this candidate's position as it would with the nearest one's.
__device__ __forceinline__ int nearest_hit_masked_bvh(const DevScene& sc, D3 pos, D3 dir, const PrimMasks& pm, bool walk, double occl_t2,
                                                      D3& hit_pos) {
  double best_t = 0.0;
  int best_o = -1;
  scan_prims(sc, pm, pos, dir, best_t, best_o);
  if (best_o >= 0) best_o = sc.unb_obj[best_o];
  if (walk) bvh_traverse(sc, pos, dir, best_t, best_o, occl_t2);
  if (best_o < 0) return 0;
  hit_pos = pos + dir * best_t;
  const ppm_prim& s = sc.gprims[best_o];
  if (s.type == PPM_SHAPE_SPHERE) {
    D3 n;
    if (!normalize(hit_pos - ld3(s.position), n)) return 2;
  }
  return 1;
}

// Shadow-ray variant for k_direct_light: only the primitives in `pm` (the conservative per-node
// classification of kernels_eye.cuh, cull_classify) and only the data `illuminated` (tracer.rs:272-290)
// consumes: whether calc_intersection returns Some, and the hit position.
// Returns 0 = no candidate among the masked primitives, 1 = hit (hit_pos set), 2 = calc_intersection is None
// because get_normal failed.
__device__ __forceinline__ int nearest_hit_masked(const DevScene& sc, D3 pos, D3 dir, const PrimMasks& pm, D3& hit_pos) {
  double best_t = 0.0;
  int best_o = -1;
  scan_prims(sc, pm, pos, dir, best_t, best_o);
  if (best_o < 0) return 0;
  hit_pos = pos + dir * best_t;
  const ppm_prim& s = sc.prims[best_o];
  if (s.type == PPM_SHAPE_SPHERE) {
    D3 n;
    if (!normalize(hit_pos - ld3(s.position), n)) return 2;       // get_normal -> None
  }
  return 1;
}

// ---------------------------------------------------------------------------
// physics.rs
// ---------------------------------------------------------------------------
__device__ __forceinline__ double relative_ior(double ior1, double ior2) { return ior1 == 0.0 ? 1.0 : ior2 / ior1; }  // :200-205

// physics.rs:214-221
__device__ __forceinline__ void specular_reflection(D3 nvec, D3 vvec, D3& rvec, double& cos1) {
  double c = -dot(vvec, nvec);
  if (c < 0.0) { rvec = nvec; cos1 = -c; return; }
  D3 r = mk3(1.0, 0.0, 0.0);
  normalize(vvec + (2.0 * c) * nvec, r);
  rvec = r; cos1 = c;
}
// physics.rs:232-261 (two draws: xi0 then the azimuth)
__device__ __forceinline__ D3 reflection_glossy(D3 nvec, D3 rvec, double pw, Philox& rng) {
  D3 uvec = mk3(1.0, 0.0, 0.0);
  if (!normalize(cross(mk3(0.00424, 1.0, 0.00764), rvec), uvec))
    normalize(cross(mk3(1.0, 0.00424, 0.00764), rvec), uvec);
  D3 vvec = cross(uvec, rvec);
  double c0 = dot(nvec, rvec);
  double xi0 = rng.range(0.0, 1.0);
  double xi1 = pow(xi0, pw * c0);
  double xi2 = 2.0 * PPM_PI * rng.range(0.0, 1.0);
  double sn, cs;
  sincos(xi2, &sn, &cs);
  double x = cs * sqrt(1.0 - xi1 * xi1);
  double y = xi1;
  double z = sn * sqrt(1.0 - xi1 * xi1);
  D3 wi = (x * uvec + y * rvec) + z * vvec;
  if (dot(nvec, wi) < 0.0) wi = ((-x) * uvec + y * rvec) - z * vvec;
  D3 o;
  if (normalize(wi, o)) return o;
  return mk3(1.0, 0.0, 0.0);
}
// physics.rs:270-285; false = None
__device__ __forceinline__ bool specular_refraction(D3 nvec, D3 vvec, double eta, D3& tvec, double& cos2) {
  double cos1 = -dot(vvec, nvec);
  cos2 = 0.0;
  if (cos1 < 0.0) return false;
  double g0 = eta * eta + cos1 * cos1 - 1.0;
  if (g0 < 0.0) return false;
  double g = sqrt(g0);
  bool ok = normalize((1.0 / eta) * (vvec + (cos1 - g) * nvec), tvec);
  cos2 = g / eta;
  // a NaN direction (eta == 0 at exactly normal incidence) is treated as None
  return ok && tvec.x == tvec.x && tvec.y == tvec.y && tvec.z == tvec.z;
}
// physics.rs:316-318 / surface.rs:512-515: (1 - c)^5 via powf
__device__ __forceinline__ double pow5(double c) { return pow(1.0 - c, 5.0); }
__device__ __forceinline__ double schlick(double f0, double c) { return f0 + (1.0 - f0) * pow5(c); }
// physics.rs:323-335 with a single threshold: 0 if p <= p0 else 1
__device__ __forceinline__ int roulette(double p0, Philox& rng) { return rng.range(0.0, 1.0) > p0 ? 1 : 0; }

// algebra.rs:225-235 + surface.rs:408-417 (uniform, not cosine weighted)
__device__ __forceinline__ D3 diffuse_reflection(D3 n, Philox& rng) {
  double phi = rng.range(0.0, 2.0 * PPM_PI);
  double xi = rng.range(-1.0, 1.0);
  double xi2 = sqrt(1.0 - xi * xi);      // xi.powf(2.0) == xi*xi exactly
  double sn, cs;
  sincos(phi, &sn, &cs);
  D3 d = mk3(1.0, 0.0, 0.0);
  normalize(mk3(xi2 * cs, xi, xi2 * sn), d);
  return dot(n, d) > 0.0 ? d : -d;
}

// ---------------------------------------------------------------------------
// surface.rs predicates
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool is3(const double c[3], double v) { return c[0] == v && c[1] == v && c[2] == v; }
// surface.rs:68-100
__device__ __forceinline__ bool surf_reflect(const ppm_material& m, double c) {
  if (m.surface == PPM_SURF_SIMPLE) return !(m.p0 == 1.0 || (c == 1.0 && is3(m.color_b, 0.0)));
  if (m.surface == PPM_SURF_TS) return m.metalness == 1.0 ? !is3(m.color_b, 0.0) : true;
  return false;
}
// surface.rs:102-133
__device__ __forceinline__ bool surf_refract(const ppm_material& m, double c) {
  if (m.surface == PPM_SURF_SIMPLE) return !(c == 0.0 && is3(m.color_b, 1.0));
  if (m.surface == PPM_SURF_TS) return m.metalness == 0.0 && m.p0 < 1.0 && !is3(m.color_a, 0.0);
  return false;
}
// surface.rs:289-310
__device__ __forceinline__ bool surf_store_photon(const ppm_material& m) {
  if (m.surface == PPM_SURF_SIMPLE) return m.p0 > 0.0;
  if (m.surface == PPM_SURF_TS) return m.metalness != 1.0 && m.p0 != 0.0;
  return true;
}
__device__ __forceinline__ double surf_power_glossy(const ppm_material& m) {   // surface.rs:381-402
  return (m.surface == PPM_SURF_SIMPLE || m.surface == PPM_SURF_TS) ? m.density_pow : 0.0;
}

// ---------------------------------------------------------------------------
// Emission: Light::generate_photon, light.rs:67-91
// ---------------------------------------------------------------------------
__device__ __forceinline__ int decide_wavelength(const double c[3], double p) {   // physics.rs:74-84
  if (p < c[0]) return PPM_WL_RED;
  if (p < c[0] + c[1]) return PPM_WL_GREEN;
  return PPM_WL_BLUE;
}
// defer = true: for an area light the direction (the LAST draws of the emission) is not sampled here;
// the function returns true and the caller samples diffuse_reflection(l.nvec) from the same stream.
template <bool DEFER>
__device__ __forceinline__ bool generate_photon_t(const ppm_light& l, Philox& rng, int& wl, D3& pos, D3& dir) {
  wl = decide_wavelength(l.color, rng.range(0.0, 1.0));       // select_wavelength, light.rs:157-160
  if (l.type == PPM_LIGHT_POINT) {
    pos = ld3(l.pos);
    for (;;) {                                                 // generate_random_dir, algebra.rs:211-223
      double x = rng.range(-1.0, 1.0), y = rng.range(-1.0, 1.0), z = rng.range(-1.0, 1.0);
      D3 v = mk3(x, y, z);
      double len = sqrt(dot(v, v));
      if (0.0 < len && len <= 1.0) { normalize(v, dir); break; }
    }
  } else {
    double t1 = rng.range(0.0, 1.0);
    double t2 = rng.range(0.0, 1.0);
    pos = (ld3(l.pos) + t1 * ld3(l.dir1)) + t2 * ld3(l.dir2);
    if (l.type == PPM_LIGHT_PARALLELOGRAM) {
      if (DEFER) { dir = ld3(l.nvec); return true; }
      dir = diffuse_reflection(ld3(l.nvec), rng);
    } else {
      dir = ld3(l.dir);
    }
  }
  return false;
}
__device__ __forceinline__ void generate_photon(const ppm_light& l, Philox& rng, int& wl, D3& pos, D3& dir) {
  generate_photon_t<false>(l, rng, wl, pos, dir);
}

// ---------------------------------------------------------------------------
// One bounce of trace_photon, tracer.rs:31-125: decides how the path goes on.
// Returns true if the photon continues with (dir, medium); false = absorbed.
// medium: material index, -1 = M_AIR (scene.rs:13-18, ior 1/1/1).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double medium_ior(const DevScene& sc, int medium, int wl) {
  return medium < 0 ? 1.0 : sc.mats[medium].ior[wl];
}
// need_diffuse: set when the photon goes on diffusely; the direction -- diffuse_reflection(is.nvec), always the
// LAST draws of the bounce -- is then left to the caller (k_trace_photons samples it together with the emission
// directions of freshly regenerated lanes, so the expensive sincos/sqrt/normalize runs on full warps).
__device__ __forceinline__ bool photon_bounce(const DevScene& sc, const Isect& is, int wl, D3 in_dir, Philox& rng,
                                              int& medium, D3& out_dir, bool& need_diffuse) {
  need_diffuse = false;
  const ppm_material& m = sc.mats[is.mat];
  if (m.surface == PPM_SURF_SIMPLE) {
    // roughness() of a Simple surface is its *diffuseness* (surface.rs:358-367)
    if (roulette(m.p0, rng) == 0) {
      // reflect_diff, tracer.rs:83-92
      if (roulette(m.color_a[wl], rng) != 0) return false;
      need_diffuse = true;
      return true;
    }
    // reflect_spec, tracer.rs:94-109
    D3 rdir; double cos1;
    specular_reflection(is.nvec, in_dir, rdir, cos1);
    double f = schlick(m.color_b[wl], cos1);
    if (roulette(f, rng) == 0) { out_dir = rdir; return true; }
    if (m.ior[wl] == 0.0) return false;
    // reflect_trans, tracer.rs:111-125
    double eta = relative_ior(medium_ior(sc, medium, wl), m.ior[wl]);
    D3 tdir; double cos2;
    if (!specular_refraction(is.nvec, in_dir, eta, tdir, cos2)) return false;
    medium = dot(tdir, is.nvec) < 0.0 ? is.mat : -1;
    out_dir = tdir;
    return true;
  }
  if (m.surface == PPM_SURF_TS) {
    // Surface::next_direction, surface.rs:211-258
    double eta = relative_ior(medium_ior(sc, medium, wl), m.ior[wl]);
    D3 rdir0; double cos1;
    specular_reflection(is.nvec, in_dir, rdir0, cos1);
    D3 rdir = reflection_glossy(is.nvec, rdir0, m.density_pow, rng);
    D3 hvec = mk3(1.0, 0.0, 0.0);
    normalize(rdir - in_dir, hvec);
    D3 tdir; double cos2;
    bool has_t = specular_refraction(hvec, in_dir, eta, tdir, cos2);
    double c = cos1 < cos2 ? cos1 : cos2;
    double f = schlick(m.color_b[wl], c);
    if (roulette(f, rng) == 0) { out_dir = rdir; return true; }
    if (roulette(m.color_a[wl], rng) == 1) return false;
    if (roulette(m.p0, rng) == 0) { need_diffuse = true; return true; }
    if (!has_t) return false;
    medium = is.mat;              // `if m == true { m0 } else { &is1.mate }`, tracer.rs:68
    out_dir = tdir;
    return true;
  }
  return false;                   // Nothing / DisneyBRDF / Brady: `_ => vec![]`
}

// ---------------------------------------------------------------------------
// Eye path node: the per-node part of trace_ray, tracer.rs:129-177, turned
// top-down.  bsdf() (surface.rs:135-206) is linear in (di, si, ti), so
//   L = emittance/(2 pi) + kd (.) di + ks (.) si + kt (.) ti
// with per-channel coefficients known before the children are traced.
// ---------------------------------------------------------------------------
struct EyeNode {
  D3 kd, ks, kt;      // coefficients of di, si, ti
  D3 rdir, tdir;
  bool reflect, refract;
  int t_medium;
};
// classic = trace_ray_classic (tracer.rs:221-259, the `rtc` binary): mirror direction without the
// glossy lobe (no random draws), refraction about the geometric normal, Fresnel from cos1, and the
// next medium chosen by the sign of tdir.nvec.
__device__ __forceinline__ void eye_node(const DevScene& sc, const Isect& is, D3 in_dir, int medium, Philox& rng, EyeNode& nd,
                                         bool classic = false) {
  const ppm_material& m = sc.mats[is.mat];
  if (m.surface == PPM_SURF_SIMPLE && m.p0 == 1.0) {
    // Purely diffuse Simple surface: ks = (1 - p0) f = 0 and kt = (1 - p0)(1 - metalness) f2 = 0 whatever the
    // Fresnel term, so neither child is ever traced (tracer.rs:152-171 multiply their radiance by zero) and
    // kd does not depend on the directions: skip the glossy lobe, the refraction and both pow() calls.  The
    // node's random draws come from its own Philox stream, so not drawing them changes nothing else.
    nd.kd = m.p0 * (ld3(m.color_a) * (1.0 / PPM_PI));
    nd.ks = nd.kt = mk3(0.0, 0.0, 0.0);
    nd.rdir = nd.tdir = mk3(1.0, 0.0, 0.0);
    nd.reflect = nd.refract = false;
    nd.t_medium = -1;
    return;
  }
  D3 rdir0; double cos1;
  specular_reflection(is.nvec, in_dir, rdir0, cos1);
  nd.rdir = classic ? rdir0 : reflection_glossy(is.nvec, rdir0, surf_power_glossy(m), rng);
  nd.reflect = surf_reflect(m, cos1);
  // relative_ior_average, physics.rs:192-196
  double a1 = medium < 0 ? (1.0 + 1.0 + 1.0) / 3.0 : (sc.mats[medium].ior[0] + sc.mats[medium].ior[1] + sc.mats[medium].ior[2]) / 3.0;
  double a2 = (m.ior[0] + m.ior[1] + m.ior[2]) / 3.0;
  double eta = relative_ior(a1, a2);
  D3 hvec = is.nvec;
  bool hv = classic ? true : normalize(nd.rdir - in_dir, hvec);
  double cos2;
  bool has_t = specular_refraction(hvec, in_dir, eta, nd.tdir, cos2);
  nd.refract = hv && has_t && surf_refract(m, cos1);
  nd.t_medium = is.io == 0 ? is.mat : -1;          // tracer.rs:164-167
  if (classic && has_t) nd.t_medium = dot(nd.tdir, is.nvec) < 0.0 ? is.mat : -1;   // tracer.rs:251
  double c = classic ? cos1 : (cos1 < cos2 ? cos1 : cos2);                          // tracer.rs:258 vs :173
  const double ONE_PI = 1.0 / PPM_PI;
  nd.kd = nd.ks = nd.kt = mk3(0.0, 0.0, 0.0);
  if (m.surface == PPM_SURF_SIMPLE) {
    // diffuseness*(reflectance*ONE_PI*di) + (1-diffuseness)*(f*si + (1-metalness)*f2*ti)
    double c2 = pow5(c);
    D3 spec = ld3(m.color_b);
    D3 f = mk3(spec.x + (1.0 - spec.x) * c2, spec.y + (1.0 - spec.y) * c2, spec.z + (1.0 - spec.z) * c2);
    D3 f2 = mk3(1.0 - f.x, 1.0 - f.y, 1.0 - f.z);
    D3 refl = ld3(m.color_a);
    nd.kd = m.p0 * (refl * ONE_PI);
    nd.ks = (1.0 - m.p0) * f;
    nd.kt = (1.0 - m.p0) * ((1.0 - m.metalness) * f2);
  } else if (m.surface == PPM_SURF_TS) {
    double c2 = pow5(c);
    D3 spec = ld3(m.color_b);
    D3 f = mk3(spec.x + (1.0 - spec.x) * c2, spec.y + (1.0 - spec.y) * c2, spec.z + (1.0 - spec.z) * c2);
    nd.ks = f;
    if (m.metalness == 0.0) {
      D3 fa = cmul(mk3(1.0 - f.x, 1.0 - f.y, 1.0 - f.z), ld3(m.color_a));
      nd.kd = fa * (m.p0 * ONE_PI);
      nd.kt = fa * (1.0 - m.p0);
    }
  }
}

#endif  // PPM_DEV_CORE_CUH_
