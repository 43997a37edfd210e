// ppm_oracle.cpp -- CPU restatement of the eijian/ppmpa render hot path.
//
// *** TEST INFRASTRUCTURE ONLY. ***  This file is the parity oracle: only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load it.  The product (ppmpa_b200/, include/ppm.h) never
// links, imports or calls anything in oracle/.
//
// It follows the reference's Rust sources function by function (file:line
// cited at each function, paths relative to the reference tree), keeps the
// reference's structure (recursion, Vec + stable sort for the nearest hit,
// bottom-up bsdf combination) and its f64 operation order, and is compiled
// with -ffp-contract=off so that no FMA is formed (rustc/LLVM never contracts).
//
// Pinning status
//   * PINNED by the reference's own known answers (tests/test_oracle_golden.py):
//     algebra.rs:268,275  geometry.rs:217-218,227,240-242  tracer.rs:369-372
//     physics.rs:372-377,409-415 (the still-valid vectors).
//   * PARITY UNPINNED at two third-party crate boundaries that are not in the
//     reference tree and cannot be built here (no cargo/rustc, no crate cache):
//       - kdtree ^0.5.1 (Cargo.toml:13): `within(point, r2, squared_euclidean)`
//         is restated from its published semantics -- squared_euclidean =
//         fold(0,+) of (a_i-b_i)^2 in index order; membership d2 <= r2
//         (inclusive); results ascending by distance.
//       - rand ^0.6 (Cargo.toml:10): thread_rng() is OS seeded, so nothing
//         downstream of a random draw is reproducible in the reference itself.
//         The oracle takes an injectable generator (Rng below); the engine
//         and the oracle both use Philox4x32-10 with the same counter layout
//         so that they consume identical draws in identical order.
//   * The reference itself cannot be compiled in this image (Rust toolchain
//     absent) so there is no oracle/_ref.
#include "../include/ppm.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

namespace {

typedef double Flt;

// src/ray/mod.rs:13-17
const Flt NEARLY0 = 0.0001;
const Flt PI = 3.14159265358979323846264338327950288;  // f64::consts::PI
const Flt PI2 = PI * 2.0;
const Flt PI4 = PI * 4.0;

// ---------------------------------------------------------------------------
// algebra.rs
// ---------------------------------------------------------------------------
struct V3 {
  Flt v[3];
};
inline V3 mk(Flt x, Flt y, Flt z) { V3 r = {{x, y, z}}; return r; }
inline V3 from(const double a[3]) { return mk(a[0], a[1], a[2]); }
inline void to(const V3& a, double o[3]) { o[0] = a.v[0]; o[1] = a.v[1]; o[2] = a.v[2]; }
inline V3 neg(const V3& a) { return mk(-a.v[0], -a.v[1], -a.v[2]); }                    // algebra.rs:54-61
inline V3 add(const V3& a, const V3& b) { return mk(a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]); }  // :63-72
inline V3 sub(const V3& a, const V3& b) { return mk(a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]); }  // :74-83
inline V3 mul(const V3& a, Flt s) { return mk(a.v[0] * s, a.v[1] * s, a.v[2] * s); }    // :85-94
inline V3 mul(Flt s, const V3& a) { return mk(s * a.v[0], s * a.v[1], s * a.v[2]); }    // :96-105
inline Flt dot(const V3& a, const V3& b) {                                              // :146-148
  return a.v[0] * b.v[0] + a.v[1] * b.v[1] + a.v[2] * b.v[2];
}
inline Flt square(const V3& a) { return dot(a, a); }                                    // :127-129
inline Flt norm(const V3& a) { return std::sqrt(square(a)); }                           // :134-136
inline bool normalize(const V3& a, V3* out) {                                           // :151-158
  Flt n = norm(a);
  if (n == 0.0) return false;
  *out = mul(a, 1.0 / n);
  return true;
}
inline V3 cross(const V3& a, const V3& b) {                                             // :174-180
  return mk(a.v[1] * b.v[2] - b.v[1] * a.v[2],
            a.v[2] * b.v[0] - b.v[2] * a.v[0],
            a.v[0] * b.v[1] - b.v[0] * a.v[1]);
}
const V3 V_O = {{0.0, 0.0, 0.0}};
const V3 V_EX = {{1.0, 0.0, 0.0}};

// ---------------------------------------------------------------------------
// RNG: stands in for rand::thread_rng() (injectable; see header comment).
// gen_range(lo, hi) is half-open like rand 0.6's; it is DEFINED here as
// lo + (hi - lo) * u with u a 53-bit uniform in [0,1).
// ---------------------------------------------------------------------------
struct Rng {
  virtual ~Rng() {}
  virtual Flt next01() = 0;
  Flt gen_range(Flt lo, Flt hi) { return lo + (hi - lo) * next01(); }
};

// Philox4x32-10 (Salmon et al., SC'11), counter layout shared with the engine:
//   key = (seed_lo, seed_hi ^ pass)
//   ctr = (index_lo, index_hi, sub, (domain << 24) | block)
// draw k of a stream uses block k>>1, words (0,1) for even k and (2,3) for odd
// k;  u = (((hi32 << 32) | lo32) >> 11) * 2^-53.
inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
enum { DOMAIN_PHOTON = 1, DOMAIN_EYE = 2 };
struct PhiloxRng : Rng {
  uint32_t k0, k1, c0, c1, c2, dom;
  uint32_t k;
  uint32_t cache[4];
  PhiloxRng(uint64_t seed, uint32_t pass, uint32_t domain, uint64_t index, uint32_t sub)
      : k0((uint32_t)seed), k1((uint32_t)(seed >> 32) ^ pass), c0((uint32_t)index),
        c1((uint32_t)(index >> 32)), c2(sub), dom(domain), k(0) {}
  Flt next01() override {
    if ((k & 1u) == 0) {
      cache[0] = c0; cache[1] = c1; cache[2] = c2; cache[3] = (dom << 24) | (k >> 1);
      philox4x32_10(cache, k0, k1);
    }
    uint32_t lo = cache[(k & 1u) * 2], hi = cache[(k & 1u) * 2 + 1];
    ++k;
    uint64_t bits = (((uint64_t)hi << 32) | lo) >> 11;
    return (Flt)bits * (1.0 / 9007199254740992.0);
  }
};
struct SeqRng : Rng {  // replays a caller-supplied sequence (tests)
  const double* s; int64_t n, i;
  SeqRng(const double* s_, int64_t n_) : s(s_), n(n_), i(0) {}
  Flt next01() override { return (i < n) ? s[i++] : 0.5; }
};

// algebra.rs:211-223
V3 generate_random_dir(Rng& rng) {
  for (;;) {
    Flt x = rng.gen_range(-1.0, 1.0);
    Flt y = rng.gen_range(-1.0, 1.0);
    Flt z = rng.gen_range(-1.0, 1.0);
    V3 v = mk(x, y, z);
    Flt len = norm(v);
    if (0.0 < len && len <= 1.0) { V3 o; normalize(v, &o); return o; }
  }
}
// algebra.rs:225-235
V3 generate_random_dir_by_angle(Rng& rng) {
  Flt phi = rng.gen_range(0.0, 2.0 * PI);
  Flt xi = rng.gen_range(-1.0, 1.0);
  Flt xi2 = std::sqrt(1.0 - std::pow(xi, 2.0));
  Flt x = xi2 * std::cos(phi);
  Flt y = xi;
  Flt z = xi2 * std::sin(phi);
  V3 o = V_EX;
  normalize(mk(x, y, z), &o);
  return o;
}

// ---------------------------------------------------------------------------
// geometry.rs
// ---------------------------------------------------------------------------
struct Ray { V3 pos, dir; };
inline V3 target(const Ray& r, Flt t) { return add(r.pos, mul(r.dir, t)); }  // geometry.rs:56-58

// geometry.rs:149-165
bool method_moller(Flt l, const V3& p0, const V3& d1, const V3& d2, const V3& p, const V3& d,
                   Flt* uo, Flt* vo, Flt* to_) {
  V3 re2 = cross(d, d2);
  Flt det_a = dot(re2, d1);
  V3 pp = sub(p, p0);
  V3 te1 = cross(pp, d1);
  Flt u = dot(re2, pp) / det_a;
  Flt v = dot(te1, d) / det_a;
  Flt t = dot(te1, d2) / det_a;
  if (det_a == 0.0 || u < 0.0 || u > 1.0 || v < 0.0 || v > 1.0 || u + v > l) return false;
  *uo = u; *vo = v; *to_ = t;
  return true;
}
// geometry.rs:170-177
void distance_plain(const Ray& r, const V3& n, Flt d, std::vector<Flt>& out) {
  Flt cos0 = dot(n, r.dir);
  if (cos0 == 0.0) return;
  out.push_back((d + dot(n, r.pos)) / -cos0);
}
// geometry.rs:179-193
void distance_sphere(const Ray& r, const V3& c, Flt rad, std::vector<Flt>& out) {
  V3 o = sub(c, r.pos);
  Flt t0 = dot(o, r.dir);
  Flt t1 = rad * rad - (square(o) - (t0 * t0));
  Flt t2 = std::sqrt(t1);
  if (t1 <= 0.0) return;
  if (t2 == 0.0) { out.push_back(t0); }
  else { out.push_back(t0 - t2); out.push_back(t0 + t2); }
}
// geometry.rs:195-202
void distance_polygon(Flt l, const Ray& r, const V3& p, const V3& d1, const V3& d2, std::vector<Flt>& out) {
  Flt u, v, t;
  if (method_moller(l, p, d1, d2, r.pos, r.dir, &u, &v, &t)) out.push_back(t);
}
// geometry.rs:132-145
void shape_distance(const ppm_prim& s, const Ray& r, std::vector<Flt>& out) {
  switch (s.type) {
    case PPM_SHAPE_PLAIN: distance_plain(r, from(s.nvec), s.scalar, out); break;
    case PPM_SHAPE_SPHERE: distance_sphere(r, from(s.position), s.scalar, out); break;
    case PPM_SHAPE_POLYGON: distance_polygon(1.0, r, from(s.position), from(s.dir1), from(s.dir2), out); break;
    case PPM_SHAPE_PARALLELOGRAM: distance_polygon(2.0, r, from(s.position), from(s.dir1), from(s.dir2), out); break;
    default: break;  // Point
  }
}
// geometry.rs:117-130
bool shape_normal(const ppm_prim& s, const V3& p, V3* n) {
  switch (s.type) {
    case PPM_SHAPE_PLAIN:
    case PPM_SHAPE_POLYGON:
    case PPM_SHAPE_PARALLELOGRAM: *n = from(s.nvec); return true;
    case PPM_SHAPE_SPHERE: return normalize(sub(p, from(s.position)), n);
    default: return false;
  }
}

// ---------------------------------------------------------------------------
// physics.rs
// ---------------------------------------------------------------------------
inline Flt clip_color(Flt a) { return a < 0.0 ? 0.0 : a; }       // physics.rs:164-170
void color_normalize(const double c[3], double o[3]) {           // physics.rs:61-71
  Flt r1 = clip_color(c[0]), g1 = clip_color(c[1]), b1 = clip_color(c[2]);
  Flt mag = r1 + g1 + b1;
  if (mag == 0.0) { o[0] = 1.0 / 3.0; o[1] = 1.0 / 3.0; o[2] = 1.0 / 3.0; }
  else { o[0] = r1 / mag; o[1] = g1 / mag; o[2] = b1 / mag; }
}
int decide_wavelength(const double c[3], Flt p) {                // physics.rs:74-84
  if (p < c[0]) return PPM_WL_RED;
  if (p < c[0] + c[1]) return PPM_WL_GREEN;
  return PPM_WL_BLUE;
}
inline Flt relative_ior(Flt ior1, Flt ior2) { return ior1 == 0.0 ? 1.0 : ior2 / ior1; }  // physics.rs:200-205
inline Flt relative_ior_wavelength(const double i1[3], const double i2[3], int wl) {      // :183-189
  return relative_ior(i1[wl], i2[wl]);
}
inline Flt relative_ior_average(const double i1[3], const double i2[3]) {                 // :192-196
  Flt a1 = (i1[0] + i1[1] + i1[2]) / 3.0;
  Flt a2 = (i2[0] + i2[1] + i2[2]) / 3.0;
  return relative_ior(a1, a2);
}
// physics.rs:214-221
void specular_reflection(const V3& nvec, const V3& vvec, V3* rvec, Flt* cos1) {
  Flt c = -dot(vvec, nvec);
  if (c < 0.0) { *rvec = nvec; *cos1 = -c; return; }
  V3 r = V_EX;
  normalize(add(vvec, mul(2.0 * c, nvec)), &r);  // .unwrap(): zero only for degenerate input
  *rvec = r; *cos1 = c;
}
// physics.rs:232-261
V3 reflection_glossy(const V3& nvec, const V3& rvec, Flt pw, Rng& rng) {
  V3 uvec = V_EX;
  if (!normalize(cross(mk(0.00424, 1.0, 0.00764), rvec), &uvec)) {
    normalize(cross(mk(1.0, 0.00424, 0.00764), rvec), &uvec);
  }
  V3 vvec = cross(uvec, rvec);
  Flt c0 = dot(nvec, rvec);
  Flt xi0 = rng.gen_range(0.0, 1.0);
  Flt xi1 = std::pow(xi0, pw * c0);
  Flt xi2 = 2.0 * PI * rng.gen_range(0.0, 1.0);
  Flt x = std::cos(xi2) * std::sqrt(1.0 - xi1 * xi1);
  Flt y = xi1;
  Flt z = std::sin(xi2) * std::sqrt(1.0 - xi1 * xi1);
  V3 wi = add(add(mul(x, uvec), mul(y, rvec)), mul(z, vvec));
  if (dot(nvec, wi) < 0.0) {
    wi = sub(add(mul(-x, uvec), mul(y, rvec)), mul(z, vvec));
  }
  V3 o;
  if (normalize(wi, &o)) return o;
  return V_EX;
}
// physics.rs:270-285.  returns false for None
bool specular_refraction(const V3& nvec, const V3& vvec, Flt eta, V3* tvec, Flt* cos2) {
  Flt cos1 = -dot(vvec, nvec);
  if (cos1 < 0.0) { *cos2 = 0.0; return false; }
  Flt sq_eta = eta * eta;
  Flt sq_cos = cos1 * cos1;
  Flt g0 = sq_eta + sq_cos - 1.0;
  if (g0 < 0.0) { *cos2 = 0.0; return false; }
  Flt g = std::sqrt(g0);
  bool ok = normalize(mul(1.0 / eta, add(vvec, mul(cos1 - g, nvec))), tvec);
  *cos2 = g / eta;
  return ok;
}
inline Flt schlick(Flt f0, Flt c) { return f0 + (1.0 - f0) * std::pow(1.0 - c, 5.0); }  // physics.rs:316-318
// physics.rs:329-335
size_t check_under(const Flt* ps, size_t n, Flt p) {
  size_t i = 0;
  while (i < n && p > ps[i]) i += 1;
  return i;
}
// physics.rs:323-327
size_t russian_roulette(const Flt* ps, size_t n, Rng& rng) {
  Flt p = rng.gen_range(0.0, 1.0);
  return check_under(ps, n, p);
}
inline size_t russian_roulette1(Flt p0, Rng& rng) { return russian_roulette(&p0, 1, rng); }

// ---------------------------------------------------------------------------
// surface.rs
// ---------------------------------------------------------------------------
const Flt ONE_PI = 1.0 / PI;            // surface.rs:13, tracer.rs:24
const Flt SR_HALF = 1.0 / (2.0 * PI);   // tracer.rs:25

inline bool is3(const double c[3], Flt v) { return c[0] == v && c[1] == v && c[2] == v; }

Flt density_pow_of(Flt rough) {  // surface.rs:52,63
  return 1.0 / (std::pow(10.0, 5.0 * (1.0 - std::sqrt(rough))) + 1.0);
}
// surface.rs:68-100
bool surf_reflect(const ppm_material& m, Flt c) {
  switch (m.surface) {
    case PPM_SURF_SIMPLE:
      return (m.p0 == 1.0 || (c == 1.0 && is3(m.color_b, 0.0))) == false;
    case PPM_SURF_TS:
      if (m.metalness == 0.0) return true;
      if (m.metalness == 1.0) return !is3(m.color_b, 0.0);
      return true;
    default: return false;
  }
}
// surface.rs:102-133
bool surf_refract(const ppm_material& m, Flt c) {
  switch (m.surface) {
    case PPM_SURF_SIMPLE:
      return (c == 0.0 && is3(m.color_b, 1.0)) == false;
    case PPM_SURF_TS:
      if (m.metalness == 0.0) {
        if (m.p0 < 1.0 && !is3(m.color_a, 0.0)) return true;
      }
      return false;
    default: return false;
  }
}
// surface.rs:512-515
void reflection_index(const double col[3], Flt c, Flt out[3]) {
  Flt c2 = std::pow(1.0 - c, 5.0);
  for (int i = 0; i < 3; ++i) out[i] = col[i] + (1.0 - col[i]) * c2;
}
struct Rad { Flt c[3]; };
const Rad RAD0 = {{0.0, 0.0, 0.0}};
inline Rad radd(const Rad& a, const Rad& b) { Rad r = {{a.c[0] + b.c[0], a.c[1] + b.c[1], a.c[2] + b.c[2]}}; return r; }
inline Rad rmul(const Rad& a, Flt s) { Rad r = {{a.c[0] * s, a.c[1] * s, a.c[2] * s}}; return r; }     // Radiance * Flt
inline Rad rmul(Flt s, const Rad& a) { Rad r = {{s * a.c[0], s * a.c[1], s * a.c[2]}}; return r; }     // Flt * Radiance
inline Rad cmulr(const Flt col[3], const Rad& a) { Rad r = {{col[0] * a.c[0], col[1] * a.c[1], col[2] * a.c[2]}}; return r; }  // Color * Radiance

// surface.rs:135-206
Rad surf_bsdf(const ppm_material& m, const V3& /*nvec*/, const V3& edir, const V3& rdir,
              Flt cos0, const Rad& di, const Rad& si, const Rad& ti) {
  switch (m.surface) {
    case PPM_SURF_SIMPLE: {
      Flt f[3], f2[3];
      reflection_index(m.color_b, cos0, f);
      for (int i = 0; i < 3; ++i) f2[i] = 1.0 - f[i];                       // -f  (physics.rs:131-138)
      Flt refl_pi[3] = {m.color_a[0] * ONE_PI, m.color_a[1] * ONE_PI, m.color_a[2] * ONE_PI};
      Rad a = rmul(m.p0, cmulr(refl_pi, di));                               // diffuseness * (refl*ONE_PI*di)
      Flt mf2[3] = {(1.0 - m.metalness) * f2[0], (1.0 - m.metalness) * f2[1], (1.0 - m.metalness) * f2[2]};
      Rad b = rmul(1.0 - m.p0, radd(cmulr(f, si), cmulr(mf2, ti)));
      return radd(a, b);
    }
    case PPM_SURF_TS: {
      V3 lvec = rdir, vvec = neg(edir), hvec = V_EX;
      normalize(add(lvec, vvec), &hvec);   // .unwrap() in the reference; value unused
      Flt f[3], f2[3];
      reflection_index(m.color_b, cos0, f);
      for (int i = 0; i < 3; ++i) f2[i] = 1.0 - f[i];
      Rad i_de = RAD0;
      if (m.metalness == 0.0) {
        Flt fa[3] = {f2[0] * m.color_a[0], f2[1] * m.color_a[1], f2[2] * m.color_a[2]};
        Rad inner = radd(rmul(m.p0 * ONE_PI, di), rmul(1.0 - m.p0, ti));
        i_de = cmulr(fa, inner);
      }
      Rad i_mt = cmulr(f, si);
      return radd(i_de, i_mt);
    }
    default: return RAD0;
  }
}
// surface.rs:408-417
V3 diffuse_reflection(const V3& n, Rng& rng) {
  V3 d = generate_random_dir_by_angle(rng);
  Flt c = dot(n, d);
  return c > 0.0 ? d : neg(d);
}
// surface.rs:211-258.  returns 0 = None, 1 = Some(dir, true), 2 = Some(dir, false)
int surf_next_direction(const ppm_material& m, Flt eta, const V3& nvec, const V3& vvec, int wl,
                        Rng& rng, V3* out) {
  V3 rdir0; Flt cos1;
  specular_reflection(nvec, vvec, &rdir0, &cos1);
  V3 rdir = reflection_glossy(nvec, rdir0, m.density_pow, rng);   // power_glossy(), surface.rs:381-402
  V3 hvec = V_EX;
  normalize(sub(rdir, vvec), &hvec);
  V3 tdir; Flt cos2;
  bool has_t = specular_refraction(hvec, vvec, eta, &tdir, &cos2);
  Flt c = cos1 < cos2 ? cos1 : cos2;
  if (m.surface != PPM_SURF_TS) return 0;
  Flt f = schlick(m.color_b[wl], c);
  if (russian_roulette1(f, rng) == 0) { *out = rdir; return 1; }
  if (russian_roulette1(m.color_a[wl], rng) == 1) return 0;
  if (russian_roulette1(m.p0, rng) == 0) { *out = diffuse_reflection(nvec, rng); return 1; }
  if (has_t) { *out = tdir; return 2; }
  return 0;
}
// surface.rs:289-310
bool surf_store_photon(const ppm_material& m) {
  switch (m.surface) {
    case PPM_SURF_SIMPLE: return m.p0 > 0.0;
    case PPM_SURF_TS: return m.metalness != 1.0 && m.p0 != 0.0;
    default: return true;
  }
}
// surface.rs:358-379 (Simple returns *diffuseness*)
Flt surf_roughness(const ppm_material& m) {
  switch (m.surface) {
    case PPM_SURF_SIMPLE: return m.p0;
    case PPM_SURF_TS: return m.roughness;
    default: return 0.0;
  }
}
Flt surf_albedo_diff(const ppm_material& m, int wl) {  // surface.rs:312-333
  return (m.surface == PPM_SURF_SIMPLE || m.surface == PPM_SURF_TS) ? m.color_a[wl] : 0.0;
}
Flt surf_albedo_spec(const ppm_material& m, int wl) {  // surface.rs:335-356
  return (m.surface == PPM_SURF_SIMPLE || m.surface == PPM_SURF_TS) ? m.color_b[wl] : 0.0;
}
Flt surf_power_glossy(const ppm_material& m) {         // surface.rs:381-402
  return (m.surface == PPM_SURF_SIMPLE || m.surface == PPM_SURF_TS) ? m.density_pow : 0.0;
}

// ---------------------------------------------------------------------------
// scene.rs:13-18  M_AIR
// ---------------------------------------------------------------------------
ppm_material make_air() {
  ppm_material m;
  std::memset(&m, 0, sizeof m);
  for (int i = 0; i < 3; ++i) { m.transmittance[i] = 1.0; m.ior[i] = 1.0; }
  m.surface = PPM_SURF_NOTHING;
  return m;
}
const ppm_material M_AIR = make_air();

struct Scene {
  const ppm_prim* prims; int nprims;
  const ppm_material* mats; int nmats;
  const ppm_light* lights; int nlights;
};

// ---------------------------------------------------------------------------
// optics.rs
// ---------------------------------------------------------------------------
struct Photon { int wl; Ray ray; };
// optics.rs:224-233
Rad photon_to_radiance(const V3& n, Flt pw, const Photon& ph) {
  Flt cos0 = dot(n, ph.ray.dir);
  Flt pw2 = cos0 < 0.0 ? pw * -cos0 : 0.0;
  Rad r = RAD0;
  r.c[ph.wl] = pw2;
  return r;
}

// ---------------------------------------------------------------------------
// light.rs
// ---------------------------------------------------------------------------
const Flt PARA_DIV = 0.2;                 // light.rs:162
const Flt TS5[5] = {0.1, 0.3, 0.5, 0.7, 0.9};  // TSS[i*5+j] = (TS5[i], TS5[j]), light.rs:164-170

int select_wavelength(const double c[3], Rng& rng) {  // light.rs:157-160
  return decide_wavelength(c, rng.gen_range(0.0, 1.0));
}
// light.rs:67-91
Photon generate_photon(const ppm_light& l, Rng& rng) {
  Photon ph;
  switch (l.type) {
    case PPM_LIGHT_POINT: {
      ph.wl = select_wavelength(l.color, rng);
      ph.ray.pos = from(l.pos);
      ph.ray.dir = generate_random_dir(rng);
      break;
    }
    case PPM_LIGHT_PARALLELOGRAM: {
      ph.wl = select_wavelength(l.color, rng);
      Flt t1 = rng.gen_range(0.0, 1.0);
      Flt t2 = rng.gen_range(0.0, 1.0);
      V3 d = diffuse_reflection(from(l.nvec), rng);
      ph.ray.pos = add(add(from(l.pos), mul(t1, from(l.dir1))), mul(t2, from(l.dir2)));
      ph.ray.dir = d;
      break;
    }
    default: {  // SunLight
      ph.wl = select_wavelength(l.color, rng);
      Flt t1 = rng.gen_range(0.0, 1.0);
      Flt t2 = rng.gen_range(0.0, 1.0);
      ph.ray.pos = add(add(from(l.pos), mul(t1, from(l.dir1))), mul(t2, from(l.dir2)));
      ph.ray.dir = from(l.dir);
      break;
    }
  }
  return ph;
}
// light.rs:93-129
void light_get_direction(const ppm_light& l, const V3& p, std::vector<V3>& out) {
  switch (l.type) {
    case PPM_LIGHT_POINT: out.push_back(sub(from(l.pos), p)); break;
    case PPM_LIGHT_PARALLELOGRAM: {
      V3 nv = from(l.nvec);
      for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j) {
          // gen_pos, light.rs:152-154
          V3 gp = add(add(from(l.pos), mul(TS5[i], from(l.dir1))), mul(TS5[j], from(l.dir2)));
          V3 d = sub(gp, p);
          if (dot(nv, d) < 0.0) out.push_back(d);
        }
      break;
    }
    default: {
      V3 d = sub(from(l.pos), p);
      Flt cos0 = dot(from(l.nvec), d);
      if (cos0 > 0.0) break;
      V3 dt2 = neg(from(l.dir));
      Flt u, v, t;
      if (method_moller(2.0, from(l.pos), from(l.dir1), from(l.dir2), p, dt2, &u, &v, &t))
        out.push_back(mul(t, dt2));
      break;
    }
  }
}
// light.rs:131-150.  NOTE rs starts with RADIANCE0 (the off-by-one of SURVEY B-6)
void light_get_radiance(const ppm_light& l, const std::vector<Flt>& ds, std::vector<Rad>& rs) {
  rs.push_back(RAD0);
  for (size_t i = 0; i < ds.size(); ++i) {
    Flt d = ds[i];
    Rad r;
    switch (l.type) {
      case PPM_LIGHT_POINT: {
        Flt l0 = l.flux / (PI4 * d);
        r.c[0] = l.color[0] * l0; r.c[1] = l.color[1] * l0; r.c[2] = l.color[2] * l0;
        break;
      }
      case PPM_LIGHT_PARALLELOGRAM: {
        Flt l0 = (2.0 * l.flux * PARA_DIV * PARA_DIV) / (PI4 * d);
        r.c[0] = l.color[0] * l0; r.c[1] = l.color[1] * l0; r.c[2] = l.color[2] * l0;
        break;
      }
      default:
        r.c[0] = l.color[0] * l.flux; r.c[1] = l.color[1] * l.flux; r.c[2] = l.color[2] * l.flux;
        break;
    }
    rs.push_back(r);
  }
}

// ---------------------------------------------------------------------------
// tracer.rs
// ---------------------------------------------------------------------------
const int MAX_TRACE = 10;  // tracer.rs:27

struct Intersection {      // tracer.rs:299-304
  V3 pos, nvec;
  const ppm_material* mate;
  int io;                  // 0 In, 1 Out
  int obj; Flt t;          // extra: what the bit-exact tier compares
};
// tracer.rs:306-350 (+ calc_distance :352-359)
bool calc_intersection(const Ray& r, const Scene& sc, Intersection* out) {
  std::vector<std::pair<Flt, int>> iss1;
  std::vector<Flt> ts;
  for (int o = 0; o < sc.nprims; ++o) {
    ts.clear();
    shape_distance(sc.prims[o], r, ts);
    for (size_t i = 0; i < ts.size(); ++i) {
      if (ts[i] < NEARLY0) continue;
      iss1.push_back(std::make_pair(ts[i], o));
    }
  }
  if (iss1.empty()) return false;
  std::stable_sort(iss1.begin(), iss1.end(),
                   [](const std::pair<Flt, int>& a, const std::pair<Flt, int>& b) { return a.first < b.first; });
  Flt t = iss1[0].first;
  int obj = iss1[0].second;
  V3 p = target(r, t);
  V3 n;
  if (!shape_normal(sc.prims[obj], p, &n)) return false;
  out->pos = p; out->obj = obj; out->t = t;
  out->mate = &sc.mats[sc.prims[obj].material];
  if (dot(n, r.dir) > 0.0) { out->nvec = neg(n); out->io = 1; }
  else { out->nvec = n; out->io = 0; }
  return true;
}

struct Rec { Photon ph; uint64_t tag; };
void trace_photon(bool uc, const ppm_material* m0, const Scene& sc, int l, const Photon& ph, Rng& rng,
                  uint64_t index, std::vector<Rec>& out);

// tracer.rs:83-92
void reflect_diff(bool uc, const ppm_material* m0, const Scene& sc, int l, const Photon& ph,
                  const Intersection& is, Rng& rng, uint64_t index, std::vector<Rec>& out) {
  size_t i = russian_roulette1(surf_albedo_diff(*is.mate, ph.wl), rng);
  if (i == 0) {
    V3 dr = diffuse_reflection(is.nvec, rng);
    Photon nx = {ph.wl, {is.pos, dr}};
    trace_photon(uc, m0, sc, l + 1, nx, rng, index, out);
  }
}
// tracer.rs:111-125
void reflect_trans(bool uc, const ppm_material* m0, const Scene& sc, int l, const Photon& ph,
                   const Intersection& is, Rng& rng, uint64_t index, std::vector<Rec>& out) {
  Flt eta = relative_ior_wavelength(m0->ior, is.mate->ior, ph.wl);
  V3 tdir; Flt cos2;
  if (specular_refraction(is.nvec, ph.ray.dir, eta, &tdir, &cos2)) {
    const ppm_material* m02 = dot(tdir, is.nvec) < 0.0 ? is.mate : &M_AIR;
    Photon nx = {ph.wl, {is.pos, tdir}};
    trace_photon(uc, m02, sc, l + 1, nx, rng, index, out);
  }
}
// tracer.rs:94-109
void reflect_spec(bool uc, const ppm_material* m0, const Scene& sc, int l, const Photon& ph,
                  const Intersection& is, Rng& rng, uint64_t index, std::vector<Rec>& out) {
  V3 rdir; Flt cos1;
  specular_reflection(is.nvec, ph.ray.dir, &rdir, &cos1);
  Flt f = schlick(surf_albedo_spec(*is.mate, ph.wl), cos1);
  size_t j = russian_roulette1(f, rng);
  if (j == 0) {
    Photon nx = {ph.wl, {is.pos, rdir}};
    trace_photon(uc, m0, sc, l + 1, nx, rng, index, out);
  } else {
    if (is.mate->ior[ph.wl] == 0.0) return;
    reflect_trans(uc, m0, sc, l, ph, is, rng, index, out);
  }
}
// tracer.rs:31-81.  Records are appended AFTER the recursive call (deepest first).
void trace_photon(bool uc, const ppm_material* m0, const Scene& sc, int l, const Photon& ph, Rng& rng,
                  uint64_t index, std::vector<Rec>& out) {
  if (l >= MAX_TRACE) return;
  Intersection is1;
  if (!calc_intersection(ph.ray, sc, &is1)) return;
  const ppm_material& sf = *is1.mate;
  switch (sf.surface) {
    case PPM_SURF_SIMPLE: {
      if (russian_roulette1(surf_roughness(sf), rng) == 0) reflect_diff(uc, m0, sc, l, ph, is1, rng, index, out);
      else reflect_spec(uc, m0, sc, l, ph, is1, rng, index, out);
      break;
    }
    case PPM_SURF_TS: {
      Flt eta = relative_ior_wavelength(m0->ior, is1.mate->ior, ph.wl);
      V3 dir;
      int m = surf_next_direction(sf, eta, is1.nvec, ph.ray.dir, ph.wl, rng, &dir);
      if (m != 0) {
        const ppm_material* mate = (m == 1) ? m0 : is1.mate;
        Photon nx = {ph.wl, {is1.pos, dir}};
        trace_photon(uc, mate, sc, l + 1, nx, rng, index, out);
      }
      break;
    }
    default: break;
  }
  if ((uc == false || l > 0) && surf_store_photon(sf) == true) {
    Rec rec;
    rec.ph.wl = ph.wl; rec.ph.ray.pos = is1.pos; rec.ph.ray.dir = ph.ray.dir;
    rec.tag = (index << 4) | (uint64_t)l;
    out.push_back(rec);
  }
}

// --- photon map: photonmap.rs:16-29 + kdtree `within` (restated, see header) ---
struct PhotonMap {
  Flt power, radius;            // radius = SQUARED radius (ppmpa.rs:60-63,79)
  std::vector<Photon> phs;
  // uniform grid accelerator (cell edge slightly > r): same neighbour sets as brute force
  Flt cell, org[3];
  std::unordered_map<uint64_t, std::pair<uint32_t, uint32_t>> cells;  // key -> [begin,end) in order
  std::vector<uint32_t> order;
  static uint64_t key(int64_t x, int64_t y, int64_t z) {
    return ((uint64_t)(x & 0x1FFFFF) << 42) | ((uint64_t)(y & 0x1FFFFF) << 21) | (uint64_t)(z & 0x1FFFFF);
  }
  int64_t coord(Flt p, int ax) const { return (int64_t)std::floor((p - org[ax]) / cell); }
  void build() {
    size_t n = phs.size();
    cell = std::sqrt(radius) * 1.001;
    if (!(cell > 0.0)) cell = 1.0;
    org[0] = org[1] = org[2] = 0.0;
    for (size_t i = 0; i < n; ++i)
      for (int a = 0; a < 3; ++a) if (i == 0 || phs[i].ray.pos.v[a] < org[a]) org[a] = phs[i].ray.pos.v[a];
    std::vector<std::pair<uint64_t, uint32_t>> ks(n);
    for (size_t i = 0; i < n; ++i) {
      const V3& p = phs[i].ray.pos;
      ks[i] = std::make_pair(key(coord(p.v[0], 0), coord(p.v[1], 1), coord(p.v[2], 2)), (uint32_t)i);
    }
    std::sort(ks.begin(), ks.end());
    order.resize(n);
    cells.clear();
    cells.reserve(n / 4 + 16);
    size_t b = 0;
    for (size_t i = 0; i < n; ++i) {
      order[i] = ks[i].second;
      if (i + 1 == n || ks[i + 1].first != ks[i].first) {
        cells[ks[i].first] = std::make_pair((uint32_t)b, (uint32_t)(i + 1));
        b = i + 1;
      }
    }
  }
  // kdtree::distance::squared_euclidean: fold(0,+) of (a_i - b_i)^2
  static Flt sqdist(const Flt a[3], const Flt b[3]) {
    Flt acc = 0.0;
    for (int i = 0; i < 3; ++i) acc = acc + (a[i] - b[i]) * (a[i] - b[i]);
    return acc;
  }
  // KdTree::within: all points with distance <= radius, ascending by distance
  void within(const V3& q, std::vector<std::pair<Flt, uint32_t>>& out) const {
    out.clear();
    if (phs.empty()) return;
    int64_t cx = coord(q.v[0], 0), cy = coord(q.v[1], 1), cz = coord(q.v[2], 2);
    for (int64_t x = cx - 1; x <= cx + 1; ++x)
      for (int64_t y = cy - 1; y <= cy + 1; ++y)
        for (int64_t z = cz - 1; z <= cz + 1; ++z) {
          auto it = cells.find(key(x, y, z));
          if (it == cells.end()) continue;
          for (uint32_t j = it->second.first; j < it->second.second; ++j) {
            uint32_t i = order[j];
            Flt d = sqdist(q.v, phs[i].ray.pos.v);
            if (d <= radius) out.push_back(std::make_pair(d, i));
          }
        }
    std::sort(out.begin(), out.end());
  }
  void within_brute(const V3& q, std::vector<std::pair<Flt, uint32_t>>& out) const {
    out.clear();
    for (size_t i = 0; i < phs.size(); ++i) {
      Flt d = sqdist(q.v, phs[i].ray.pos.v);
      if (d <= radius) out.push_back(std::make_pair(d, (uint32_t)i));
    }
    std::sort(out.begin(), out.end());
  }
};

// tracer.rs:198-204
const Flt K_CONE = 1.1;
const Flt FAC_K = 1.0 - 2.0 / (3.0 * K_CONE);
Flt filter_cone(Flt d, Flt rmax) {
  Flt d2 = std::sqrt(d / rmax) / K_CONE;
  return d2 > 1.0 ? 0.0 : (1.0 - d2) / FAC_K;
}
// tracer.rs:206-216 (CORR = 0.5 as in the code, not the stale 0.355 of the test)
const Flt G_ALPHA = 0.918;
const Flt G_BETA = 1.953;
const Flt E_BETA = 1.0 - 0.14184788965323;
const Flt CORR = 0.5;
Flt filter_gauss(Flt d, Flt rmax) {
  Flt e_r = 1.0 - std::exp(-G_BETA * d / (rmax * 2.0));
  return e_r > E_BETA ? 0.0 : G_ALPHA * (1.0 - e_r / E_BETA) + CORR;
}
// tracer.rs:179-195
Rad estimate_radiance(Flt radius, int pfilter, const PhotonMap& pmap, const V3& pos, const V3& nvec,
                      uint32_t* count, std::vector<std::pair<Flt, uint32_t>>& scratch) {
  pmap.within(pos, scratch);
  if (count) *count = (uint32_t)scratch.size();
  if (scratch.empty()) return RAD0;
  Rad rad = RAD0;
  for (size_t i = 0; i < scratch.size(); ++i) {
    Flt d = scratch[i].first;
    Flt wt = pfilter == PPM_FILTER_NONE ? 1.0 : (pfilter == PPM_FILTER_CONE ? filter_cone(d, radius) : filter_gauss(d, radius));
    rad = radd(rad, photon_to_radiance(nvec, wt * pmap.power, pmap.phs[scratch[i].second]));
  }
  return rmul(rad, ONE_PI / radius);
}

// tracer.rs:272-290
void illuminated(const Scene& sc, const V3& p, const V3& n, const std::vector<V3>& lds,
                 std::vector<Flt>& dists, std::vector<Flt>& coss) {
  for (size_t i = 0; i < lds.size(); ++i) {
    V3 ld3;
    if (!normalize(lds[i], &ld3)) continue;
    Flt cos0 = dot(n, ld3);
    if (cos0 < 0.0) continue;
    Ray lray = {p, ld3};
    Intersection is;
    if (!calc_intersection(lray, sc, &is)) continue;
    Flt sq_ldist = square(lds[i]);
    Flt sq_odist = square(sub(is.pos, p));
    if (sq_ldist - sq_odist > 0.002) continue;
    dists.push_back(sq_ldist);
    coss.push_back(cos0 * cos0);
  }
}
// tracer.rs:263-270
Rad get_radiance_from_light(const Scene& sc, const V3& p, const V3& n, const ppm_light& l) {
  std::vector<V3> dirs;
  light_get_direction(l, p, dirs);
  std::vector<Flt> dists, coss;
  illuminated(sc, p, n, dirs, dists, coss);
  std::vector<Rad> rs;
  light_get_radiance(l, dists, rs);
  Rad rad = RAD0;
  size_t k = std::min(rs.size(), coss.size());   // zip stops at the shorter
  for (size_t i = 0; i < k; ++i) rad = radd(rad, rmul(rs[i], coss[i]));
  return rad;
}

struct EyeCtx {
  const Scene* sc; const PhotonMap* pmap; int pfilter; Flt radius; bool uc;
  uint64_t seed; uint32_t pass; uint64_t pixel;
  std::vector<std::pair<Flt, uint32_t>> scratch;
  uint64_t n_nodes, n_gather, sum_k;
};
// tracer.rs:129-177.  node = position in the recursion tree (root 1, reflect
// child 2n, refract child 2n+1): the two glossy draws of a node come from the
// stream (seed, pass, EYE, pixel, node) so that draw order is traversal-independent.
Rad trace_ray(EyeCtx& cx, const ppm_material* m0, int l, const Ray& r, uint32_t node) {
  if (l >= MAX_TRACE) return RAD0;
  Intersection is1;
  if (!calc_intersection(r, *cx.sc, &is1)) return RAD0;
  cx.n_nodes++;
  Rad di = RAD0;
  if (cx.uc) {
    for (int i = 0; i < cx.sc->nlights; ++i)
      di = radd(di, get_radiance_from_light(*cx.sc, is1.pos, is1.nvec, cx.sc->lights[i]));
  }
  uint32_t cnt = 0;
  di = radd(di, estimate_radiance(cx.radius, cx.pfilter, *cx.pmap, is1.pos, is1.nvec, &cnt, cx.scratch));
  cx.n_gather++; cx.sum_k += cnt;
  const ppm_material& mate = *is1.mate;

  PhiloxRng rng(cx.seed, cx.pass, DOMAIN_EYE, cx.pixel, node);
  V3 rdir0; Flt cos1;
  specular_reflection(is1.nvec, r.dir, &rdir0, &cos1);
  V3 rdir = reflection_glossy(is1.nvec, rdir0, surf_power_glossy(mate), rng);
  Rad si = RAD0;
  if (surf_reflect(mate, cos1)) {
    Ray nr = {is1.pos, rdir};
    si = trace_ray(cx, m0, l + 1, nr, node * 2);
  }
  Flt eta = relative_ior_average(m0->ior, mate.ior);
  V3 hvec = V_EX;
  normalize(sub(rdir, r.dir), &hvec);
  V3 tdir; Flt cos2;
  bool has_t = specular_refraction(hvec, r.dir, eta, &tdir, &cos2);
  Rad ti = RAD0;
  if (has_t && surf_refract(mate, cos1)) {
    const ppm_material* m02 = is1.io == 0 ? &mate : &M_AIR;
    Ray nr = {is1.pos, tdir};
    ti = trace_ray(cx, m02, l + 1, nr, node * 2 + 1);
  }
  Flt c = cos1 < cos2 ? cos1 : cos2;
  Rad em = {{mate.emittance[0] * SR_HALF, mate.emittance[1] * SR_HALF, mate.emittance[2] * SR_HALF}};
  return radd(em, surf_bsdf(mate, is1.nvec, r.dir, rdir, c, di, si, ti));
}

// tracer.rs:221-259 (the `rtc` renderer): no photon map, no glossy lobe, Fresnel from cos1
Rad trace_ray_classic(const Scene& sc, const double ambient[3], const ppm_material* m0, int l, const Ray& r) {
  if (l >= 10) return RAD0;
  Intersection is1;
  if (!calc_intersection(r, sc, &is1)) return RAD0;
  const ppm_material& mate = *is1.mate;
  V3 rdir; Flt cos1;
  specular_reflection(is1.nvec, r.dir, &rdir, &cos1);
  Rad di = RAD0;
  for (int i = 0; i < sc.nlights; ++i) di = radd(di, get_radiance_from_light(sc, is1.pos, is1.nvec, sc.lights[i]));
  Rad amb = {{ambient[0], ambient[1], ambient[2]}};
  di = radd(di, amb);
  Rad si = RAD0;
  if (surf_reflect(mate, cos1)) {
    Ray nr = {is1.pos, rdir};
    si = trace_ray_classic(sc, ambient, m0, l + 1, nr);
  }
  Flt eta = relative_ior_average(m0->ior, mate.ior);
  V3 tdir; Flt cos2;
  bool has_t = specular_refraction(is1.nvec, r.dir, eta, &tdir, &cos2);
  Rad ti = RAD0;
  if (has_t && surf_refract(mate, cos1)) {
    const ppm_material* m02 = dot(tdir, is1.nvec) < 0.0 ? &mate : &M_AIR;
    Ray nr = {is1.pos, tdir};
    ti = trace_ray_classic(sc, ambient, m02, l + 1, nr);
  }
  Rad em = {{mate.emittance[0] * SR_HALF, mate.emittance[1] * SR_HALF, mate.emittance[2] * SR_HALF}};
  return radd(em, surf_bsdf(mate, is1.nvec, r.dir, rdir, cos1, di, si, ti));
}

// camera.rs:58-75
Ray generate_ray(const ppm_camera& cam, Flt y, Flt x, Rng& rng) {
  V3 blur = V_O;
  if (cam.blur) {
    Flt r1 = rng.gen_range(-0.5, 0.5);
    Flt r2 = rng.gen_range(-0.5, 0.5);
    blur = add(mul(r1, from(cam.eex)), mul(r2, from(cam.eey)));
  }
  Flt r3 = 0.0, r4 = 0.0;
  if (cam.progressive && cam.antialias) {
    r3 = rng.gen_range(-0.5, 0.5);
    r4 = rng.gen_range(-0.5, 0.5);
  }
  V3 eyepos = add(from(cam.eye_pos), blur);
  V3 eyedir = sub(add(add(from(cam.origin), mul(x + r3, from(cam.esx))), mul(y + r4, from(cam.esy))), blur);
  Ray r;
  r.pos = eyepos;
  r.dir = V_EX;
  normalize(eyedir, &r.dir);
  return r;
}

// camera.rs:151-168
bool camera_finalize(ppm_camera& c) {
  V3 eyepos = from(c.eye_pos), tgt = from(c.target_pos), upper = from(c.upper_dir);
  V3 ez, ex, ey;
  if (!normalize(sub(tgt, eyepos), &ez)) return false;
  if (!normalize(cross(upper, ez), &ex)) return false;
  if (!normalize(cross(ex, ez), &ey)) return false;
  const Flt SENSOR_SIZE = 35.0 / 1000.0;
  Flt step = (c.focus * SENSOR_SIZE / c.focal_len) / (Flt)c.xreso;
  V3 esx = mul(step, ex), esy = mul(step, ey);
  Flt ea = c.focal_len / c.f_number;
  V3 eex = mul(ea, ex), eey = mul(ea, ey);
  Flt lx = (Flt)(c.xreso / 2), ly = (Flt)(c.yreso / 2);
  V3 orig = sub(sub(mul(c.focus, ez), mul(lx - 0.5, esx)), mul(ly - 0.5, esy));
  c.photon_power = c.blur ? c.iso_sens / 100.0 * 4.9 / c.f_number * c.shut_speed / (1.0 / 250.0) : 1.0;
  to(ez, c.eye_dir); to(orig, c.origin); to(esx, c.esx); to(esy, c.esy); to(eex, c.eex); to(eey, c.eey);
  return true;
}

Scene mk_scene(const ppm_prim* prims, int np, const ppm_material* mats, int nm, const ppm_light* lights, int nl) {
  Scene s = {prims, np, mats, nm, lights, nl};
  return s;
}

void trace_all_photons(const Scene& sc, uint64_t seed, uint32_t pass, bool uc, const int64_t* n_per_light,
                       std::vector<Rec>& recs) {
  uint64_t index = 0;
  for (int li = 0; li < sc.nlights; ++li) {
    for (int64_t i = 0; i < n_per_light[li]; ++i, ++index) {
      PhiloxRng rng(seed, pass, DOMAIN_PHOTON, index, 0);
      Photon ph = generate_photon(sc.lights[li], rng);       // ppmpa.rs:89 / pm.rs:63
      trace_photon(uc, &M_AIR, sc, 0, ph, rng, index, recs);  // ppmpa.rs:89 / pm.rs:64
    }
  }
}

}  // namespace

// ===========================================================================
// C interface for the tests (ctypes).  Everything is prefixed orc_.
// ===========================================================================
extern "C" {

// ---- known-answer probes --------------------------------------------------
int orc_normalize(const double v[3], double out[3]) {
  V3 o;
  if (!normalize(from(v), &o)) return 0;
  to(o, out);
  return 1;
}
void orc_cross(const double a[3], const double b[3], double out[3]) { to(cross(from(a), from(b)), out); }
double orc_dot(const double a[3], const double b[3]) { return dot(from(a), from(b)); }
void orc_scale(const double a[3], double s, double out[3]) { to(mul(from(a), s), out); }
void orc_ray_target(const double pos[3], const double dir[3], double t, double out[3]) {
  Ray r = {from(pos), from(dir)};
  to(target(r, t), out);
}
// Shape::new_polygon / new_parallelogram, geometry.rs:93-115
int orc_new_polygon(const double p0[3], const double p1[3], const double p2[3], int parallelogram, ppm_prim* out) {
  V3 d1 = sub(from(p1), from(p0)), d2 = sub(from(p2), from(p0)), n;
  if (!normalize(cross(d1, d2), &n)) return 0;
  std::memset(out, 0, sizeof *out);
  out->type = parallelogram ? PPM_SHAPE_PARALLELOGRAM : PPM_SHAPE_POLYGON;
  to(from(p0), out->position); to(n, out->nvec); to(d1, out->dir1); to(d2, out->dir2);
  return 1;
}
int orc_shape_normal(const ppm_prim* s, const double p[3], double out[3]) {
  V3 n;
  if (!shape_normal(*s, from(p), &n)) return 0;
  to(n, out);
  return 1;
}
double orc_filter_cone(double d, double rmax) { return filter_cone(d, rmax); }
double orc_filter_gauss(double d, double rmax) { return filter_gauss(d, rmax); }
void orc_color_normalize(const double c[3], double out[3]) { color_normalize(c, out); }
int orc_decide_wavelength(const double c[3], double p) { return decide_wavelength(c, p); }
int orc_check_under(const double* ps, int n, double p) { return (int)check_under(ps, (size_t)n, p); }
double orc_schlick(double f0, double c) { return schlick(f0, c); }
double orc_density_pow(double rough) { return density_pow_of(rough); }
double orc_relative_ior_average(const double a[3], const double b[3]) { return relative_ior_average(a, b); }
int orc_specular_refraction(const double n[3], const double v[3], double eta, double t[3], double* cos2) {
  V3 tv = V_O;
  bool ok = specular_refraction(from(n), from(v), eta, &tv, cos2);
  to(tv, t);
  return ok ? 1 : 0;
}
void orc_specular_reflection(const double n[3], const double v[3], double r[3], double* cos1) {
  V3 rv;
  specular_reflection(from(n), from(v), &rv, cos1);
  to(rv, r);
}
int orc_camera_finalize(ppm_camera* c) { return camera_finalize(*c) ? 1 : 0; }
// uniform draws of the shared Philox layout (for the engine's RNG parity test)
void orc_philox_draws(uint64_t seed, uint32_t pass, uint32_t domain, uint64_t index, uint32_t sub, int n, double* out) {
  PhiloxRng rng(seed, pass, domain, index, sub);
  for (int i = 0; i < n; ++i) out[i] = rng.next01();
}
// util/iterator.rb:34-38
void orc_radius_schedule(double r0, int n, double* out) {
  const double ALPHA = 0.5;
  double r = r0;
  for (int i = 0; i < n; ++i) {
    out[i] = r;
    r = std::sqrt(((i + 1) + ALPHA) / ((i + 1) + 1.0)) * r;
  }
}
// camera.rs:92-100
void orc_radiance_to_rgb(double max_radiance, const double rad[3], int32_t rgb[3]) {
  for (int i = 0; i < 3; ++i) {
    double d2 = rad[i] / max_radiance;
    double r2 = d2 > 1.0 ? 1.0 : d2;
    rgb[i] = (int32_t)std::floor(std::pow(r2, 1.0 / 2.2) * 255.0);
  }
}
// averager2.rb:84-94
int32_t orc_averager_clip(double c, uint32_t nfile, double max_radiance) {
  double c2 = c / nfile / max_radiance;
  double r = std::pow(c2 > 1.0 ? 1.0 : c2, 1.0 / 2.2) * 255.0;
  return (int32_t)r;
}

// ---- calc_intersection ------------------------------------------------------
void orc_intersect(const ppm_prim* prims, int np, const ppm_material* mats, int nm, const double* rays6, int64_t n,
                   int32_t* hit_idx, double* t, double* pos3, double* nrm3, int32_t* io) {
  Scene sc = mk_scene(prims, np, mats, nm, nullptr, 0);
  for (int64_t i = 0; i < n; ++i) {
    Ray r = {mk(rays6[i * 6], rays6[i * 6 + 1], rays6[i * 6 + 2]), mk(rays6[i * 6 + 3], rays6[i * 6 + 4], rays6[i * 6 + 5])};
    Intersection is;
    if (calc_intersection(r, sc, &is)) {
      hit_idx[i] = is.obj;
      if (t) t[i] = is.t;
      if (pos3) to(is.pos, pos3 + i * 3);
      if (nrm3) to(is.nvec, nrm3 + i * 3);
      if (io) io[i] = is.io;
    } else {
      hit_idx[i] = -1;
      if (t) t[i] = 0.0;
      if (pos3) pos3[i * 3] = pos3[i * 3 + 1] = pos3[i * 3 + 2] = 0.0;
      if (nrm3) nrm3[i * 3] = nrm3[i * 3 + 1] = nrm3[i * 3 + 2] = 0.0;
      if (io) io[i] = 0;
    }
  }
}

// ---- emission + photon tracing ---------------------------------------------
void orc_emit_photons(const ppm_light* lights, int nl, uint64_t seed, uint32_t pass, const int64_t* n_per_light,
                      ppm_photon* out) {
  uint64_t index = 0;
  for (int li = 0; li < nl; ++li)
    for (int64_t i = 0; i < n_per_light[li]; ++i, ++index) {
      PhiloxRng rng(seed, pass, DOMAIN_PHOTON, index, 0);
      Photon ph = generate_photon(lights[li], rng);
      to(ph.ray.pos, out[index].pos); to(ph.ray.dir, out[index].dir);
      out[index].wl = ph.wl; out[index]._pad = 0;
    }
}
// returns the number of stored records (may exceed cap; only cap are written).
// Records come out in the reference's order (per photon, deepest hit first).
uint64_t orc_trace_photons(const ppm_prim* prims, int np, const ppm_material* mats, int nm, const ppm_light* lights, int nl,
                           uint64_t seed, uint32_t pass, int uc, const int64_t* n_per_light,
                           ppm_photon* out, uint64_t* tags, uint64_t cap) {
  Scene sc = mk_scene(prims, np, mats, nm, lights, nl);
  std::vector<Rec> recs;
  trace_all_photons(sc, seed, pass, uc != 0, n_per_light, recs);
  for (size_t i = 0; i < recs.size() && i < cap; ++i) {
    to(recs[i].ph.ray.pos, out[i].pos); to(recs[i].ph.ray.dir, out[i].dir);
    out[i].wl = recs[i].ph.wl; out[i]._pad = 0;
    if (tags) tags[i] = recs[i].tag;
  }
  return (uint64_t)recs.size();
}
// one photon path with an injected draw sequence (tests of the RR logic)
uint64_t orc_trace_one_photon_seq(const ppm_prim* prims, int np, const ppm_material* mats, int nm,
                                  const ppm_photon* start, int uc, const double* draws, int64_t ndraws,
                                  ppm_photon* out, uint64_t cap) {
  Scene sc = mk_scene(prims, np, mats, nm, nullptr, 0);
  SeqRng rng(draws, ndraws);
  Photon ph = {start->wl, {from(start->pos), from(start->dir)}};
  std::vector<Rec> recs;
  trace_photon(uc != 0, &M_AIR, sc, 0, ph, rng, 0, recs);
  for (size_t i = 0; i < recs.size() && i < cap; ++i) {
    to(recs[i].ph.ray.pos, out[i].pos); to(recs[i].ph.ray.dir, out[i].dir);
    out[i].wl = recs[i].ph.wl; out[i]._pad = 0;
  }
  return (uint64_t)recs.size();
}

// ---- photon map handle -------------------------------------------------------
void* orc_map_build(const ppm_photon* ph, uint64_t n, double power, double radius2) {
  PhotonMap* m = new PhotonMap();
  m->power = power; m->radius = radius2;
  m->phs.resize(n);
  for (uint64_t i = 0; i < n; ++i) {
    m->phs[i].wl = ph[i].wl; m->phs[i].ray.pos = from(ph[i].pos); m->phs[i].ray.dir = from(ph[i].dir);
  }
  m->build();
  return m;
}
void orc_map_free(void* m) { delete (PhotonMap*)m; }
// neighbour set of one query: indices ascending by (d2, index); returns the count
uint32_t orc_within(void* map, const double q[3], int brute, uint32_t* idx, double* d2, uint32_t cap) {
  PhotonMap* m = (PhotonMap*)map;
  std::vector<std::pair<Flt, uint32_t>> out;
  if (brute) m->within_brute(from(q), out); else m->within(from(q), out);
  for (size_t i = 0; i < out.size() && i < cap; ++i) { if (idx) idx[i] = out[i].second; if (d2) d2[i] = out[i].first; }
  return (uint32_t)out.size();
}
static void gather_range(const PhotonMap* m, const double* pos3, const double* nrm3, int64_t b, int64_t e, int filter,
                         double* rgb3, uint32_t* counts) {
  std::vector<std::pair<Flt, uint32_t>> scratch;
  for (int64_t i = b; i < e; ++i) {
    uint32_t c = 0;
    Rad r = estimate_radiance(m->radius, filter, *m, mk(pos3[i * 3], pos3[i * 3 + 1], pos3[i * 3 + 2]),
                              mk(nrm3[i * 3], nrm3[i * 3 + 1], nrm3[i * 3 + 2]), &c, scratch);
    rgb3[i * 3] = r.c[0]; rgb3[i * 3 + 1] = r.c[1]; rgb3[i * 3 + 2] = r.c[2];
    if (counts) counts[i] = c;
  }
}
// estimate_radiance over a batch of (pos, normal); nthreads > 1 splits the batch
void orc_gather(void* map, const double* pos3, const double* nrm3, int64_t n, int filter, double* rgb3, uint32_t* counts,
                int nthreads) {
  const PhotonMap* m = (const PhotonMap*)map;
  if (nthreads <= 1) { gather_range(m, pos3, nrm3, 0, n, filter, rgb3, counts); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) {
    int64_t b = n * t / nthreads, e = n * (t + 1) / nthreads;
    th.emplace_back(gather_range, m, pos3, nrm3, b, e, filter, rgb3, counts);
  }
  for (auto& x : th) x.join();
}

// k-NN estimate.  NO REFERENCE IMPLEMENTATION EXISTS (n_sample_photon is dead code, photonmap.rs:18,
// camera.rs:181); this is the brute-force statement of the semantics defined in SURVEY.md 8c /
// include/ppm.h: k nearest within r; r_k^2 = k-th smallest d2 replaces r^2 (ties included); fewer
// than k (or r_k^2 == 0) -> fixed radius.
void orc_gather_knn(void* map, const double* pos3, const double* nrm3, int64_t n, uint32_t k, int filter, double* rgb3,
                    double* r2k, uint32_t* counts) {
  const PhotonMap* m = (const PhotonMap*)map;
  std::vector<std::pair<Flt, uint32_t>> nb;
  for (int64_t i = 0; i < n; ++i) {
    V3 q = mk(pos3[i * 3], pos3[i * 3 + 1], pos3[i * 3 + 2]), nv = mk(nrm3[i * 3], nrm3[i * 3 + 1], nrm3[i * 3 + 2]);
    m->within(q, nb);
    Flt rk = m->radius;
    if (nb.size() >= k && nb[k - 1].first > 0.0) rk = nb[k - 1].first;
    Rad rad = RAD0;
    uint32_t used = 0;
    for (size_t j = 0; j < nb.size() && nb[j].first <= rk; ++j, ++used) {
      Flt d = nb[j].first;
      Flt wt = filter == PPM_FILTER_NONE ? 1.0 : (filter == PPM_FILTER_CONE ? filter_cone(d, rk) : filter_gauss(d, rk));
      rad = radd(rad, photon_to_radiance(nv, wt * m->power, m->phs[nb[j].second]));
    }
    rad = rmul(rad, ONE_PI / rk);
    rgb3[i * 3] = rad.c[0]; rgb3[i * 3 + 1] = rad.c[1]; rgb3[i * 3 + 2] = rad.c[2];
    if (r2k) r2k[i] = rk;
    if (counts) counts[i] = used;
  }
}

// ---- camera rays + trace_ray --------------------------------------------------
void orc_generate_rays(const ppm_camera* cam, uint64_t seed, uint32_t pass, double* rays6) {
  for (int y = 0; y < cam->yreso; ++y)
    for (int x = 0; x < cam->xreso; ++x) {
      uint64_t pix = (uint64_t)y * cam->xreso + x;
      PhiloxRng rng(seed, pass, DOMAIN_EYE, pix, 0);
      Ray r = generate_ray(*cam, (Flt)y, (Flt)x, rng);   // screen_map is (y, x), camera.rs:170-175
      to(r.pos, rays6 + pix * 6); to(r.dir, rays6 + pix * 6 + 3);
    }
}
static void trace_rays_range(const Scene* sc, const PhotonMap* m, int pfilter, const double* rays6, int64_t b, int64_t e,
                             int64_t first_pixel, uint64_t seed, uint32_t pass, int uc, double* rgb3, uint64_t* stats) {
  EyeCtx cx;
  cx.sc = sc; cx.pmap = m; cx.pfilter = pfilter; cx.radius = m->radius; cx.uc = uc != 0;
  cx.seed = seed; cx.pass = pass; cx.n_nodes = cx.n_gather = cx.sum_k = 0;
  for (int64_t i = b; i < e; ++i) {
    Ray r = {mk(rays6[i * 6], rays6[i * 6 + 1], rays6[i * 6 + 2]), mk(rays6[i * 6 + 3], rays6[i * 6 + 4], rays6[i * 6 + 5])};
    cx.pixel = (uint64_t)(first_pixel + i);
    Rad c = trace_ray(cx, &M_AIR, 0, r, 1);
    rgb3[i * 3] = c.c[0]; rgb3[i * 3 + 1] = c.c[1]; rgb3[i * 3 + 2] = c.c[2];
  }
  if (stats) { stats[0] = cx.n_nodes; stats[1] = cx.n_gather; stats[2] = cx.sum_k; }
}
void orc_trace_rays(const ppm_prim* prims, int np, const ppm_material* mats, int nm, const ppm_light* lights, int nl,
                    void* map, int pfilter, const double* rays6, int64_t n, int64_t first_pixel,
                    uint64_t seed, uint32_t pass, int uc, double* rgb3, int nthreads, uint64_t* stats3) {
  Scene sc = mk_scene(prims, np, mats, nm, lights, nl);
  const PhotonMap* m = (const PhotonMap*)map;
  if (nthreads <= 1) { trace_rays_range(&sc, m, pfilter, rays6, 0, n, first_pixel, seed, pass, uc, rgb3, stats3); return; }
  std::vector<std::thread> th;
  std::vector<uint64_t> st((size_t)nthreads * 3, 0);
  for (int t = 0; t < nthreads; ++t) {
    int64_t b = n * t / nthreads, e = n * (t + 1) / nthreads;
    th.emplace_back(trace_rays_range, &sc, m, pfilter, rays6, b, e, first_pixel, seed, pass, uc, rgb3, &st[(size_t)t * 3]);
  }
  for (auto& x : th) x.join();
  if (stats3) { stats3[0] = stats3[1] = stats3[2] = 0; for (int t = 0; t < nthreads; ++t) for (int k = 0; k < 3; ++k) stats3[k] += st[(size_t)t * 3 + k]; }
}
// trace_ray_classic over a batch of rays (rtc.rs:29-30)
void orc_trace_rays_classic(const ppm_prim* prims, int np, const ppm_material* mats, int nm, const ppm_light* lights, int nl,
                            const double ambient[3], const double* rays6, int64_t n, double* rgb3) {
  Scene sc = mk_scene(prims, np, mats, nm, lights, nl);
  for (int64_t i = 0; i < n; ++i) {
    Ray r = {mk(rays6[i * 6], rays6[i * 6 + 1], rays6[i * 6 + 2]), mk(rays6[i * 6 + 3], rays6[i * 6 + 4], rays6[i * 6 + 5])};
    Rad c = trace_ray_classic(sc, ambient, &M_AIR, 0, r);
    rgb3[i * 3] = c.c[0]; rgb3[i * 3 + 1] = c.c[1]; rgb3[i * 3 + 2] = c.c[2];
  }
}
// direct light probe: get_radiance_from_light summed over lights (tracer.rs:136-141)
void orc_direct_light(const ppm_prim* prims, int np, const ppm_material* mats, int nm, const ppm_light* lights, int nl,
                      const double* pos3, const double* nrm3, int64_t n, double* rgb3) {
  Scene sc = mk_scene(prims, np, mats, nm, lights, nl);
  for (int64_t i = 0; i < n; ++i) {
    Rad rad = RAD0;
    for (int li = 0; li < nl; ++li)
      rad = radd(rad, get_radiance_from_light(sc, mk(pos3[i * 3], pos3[i * 3 + 1], pos3[i * 3 + 2]),
                                              mk(nrm3[i * 3], nrm3[i * 3 + 1], nrm3[i * 3 + 2]), lights[li]));
    rgb3[i * 3] = rad.c[0]; rgb3[i * 3 + 1] = rad.c[1]; rgb3[i * 3 + 2] = rad.c[2];
  }
}

// ---- one whole pass = `ppmpa` main (ppmpa.rs:21-46,74-84), rows [row0,row1) ----
// times_s[0] photon trace, [1] map build, [2] eye trace (incl. gather); stats3 as above.
// Single threaded, exactly like one reference process.
int orc_render_pass(const ppm_prim* prims, int np, const ppm_material* mats, int nm, const ppm_light* lights, int nl,
                    const ppm_camera* cam, uint64_t seed, uint32_t pass, int64_t nphoton, double radius2, int uc,
                    int row0, int row1, double* rgb3, double* times_s, uint64_t* stats4) {
  typedef std::chrono::steady_clock clk;
  Scene sc = mk_scene(prims, np, mats, nm, lights, nl);
  Flt flux = 0.0;
  for (int i = 0; i < nl; ++i) flux = flux + lights[i].flux;           // ppmpa.rs:30
  Flt power = flux / (Flt)nphoton;
  std::vector<int64_t> ns(nl);
  for (int i = 0; i < nl; ++i) ns[i] = (int64_t)std::round(lights[i].flux / power);  // ppmpa.rs:70-72
  auto t0 = clk::now();
  std::vector<Rec> recs;
  trace_all_photons(sc, seed, pass, uc != 0, ns.data(), recs);
  auto t1 = clk::now();
  PhotonMap m;
  m.power = power; m.radius = radius2;
  m.phs.resize(recs.size());
  for (size_t i = 0; i < recs.size(); ++i) m.phs[i] = recs[i].ph;
  m.build();
  auto t2 = clk::now();
  EyeCtx cx;
  cx.sc = &sc; cx.pmap = &m; cx.pfilter = cam->pfilter; cx.radius = radius2; cx.uc = uc != 0;
  cx.seed = seed; cx.pass = pass; cx.n_nodes = cx.n_gather = cx.sum_k = 0;
  for (int y = row0; y < row1; ++y)
    for (int x = 0; x < cam->xreso; ++x) {
      uint64_t pix = (uint64_t)y * cam->xreso + x;
      PhiloxRng rng(seed, pass, DOMAIN_EYE, pix, 0);
      Ray r = generate_ray(*cam, (Flt)y, (Flt)x, rng);
      cx.pixel = pix;
      Rad c = trace_ray(cx, &M_AIR, 0, r, 1);
      size_t o = ((size_t)(y - row0) * cam->xreso + x) * 3;
      rgb3[o] = c.c[0]; rgb3[o + 1] = c.c[1]; rgb3[o + 2] = c.c[2];
    }
  auto t3 = clk::now();
  if (times_s) {
    times_s[0] = std::chrono::duration<double>(t1 - t0).count();
    times_s[1] = std::chrono::duration<double>(t2 - t1).count();
    times_s[2] = std::chrono::duration<double>(t3 - t2).count();
  }
  if (stats4) { stats4[0] = recs.size(); stats4[1] = cx.n_nodes; stats4[2] = cx.n_gather; stats4[3] = cx.sum_k; }
  return 0;
}
// NPARA-style pass parallelism (util/iterator.rb:18,111-117): `nthreads` independent
// single-threaded passes (pass0 .. pass0+nthreads-1) run concurrently; rows [row0,row1) each.
// out_rgb3 may be NULL (timing only).  times_s[nthreads][3], stats[nthreads][4].
int orc_render_passes_parallel(const ppm_prim* prims, int np, const ppm_material* mats, int nm, const ppm_light* lights, int nl,
                               const ppm_camera* cam, uint64_t seed, uint32_t pass0, int nthreads, int64_t nphoton,
                               const double* radius2_per_pass, int uc, int row0, int row1, double* times_s, uint64_t* stats) {
  std::vector<std::thread> th;
  std::vector<std::vector<double>> img((size_t)nthreads);
  for (int t = 0; t < nthreads; ++t) {
    img[(size_t)t].resize((size_t)(row1 - row0) * cam->xreso * 3);
    th.emplace_back([=, &img]() {
      orc_render_pass(prims, np, mats, nm, lights, nl, cam, seed, pass0 + (uint32_t)t, nphoton, radius2_per_pass[t], uc,
                      row0, row1, img[(size_t)t].data(), times_s ? times_s + t * 3 : nullptr, stats ? stats + t * 4 : nullptr);
    });
  }
  for (auto& x : th) x.join();
  return 0;
}

// The same, with a row band PER THREAD (row0s[t] .. row1s[t]): bench.py spreads the bands evenly over the image so
// that the extrapolated pass time is a stratified sample of the image, not its centre rows.
int orc_render_passes_bands(const ppm_prim* prims, int np, const ppm_material* mats, int nm, const ppm_light* lights, int nl,
                            const ppm_camera* cam, uint64_t seed, const uint32_t* pass_ids, int nthreads, int64_t nphoton,
                            const double* radius2_per_pass, int uc, const int* row0s, const int* row1s, double* times_s, uint64_t* stats) {
  std::vector<std::thread> th;
  std::vector<std::vector<double>> img((size_t)nthreads);
  for (int t = 0; t < nthreads; ++t) {
    img[(size_t)t].resize((size_t)(row1s[t] - row0s[t]) * cam->xreso * 3);
    th.emplace_back([=, &img]() {
      orc_render_pass(prims, np, mats, nm, lights, nl, cam, seed, pass_ids[t], nphoton, radius2_per_pass[t], uc,
                      row0s[t], row1s[t], img[(size_t)t].data(), times_s ? times_s + t * 3 : nullptr, stats ? stats + t * 4 : nullptr);
    });
  }
  for (auto& x : th) x.join();
  return 0;
}

}  // extern "C"
