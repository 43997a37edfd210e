// cull_table.cuh -- host side of the shadow-ray culling of k_direct_light: the per-scene table (kernels_eye.cuh: DevCull).
// Part of the single translation unit engine.cu.
#ifndef PPM_CULL_TABLE_CUH_
#define PPM_CULL_TABLE_CUH_

#include "kernels_eye.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

// a polygon / parallelogram lying in the plane of light l (unit normal nl): every vertex within 1e-12 (relative to the
// scene scale) of the plane through the quad -- the emitter's own geometry
static inline bool cull_in_light_plane(const ppm_light& l, const double nl[3], const ppm_prim& s) {
  if (s.type != PPM_SHAPE_POLYGON && s.type != PPM_SHAPE_PARALLELOGRAM) return false;
  auto len3 = [](const double* a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); };
  const double scale = 1.0 + len3(l.pos) + len3(s.position) + len3(s.dir1) + len3(s.dir2);
  for (int j = 0; j < 4; ++j) {
    double h = 0.0;
    for (int k = 0; k < 3; ++k)
      h += nl[k] * ((s.position[k] + ((j & 1) ? s.dir1[k] : 0.0) + ((j & 2) ? s.dir2[k] : 0.0)) - l.pos[k]);
    if (!(std::fabs(h) <= 1e-12 * scale)) return false;
  }
  return true;
}

// Per-scene table for the conservative shadow-ray culling of k_direct_light (kernels_eye.cuh).
// Everything here is a bound with a 1e-6 safety margin, never a quantity that enters a result.
static inline void build_cull(const DevScene& sc, DevCull& cu) {
  std::memset(&cu, 0, sizeof cu);
  auto len3 = [](const double* a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); };
  auto quad_sphere = [&](const double* p0, const double* d1, const double* d2, double* c, double* r) {
    double s[3], d[3];
    for (int k = 0; k < 3; ++k) { s[k] = d1[k] + d2[k]; d[k] = d1[k] - d2[k]; c[k] = p0[k] + 0.5 * s[k]; }
    double rr = 0.5 * std::max(len3(s), len3(d));
    *r = rr * (1.0 + 1e-6) + 1e-6 * (1.0 + len3(c));
  };
  for (int o = 0; o < sc.nprims; ++o) {
    const ppm_prim& s = sc.prims[o];
    CullPrim& cp = cu.prim[o];
    if (s.type == PPM_SHAPE_PLAIN) {
      cp.kind = 1;
      double nl = std::max(1.0, len3(s.nvec));
      double scale = nl * (1.0 + std::fabs(s.scalar));
      cp.c[0] = 1e-6 * scale;    // D: sign margin on dist + n.p
      cp.c[1] = 1e-2 * scale;    // gap: the light must be closer to the plane than the node by this much
    } else if (s.type == PPM_SHAPE_SPHERE) {
      cp.kind = 2;
      for (int k = 0; k < 3; ++k) cp.c[k] = s.position[k];
      cp.r = std::fabs(s.scalar) * (1.0 + 1e-6) + 1e-6 * (1.0 + len3(s.position));
    } else if (s.type == PPM_SHAPE_POLYGON || s.type == PPM_SHAPE_PARALLELOGRAM) {
      cp.kind = 2;               // the triangle u + v <= 1 is a subset of its parallelogram
      quad_sphere(s.position, s.dir1, s.dir2, cp.c, &cp.r);
      cp.nvtx = 4;
      for (int j = 0; j < 4; ++j)
        for (int k = 0; k < 3; ++k)
          cp.vtx[j][k] = s.position[k] + ((j == 1 || j == 2) ? s.dir1[k] : 0.0) + ((j >= 2) ? s.dir2[k] : 0.0);
      for (int j = 0; j < 4 && cp.nvtx; ++j)
        for (int k = 0; k < 3; ++k)
          if (!(std::fabs(cp.vtx[j][k]) < 1e150)) cp.nvtx = 0;
    } else {
      cp.kind = 0;               // Point: calc_distance never yields a root
    }
    if (cp.kind == 2 && !(cp.r < 1e150)) cp.kind = 3;   // non-finite geometry: always tested
  }
  for (int li = 0; li < sc.nlights; ++li) {
    const ppm_light& l = sc.lights[li];
    CullLight& cl = cu.light[li];
    if (l.type != PPM_LIGHT_PARALLELOGRAM) continue;
    quad_sphere(l.pos, l.dir1, l.dir2, cl.c, &cl.r);
    {
      // scales of the "which side of the light's plane" tests (light_never_tested, k_direct_light): the spread of
      // nvec . (sample - p) over the samples and the magnitudes that bound its rounding error
      auto l1 = [](const double* a) { return std::fabs(a[0]) + std::fabs(a[1]) + std::fabs(a[2]); };
      auto dt = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
      cl.side_e = std::fabs(dt(l.nvec, l.dir1)) + std::fabs(dt(l.nvec, l.dir2));
      cl.ext = l1(l.pos) + l1(l.dir1) + l1(l.dir2);
      cl.ln1 = l1(l.nvec);
      if (!(cl.side_e < 1e150) || !(cl.ext < 1e150) || !(cl.ln1 < 1e150)) cl.side_e = cl.ext = cl.ln1 = 1e300;   // never certain
    }
    for (int j = 0; j < 4; ++j)
      for (int k = 0; k < 3; ++k)
        cl.corner[j][k] = l.pos[k] + ((j == 1 || j == 2) ? l.dir1[k] : 0.0) + ((j >= 2) ? l.dir2[k] : 0.0);
    if (!(cl.r < 1e150)) { cl.r = 1e300; }              // r^2 overflows -> the cone test is off, planes below stay valid or NaN
    // unit normal of the light's plane and the polygons / parallelograms lying in it (the emitter's own
    // geometry): every vertex within 1e-12 (relative to the scene scale) of the plane through the quad
    {
      const double cx[3] = {l.dir1[1] * l.dir2[2] - l.dir2[1] * l.dir1[2], l.dir1[2] * l.dir2[0] - l.dir2[2] * l.dir1[0],
                            l.dir1[0] * l.dir2[1] - l.dir2[0] * l.dir1[1]};
      const double cn = len3(cx);
      if (cn > 0.0 && cn < 1e150) {
        for (int k = 0; k < 3; ++k) cl.nl[k] = cx[k] / cn;
        for (int o = 0; o < sc.nprims; ++o) {
          const ppm_prim& s = sc.prims[o];
          if (cull_in_light_plane(l, cl.nl, s)) cl.coplanar |= 1ull << o;
        }
      }
    }
    for (int o = 0; o < sc.nprims; ++o) {
      const ppm_prim& s = sc.prims[o];
      if (s.type != PPM_SHAPE_PLAIN) continue;
      double hmin = 0.0, hmax = 0.0;
      for (int j = 0; j < 4; ++j) {
        double h = s.scalar;
        for (int k = 0; k < 3; ++k) h += s.nvec[k] * (l.pos[k] + ((j & 1) ? l.dir1[k] : 0.0) + ((j & 2) ? l.dir2[k] : 0.0));
        if (j == 0 || h < hmin) hmin = h;
        if (j == 0 || h > hmax) hmax = h;
        if (!(h == h)) { hmin = -1e300; hmax = 1e300; break; }   // NaN: the plane is always tested
      }
      cl.hmin[o] = hmin; cl.hmax[o] = hmax;
    }
  }
}

#endif
