// engine.cu -- CUDA engine behind include/ppm.h (sm_100a, f64, no FMA).
//
// This file holds the context, the host-side orchestration (streams, lanes, buffers) and the C ABI.
// The kernels live in headers of the same translation unit (one per hot-path stage of SURVEY.md 8a):
//   kernels_photon.cuh  k_intersect (calc_intersection probe, tracer.rs:306-350), k_emit (light.rs:67-91),
//                       k_trace_photons (tracer.rs:31-125), k_import / k_export
//   kernels_map.cuh     k_bbox, k_axis_hist, k_cell_key, k_scatter: the photon map as a radix-sorted uniform
//                       grid (replaces the kd-tree of photonmap.rs:23-29)
//   kernels_gather.cuh  k_query_key, k_gather (estimate_radiance, tracer.rs:179-216), k_knn_*, k_within
//   kernels_eye.cuh     k_gen_rays (camera.rs:58-75), k_eye_expand (trace_ray, tracer.rs:129-177),
//                       k_direct_light (tracer.rs:263-290), k_combine (surface.rs:135-206, averager2.rb:49-62)
//   dev_core.cuh        f64 math in reference order, Philox, nearest_hit, BSDF sampling
#include "dev_core.cuh"
#include "kernels_photon.cuh"
#include "kernels_map.cuh"
#include "kernels_gather.cuh"
#include "kernels_eye.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

// ===========================================================================
// context
// ===========================================================================
struct DBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes * 2 + 256;   // geometric growth: steady state never reallocates
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T* as() const { return (T*)p; }
};

struct ppm_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  std::string err;
  bool have_scene = false, have_camera = false, have_map = false;
  DevScene scene;
  DBuf dl_dbg, dl_masks;
  const unsigned long long* dl_masks_cur = nullptr;   // masks of the current node list (render_pass)
  DBuf cull;                      // DevCull: per-scene table for the shadow-ray culling of k_direct_light
  ppm_camera cam;
  // unsorted records
  DBuf r_pos, r_dir, r_wl, r_tag, counter;
  uint64_t n_rec = 0;
  int tag_bits = 38;              // bits of the largest record tag (photon << 4 | depth) of the current photon set
  double power = 0.0;
  // map
  DBuf keys, keys2, vals, vals2, cub_tmp, cell_start, hist, bbox, axis_hist;
  DBuf m_P, m_D, m_orig;
  DBuf q_key, q_key2, q_idx, q_idx2, heavy;
  DBuf knn_lo, knn_hi, knn_thr, knn_cnt;
  Grid grid;
  double r2 = 0.0;
  // staging for h_or_d arguments
  DBuf st_in0, st_in1, st_out0, st_out1, st_out2, st_out3, st_out4;
  // eye path
  DBuf e_head, e_prev, e_pos, e_nrm, e_w, e_emit, e_direct, e_photon, e_rays;
  uint64_t eye_cap = 0;            // capacity of the gather-node pool
  DBuf pass_img, accum, npass, stats;
  uint64_t accum_pixels = 0;
  // last pass stats
  double ms[8] = {0};
  uint64_t counters[8] = {0};
  // Two streams: the photon branch (trace -> map build) and the eye branch (expand -> direct
  // light) of a pass are independent until the gather, so render_pass runs them concurrently.
  cudaStream_t stream2 = nullptr;
  DBuf cub_tmp2;
  enum { EV_A0, EV_A1, EV_A2, EV_A3, EV_A4, EV_A5, EV_A6, EV_A7, EV_A8, EV_B0, EV_B1, EV_B2, EV_B3, EV_COUNT };
  cudaEvent_t ev[EV_COUNT] = {nullptr};
  bool timed = false;             // record the phase events (render_pass only)
  uint64_t launches = 0;
  std::vector<ppm_ctx*> twins;    // further lanes on the same GPU (ppm_render_passes), owned
};

#define PPM_DEFAULT_LANES 2
#define PPM_MAX_LANES 8

namespace {

#define CK(ctx, call)                                                                      \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                    \
      return PPM_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)
#define KCHECK(ctx)                                                                        \
  do {                                                                                     \
    (ctx)->launches++;                                                                     \
    cudaError_t e__ = cudaGetLastError();                                                  \
    if (e__ != cudaSuccess) {                                                              \
      (ctx)->err = std::string("kernel launch: ") + cudaGetErrorString(e__);               \
      return PPM_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

int fail(ppm_ctx* c, int code, const std::string& m) { if (c) c->err = m; return code; }

bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
// input: returns a device pointer holding `bytes` of user data
int stage_in(ppm_ctx* c, const void* user, size_t bytes, DBuf& scratch, const void** dev) {
  if (is_device_ptr(user)) { *dev = user; return PPM_OK; }
  CK(c, scratch.ensure(bytes));
  CK(c, cudaMemcpyAsync(scratch.p, user, bytes, cudaMemcpyHostToDevice, c->stream));
  *dev = scratch.p;
  return PPM_OK;
}
// output: returns a device pointer to write; finish_out copies back if user is host memory
int stage_out(ppm_ctx* c, void* user, size_t bytes, DBuf& scratch, void** dev) {
  if (!user) { *dev = nullptr; return PPM_OK; }
  if (is_device_ptr(user)) { *dev = user; return PPM_OK; }
  CK(c, scratch.ensure(bytes));
  *dev = scratch.p;
  return PPM_OK;
}
int finish_out(ppm_ctx* c, void* user, size_t bytes, void* dev) {
  if (!user || user == dev) return PPM_OK;
  CK(c, cudaMemcpyAsync(user, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
  return PPM_OK;
}
inline unsigned nblk(int64_t n, int b) { return (unsigned)((n + b - 1) / b); }

RecBuf recbuf(ppm_ctx* c) {
  RecBuf r;
  r.pos3 = c->r_pos.as<double>(); r.dir3 = c->r_dir.as<double>(); r.wl = c->r_wl.as<uint8_t>(); r.tag = c->r_tag.as<uint64_t>();
  return r;
}
int ensure_records(ppm_ctx* c, uint64_t cap) {
  CK(c, c->r_pos.ensure(cap * 24)); CK(c, c->r_dir.ensure(cap * 24));
  CK(c, c->r_wl.ensure(cap)); CK(c, c->r_tag.ensure(cap * 8));
  CK(c, c->counter.ensure(64));
  return PPM_OK;
}
MapSoA mapsoa(ppm_ctx* c) {
  MapSoA m;
  m.P = c->m_P.as<double2>(); m.D = c->m_D.as<double2>(); m.orig = c->m_orig.as<uint32_t>();
  return m;
}
// bits needed for the tags (index << 4 | depth) of `count` photons; cell bits go above them in the sort key
int tag_bits_for(uint64_t count) {
  int b = 1;
  while (b < 38 && (1ull << b) < count) ++b;
  return std::min(38, b + 4);
}
int light_split(ppm_ctx* c, const int64_t* n_per_light, LightSplit* ls, int64_t* total) {
  int64_t acc = 0;
  for (int i = 0; i < c->scene.nlights; ++i) {
    if (n_per_light[i] < 0) return fail(c, PPM_ERR_ARG, "negative photon count");
    ls->first[i] = acc; acc += n_per_light[i];
  }
  for (int i = c->scene.nlights; i <= PPM_MAX_LIGHTS; ++i) ls->first[i] = acc;
  *total = acc;
  return PPM_OK;
}

// Per-scene table for the conservative shadow-ray culling of k_direct_light (kernels_eye.cuh).
// Everything here is a bound with a 1e-6 safety margin, never a quantity that enters a result.
void build_cull(const DevScene& sc, DevCull& cu) {
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
          if (s.type != PPM_SHAPE_POLYGON && s.type != PPM_SHAPE_PARALLELOGRAM) continue;
          bool in_plane = true;
          double scale = 1.0 + len3(l.pos) + len3(s.position) + len3(s.dir1) + len3(s.dir2);
          for (int j = 0; j < 4 && in_plane; ++j) {
            double h = 0.0;
            for (int k = 0; k < 3; ++k)
              h += cl.nl[k] * ((s.position[k] + ((j & 1) ? s.dir1[k] : 0.0) + ((j & 2) ? s.dir2[k] : 0.0)) - l.pos[k]);
            if (!(std::fabs(h) <= 1e-12 * scale)) in_plane = false;
          }
          if (in_plane) cl.coplanar |= 1ull << o;
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
int upload_cull(ppm_ctx* c) {
  static_assert(sizeof(DevCull) < (1 << 16), "cull table");
  DevCull cu;
  build_cull(c->scene, cu);
  CK(c, c->cull.ensure(sizeof cu));
  CK(c, cudaMemcpy(c->cull.p, &cu, sizeof cu, cudaMemcpyHostToDevice));
  return PPM_OK;
}
// PPM_DL_CULL=0 switches the culling off (parity tests compare both settings bit for bit)
const DevCull* cull_arg(ppm_ctx* c) {
  const char* e = std::getenv("PPM_DL_CULL");
  if (e && e[0] == '0') return nullptr;
  return c->cull.as<DevCull>();
}

// k_dl_classify: per-node culling masks for every light (nullptr result = culling off: PPM_DL_CULL=0 or 64 primitives)
int launch_dl_classify(ppm_ctx* c, cudaStream_t st, const double* dpos, int64_t n, const unsigned long long** masks_out) {
  *masks_out = nullptr;
  const DevCull* cull = cull_arg(c);
  if (!cull || c->scene.nprims > 63 || c->scene.nlights <= 0 || n <= 0) return PPM_OK;
  CK(c, c->dl_masks.ensure((size_t)n * (size_t)c->scene.nlights * 8));
  k_dl_classify<<<nblk(n, 128), 128, 0, st>>>(c->scene, cull, dpos, n, c->dl_masks.as<unsigned long long>());
  KCHECK(c);
  *masks_out = c->dl_masks.as<unsigned long long>();
  return PPM_OK;
}
int launch_direct_light(ppm_ctx* c, cudaStream_t st, const double* dpos, const double* dnrm, int64_t n, double* dout,
                        const unsigned long long* masks, const uint32_t* order = nullptr) {
  unsigned long long* dbg = nullptr;
  const bool stats = std::getenv("PPM_DL_STATS") != nullptr;
  if (stats) {
    CK(c, c->dl_dbg.ensure(160));
    CK(c, cudaMemsetAsync(c->dl_dbg.p, 0, 160, st));
    dbg = c->dl_dbg.as<unsigned long long>();
  }
  k_direct_light<<<nblk(n, 128), 128, 0, st>>>(c->scene, masks, order, dpos, dnrm, n, dout, dbg);
  KCHECK(c);
  if (stats) {
    unsigned long long h[20];
    CK(c, cudaMemcpyAsync(h, dbg, 160, cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    if (h[0]) std::fprintf(stderr, "[ppm direct light] nodes=%llu tested prims/node: own %.3f, warp union %.3f; certificate %.1f%%\n", h[0],
                           (double)h[1] / h[0], (double)h[2] / h[0], 100.0 * h[3] / h[0]);
    if (h[0]) {
      std::fprintf(stderr, "[ppm direct light] nodes by tested prims 0..7+: own");
      for (int k = 0; k < 8; ++k) std::fprintf(stderr, " %.1f%%", 100.0 * h[4 + k] / h[0]);
      std::fprintf(stderr, " | warp union");
      for (int k = 0; k < 8; ++k) std::fprintf(stderr, " %.1f%%", 100.0 * h[12 + k] / h[0]);
      std::fprintf(stderr, "\n");
    }
  }
  return PPM_OK;
}

// -- internal (device-resident) building blocks shared by the probes and render_pass --
int trace_photons_launch(ppm_ctx* c, uint64_t seed, uint32_t pass, int uc, const int64_t* n_per_light, uint64_t* cap_out) {
  LightSplit ls;
  int64_t total = 0;
  int rc = light_split(c, n_per_light, &ls, &total);
  if (rc) return rc;
  uint64_t cap = (uint64_t)total * PPM_MAX_TRACE;
  if (cap == 0) cap = 1;
  c->tag_bits = tag_bits_for((uint64_t)total);
  rc = ensure_records(c, cap);
  if (rc) return rc;
  CK(c, cudaMemsetAsync(c->counter.p, 0, 16, c->stream));     // [0] record counter, [1] photon ticket
  if (total > 0) {
    // persistent grid: enough CTAs to fill every SM (resident CTAs are limited by registers),
    // never more threads than photons
    const unsigned blocks = (unsigned)std::min<int64_t>((total + 127) / 128, (int64_t)c->sm_count * 8);
    k_trace_photons<<<blocks, 128, 0, c->stream>>>(c->scene, ls, seed, pass, uc, total, recbuf(c),
                                                   c->counter.as<unsigned long long>(), cap,
                                                   c->counter.as<unsigned long long>() + 1);
    KCHECK(c);
  }
  *cap_out = cap;
  return PPM_OK;
}
int trace_photons_finish(ppm_ctx* c, uint64_t cap, double power) {
  unsigned long long n = 0;
  CK(c, cudaMemcpyAsync(&n, c->counter.p, 8, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  if (n > cap) return fail(c, PPM_ERR_CAPACITY, "photon record capacity exceeded");
  c->n_rec = n; c->power = power; c->have_map = false;
  return PPM_OK;
}
int do_trace_photons(ppm_ctx* c, uint64_t seed, uint32_t pass, int uc, const int64_t* n_per_light, double power) {
  uint64_t cap = 0;
  int rc = trace_photons_launch(c, seed, pass, uc, n_per_light, &cap);
  if (rc) return rc;
  return trace_photons_finish(c, cap, power);
}

struct HostTrace {
  bool on; std::chrono::steady_clock::time_point t0; std::string log; cudaStream_t st;
  explicit HostTrace(cudaStream_t s) : on(std::getenv("PPM_TRACE") != nullptr), t0(std::chrono::steady_clock::now()), st(s) {}
  void mark(const char* what) {
    if (!on) return;
    cudaStreamSynchronize(st);
    auto t1 = std::chrono::steady_clock::now();
    char b[96];
    std::snprintf(b, sizeof b, " %s=%.3f", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    log += b; t0 = t1;
  }
  ~HostTrace() { if (on) std::fprintf(stderr, "[ppm trace]%s\n", log.c_str()); }
};

int do_map_build(ppm_ctx* c, double radius2) {
  HostTrace tr(c->stream);
  if (!(radius2 > 0.0)) return fail(c, PPM_ERR_ARG, "radius2 must be > 0");
  const uint64_t n = c->n_rec;
  c->r2 = radius2;
  Grid g;
  std::memset(&g, 0, sizeof g);
  double cell = std::sqrt(radius2) * (1.0 + 1.0 / 1024.0);   // edge slightly > r: the 27-cell walk can never miss
  g.nx = g.ny = g.nz = 1; g.inv_cell = 1.0 / cell; g.ncells = 1;
  if (n > 0) {
    CK(c, c->bbox.ensure(48));
    unsigned long long init[6] = {~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull};
    CK(c, cudaMemcpyAsync(c->bbox.p, init, 48, cudaMemcpyHostToDevice, c->stream));
    k_bbox<<<std::min<unsigned>(nblk((int64_t)n, 256), 148 * 8), 256, 0, c->stream>>>(c->r_pos.as<double>(), n,
                                                                                     c->bbox.as<unsigned long long>());
    KCHECK(c);
    unsigned long long mm[6];
    CK(c, cudaMemcpyAsync(mm, c->bbox.p, 48, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    double lo[3], hi[3];
    for (int k = 0; k < 3; ++k) { lo[k] = dec_ord(mm[k]); hi[k] = dec_ord(mm[3 + k]); }
    tr.mark("bbox");
    for (int k = 0; k < 3; ++k)
      if (!(lo[k] == lo[k]) || !(hi[k] == hi[k]) || std::isinf(lo[k]) || std::isinf(hi[k]))
        return fail(c, PPM_ERR_ARG, "photon positions are not finite");
    // Region covered by the dense grid.  The reference leaks a few photons (~1e-4) through
    // wall corners (a bounce closer than NEARLY0 to the next wall skips it), and those land
    // tens of metres outside the room on the infinite planes, so the raw bounding box is
    // erratic and mostly empty.  The grid therefore covers a per-axis TRIMMED range (at most
    // n/1024 photons cut on each side, found with device histograms); photons and queries
    // outside are clamped into the boundary cells, which keeps the 27-cell walk exact
    // (clamping never increases the cell distance of two points).
    const uint64_t trim = n >= 4096 ? n / 1024 : 0;
    if (trim > 0) {
      CK(c, c->axis_hist.ensure(3 * AXIS_BINS * 4));
      std::vector<uint32_t> hh(3 * AXIS_BINS);
      for (int iter = 0; iter < 4; ++iter) {
        AxisRange ar;
        double w[3];
        bool fine = true;
        for (int k = 0; k < 3; ++k) {
          w[k] = (hi[k] - lo[k]) / (double)AXIS_BINS;
          if (!(w[k] > 0.0)) w[k] = 1.0;
          ar.lo[k] = lo[k]; ar.inv_w[k] = 1.0 / w[k];
          if (w[k] > 0.5 * cell) fine = false;
        }
        if (iter > 0 && fine) break;
        CK(c, cudaMemsetAsync(c->axis_hist.p, 0, 3 * AXIS_BINS * 4, c->stream));
        k_axis_hist<<<std::min<unsigned>(nblk((int64_t)n, 256 * 8), 148 * 4), 256, 0, c->stream>>>(c->r_pos.as<double>(), n, ar,
                                                                                              c->axis_hist.as<uint32_t>());
        KCHECK(c);
        CK(c, cudaMemcpyAsync(hh.data(), c->axis_hist.p, 3 * AXIS_BINS * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
        for (int k = 0; k < 3; ++k) {
          const uint32_t* h = hh.data() + k * AXIS_BINS;
          uint64_t acc = 0;
          int a = 0, b = AXIS_BINS - 1;
          while (a < AXIS_BINS - 1 && acc + h[a] <= trim) acc += h[a++];
          acc = 0;
          while (b > a && acc + h[b] <= trim) acc += h[b--];
          double nlo = lo[k] + (double)a * w[k], nhi = lo[k] + (double)(b + 1) * w[k];
          lo[k] = std::max(lo[k], nlo); hi[k] = std::min(hi[k], nhi);
        }
        if (fine) break;
      }
      tr.mark("trim");
    }
    // The origin is padded by half a cell: surfaces that bound the photon cloud (the room's
    // walls) then sit mid-cell, so the +-1 ulp noise of hit points on such a plane cannot
    // straddle a cell boundary (which would split every warp of queries on that wall).
    const double CELL_CAP = 67108864.0;   // 2^26 cells
    for (;;) {
      double dims[3];
      for (int k = 0; k < 3; ++k) dims[k] = std::floor((hi[k] - (lo[k] - 0.5 * cell)) / cell) + 2.0;
      if (dims[0] * dims[1] * dims[2] <= CELL_CAP && dims[0] < 2e9 && dims[1] < 2e9 && dims[2] < 2e9) {
        g.nx = (int32_t)dims[0]; g.ny = (int32_t)dims[1]; g.nz = (int32_t)dims[2];
        break;
      }
      cell *= 2.0;
    }
    for (int k = 0; k < 3; ++k) g.org[k] = lo[k] - 0.5 * cell;
    g.inv_cell = 1.0 / cell;
    g.ncells = (uint32_t)g.nx * (uint32_t)g.ny * (uint32_t)g.nz;
  }
  c->grid = g;
  {
    // cudaFree/cudaMalloc stall for 10-400 ms on this platform: size the cell tables for at
    // least 2^24 cells up front (64 MB per table) so the shrinking radius does not regrow them every few passes
    size_t want = std::max<size_t>((size_t)g.ncells + 1, (size_t)1 << 24) * 4;
    CK(c, c->hist.ensure(want));
    CK(c, c->cell_start.ensure(want));
  }
  tr.mark("alloc_cells");
  CK(c, cudaMemsetAsync(c->hist.p, 0, ((size_t)g.ncells + 1) * 4, c->stream));
  tr.mark("memset");
  size_t nn = n ? n : 1;
  CK(c, c->keys.ensure(nn * 8)); CK(c, c->keys2.ensure(nn * 8));
  CK(c, c->vals.ensure(nn * 4)); CK(c, c->vals2.ensure(nn * 4));
  CK(c, c->m_P.ensure(nn * 32)); CK(c, c->m_D.ensure(nn * 32)); CK(c, c->m_orig.ensure(nn * 4));
  tr.mark("alloc_map");
  if (n > 0) {
    k_cell_key<<<nblk((int64_t)n, 256), 256, 0, c->stream>>>(g, c->r_pos.as<double>(), c->r_tag.as<uint64_t>(), n, c->tag_bits,
                                                            c->keys.as<uint64_t>(), c->vals.as<uint32_t>(), c->hist.as<uint32_t>());
    KCHECK(c);
    int cell_bits = 1;
    while ((1ull << cell_bits) < (unsigned long long)g.ncells) ++cell_bits;
    int end_bit = std::min(64, c->tag_bits + cell_bits);
    size_t tmp = 0;
    CK(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp, c->keys.as<uint64_t>(), c->keys2.as<uint64_t>(), c->vals.as<uint32_t>(),
                                          c->vals2.as<uint32_t>(), (int64_t)n, 0, end_bit, c->stream));
    CK(c, c->cub_tmp.ensure(tmp));
    tr.mark("key");
    CK(c, cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp, c->keys.as<uint64_t>(), c->keys2.as<uint64_t>(), c->vals.as<uint32_t>(),
                                          c->vals2.as<uint32_t>(), (int64_t)n, 0, end_bit, c->stream));
    tr.mark("sort");
    k_scatter<<<nblk((int64_t)n, 256), 256, 0, c->stream>>>(recbuf(c), c->vals2.as<uint32_t>(), n, mapsoa(c));
    KCHECK(c);
    tr.mark("scatter");
  }
  {
    size_t tmp = 0;
    CK(c, cub::DeviceScan::ExclusiveSum(nullptr, tmp, c->hist.as<uint32_t>(), c->cell_start.as<uint32_t>(), (int64_t)g.ncells + 1, c->stream));
    CK(c, c->cub_tmp.ensure(tmp));
    CK(c, cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp, c->hist.as<uint32_t>(), c->cell_start.as<uint32_t>(), (int64_t)g.ncells + 1, c->stream));
  }
  tr.mark("scan");
  c->have_map = true;
  return PPM_OK;
}

// key the queries by cell and sort them (stable: ties keep query order -> deterministic)
int gather_sort_queries(ppm_ctx* c, const double* dpos, int64_t n) {
  if (n >= (1ll << 32)) return fail(c, PPM_ERR_CAPACITY, "at most 2^32-1 gather queries per call");
  CK(c, c->q_key.ensure((size_t)n * 4)); CK(c, c->q_key2.ensure((size_t)n * 4));
  CK(c, c->q_idx.ensure((size_t)n * 4)); CK(c, c->q_idx2.ensure((size_t)n * 4));
  k_query_key<<<nblk(n, 256), 256, 0, c->stream>>>(c->grid, dpos, n, c->q_key.as<uint32_t>(), c->q_idx.as<uint32_t>());
  KCHECK(c);
  unsigned long long maxkey = c->grid.ncells;
  int bits = 1;
  while (bits < 32 && (1ull << bits) < maxkey) ++bits;
  size_t tmp = 0;
  CK(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp, c->q_key.as<uint32_t>(), c->q_key2.as<uint32_t>(), c->q_idx.as<uint32_t>(),
                                        c->q_idx2.as<uint32_t>(), n, 0, bits, c->stream));
  CK(c, c->cub_tmp.ensure(tmp));
  CK(c, cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp, c->q_key.as<uint32_t>(), c->q_key2.as<uint32_t>(), c->q_idx.as<uint32_t>(),
                                        c->q_idx2.as<uint32_t>(), n, 0, bits, c->stream));
  return PPM_OK;
}
// warp-cooperative gather over the sorted queries; mode 0 fixed radius, 1 per-query radius, 2 count only
int gather_launch(ppm_ctx* c, const double* dpos, const double* dnrm, int64_t n, int filter, int mode, const double* r2q,
                  double* drgb, uint32_t* dcounts, unsigned long long* dsumk) {
  const int B = GATHER_WARPS * 32;
  const uint32_t* cs = c->cell_start.as<uint32_t>();
  const uint32_t* qk = c->q_key2.as<uint32_t>();
  const uint32_t* qx = c->q_idx2.as<uint32_t>();
  // heavy groups: warps publish them as parts in a device-side list; k_gather_heavy is launched only if there are any
  HeavyList hl;
  std::memset(&hl, 0, sizeof hl);
  if (!(std::getenv("PPM_GATHER_HEAVY") && std::getenv("PPM_GATHER_HEAVY")[0] == '0')) {
    const uint32_t cap = 1u << 15;                     // parts: 32 MB of partial sums per context
    const size_t off_groups = 64, off_parts = off_groups + (size_t)(cap / 2) * sizeof(HeavyGroup),
                 off_partials = off_parts + (size_t)cap * sizeof(HeavyPart), bytes = off_partials + (size_t)cap * sizeof(HeavyPartial);
    CK(c, c->heavy.ensure(bytes));
    char* base = c->heavy.as<char>();
    hl.ctr = (unsigned int*)base; hl.groups = (HeavyGroup*)(base + off_groups); hl.parts = (HeavyPart*)(base + off_parts);
    hl.partials = (HeavyPartial*)(base + off_partials);
    hl.cap_parts = cap;
    CK(c, cudaMemsetAsync(hl.ctr, 0, 16, c->stream));  // [0] reservations, [1] ticket of k_gather_heavy, [2] groups, [3] parts published
  }
  if (c->timed) cudaEventRecord(c->ev[ppm_ctx::EV_A4], c->stream);   // start of k_gather
#define GATHER_LAUNCH(F, M) k_gather<F, M><<<nblk(n, B), B, 0, c->stream>>>(c->grid, cs, mapsoa(c), qk, qx, dpos, dnrm, n, c->power, c->r2, r2q, drgb, dcounts, dsumk, hl)
  if (mode == 2) GATHER_LAUNCH(PPM_FILTER_NONE, 2);
  else if (mode == 1) {
    switch (filter) {
      case PPM_FILTER_NONE: GATHER_LAUNCH(PPM_FILTER_NONE, 1); break;
      case PPM_FILTER_CONE: GATHER_LAUNCH(PPM_FILTER_CONE, 1); break;
      default:              GATHER_LAUNCH(PPM_FILTER_GAUSS, 1); break;
    }
  } else {
    switch (filter) {
      case PPM_FILTER_NONE: GATHER_LAUNCH(PPM_FILTER_NONE, 0); break;
      case PPM_FILTER_CONE: GATHER_LAUNCH(PPM_FILTER_CONE, 0); break;
      default:              GATHER_LAUNCH(PPM_FILTER_GAUSS, 0); break;
    }
  }
#undef GATHER_LAUNCH
  KCHECK(c);
  if (c->timed) cudaEventRecord(c->ev[ppm_ctx::EV_A5], c->stream);   // end of k_gather (re-recorded after k_gather_heavy)
  if (hl.ctr) {
    unsigned int nparts = 0;
    CK(c, cudaMemcpyAsync(&nparts, hl.ctr + 3, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    if (nparts > 0) {
      const unsigned grid = std::min<unsigned>((nparts + GATHER_WARPS - 1) / GATHER_WARPS, (unsigned)c->sm_count * 16u);
#define HEAVY_LAUNCH(F, M) k_gather_heavy<F, M><<<grid, B, 0, c->stream>>>(c->grid, cs, mapsoa(c), qx, dpos, dnrm, n, c->power, c->r2, r2q, drgb, dcounts, dsumk, hl)
      if (mode == 2) HEAVY_LAUNCH(PPM_FILTER_NONE, 2);
      else if (mode == 1) {
        switch (filter) {
          case PPM_FILTER_NONE: HEAVY_LAUNCH(PPM_FILTER_NONE, 1); break;
          case PPM_FILTER_CONE: HEAVY_LAUNCH(PPM_FILTER_CONE, 1); break;
          default:              HEAVY_LAUNCH(PPM_FILTER_GAUSS, 1); break;
        }
      } else {
        switch (filter) {
          case PPM_FILTER_NONE: HEAVY_LAUNCH(PPM_FILTER_NONE, 0); break;
          case PPM_FILTER_CONE: HEAVY_LAUNCH(PPM_FILTER_CONE, 0); break;
          default:              HEAVY_LAUNCH(PPM_FILTER_GAUSS, 0); break;
        }
      }
#undef HEAVY_LAUNCH
      KCHECK(c);
      if (c->timed) cudaEventRecord(c->ev[ppm_ctx::EV_A5], c->stream);
    }
  }
  return PPM_OK;
}
int launch_gather(ppm_ctx* c, const double* dpos, const double* dnrm, int64_t n, int filter, double* drgb, uint32_t* dcounts,
                  unsigned long long* dsumk) {
  if (n <= 0) return PPM_OK;
  if (filter < PPM_FILTER_NONE || filter > PPM_FILTER_GAUSS) return fail(c, PPM_ERR_ARG, "bad filter");
  int rc = gather_sort_queries(c, dpos, n);
  if (rc) return rc;
  return gather_launch(c, dpos, dnrm, n, filter, 0, nullptr, drgb, dcounts, dsumk);
}
// k-NN estimate (no reference implementation exists: n_sample_photon is dead code, photonmap.rs:18,
// camera.rs:181; semantics defined in SURVEY.md 8c): for every query the k nearest photons within r;
// if k are found, r_k^2 = the k-th smallest d2 replaces r^2 in the membership test, the filter and the
// normaliser; otherwise the fixed radius is used.  r_k^2 is found EXACTLY by bisection on the bit
// pattern of d2 with the count-only gather (<= 63 steps).
int launch_gather_knn(ppm_ctx* c, const double* dpos, const double* dnrm, int64_t n, uint32_t k, int filter, double* drgb,
                      double* dr2k, uint32_t* dcounts) {
  if (n <= 0) return PPM_OK;
  int rc = gather_sort_queries(c, dpos, n);
  if (rc) return rc;
  CK(c, c->knn_lo.ensure((size_t)n * 8)); CK(c, c->knn_hi.ensure((size_t)n * 8)); CK(c, c->knn_thr.ensure((size_t)n * 8));
  CK(c, c->knn_cnt.ensure((size_t)n * 4)); CK(c, c->counter.ensure(64));
  unsigned long long* lo = c->knn_lo.as<unsigned long long>();
  unsigned long long* hi = c->knn_hi.as<unsigned long long>();
  double* thr = c->knn_thr.as<double>();
  uint32_t* cnt = c->knn_cnt.as<uint32_t>();
  unsigned int* nact = c->counter.as<unsigned int>() + 8;
  // photons within the fixed radius
  if ((rc = gather_launch(c, dpos, dnrm, n, PPM_FILTER_NONE, 0, nullptr, drgb, cnt, nullptr))) return rc;
  CK(c, cudaMemsetAsync(nact, 0, 4, c->stream));
  k_knn_init<<<nblk(n, 256), 256, 0, c->stream>>>(n, c->r2, cnt, k, lo, hi, thr, nact);
  KCHECK(c);
  for (int it = 0; it < 70; ++it) {
    unsigned int active = 0;
    CK(c, cudaMemcpyAsync(&active, nact, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    if (!active) break;
    if ((rc = gather_launch(c, dpos, dnrm, n, PPM_FILTER_NONE, 2, thr, nullptr, cnt, nullptr))) return rc;
    CK(c, cudaMemsetAsync(nact, 0, 4, c->stream));
    k_knn_step<<<nblk(n, 256), 256, 0, c->stream>>>(n, cnt, k, lo, hi, thr, nact);
    KCHECK(c);
  }
  k_knn_finish<<<nblk(n, 256), 256, 0, c->stream>>>(n, c->r2, thr);
  KCHECK(c);
  if ((rc = gather_launch(c, dpos, dnrm, n, filter, 1, thr, drgb, dcounts, nullptr))) return rc;
  if (dr2k) CK(c, cudaMemcpyAsync(dr2k, thr, (size_t)n * 8, cudaMemcpyDeviceToDevice, c->stream));
  return PPM_OK;
}

// Eye branch, part 1 (stream `st`): expand the eye paths into the gather-node list and
// compute the classic direct light at every node.  drays == NULL generates camera rays.
int eye_front(ppm_ctx* c, cudaStream_t st, DBuf& tmpbuf, const double* drays, int64_t n, int64_t first_pixel, uint64_t seed,
              uint32_t pass, int uc, uint32_t* nn_out, int classic = 0, bool defer_direct = false) {
  (void)tmpbuf;
  CK(c, c->e_head.ensure((size_t)n * 4));
  CK(c, c->e_emit.ensure((size_t)n * 24)); CK(c, c->stats.ensure(64));
  unsigned long long* dstats = c->stats.as<unsigned long long>();
  if (c->eye_cap < (uint64_t)n * 2) c->eye_cap = (uint64_t)n * 2;      // first guess: two gather nodes per pixel
  if (c->timed) cudaEventRecord(c->ev[ppm_ctx::EV_B0], st);
  uint32_t nn = 0;
  for (int attempt = 0;; ++attempt) {
    const size_t cap = (size_t)c->eye_cap;
    if (cap >= 0xFFFFFFFFull) return fail(c, PPM_ERR_CAPACITY, "more than 2^32-1 gather nodes");
    CK(c, c->e_pos.ensure(cap * 24)); CK(c, c->e_nrm.ensure(cap * 24)); CK(c, c->e_w.ensure(cap * 24));
    CK(c, c->e_prev.ensure(cap * 4));
    CK(c, c->e_direct.ensure(cap * 24)); CK(c, c->e_photon.ensure(cap * 24));
    CK(c, cudaMemsetAsync(c->stats.p, 0, 64, st));
    EyeNodes nodes = {c->e_pos.as<double>(), c->e_nrm.as<double>(), c->e_w.as<double>(), c->e_prev.as<uint32_t>()};
    k_eye_expand<<<nblk(n, 128), 128, 0, st>>>(c->scene, c->cam, drays, n, first_pixel, seed, pass, nodes, (uint32_t)cap,
                                              c->e_head.as<uint32_t>(), c->e_emit.as<double>(), dstats + 2, dstats, classic);
    KCHECK(c);
    unsigned long long made = 0;
    CK(c, cudaMemcpyAsync(&made, dstats + 2, 8, cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    if (made <= cap) { nn = (uint32_t)made; break; }
    if (attempt >= 2) return fail(c, PPM_ERR_CAPACITY, "gather-node pool keeps overflowing");
    c->eye_cap = made + made / 4;                           // pool overflow: grow and walk again
  }
  EyeNodes nodes = {c->e_pos.as<double>(), c->e_nrm.as<double>(), c->e_w.as<double>(), c->e_prev.as<uint32_t>()};
  cudaEventRecord(c->ev[ppm_ctx::EV_B1], st);
  *nn_out = nn;
  c->dl_masks_cur = nullptr;
  if (uc && nn) {                                            // culling masks: right after the expansion, beside the photon branch
    int rc = launch_dl_classify(c, st, nodes.pos3, nn, &c->dl_masks_cur);
    if (rc) return rc;
  }
  if (defer_direct) return PPM_OK;                           // render_pass launches the direct light in cell-sorted order
  cudaEventRecord(c->ev[ppm_ctx::EV_B3], st);
  if (uc && nn) {
    int rc = launch_direct_light(c, st, nodes.pos3, nodes.nrm3, nn, c->e_direct.as<double>(), c->dl_masks_cur);
    if (rc) return rc;
  }
  cudaEventRecord(c->ev[ppm_ctx::EV_B2], st);
  return PPM_OK;
}
// Eye branch, part 2 (main stream; the photon map must be built): gather at every node.
// Needs the node list (event EV_B1) but not the direct light, so in render_pass it runs
// concurrently with k_direct_light.
int eye_gather(ppm_ctx* c, uint32_t nn, int uc_sorted_direct = 0) {
  if (c->timed) cudaEventRecord(c->ev[ppm_ctx::EV_A3], c->stream);
  if (nn) {
    if (c->cam.pfilter < PPM_FILTER_NONE || c->cam.pfilter > PPM_FILTER_GAUSS) return fail(c, PPM_ERR_ARG, "bad filter");
    int rc = gather_sort_queries(c, c->e_pos.as<double>(), nn);
    if (rc) return rc;
    if (uc_sorted_direct) {
      // direct light on stream2, over the nodes in the cell-sorted order the gather uses (coherent culling masks);
      // it runs concurrently with k_gather on the main stream
      cudaEventRecord(c->ev[ppm_ctx::EV_A8], c->stream);
      cudaStreamWaitEvent(c->stream2, c->ev[ppm_ctx::EV_A8], 0);
      cudaEventRecord(c->ev[ppm_ctx::EV_B3], c->stream2);
      rc = launch_direct_light(c, c->stream2, c->e_pos.as<double>(), c->e_nrm.as<double>(), nn, c->e_direct.as<double>(),
                               c->dl_masks_cur, c->q_idx2.as<uint32_t>());
      if (rc) return rc;
      cudaEventRecord(c->ev[ppm_ctx::EV_B2], c->stream2);
    }
    rc = gather_launch(c, c->e_pos.as<double>(), c->e_nrm.as<double>(), nn, c->cam.pfilter, 0, nullptr, c->e_photon.as<double>(), nullptr,
                       c->stats.as<unsigned long long>() + 1);
    if (rc) return rc;
  } else if (uc_sorted_direct) {
    cudaEventRecord(c->ev[ppm_ctx::EV_B3], c->stream2);
    cudaEventRecord(c->ev[ppm_ctx::EV_B2], c->stream2);
  }
  c->counters[3] = nn;
  return PPM_OK;
}
// Eye branch, part 3 (main stream, after direct light AND gather): combine per pixel,
// optionally add into the accumulator.
int eye_combine(ppm_ctx* c, int64_t n, int64_t first_pixel, int uc, double* dout, double* daccum, int classic = 0) {
  const D3 amb = {c->cam.ambient[0], c->cam.ambient[1], c->cam.ambient[2]};
  k_combine<<<nblk(n, 256), 256, 0, c->stream>>>(c->e_head.as<uint32_t>(), c->e_prev.as<uint32_t>(), c->e_w.as<double>(),
                                                (uc || classic) ? c->e_direct.as<double>() : nullptr,
                                                classic ? nullptr : c->e_photon.as<double>(), c->e_emit.as<double>(), n, dout, daccum,
                                                first_pixel, amb);
  KCHECK(c);
  return PPM_OK;
}
// serial version on the main stream (ppm_trace_rays)
int do_trace_rays(ppm_ctx* c, const double* drays, int64_t n, int64_t first_pixel, uint64_t seed, uint32_t pass, int uc,
                  double* dout, double* daccum, int classic = 0) {
  if (n <= 0) return PPM_OK;
  uint32_t nn = 0;
  int rc = eye_front(c, c->stream, c->cub_tmp, drays, n, first_pixel, seed, pass, classic ? 1 : uc, &nn, classic);
  if (rc) return rc;
  if (!classic && (rc = eye_gather(c, nn))) return rc;
  return eye_combine(c, n, first_pixel, uc, dout, daccum, classic);
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int ppm_create(int device, ppm_ctx** out) {
  if (!out) return PPM_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return PPM_ERR_NODEVICE; }
  if (device < 0 || device >= ndev) return PPM_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return PPM_ERR_CUDA;
  ppm_ctx* c = new ppm_ctx();
  c->device = device;
  cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (c->sm_count <= 0) c->sm_count = 148;
  // The main stream gets the highest priority: in render_pass its small photon-branch kernels
  // must be dispatched ahead of the remaining blocks of the long eye-branch kernels on stream2.
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { delete c; return PPM_ERR_CUDA; }
  if (cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, prio_lo) != cudaSuccess) { cudaStreamDestroy(c->stream); delete c; return PPM_ERR_CUDA; }
  for (int i = 0; i < ppm_ctx::EV_COUNT; ++i) cudaEventCreate(&c->ev[i]);
  std::memset(&c->scene, 0, sizeof c->scene);
  ppm_camera_default(&c->cam);
  *out = c;
  return PPM_OK;
}

void ppm_destroy(ppm_ctx* c) {
  if (!c) return;
  for (ppm_ctx* t : c->twins) ppm_destroy(t);
  c->twins.clear();
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  DBuf* all[] = {&c->r_pos, &c->r_dir, &c->r_wl, &c->r_tag, &c->counter, &c->keys, &c->keys2, &c->vals, &c->vals2, &c->cub_tmp,
                 &c->cell_start, &c->hist, &c->bbox, &c->axis_hist, &c->m_P, &c->m_D, &c->m_orig, &c->q_key, &c->q_key2, &c->q_idx, &c->q_idx2, &c->knn_lo, &c->knn_hi, &c->knn_thr, &c->knn_cnt,
                 &c->st_in0, &c->st_in1, &c->st_out0, &c->st_out1, &c->st_out2, &c->st_out3, &c->st_out4,
                 &c->e_head, &c->e_prev, &c->e_pos, &c->e_nrm, &c->e_w, &c->e_emit, &c->e_direct, &c->e_photon, &c->e_rays,
                 &c->pass_img, &c->accum, &c->npass, &c->stats, &c->cub_tmp2, &c->cull, &c->dl_dbg, &c->dl_masks, &c->heavy};
  for (DBuf* b : all) b->release();
  for (int i = 0; i < ppm_ctx::EV_COUNT; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  cudaStreamSynchronize(c->stream2);
  cudaStreamDestroy(c->stream2);
  cudaStreamDestroy(c->stream);
  delete c;
}

const char* ppm_last_error(const ppm_ctx* c) { return c ? c->err.c_str() : "null context"; }
void* ppm_stream(ppm_ctx* c) { return c ? (void*)c->stream : nullptr; }

int ppm_scene_set(ppm_ctx* c, const ppm_prim* prims, int32_t nprims, const ppm_material* mats, int32_t nmats,
                  const ppm_light* lights, int32_t nlights) {
  if (!c) return PPM_ERR_ARG;
  if (!prims || !mats || nprims <= 0 || nmats <= 0 || nlights < 0 || (nlights > 0 && !lights)) return fail(c, PPM_ERR_ARG, "null/empty scene arrays");
  if (nprims > PPM_MAX_PRIMS || nmats > PPM_MAX_MATS || nlights > PPM_MAX_LIGHTS) return fail(c, PPM_ERR_CAPACITY, "scene exceeds 64 prims / 48 materials / 8 lights");
  for (int i = 0; i < nprims; ++i) {
    if (prims[i].material < 0 || prims[i].material >= nmats) return fail(c, PPM_ERR_ARG, "primitive material index out of range");
    if (prims[i].type < PPM_SHAPE_POINT || prims[i].type > PPM_SHAPE_PARALLELOGRAM) return fail(c, PPM_ERR_ARG, "bad shape type");
  }
  for (int i = 0; i < nlights; ++i)
    if (lights[i].type < PPM_LIGHT_POINT || lights[i].type > PPM_LIGHT_SUN) return fail(c, PPM_ERR_ARG, "bad light type");
  std::memset(&c->scene, 0, sizeof c->scene);
  c->scene.nprims = nprims; c->scene.nmats = nmats; c->scene.nlights = nlights;
  std::memcpy(c->scene.prims, prims, sizeof(ppm_prim) * nprims);
  std::memcpy(c->scene.mats, mats, sizeof(ppm_material) * nmats);
  if (nlights) std::memcpy(c->scene.lights, lights, sizeof(ppm_light) * nlights);
  c->scene.types.nwords = nprims > 32 ? 2 : 1;
  for (int o = 0; o < nprims; ++o) {
    const unsigned long long bit = 1ull << o;
    if (prims[o].type == PPM_SHAPE_PLAIN) c->scene.types.plain |= bit;
    else if (prims[o].type == PPM_SHAPE_SPHERE) c->scene.types.sphere |= bit;
    else if (prims[o].type == PPM_SHAPE_POLYGON) c->scene.types.poly |= bit;
    else if (prims[o].type == PPM_SHAPE_PARALLELOGRAM) c->scene.types.para |= bit;
  }
  CK(c, cudaSetDevice(c->device));
  int rc = upload_cull(c);
  if (rc) return rc;
  c->have_scene = true;
  return PPM_OK;
}

int ppm_camera_set(ppm_ctx* c, const ppm_camera* cam) {
  if (!c || !cam) return PPM_ERR_ARG;
  if (cam->xreso <= 0 || cam->yreso <= 0) return fail(c, PPM_ERR_ARG, "bad resolution");
  if (cam->pfilter < PPM_FILTER_NONE || cam->pfilter > PPM_FILTER_GAUSS) return fail(c, PPM_ERR_ARG, "bad photon filter");
  c->cam = *cam;
  c->have_camera = true;
  return PPM_OK;
}

int ppm_intersect(ppm_ctx* c, const double* rays6, int64_t n, int32_t* hit_idx, double* t, double* pos3, double* nrm3, int32_t* io) {
  if (!c) return PPM_ERR_ARG;
  if (!c->have_scene) return fail(c, PPM_ERR_STATE, "scene not set");
  if (n < 0 || (n > 0 && (!rays6 || !hit_idx))) return fail(c, PPM_ERR_ARG, "null rays / hit_idx");
  if (n == 0) return PPM_OK;
  CK(c, cudaSetDevice(c->device));
  const void* drays; void *dh, *dt, *dp, *dn, *di;
  int rc;
  if ((rc = stage_in(c, rays6, (size_t)n * 48, c->st_in0, &drays))) return rc;
  if ((rc = stage_out(c, hit_idx, (size_t)n * 4, c->st_out0, &dh))) return rc;
  if ((rc = stage_out(c, t, (size_t)n * 8, c->st_out1, &dt))) return rc;
  if ((rc = stage_out(c, pos3, (size_t)n * 24, c->st_out2, &dp))) return rc;
  if ((rc = stage_out(c, nrm3, (size_t)n * 24, c->st_out3, &dn))) return rc;
  if ((rc = stage_out(c, io, (size_t)n * 4, c->st_out4, &di))) return rc;
  k_intersect<<<nblk(n, 128), 128, 0, c->stream>>>(c->scene, (const double*)drays, n, (int32_t*)dh, (double*)dt, (double*)dp,
                                                  (double*)dn, (int32_t*)di);
  KCHECK(c);
  if ((rc = finish_out(c, hit_idx, (size_t)n * 4, dh))) return rc;
  if ((rc = finish_out(c, t, (size_t)n * 8, dt))) return rc;
  if ((rc = finish_out(c, pos3, (size_t)n * 24, dp))) return rc;
  if ((rc = finish_out(c, nrm3, (size_t)n * 24, dn))) return rc;
  if ((rc = finish_out(c, io, (size_t)n * 4, di))) return rc;
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

int ppm_emit_photons(ppm_ctx* c, uint64_t seed, uint32_t pass, const int64_t* n_per_light, ppm_photon* out) {
  if (!c) return PPM_ERR_ARG;
  if (!c->have_scene || c->scene.nlights == 0) return fail(c, PPM_ERR_STATE, "scene with lights not set");
  if (!n_per_light || !out) return fail(c, PPM_ERR_ARG, "null argument");
  CK(c, cudaSetDevice(c->device));
  LightSplit ls; int64_t total;
  int rc = light_split(c, n_per_light, &ls, &total);
  if (rc) return rc;
  if (total == 0) return PPM_OK;
  void* d;
  if ((rc = stage_out(c, out, (size_t)total * sizeof(ppm_photon), c->st_out0, &d))) return rc;
  k_emit<<<nblk(total, 128), 128, 0, c->stream>>>(c->scene, ls, seed, pass, total, (ppm_photon*)d);
  KCHECK(c);
  if ((rc = finish_out(c, out, (size_t)total * sizeof(ppm_photon), d))) return rc;
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

int ppm_trace_photons(ppm_ctx* c, uint64_t seed, uint32_t pass, int uc, const int64_t* n_per_light, double power, uint64_t* n_stored) {
  if (!c) return PPM_ERR_ARG;
  if (!c->have_scene || c->scene.nlights == 0) return fail(c, PPM_ERR_STATE, "scene with lights not set");
  if (!n_per_light) return fail(c, PPM_ERR_ARG, "null n_per_light");
  CK(c, cudaSetDevice(c->device));
  int rc = do_trace_photons(c, seed, pass, uc, n_per_light, power);
  if (rc) return rc;
  if (n_stored) *n_stored = c->n_rec;
  return PPM_OK;
}

int ppm_photons_count(ppm_ctx* c, uint64_t* n, double* power) {
  if (!c) return PPM_ERR_ARG;
  if (n) *n = c->n_rec;
  if (power) *power = c->power;
  return PPM_OK;
}

int ppm_photons_export(ppm_ctx* c, ppm_photon* out, uint64_t cap, uint64_t* tags) {
  if (!c) return PPM_ERR_ARG;
  if (c->n_rec == 0) return PPM_OK;
  if (!out) return fail(c, PPM_ERR_ARG, "null output");
  if (cap < c->n_rec) return fail(c, PPM_ERR_CAPACITY, "export buffer too small");
  CK(c, cudaSetDevice(c->device));
  void* d; int rc;
  size_t bytes = (size_t)c->n_rec * sizeof(ppm_photon);
  if ((rc = stage_out(c, out, bytes, c->st_out0, &d))) return rc;
  k_export<<<nblk((int64_t)c->n_rec, 256), 256, 0, c->stream>>>(recbuf(c), c->n_rec, (ppm_photon*)d);
  KCHECK(c);
  if ((rc = finish_out(c, out, bytes, d))) return rc;
  if (tags) CK(c, cudaMemcpyAsync(tags, c->r_tag.p, (size_t)c->n_rec * 8, is_device_ptr(tags) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

int ppm_photons_import(ppm_ctx* c, const ppm_photon* in, uint64_t n, double power) {
  if (!c) return PPM_ERR_ARG;
  if (n > 0 && !in) return fail(c, PPM_ERR_ARG, "null input");
  if (n >= (1ull << 32)) return fail(c, PPM_ERR_CAPACITY, "at most 2^32-1 photon records");
  CK(c, cudaSetDevice(c->device));
  int rc = ensure_records(c, n ? n : 1);
  if (rc) return rc;
  if (n) {
    const void* d;
    if ((rc = stage_in(c, in, (size_t)n * sizeof(ppm_photon), c->st_in0, &d))) return rc;
    k_import<<<nblk((int64_t)n, 256), 256, 0, c->stream>>>((const ppm_photon*)d, n, recbuf(c));
    KCHECK(c);
    CK(c, cudaStreamSynchronize(c->stream));
  }
  c->n_rec = n; c->power = power; c->have_map = false;
  c->tag_bits = tag_bits_for(n);
  return PPM_OK;
}

int ppm_map_build(ppm_ctx* c, double radius2) {
  if (!c) return PPM_ERR_ARG;
  CK(c, cudaSetDevice(c->device));
  int rc = do_map_build(c, radius2);
  if (rc) return rc;
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

int ppm_within(ppm_ctx* c, const double* q3, int64_t nq, uint32_t* idx, uint32_t* count, uint32_t cap) {
  if (!c) return PPM_ERR_ARG;
  if (!c->have_map) return fail(c, PPM_ERR_STATE, "photon map not built");
  if (nq < 0 || (nq > 0 && (!q3 || !count || (cap > 0 && !idx)))) return fail(c, PPM_ERR_ARG, "null argument");
  if (nq == 0) return PPM_OK;
  CK(c, cudaSetDevice(c->device));
  const void* dq; void *di, *dc; int rc;
  size_t ib = (size_t)nq * (cap ? cap : 1) * 4;
  if ((rc = stage_in(c, q3, (size_t)nq * 24, c->st_in0, &dq))) return rc;
  CK(c, c->st_out0.ensure(ib));
  di = c->st_out0.p;
  CK(c, cudaMemsetAsync(di, 0, ib, c->stream));      // slots beyond a query's count stay 0
  if ((rc = stage_out(c, count, (size_t)nq * 4, c->st_out1, &dc))) return rc;
  k_within<<<nblk(nq, 128), 128, 0, c->stream>>>(c->grid, c->cell_start.as<uint32_t>(), mapsoa(c), (const double*)dq, nq, c->r2,
                                                (uint32_t*)di, (uint32_t*)dc, cap);
  KCHECK(c);
  if ((rc = finish_out(c, count, (size_t)nq * 4, dc))) return rc;
  if (cap) {
    // sort each neighbour list ascending by photon index on the host (probe only)
    std::vector<uint32_t> h((size_t)nq * cap), hc((size_t)nq);
    CK(c, cudaMemcpyAsync(h.data(), di, ib, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpyAsync(hc.data(), dc, (size_t)nq * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    for (int64_t q = 0; q < nq; ++q) {
      uint32_t k = std::min(hc[(size_t)q], cap);
      std::sort(h.begin() + (size_t)q * cap, h.begin() + (size_t)q * cap + k);
    }
    if (is_device_ptr(idx)) CK(c, cudaMemcpy(idx, h.data(), ib, cudaMemcpyHostToDevice));
    else std::memcpy(idx, h.data(), ib);
  }
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

int ppm_gather(ppm_ctx* c, const double* pos3, const double* nrm3, int64_t n, int filter, double* rgb3, uint32_t* counts) {
  if (!c) return PPM_ERR_ARG;
  if (!c->have_map) return fail(c, PPM_ERR_STATE, "photon map not built");
  if (n < 0 || (n > 0 && (!pos3 || !nrm3 || !rgb3))) return fail(c, PPM_ERR_ARG, "null argument");
  if (filter < PPM_FILTER_NONE || filter > PPM_FILTER_GAUSS) return fail(c, PPM_ERR_ARG, "bad filter");
  if (n == 0) return PPM_OK;
  CK(c, cudaSetDevice(c->device));
  const void *dp, *dn; void *dr, *dc; int rc;
  if ((rc = stage_in(c, pos3, (size_t)n * 24, c->st_in0, &dp))) return rc;
  if ((rc = stage_in(c, nrm3, (size_t)n * 24, c->st_in1, &dn))) return rc;
  if ((rc = stage_out(c, rgb3, (size_t)n * 24, c->st_out0, &dr))) return rc;
  if ((rc = stage_out(c, counts, (size_t)n * 4, c->st_out1, &dc))) return rc;
  if ((rc = launch_gather(c, (const double*)dp, (const double*)dn, n, filter, (double*)dr, (uint32_t*)dc, nullptr))) return rc;
  if ((rc = finish_out(c, rgb3, (size_t)n * 24, dr))) return rc;
  if ((rc = finish_out(c, counts, (size_t)n * 4, dc))) return rc;
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

int ppm_gather_knn(ppm_ctx* c, const double* pos3, const double* nrm3, int64_t n, uint32_t k, int filter, double* rgb3,
                   double* r2k, uint32_t* counts) {
  if (!c) return PPM_ERR_ARG;
  if (!c->have_map) return fail(c, PPM_ERR_STATE, "photon map not built");
  if (n < 0 || k == 0 || (n > 0 && (!pos3 || !nrm3 || !rgb3))) return fail(c, PPM_ERR_ARG, "null argument or k == 0");
  if (filter < PPM_FILTER_NONE || filter > PPM_FILTER_GAUSS) return fail(c, PPM_ERR_ARG, "bad filter");
  if (n == 0) return PPM_OK;
  CK(c, cudaSetDevice(c->device));
  const void *dp, *dn; void *dr, *dk, *dc; int rc;
  if ((rc = stage_in(c, pos3, (size_t)n * 24, c->st_in0, &dp))) return rc;
  if ((rc = stage_in(c, nrm3, (size_t)n * 24, c->st_in1, &dn))) return rc;
  if ((rc = stage_out(c, rgb3, (size_t)n * 24, c->st_out0, &dr))) return rc;
  if ((rc = stage_out(c, r2k, (size_t)n * 8, c->st_out1, &dk))) return rc;
  if ((rc = stage_out(c, counts, (size_t)n * 4, c->st_out2, &dc))) return rc;
  if ((rc = launch_gather_knn(c, (const double*)dp, (const double*)dn, n, k, filter, (double*)dr, (double*)dk, (uint32_t*)dc))) return rc;
  if ((rc = finish_out(c, rgb3, (size_t)n * 24, dr))) return rc;
  if ((rc = finish_out(c, r2k, (size_t)n * 8, dk))) return rc;
  if ((rc = finish_out(c, counts, (size_t)n * 4, dc))) return rc;
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

int ppm_direct_light(ppm_ctx* c, const double* pos3, const double* nrm3, int64_t n, double* rgb3) {
  if (!c) return PPM_ERR_ARG;
  if (!c->have_scene) return fail(c, PPM_ERR_STATE, "scene not set");
  if (n < 0 || (n > 0 && (!pos3 || !nrm3 || !rgb3))) return fail(c, PPM_ERR_ARG, "null argument");
  if (n == 0) return PPM_OK;
  CK(c, cudaSetDevice(c->device));
  const void *dp, *dn; void* dr; int rc;
  if ((rc = stage_in(c, pos3, (size_t)n * 24, c->st_in0, &dp))) return rc;
  if ((rc = stage_in(c, nrm3, (size_t)n * 24, c->st_in1, &dn))) return rc;
  if ((rc = stage_out(c, rgb3, (size_t)n * 24, c->st_out0, &dr))) return rc;
  const unsigned long long* masks = nullptr;
  if ((rc = launch_dl_classify(c, c->stream, (const double*)dp, n, &masks))) return rc;
  if ((rc = launch_direct_light(c, c->stream, (const double*)dp, (const double*)dn, n, (double*)dr, masks))) return rc;
  if ((rc = finish_out(c, rgb3, (size_t)n * 24, dr))) return rc;
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

int ppm_generate_rays(ppm_ctx* c, uint64_t seed, uint32_t pass, double* rays6) {
  if (!c) return PPM_ERR_ARG;
  if (!c->have_camera) return fail(c, PPM_ERR_STATE, "camera not set");
  if (!rays6) return fail(c, PPM_ERR_ARG, "null output");
  CK(c, cudaSetDevice(c->device));
  int64_t n = (int64_t)c->cam.xreso * c->cam.yreso;
  void* d; int rc;
  if ((rc = stage_out(c, rays6, (size_t)n * 48, c->st_out0, &d))) return rc;
  k_gen_rays<<<nblk(n, 128), 128, 0, c->stream>>>(c->cam, seed, pass, n, (double*)d);
  KCHECK(c);
  if ((rc = finish_out(c, rays6, (size_t)n * 48, d))) return rc;
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

int ppm_trace_rays_classic(ppm_ctx* c, const double* rays6, int64_t n, int64_t first_pixel, uint64_t seed, uint32_t pass, double* rgb3) {
  if (!c) return PPM_ERR_ARG;
  if (!c->have_scene) return fail(c, PPM_ERR_STATE, "scene not set");
  if (n < 0 || first_pixel < 0 || (n > 0 && (!rays6 || !rgb3))) return fail(c, PPM_ERR_ARG, "null argument");
  if (n == 0) return PPM_OK;
  CK(c, cudaSetDevice(c->device));
  const void* dr; void* dout; int rc;
  if ((rc = stage_in(c, rays6, (size_t)n * 48, c->st_in0, &dr))) return rc;
  if ((rc = stage_out(c, rgb3, (size_t)n * 24, c->st_out0, &dout))) return rc;
  if ((rc = do_trace_rays(c, (const double*)dr, n, first_pixel, seed, pass, 1, (double*)dout, nullptr, 1))) return rc;
  if ((rc = finish_out(c, rgb3, (size_t)n * 24, dout))) return rc;
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

int ppm_trace_rays(ppm_ctx* c, const double* rays6, int64_t n, int64_t first_pixel, uint64_t seed, uint32_t pass, int uc, double* rgb3) {
  if (!c) return PPM_ERR_ARG;
  if (!c->have_scene) return fail(c, PPM_ERR_STATE, "scene not set");
  if (!c->have_map) return fail(c, PPM_ERR_STATE, "photon map not built");
  if (n < 0 || first_pixel < 0 || (n > 0 && (!rays6 || !rgb3))) return fail(c, PPM_ERR_ARG, "null argument");
  if (n == 0) return PPM_OK;
  CK(c, cudaSetDevice(c->device));
  const void* dr; void* dout; int rc;
  if ((rc = stage_in(c, rays6, (size_t)n * 48, c->st_in0, &dr))) return rc;
  if ((rc = stage_out(c, rgb3, (size_t)n * 24, c->st_out0, &dout))) return rc;
  if ((rc = do_trace_rays(c, (const double*)dr, n, first_pixel, seed, pass, uc, (double*)dout, nullptr))) return rc;
  if ((rc = finish_out(c, rgb3, (size_t)n * 24, dout))) return rc;
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

static int ensure_accum(ppm_ctx* c) {
  uint64_t npix = (uint64_t)c->cam.xreso * (uint64_t)c->cam.yreso;
  if (c->accum_pixels == npix && c->accum.p) return PPM_OK;
  // sum image and pass counter live in ONE allocation so a single reduce covers both
  CK(c, c->accum.ensure((size_t)(npix * 3 + 1) * 8));
  CK(c, cudaMemsetAsync(c->accum.p, 0, (size_t)(npix * 3 + 1) * 8, c->stream));
  c->accum_pixels = npix;
  return PPM_OK;
}

int ppm_render_pass(ppm_ctx* c, uint64_t seed, uint32_t pass, int64_t nphoton, double radius2, int uc) {
  if (!c) return PPM_ERR_ARG;
  if (!c->have_scene || c->scene.nlights == 0) return fail(c, PPM_ERR_STATE, "scene with lights not set");
  if (!c->have_camera) return fail(c, PPM_ERR_STATE, "camera not set");
  if (nphoton <= 0 || !(radius2 > 0.0)) return fail(c, PPM_ERR_ARG, "nphoton and radius2 must be positive");
  CK(c, cudaSetDevice(c->device));
  int rc;
  double power;
  int64_t ns[PPM_MAX_LIGHTS];
  if ((rc = ppm_photon_budget(c->scene.lights, c->scene.nlights, nphoton, &power, ns))) return fail(c, rc, "photon budget");
  if ((rc = ensure_accum(c))) return rc;
  const int64_t npix = (int64_t)c->cam.xreso * c->cam.yreso;
  CK(c, c->pass_img.ensure((size_t)npix * 24));
  const uint64_t l0 = c->launches;
  typedef ppm_ctx E;
  c->timed = true;
  // photon branch on the main stream (async), eye branch on stream2, joined before the gather
  cudaEventRecord(c->ev[E::EV_A0], c->stream);
  CK(c, cudaStreamWaitEvent(c->stream2, c->ev[E::EV_A0], 0));
  uint64_t cap = 0;
  uint32_t nn = 0;
  rc = trace_photons_launch(c, seed, pass, uc, ns, &cap);
  cudaEventRecord(c->ev[E::EV_A1], c->stream);
  if (!rc) rc = eye_front(c, c->stream2, c->cub_tmp2, nullptr, npix, 0, seed, pass, uc, &nn, 0, /*defer_direct=*/uc != 0);
  if (!rc) rc = trace_photons_finish(c, cap, power);
  if (!rc) rc = do_map_build(c, radius2);
  cudaEventRecord(c->ev[E::EV_A2], c->stream);
  if (!rc) {
    cudaStreamWaitEvent(c->stream, c->ev[E::EV_B1], 0);      // node list ready
    rc = eye_gather(c, nn, uc);                               // k_gather runs concurrently with k_direct_light (stream2)
  }
  if (!rc) {
    cudaStreamWaitEvent(c->stream, c->ev[E::EV_B2], 0);      // direct light done
    if (c->timed) cudaEventRecord(c->ev[E::EV_A7], c->stream);
    rc = eye_combine(c, npix, 0, uc, c->pass_img.as<double>(), c->accum.as<double>());
  }
  c->timed = false;
  if (rc) { cudaStreamSynchronize(c->stream2); cudaStreamSynchronize(c->stream); return rc; }
  k_bump<<<1, 1, 0, c->stream>>>(c->accum.as<double>() + (size_t)npix * 3);
  KCHECK(c);
  cudaEventRecord(c->ev[E::EV_A6], c->stream);
  unsigned long long st[2];
  CK(c, cudaMemcpyAsync(st, c->stats.p, 16, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  auto el = [&](int a, int b) { float f = 0.f; cudaEventElapsedTime(&f, c->ev[a], c->ev[b]); return (double)f; };
  c->ms[0] = el(E::EV_A0, E::EV_A1);                 // photon trace
  c->ms[1] = el(E::EV_A1, E::EV_A2);                 // map build (includes waiting for host readbacks)
  c->ms[2] = el(E::EV_B0, E::EV_B1);                 // eye expand (stream2, concurrent with the photon branch)
  c->ms[3] = el(E::EV_B3, E::EV_B2);                 // direct light (stream2)
  c->ms[4] = el(E::EV_A3, E::EV_A5);                 // gather: query sort + kernel
  c->ms[5] = el(E::EV_A7, E::EV_A6);                 // combine + accumulate
  c->ms[6] = el(E::EV_A0, E::EV_A6);                 // whole pass
  c->ms[7] = nn ? el(E::EV_A4, E::EV_A5) : 0.0;      // k_gather alone
  if (std::getenv("PPM_TRACE"))
    std::fprintf(stderr, "[ppm timeline ms since A0] A1=%.3f A2=%.3f B0=%.3f B1=%.3f B2=%.3f A3=%.3f A4=%.3f A5=%.3f A7=%.3f A6=%.3f\n",
                 el(E::EV_A0, E::EV_A1), el(E::EV_A0, E::EV_A2), el(E::EV_A0, E::EV_B0), el(E::EV_A0, E::EV_B1), el(E::EV_A0, E::EV_B2),
                 el(E::EV_A0, E::EV_A3), nn ? el(E::EV_A0, E::EV_A4) : 0.0, el(E::EV_A0, E::EV_A5), el(E::EV_A0, E::EV_A7), el(E::EV_A0, E::EV_A6));
  int64_t emitted = 0;
  for (int i = 0; i < c->scene.nlights; ++i) emitted += ns[i];
  c->counters[0] = (uint64_t)emitted; c->counters[1] = c->n_rec; c->counters[2] = st[0];
  c->counters[4] = st[1]; c->counters[5] = c->launches - l0;
  return PPM_OK;
}

// A batch of passes with cross-pass overlap.  Passes are independent, and within one pass the
// FP64-bound kernels (direct light, gather) and the latency-bound ones (photon tracing, eye-path
// expansion, map build) cannot fill the GPU together.  Two lanes -- this context and an internal
// twin on the same GPU, each with its own streams and buffers, each driven by its own host
// thread -- render alternating passes so that different phases of two passes co-schedule.  The
// twin's accumulator is merged afterwards, so ppm_accum_read / ppm_image_mean see every pass.
int ppm_render_passes(ppm_ctx* c, uint64_t seed, uint32_t first_pass, uint32_t pass_stride, int32_t npass, int64_t nphoton,
                      const double* radius2, int uc) {
  if (!c) return PPM_ERR_ARG;
  if (npass < 0 || (npass > 0 && !radius2)) return fail(c, PPM_ERR_ARG, "bad pass batch");
  if (npass == 0) return PPM_OK;
  if (!c->have_scene || !c->have_camera) return fail(c, PPM_ERR_STATE, "scene and camera must be set");
  // lanes: this context plus (lanes - 1) internal twins on the same GPU; PPM_LANES overrides the default
  int lanes = PPM_DEFAULT_LANES;
  if (const char* e = std::getenv("PPM_LANES")) lanes = std::atoi(e);
  lanes = std::max(1, std::min(lanes, (int)PPM_MAX_LANES));
  lanes = std::min<int>(lanes, npass);
  double ms_sum[8] = {0}; uint64_t ct_sum[8] = {0};
  if (lanes == 1) {
    for (int32_t i = 0; i < npass; ++i) {
      int rc = ppm_render_pass(c, seed, first_pass + (uint32_t)i * pass_stride, nphoton, radius2[i], uc);
      if (rc) return rc;
      for (int k = 0; k < 8; ++k) { ms_sum[k] += c->ms[k]; ct_sum[k] += c->counters[k]; }
    }
  } else {
    std::vector<ppm_ctx*> lane(lanes, nullptr);
    lane[0] = c;
    for (int j = 1; j < lanes; ++j) {
      if ((int)c->twins.size() < j) {
        ppm_ctx* t = nullptr;
        int rc = ppm_create(c->device, &t);
        if (rc) return fail(c, rc, "cannot create another lane");
        c->twins.push_back(t);
      }
      ppm_ctx* t = c->twins[j - 1];
      t->scene = c->scene; t->cam = c->cam; t->have_camera = true;
      { int rc = upload_cull(t); if (rc) return fail(c, rc, "lane: " + t->err); }
      t->have_scene = true;
      lane[j] = t;
    }
    std::vector<int> rcs(lanes, PPM_OK);
    std::vector<std::array<double, 8>> lms(lanes);
    std::vector<std::array<uint64_t, 8>> lct(lanes);
    auto run_lane = [&](int j) {
      ppm_ctx* x = lane[j];
      cudaSetDevice(x->device);
      lms[j].fill(0.0); lct[j].fill(0);
      for (int32_t i = j; i < npass; i += lanes) {
        rcs[j] = ppm_render_pass(x, seed, first_pass + (uint32_t)i * pass_stride, nphoton, radius2[i], uc);
        if (rcs[j]) return;
        for (int k = 0; k < 8; ++k) { lms[j][k] += x->ms[k]; lct[j][k] += x->counters[k]; }
      }
    };
    std::vector<std::thread> workers;
    for (int j = 1; j < lanes; ++j) workers.emplace_back(run_lane, j);
    run_lane(0);
    for (auto& w : workers) w.join();
    if (rcs[0]) return rcs[0];
    for (int j = 1; j < lanes; ++j)
      if (rcs[j]) return fail(c, rcs[j], std::string("lane: ") + lane[j]->err);
    for (int j = 0; j < lanes; ++j)
      for (int k = 0; k < 8; ++k) { ms_sum[k] += lms[j][k]; ct_sum[k] += lct[j][k]; }
    // merge the twins' accumulators (and pass counters) into ours; last pass image follows the last pass
    CK(c, cudaSetDevice(c->device));
    const int64_t nacc = (int64_t)c->accum_pixels * 3 + 1;
    for (int j = 1; j < lanes; ++j) {
      CK(c, cudaStreamSynchronize(lane[j]->stream));
      k_accum_merge<<<nblk(nacc, 256), 256, 0, c->stream>>>(c->accum.as<double>(), lane[j]->accum.as<double>(), nacc);
      KCHECK(c);
    }
    const int last_lane = (npass - 1) % lanes;
    if (last_lane != 0)
      CK(c, cudaMemcpyAsync(c->pass_img.p, lane[last_lane]->pass_img.p, (size_t)c->accum_pixels * 24, cudaMemcpyDeviceToDevice, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
  }
  std::memcpy(c->ms, ms_sum, sizeof ms_sum);          // batch totals (sum over the passes of all lanes)
  std::memcpy(c->counters, ct_sum, sizeof ct_sum);
  return PPM_OK;
}

int ppm_last_pass_stats(ppm_ctx* c, double ms[8], uint64_t counters[8]) {
  if (!c) return PPM_ERR_ARG;
  if (ms) std::memcpy(ms, c->ms, sizeof c->ms);
  if (counters) std::memcpy(counters, c->counters, sizeof c->counters);
  return PPM_OK;
}

static int copy_out(ppm_ctx* c, void* user, const void* dev, size_t bytes) {
  CK(c, cudaMemcpyAsync(user, dev, bytes, is_device_ptr(user) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

int ppm_pass_image_read(ppm_ctx* c, double* rgb3) {
  if (!c || !rgb3) return PPM_ERR_ARG;
  if (!c->pass_img.p) return fail(c, PPM_ERR_STATE, "no pass rendered yet");
  CK(c, cudaSetDevice(c->device));
  return copy_out(c, rgb3, c->pass_img.p, (size_t)c->cam.xreso * c->cam.yreso * 24);
}

int ppm_accum_reset(ppm_ctx* c) {
  if (!c) return PPM_ERR_ARG;
  CK(c, cudaSetDevice(c->device));
  c->accum_pixels = 0;
  if (c->have_camera) return ensure_accum(c);
  return PPM_OK;
}

int ppm_accum_read(ppm_ctx* c, double* rgb3, uint32_t* n_pass) {
  if (!c) return PPM_ERR_ARG;
  if (!c->accum.p || !c->accum_pixels) return fail(c, PPM_ERR_STATE, "no accumulator yet");
  CK(c, cudaSetDevice(c->device));
  if (rgb3) { int rc = copy_out(c, rgb3, c->accum.p, (size_t)c->accum_pixels * 24); if (rc) return rc; }
  if (n_pass) {
    double np = 0.0;
    CK(c, cudaMemcpyAsync(&np, c->accum.as<double>() + c->accum_pixels * 3, 8, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    *n_pass = (uint32_t)np;
  }
  return PPM_OK;
}

int ppm_accum_device(ppm_ctx* c, void** sum_dev, void** npass_dev, uint64_t* n_doubles) {
  if (!c) return PPM_ERR_ARG;
  CK(c, cudaSetDevice(c->device));
  if (!c->have_camera) return fail(c, PPM_ERR_STATE, "camera not set");
  int rc = ensure_accum(c);
  if (rc) return rc;
  CK(c, cudaStreamSynchronize(c->stream));
  if (sum_dev) *sum_dev = c->accum.p;
  if (npass_dev) *npass_dev = c->accum.as<double>() + c->accum_pixels * 3;
  if (n_doubles) *n_doubles = c->accum_pixels * 3 + 1;
  return PPM_OK;
}

int ppm_image_mean(ppm_ctx* c, double* rgb3) {
  if (!c || !rgb3) return PPM_ERR_ARG;
  if (!c->accum.p || !c->accum_pixels) return fail(c, PPM_ERR_STATE, "no accumulator yet");
  CK(c, cudaSetDevice(c->device));
  int64_t n = (int64_t)c->accum_pixels * 3;
  void* d; int rc;
  if ((rc = stage_out(c, rgb3, (size_t)n * 8, c->st_out0, &d))) return rc;
  k_scale<<<nblk(n, 256), 256, 0, c->stream>>>(c->accum.as<double>(), c->accum.as<double>() + n, n, (double*)d);
  KCHECK(c);
  if ((rc = finish_out(c, rgb3, (size_t)n * 8, d))) return rc;
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

}  // extern "C"
