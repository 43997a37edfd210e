// engine.cu -- CUDA engine behind include/ppm.h (sm_100a, f64, no FMA).
//
// This file holds the context, the host-side orchestration (streams, lanes, buffers, CUDA graphs) and the C ABI.
// The kernels live in headers of the same translation unit (one per hot-path stage of SURVEY.md 8a):
//   pass_state.cuh      PassDev: the device-resident state of a pass (grid, radius, RNG keys, counters, stamps)
//   kernels_photon.cuh  k_intersect (calc_intersection probe, tracer.rs:306-350), k_emit (light.rs:67-91),
//                       k_trace_photons (tracer.rs:31-125), k_import / k_export
//   kernels_map.cuh     compact cell index, device-length scans, hand-written stable radix sort, counting sort of
//                       the queries (replaces the kd-tree of photonmap.rs:23-29)
//   kernels_gather.cuh  k_gather (estimate_radiance, tracer.rs:179-216), k_gather_heavy, k_knn_*, k_within
//   kernels_eye.cuh     k_gen_rays (camera.rs:58-75), k_eye_expand (trace_ray, tracer.rs:129-177),
//                       k_direct_light (tracer.rs:263-290), k_combine (surface.rs:135-206, averager2.rb:49-62)
//   dev_core.cuh        f64 math in reference order, Philox, nearest_hit, BSDF sampling
//
// A whole pass (ppmpa.rs:74-84) is a FIXED sequence of launches with fixed arguments: every size that used to be read
// back (record count, node count, occupied cells, heavy parts) stays in PassDev, grids are sized for the buffer
// capacities and kernels return early.  The sequence is captured once per lane as a CUDA graph; a batch of passes
// (util/iterator.rb:96-117) is N graph launches and ONE synchronisation at the end, where the per-pass reports
// (PassOut) are read and passes that overflowed a buffer are rendered again with larger buffers.
#include "dev_core.cuh"
#include "pass_state.cuh"
#include "kernels_photon.cuh"
#include "kernels_map.cuh"
#include "kernels_gather.cuh"
#include "kernels_eye.cuh"
#include "cull_table.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace ppmhost {   // host_bvh.cpp
bool bvh_build(const ppm_prim* prims, int64_t n, std::vector<BvhNode>& out_nodes, std::vector<BvhPrim>& out_prims, int* depth, std::string& err);
}

// ===========================================================================
// context
// ===========================================================================
struct DBuf {
  void* p = nullptr;
  size_t cap = 0;
  // grows to exactly what is asked (+ 1/16 slack): capacities are calibrated per scene, not doubled
  cudaError_t grow(size_t bytes) {
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 16 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T* as() const { return (T*)p; }
};

#define PPM_DEFAULT_LANES 2
#define PPM_MAX_LANES 8
#define PPM_CAL_PASS 0xFFFFFFFFu        // Philox pass id of the per-scene calibration (never a rendered pass)
#define PPM_CAL_PHOTONS (1 << 17)

struct CalKey {
  uint64_t scene_ver = 0, cam_ver = 0, seed = 0;
  int64_t nphoton = 0;
  int uc = 0;
  bool operator==(const CalKey& o) const { return scene_ver == o.scene_ver && cam_ver == o.cam_ver && seed == o.seed && nphoton == o.nphoton && uc == o.uc; }
};
struct GraphKey {
  uint64_t scene_ver = 0, cam_ver = 0, buf_gen = 0, rec_cap = 0, eye_cap = 0;
  int64_t nphoton = 0;
  int uc = 0, cull = 0, heavy = 0, accumulate = 0;
  bool operator==(const GraphKey& o) const {
    return scene_ver == o.scene_ver && cam_ver == o.cam_ver && buf_gen == o.buf_gen && rec_cap == o.rec_cap && eye_cap == o.eye_cap &&
           nphoton == o.nphoton && uc == o.uc && cull == o.cull && heavy == o.heavy && accumulate == o.accumulate;
  }
};

struct ppm_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;     // photon branch + gather + combine (highest priority)
  cudaStream_t stream2 = nullptr;    // eye branch: expansion, classification, query sort, direct light
  std::string err;
  bool have_scene = false, have_camera = false, have_map = false;
  uint64_t scene_ver = 0, cam_ver = 0, buf_gen = 0;
  DevScene scene;
  ppm_camera cam;
  // switches (ppm_option_set; initial values from the environment, read once at ppm_create)
  int opt_lanes = PPM_DEFAULT_LANES, opt_dl_cull = 1, opt_gather_heavy = 1, opt_dl_stats = 0, opt_graph = 1, opt_bvh = 0;
  // the scene as handed in (BVH mode keeps it to recognise "the same scene again") and its hierarchy (bvh_types.h)
  std::vector<ppm_prim> h_prims;
  std::vector<ppm_material> h_mats;
  std::vector<ppm_light> h_lights;
  DBuf bvh_nodes, bvh_prims, g_prims;
  uint64_t bvh_nnodes = 0, bvh_nprims = 0;
  int bvh_depth = 0;
  // pass state: device copy, host mirror (probe entry points), batch table and per-pass reports
  PassDev* ps = nullptr;
  PassDev hps;
  BatchDev* bt = nullptr;
  BatchDev* bt_h = nullptr;          // pinned
  PassOut* out = nullptr;
  PassOut* out_h = nullptr;          // pinned
  // unsorted records of the current photon set
  DBuf r_pos, r_dir, r_wl, r_tag, pmask, pbase;
  uint64_t n_rec = 0;
  bool rec_traced = false;           // records come from k_trace_photons (pmask valid) rather than an import
  uint64_t rec_nphoton = 0;          // emitted photons of the traced set
  double power = 0.0;
  // map build
  DBuf order0, cid, words_p, start_p, key_a, key_b, val_a, val_b, rs_table, tsum_p;
  DBuf m_P, m_D, m_orig;
  DBuf bbox, axis_hist;
  // query sort + gather
  DBuf qcell, words_q, qcnt, qstart, qrank, qpic, skey, sidx, tsum_q, heavy;
  DBuf knn_lo, knn_hi, knn_thr, knn_cnt, knn_act;
  // staging for h_or_d arguments
  DBuf st_in0, st_in1, st_out0, st_out1, st_out2, st_out3, st_out4;
  // eye path
  DBuf e_head, e_prev, e_pos, e_nrm, e_w, e_emit, e_direct, e_photon;
  DBuf dl_dbg, dl_masks, cull;
  uint64_t eye_cap = 0;              // capacity of the gather-node pool (probe entry points grow it)
  uint64_t rec_cap = 0;              // rendered pass: capacity of the record buffers
  uint64_t pass_eye_cap = 0;         // rendered pass: capacity of the gather-node pool
  Bounds cal_bounds;                 // rendered pass: grid region (per-scene calibration)
  int cal_have_bounds = 0;
  DBuf pass_img, accum;
  uint64_t accum_pixels = 0;
  // per-scene calibration: grid region, record and node capacities
  bool cal_valid = false;
  CalKey cal_key;
  // the pass as a CUDA graph
  cudaGraphExec_t gexec = nullptr;
  GraphKey gkey;
  uint64_t graph_kernels = 0;
  int prio_hi = 0, prio_lo = 0;      // stream priorities (main stream = highest)
  bool capturing = false;            // enq_pass is being recorded into a graph
  std::vector<cudaGraphNode_t> seg_before, hi_nodes;   // capture bookkeeping: nodes recorded from the main stream
  enum { EV_FORK, EV_NODES, EV_DL, EV_G0, EV_G1, EV_COUNT };
  cudaEvent_t ev[EV_COUNT] = {nullptr};
  // last pass / batch stats
  double ms[8] = {0};
  uint64_t counters[8] = {0};
  double timeline[PPM_NSTAMP] = {0};  // last pass of the last batch: stamps in ms since its begin
  uint64_t launches = 0;
  std::vector<ppm_ctx*> twins;       // further lanes on the same GPU (ppm_render_passes), owned
  void* comm = nullptr;              // ncclComm_t owned by the ctx (ppm_comm_init)
};

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
#define RC(call) do { int rc__ = (call); if (rc__) return rc__; } while (0)

int fail(ppm_ctx* c, int code, const std::string& m) { if (c) c->err = m; return code; }

// device buffer of at least `bytes`; a reallocation invalidates the captured graph (its nodes hold the old pointers)
int ens(ppm_ctx* c, DBuf& b, size_t bytes) {
  if (bytes <= b.cap && b.p) return PPM_OK;
  c->buf_gen++;
  CK(c, b.grow(bytes ? bytes : 1));
  return PPM_OK;
}

bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
// input: returns a device pointer holding `bytes` of user data
int stage_in(ppm_ctx* c, const void* user, size_t bytes, DBuf& scratch, const void** dev) {
  if (is_device_ptr(user)) { *dev = user; return PPM_OK; }
  RC(ens(c, scratch, bytes));
  CK(c, cudaMemcpyAsync(scratch.p, user, bytes, cudaMemcpyHostToDevice, c->stream));
  *dev = scratch.p;
  return PPM_OK;
}
// output: returns a device pointer to write; finish_out copies back if user is host memory
int stage_out(ppm_ctx* c, void* user, size_t bytes, DBuf& scratch, void** dev) {
  if (!user) { *dev = nullptr; return PPM_OK; }
  if (is_device_ptr(user)) { *dev = user; return PPM_OK; }
  RC(ens(c, scratch, bytes));
  *dev = scratch.p;
  return PPM_OK;
}
int finish_out(ppm_ctx* c, void* user, size_t bytes, void* dev) {
  if (!user || user == dev) return PPM_OK;
  CK(c, cudaMemcpyAsync(user, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
  return PPM_OK;
}
inline unsigned nblk(int64_t n, int b) { return (unsigned)std::max<int64_t>(1, (n + b - 1) / b); }

// (Sleeping on a blocking-sync event instead of cudaStreamSynchronize was measured for the end-of-batch and read-back
// waits: e2e -2 % on one GPU, no gain with 8 ranks x 4 host threads on 32 cores -- profiles/r2b_bench_n8.json.)
cudaError_t wait_stream(ppm_ctx*, cudaStream_t st) { return cudaStreamSynchronize(st); }

// host mirror <-> device pass state (probe entry points; a rendered pass never does this)
int push_ps(ppm_ctx* c) {
  CK(c, cudaMemcpyAsync(c->ps, &c->hps, sizeof(PassDev), cudaMemcpyHostToDevice, c->stream));
  return PPM_OK;
}
int pull_ps(ppm_ctx* c) {
  CK(c, cudaMemcpyAsync(&c->hps, c->ps, sizeof(PassDev), cudaMemcpyDeviceToHost, c->stream));
  CK(c, wait_stream(c, c->stream));
  return PPM_OK;
}

RecBuf recbuf(ppm_ctx* c) {
  RecBuf r;
  r.pos3 = c->r_pos.as<double>(); r.dir3 = c->r_dir.as<double>(); r.wl = c->r_wl.as<uint8_t>(); r.tag = c->r_tag.as<uint64_t>();
  return r;
}
MapSoA mapsoa(ppm_ctx* c) {
  MapSoA m;
  m.P = c->m_P.as<double2>(); m.D = c->m_D.as<double2>(); m.orig = c->m_orig.as<uint32_t>();
  return m;
}
CellIndex cellindex(ppm_ctx* c) {
  CellIndex ix;
  ix.words = c->words_p.as<IdxWord>(); ix.start = c->start_p.as<uint32_t>();
  return ix;
}
EyeNodes eyenodes(ppm_ctx* c) {
  EyeNodes n = {c->e_pos.as<double>(), c->e_nrm.as<double>(), c->e_w.as<double>(), c->e_prev.as<uint32_t>()};
  return n;
}
unsigned scan_tiles(uint64_t max_n) { return (unsigned)std::max<uint64_t>(1, (max_n + SCAN_TILE - 1) / SCAN_TILE); }
unsigned rs_blocks(ppm_ctx* c, uint64_t cap) { return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)c->sm_count * 2, (cap + 1023) / 1024)); }
int rs_passes(uint64_t cap) {                      // keys are compact cell ranks < number of records <= cap
  int bits = 1;
  while (bits < 32 && (1ull << bits) < cap) ++bits;
  return (bits + RS_BITS - 1) / RS_BITS;
}

// record buffers (+ per-photon depth masks of a traced set)
int ensure_records(ppm_ctx* c, uint64_t cap, uint64_t nphoton) {
  if (cap == 0) cap = 1;
  RC(ens(c, c->r_pos, cap * 24)); RC(ens(c, c->r_dir, cap * 24));
  RC(ens(c, c->r_wl, cap)); RC(ens(c, c->r_tag, cap * 8));
  if (nphoton) { RC(ens(c, c->pmask, nphoton * 4)); RC(ens(c, c->pbase, nphoton * 4)); }
  return PPM_OK;
}
// everything the map build of up to `cap` records (traced from `nphoton` photons) touches
int ensure_build(ppm_ctx* c, uint64_t cap, uint64_t nphoton) {
  if (cap == 0) cap = 1;
  if (cap >= 0xFFFFFFF0ull) return fail(c, PPM_ERR_CAPACITY, "at most 2^32-16 photon records");
  RC(ens(c, c->order0, cap * 4)); RC(ens(c, c->cid, cap * 4));
  RC(ens(c, c->words_p, (size_t)PPM_CELL_CAP_WORDS * sizeof(IdxWord)));
  RC(ens(c, c->start_p, (cap + 2) * 4));
  RC(ens(c, c->key_a, cap * 4)); RC(ens(c, c->key_b, cap * 4)); RC(ens(c, c->val_a, cap * 4)); RC(ens(c, c->val_b, cap * 4));
  const size_t table = (size_t)RS_DIGITS * rs_blocks(c, cap);
  RC(ens(c, c->rs_table, table * 4));
  const size_t tiles = std::max<size_t>(std::max<size_t>(scan_tiles(table), scan_tiles(PPM_CELL_CAP_WORDS)), scan_tiles(nphoton));
  RC(ens(c, c->tsum_p, tiles * 4));
  RC(ens(c, c->m_P, cap * 32)); RC(ens(c, c->m_D, cap * 32)); RC(ens(c, c->m_orig, cap * 4));
  return PPM_OK;
}
// everything the query sort + gather of up to `cap` queries touches
int ensure_query(ppm_ctx* c, uint64_t cap) {
  if (cap == 0) cap = 1;
  if (cap >= 0xFFFFFFF0ull) return fail(c, PPM_ERR_CAPACITY, "at most 2^32-16 gather queries per call");
  RC(ens(c, c->qcell, cap * 4)); RC(ens(c, c->words_q, (size_t)PPM_CELL_CAP_WORDS * sizeof(IdxWord)));
  RC(ens(c, c->qcnt, (cap + 2) * 4)); RC(ens(c, c->qstart, (cap + 2) * 4)); RC(ens(c, c->qrank, cap * 4)); RC(ens(c, c->qpic, cap * 4));
  RC(ens(c, c->skey, cap * 4)); RC(ens(c, c->sidx, cap * 4));
  RC(ens(c, c->tsum_q, (size_t)std::max(scan_tiles(cap + 1), scan_tiles(PPM_CELL_CAP_WORDS)) * 4));
  if (c->opt_gather_heavy) {
    const uint32_t hcap = 1u << 15;                    // parts: 32 MB of partial sums per lane
    RC(ens(c, c->heavy, (size_t)(hcap / 2) * sizeof(HeavyGroup) + (size_t)hcap * sizeof(HeavyPart) + (size_t)hcap * sizeof(HeavyPartial)));
  }
  return PPM_OK;
}
int ensure_eye(ppm_ctx* c, uint64_t npix, uint64_t cap) {
  if (cap >= 0xFFFFFFF0ull) return fail(c, PPM_ERR_CAPACITY, "more than 2^32-16 gather nodes");
  if (cap == 0) cap = 1;
  RC(ens(c, c->e_head, npix * 4)); RC(ens(c, c->e_emit, npix * 24));
  RC(ens(c, c->e_pos, cap * 24)); RC(ens(c, c->e_nrm, cap * 24)); RC(ens(c, c->e_w, cap * 24)); RC(ens(c, c->e_prev, cap * 4));
  RC(ens(c, c->e_direct, cap * 24)); RC(ens(c, c->e_photon, cap * 24));
  if (c->scene.nlights > 0) RC(ens(c, c->dl_masks, cap * (size_t)c->scene.nlights * 8));
  return PPM_OK;
}
HeavyList heavylist(ppm_ctx* c) {
  HeavyList hl;
  std::memset(&hl, 0, sizeof hl);
  if (!c->opt_gather_heavy || !c->heavy.p) return hl;
  const uint32_t hcap = 1u << 15;
  char* base = c->heavy.as<char>();
  hl.ctr = c->ps->heavy;                               // device address of PassDev::heavy
  hl.groups = (HeavyGroup*)base;
  hl.parts = (HeavyPart*)(base + (size_t)(hcap / 2) * sizeof(HeavyGroup));
  hl.partials = (HeavyPartial*)(base + (size_t)(hcap / 2) * sizeof(HeavyGroup) + (size_t)hcap * sizeof(HeavyPart));
  hl.cap_parts = hcap;
  return hl;
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

int upload_cull(ppm_ctx* c) {
  static_assert(sizeof(DevCull) < (1 << 16), "cull table");
  DevCull cu;
  build_cull(c->scene, cu);
  RC(ens(c, c->cull, sizeof cu));
  CK(c, cudaMemcpy(c->cull.p, &cu, sizeof cu, cudaMemcpyHostToDevice));
  return PPM_OK;
}
// culling is possible when switched on ("dl_cull") and bit 63 of the masks is free for the certificate
// (BVH mode: bit 62 says "walk the hierarchy"; the constant list then holds only the scene's planes)
bool cull_on(const ppm_ctx* c) { return c->opt_dl_cull && c->scene.nprims <= (c->scene.bvh_on ? 62 : 63) && c->scene.nlights > 0; }

// ---- enqueue helpers: launches only, sizes come from PassDev, nothing waits for the host ------------------------
template <class Op>
int enq_scan(ppm_ctx* c, cudaStream_t st, Op op, uint64_t max_n, uint32_t* tile_sums) {
  const unsigned tiles = scan_tiles(max_n);
  k_scan_tiles<Op><<<tiles, SCAN_THREADS, 0, st>>>(op, tile_sums);
  KCHECK(c);
  k_scan_sums<Op><<<1, 1024, 0, st>>>(op, tile_sums);
  KCHECK(c);
  k_scan_apply<Op><<<tiles, SCAN_THREADS, 0, st>>>(op, tile_sums);
  KCHECK(c);
  return PPM_OK;
}
int enq_stamp(ppm_ctx* c, cudaStream_t st, int slot) {
  k_stamp<<<1, 1, 0, st>>>(c->ps, slot);
  KCHECK(c);
  return PPM_OK;
}

// photon tracing of `total` photons; seed / pass / counters are in the pass state
int enq_trace(ppm_ctx* c, cudaStream_t st, const LightSplit& ls, int64_t total, int uc, uint64_t cap) {
  if (total <= 0) return PPM_OK;
  // persistent grid: enough CTAs to fill every SM (resident CTAs are limited by registers), never more threads than photons
  const unsigned blocks = (unsigned)std::min<int64_t>((total + 127) / 128, (int64_t)c->sm_count * 8);
  if (c->scene.bvh_on) k_trace_photons<true><<<blocks, 128, 0, st>>>(c->scene, ls, c->ps, uc, total, recbuf(c), (unsigned long long)cap, c->pmask.as<uint32_t>());
  else k_trace_photons<false><<<blocks, 128, 0, st>>>(c->scene, ls, c->ps, uc, total, recbuf(c), (unsigned long long)cap, c->pmask.as<uint32_t>());
  KCHECK(c);
  return PPM_OK;
}
// photon map of the records in the record buffers: compact cell index, tag order, stable radix sort by cell rank, SoA map.
// traced: the records come from k_trace_photons (`nphoton` emitted photons; n_map is validated on the device);
// otherwise they are an imported set in tag order and PassDev::n_map is already set.
int enq_build(ppm_ctx* c, cudaStream_t st, bool traced, uint64_t nphoton, uint64_t cap) {
  PassDev* ps = c->ps;
  IdxWord* words = c->words_p.as<IdxWord>();
  uint32_t* tsum = c->tsum_p.as<uint32_t>();
  const unsigned wide = (unsigned)c->sm_count * 8;
  k_index_clear<<<(unsigned)c->sm_count * 4, 256, 0, st>>>(ps, words);
  KCHECK(c);
  if (traced) {
    ScanMasks sm = {ps, c->pmask.as<uint32_t>(), c->pbase.as<uint32_t>(), (uint32_t)nphoton, (uint32_t)std::min<uint64_t>(cap, 0xFFFFFFF0ull)};
    RC(enq_scan(c, st, sm, nphoton, tsum));
    k_photon_place<false><<<wide, 256, 0, st>>>(ps, recbuf(c), c->pmask.as<uint32_t>(), c->pbase.as<uint32_t>(), c->order0.as<uint32_t>(),
                                                c->cid.as<uint32_t>(), words, -1);
  } else {
    k_photon_place<true><<<wide, 256, 0, st>>>(ps, recbuf(c), nullptr, nullptr, c->order0.as<uint32_t>(), c->cid.as<uint32_t>(), words, -1);
  }
  KCHECK(c);
  ScanWords sw = {ps, words, &ps->n_occ_p, nullptr};
  RC(enq_scan(c, st, sw, PPM_CELL_CAP_WORDS, tsum));
  const unsigned rsb = rs_blocks(c, cap);
  const int npass = rs_passes(cap);
  uint32_t *kin = c->key_a.as<uint32_t>(), *vin = c->val_a.as<uint32_t>(), *kout = c->key_b.as<uint32_t>(), *vout = c->val_b.as<uint32_t>();
  uint32_t* table = c->rs_table.as<uint32_t>();
  for (int p = 0; p < npass; ++p) {
    const int shift = p * RS_BITS;
    if (p == 0) k_rs_hist<true><<<rsb, RS_THREADS, 0, st>>>(ps, nullptr, shift, table, c->order0.as<uint32_t>(), c->cid.as<uint32_t>(), words, kin, vin);
    else k_rs_hist<false><<<rsb, RS_THREADS, 0, st>>>(ps, kin, shift, table, nullptr, nullptr, nullptr, nullptr, nullptr);
    KCHECK(c);
    ScanFixed sf = {table, (uint32_t)(RS_DIGITS * rsb)};
    RC(enq_scan(c, st, sf, (uint64_t)RS_DIGITS * rsb, tsum));
    k_rs_scatter<<<rsb, RS_THREADS, 0, st>>>(ps, kin, vin, shift, table, kout, vout);
    KCHECK(c);
    std::swap(kin, kout); std::swap(vin, vout);
  }
  k_map_scatter<<<wide, 256, 0, st>>>(ps, recbuf(c), kin, vin, mapsoa(c), c->start_p.as<uint32_t>());
  KCHECK(c);
  return PPM_OK;
}
// counting sort of the gather queries by grid cell (needs the grid of the pass, not the map)
int enq_query_sort(ppm_ctx* c, cudaStream_t st, const double* qpos, uint64_t cap) {
  PassDev* ps = c->ps;
  IdxWord* words = c->words_q.as<IdxWord>();
  const unsigned wide = (unsigned)c->sm_count * 8;     // (one trip per thread, 5 k CTAs: k_query_mark 97 -> 153 us, more lanes on the same hot index words)
  k_index_clear<<<(unsigned)c->sm_count * 4, 256, 0, st>>>(ps, words);
  KCHECK(c);
  k_query_mark<<<wide, 256, 0, st>>>(ps, qpos, (uint32_t)cap, c->qcell.as<uint32_t>(), words, -1);
  KCHECK(c);
  ScanWords sw = {ps, words, &ps->n_occ_q, c->qcnt.as<uint32_t>()};
  RC(enq_scan(c, st, sw, PPM_CELL_CAP_WORDS, c->tsum_q.as<uint32_t>()));
  k_query_count<<<wide, 256, 0, st>>>(ps, c->qcell.as<uint32_t>(), words, c->qcnt.as<uint32_t>(), c->qrank.as<uint32_t>(), c->qpic.as<uint32_t>());
  KCHECK(c);
  ScanU32 su = {&ps->n_occ_q, c->qcnt.as<uint32_t>(), c->qstart.as<uint32_t>()};
  RC(enq_scan(c, st, su, cap + 1, c->tsum_q.as<uint32_t>()));
  k_query_scatter<<<wide, 256, 0, st>>>(ps, c->qcell.as<uint32_t>(), c->qrank.as<uint32_t>(), c->qpic.as<uint32_t>(), c->qstart.as<uint32_t>(),
                                       c->skey.as<uint32_t>(), c->sidx.as<uint32_t>());
  KCHECK(c);
  return PPM_OK;
}
// warp-cooperative gather over the sorted queries; mode 0 fixed radius, 1 per-query radius, 2 count only.
// zero_heavy: clear the heavy-part counters first (probe entry points; k_pass_begin does it for a rendered pass)
int enq_gather(ppm_ctx* c, cudaStream_t st, const double* dpos, const double* dnrm, uint64_t cap, int filter, int mode, const double* r2q,
               double* drgb, uint32_t* dcounts, bool zero_heavy, int stamp_slot) {
  const int B = GATHER_WARPS * 32;
  HeavyList hl = heavylist(c);
  if (zero_heavy) CK(c, cudaMemsetAsync(c->ps->heavy, 0, 16, st));
  const unsigned grid = nblk((int64_t)cap, B);
  const CellIndex ix = cellindex(c);
  const MapSoA m = mapsoa(c);
  const uint32_t* qk = c->skey.as<uint32_t>();
  const uint32_t* qx = c->sidx.as<uint32_t>();
#define GATHER_LAUNCH(F, M) k_gather<F, M><<<grid, B, 0, st>>>(c->ps, ix, m, qk, qx, dpos, dnrm, r2q, drgb, dcounts, hl, stamp_slot)
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
  if (hl.ctr) {
    // always launched (fixed sequence): without published parts every warp returns at once
    const unsigned hgrid = (unsigned)c->sm_count * 16u;
#define HEAVY_LAUNCH(F, M) k_gather_heavy<F, M><<<hgrid, B, 0, st>>>(c->ps, ix, m, qx, dpos, dnrm, r2q, drgb, dcounts, hl)
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
  }
  return PPM_OK;
}
int enq_dl_classify(ppm_ctx* c, cudaStream_t st, const double* dpos, const double* dnrm, uint64_t cap, unsigned long long* masks) {
  if (c->scene.bvh_on) k_dl_classify<true><<<nblk((int64_t)cap, 128), 128, 0, st>>>(c->scene, c->cull.as<DevCull>(), c->ps, (uint32_t)cap, dpos, dnrm, masks, -1);
  else k_dl_classify<false><<<nblk((int64_t)cap, 128), 128, 0, st>>>(c->scene, c->cull.as<DevCull>(), c->ps, (uint32_t)cap, dpos, dnrm, masks, -1);
  KCHECK(c);
  return PPM_OK;
}
int enq_direct_light(ppm_ctx* c, cudaStream_t st, const double* dpos, const double* dnrm, uint64_t cap, double* dout,
                     const unsigned long long* masks, const uint32_t* order, unsigned long long* dbg, int stamp_slot) {
  if (c->scene.bvh_on) k_direct_light<true><<<nblk((int64_t)cap, 128), 128, 0, st>>>(c->scene, c->ps, (uint32_t)cap, masks, order, dpos, dnrm, dout, dbg, stamp_slot);
  else k_direct_light<false><<<nblk((int64_t)cap, 128), 128, 0, st>>>(c->scene, c->ps, (uint32_t)cap, masks, order, dpos, dnrm, dout, dbg, stamp_slot);
  KCHECK(c);
  return PPM_OK;
}
int dl_stats_begin(ppm_ctx* c, cudaStream_t st, unsigned long long** dbg) {
  *dbg = nullptr;
  if (!c->opt_dl_stats) return PPM_OK;
  RC(ens(c, c->dl_dbg, 160));
  CK(c, cudaMemsetAsync(c->dl_dbg.p, 0, 160, st));
  *dbg = c->dl_dbg.as<unsigned long long>();
  return PPM_OK;
}
int dl_stats_print(ppm_ctx* c, cudaStream_t st, unsigned long long* dbg) {
  if (!dbg) return PPM_OK;
  unsigned long long h[20];
  CK(c, cudaMemcpyAsync(h, dbg, 160, cudaMemcpyDeviceToHost, st));
  CK(c, cudaStreamSynchronize(st));
  if (h[0]) {
    std::fprintf(stderr, "[ppm direct light] nodes=%llu tested prims/node: own %.3f, warp union %.3f; certificate %.1f%%\n", h[0],
                 (double)h[1] / h[0], (double)h[2] / h[0], 100.0 * h[3] / h[0]);
    std::fprintf(stderr, "[ppm direct light] nodes by tested prims 0..7+: own");
    for (int k = 0; k < 8; ++k) std::fprintf(stderr, " %.1f%%", 100.0 * h[4 + k] / h[0]);
    std::fprintf(stderr, " | warp union");
    for (int k = 0; k < 8; ++k) std::fprintf(stderr, " %.1f%%", 100.0 * h[12 + k] / h[0]);
    std::fprintf(stderr, "\n");
  }
  return PPM_OK;
}

// ---- probe building blocks (host round trips allowed) ------------------------------------------------------------
int do_trace_photons(ppm_ctx* c, uint64_t seed, uint32_t pass, int uc, const int64_t* n_per_light, double power) {
  LightSplit ls;
  int64_t total = 0;
  RC(light_split(c, n_per_light, &ls, &total));
  if ((uint64_t)total >= (1ull << 32) - 16) return fail(c, PPM_ERR_CAPACITY, "at most 2^32-16 photons per pass");
  uint64_t cap = (uint64_t)total * PPM_MAX_TRACE;
  if (cap == 0) cap = 1;
  RC(ensure_records(c, cap, (uint64_t)std::max<int64_t>(total, 1)));
  c->hps.seed = seed; c->hps.pass = pass; c->hps.power = power;
  c->hps.n_rec = 0ull; c->hps.ticket = 0ull; c->hps.status = 0u;
  RC(push_ps(c));
  RC(enq_trace(c, c->stream, ls, total, uc, cap));
  RC(pull_ps(c));
  if (c->hps.n_rec > cap) return fail(c, PPM_ERR_CAPACITY, "photon record capacity exceeded");
  c->n_rec = c->hps.n_rec; c->power = power; c->have_map = false;
  c->rec_traced = true; c->rec_nphoton = (uint64_t)total;
  return PPM_OK;
}

// Region covered by the grid, from the records in the record buffers.  The reference leaks a few photons (~1e-4)
// through wall corners (a bounce closer than NEARLY0 to the next wall skips it), and those land tens of metres outside
// the room on the infinite planes, so the raw bounding box is erratic and mostly empty.  The region is therefore a
// per-axis TRIMMED range (at most n/1024 photons cut on each side, found with device histograms); photons and queries
// outside are clamped into the boundary cells, which keeps the 27-cell walk exact (see make_grid).
int compute_bounds(ppm_ctx* c, uint64_t n, double cell, Bounds* out, int* have) {
  std::memset(out, 0, sizeof *out);
  *have = 0;
  if (n == 0) return PPM_OK;
  RC(ens(c, c->bbox, 48));
  unsigned long long init[6] = {~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull};
  CK(c, cudaMemcpyAsync(c->bbox.p, init, 48, cudaMemcpyHostToDevice, c->stream));
  k_bbox<<<std::min<unsigned>(nblk((int64_t)n, 256), (unsigned)c->sm_count * 8), 256, 0, c->stream>>>(c->r_pos.as<double>(), n,
                                                                                                  c->bbox.as<unsigned long long>());
  KCHECK(c);
  unsigned long long mm[6];
  CK(c, cudaMemcpyAsync(mm, c->bbox.p, 48, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  double lo[3], hi[3];
  for (int k = 0; k < 3; ++k) { lo[k] = dec_ord(mm[k]); hi[k] = dec_ord(mm[3 + k]); }
  for (int k = 0; k < 3; ++k)
    if (!(lo[k] == lo[k]) || !(hi[k] == hi[k]) || std::isinf(lo[k]) || std::isinf(hi[k]))
      return fail(c, PPM_ERR_ARG, "photon positions are not finite");
  const uint64_t trim = n >= 4096 ? n / 1024 : 0;
  if (trim > 0) {
    RC(ens(c, c->axis_hist, 3 * AXIS_BINS * 4));
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
      k_axis_hist<<<std::min<unsigned>(nblk((int64_t)n, 256 * 8), (unsigned)c->sm_count * 4), 256, 0, c->stream>>>(c->r_pos.as<double>(), n, ar,
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
  }
  for (int k = 0; k < 3; ++k) { out->lo[k] = lo[k]; out->hi[k] = hi[k]; }
  *have = 1;
  return PPM_OK;
}

// probe: map of the current photon set (records in the record buffers) at squared radius radius2.
// fresh_bounds: take the grid region from these records (ppm_map_build); otherwise keep the calibrated one.
int do_map_build(ppm_ctx* c, double radius2, bool fresh_bounds) {
  if (!(radius2 > 0.0)) return fail(c, PPM_ERR_ARG, "radius2 must be > 0");
  const uint64_t n = c->n_rec;
  if (fresh_bounds) {
    Bounds b; int have = 0;
    RC(compute_bounds(c, n, std::sqrt(radius2), &b, &have));
    c->hps.bounds = b; c->hps.have_bounds = have;
  }
  c->hps.r2 = radius2; c->hps.power = c->power;
  c->hps.grid = make_grid(c->hps.bounds, c->hps.have_bounds, radius2);
  c->hps.inv_pi_r2 = (1.0 / PPM_PI) / radius2;
  c->hps.n_rec = n; c->hps.n_map = c->rec_traced ? 0u : (uint32_t)n;
  c->hps.n_occ_p = 0u; c->hps.status = 0u;
  RC(ensure_build(c, n, c->rec_traced ? c->rec_nphoton : 0));
  RC(push_ps(c));
  RC(enq_build(c, c->stream, c->rec_traced, c->rec_nphoton, std::max<uint64_t>(n, 1)));
  RC(pull_ps(c));
  if (c->hps.status) return fail(c, PPM_ERR_STATE, "photon records are inconsistent (record count does not match the depth masks)");
  c->have_map = true;
  return PPM_OK;
}

// probe: sort n queries at dpos by cell (the device-side query count is set from the host)
int probe_sort_queries(ppm_ctx* c, const double* dpos, int64_t n) {
  RC(ensure_query(c, (uint64_t)n));
  c->hps.n_nodes = (unsigned long long)n; c->hps.n_query = 0u; c->hps.n_occ_q = 0u;
  c->hps.sum_k = 0ull; c->hps.cand = 0ull; c->hps.status = 0u;
  std::memset(c->hps.sum_k_s, 0, sizeof c->hps.sum_k_s); std::memset(c->hps.cand_s, 0, sizeof c->hps.cand_s);
  c->hps.heavy[0] = c->hps.heavy[1] = c->hps.heavy[2] = c->hps.heavy[3] = 0u;
  RC(push_ps(c));
  return enq_query_sort(c, c->stream, dpos, (uint64_t)n);
}
int launch_gather(ppm_ctx* c, const double* dpos, const double* dnrm, int64_t n, int filter, double* drgb, uint32_t* dcounts) {
  if (n <= 0) return PPM_OK;
  if (filter < PPM_FILTER_NONE || filter > PPM_FILTER_GAUSS) return fail(c, PPM_ERR_ARG, "bad filter");
  RC(probe_sort_queries(c, dpos, n));
  return enq_gather(c, c->stream, dpos, dnrm, (uint64_t)n, filter, 0, nullptr, drgb, dcounts, true, -1);
}
// k-NN estimate (no reference implementation exists: n_sample_photon is dead code, photonmap.rs:18,
// camera.rs:181; semantics defined in SURVEY.md 8c): for every query the k nearest photons within r;
// if k are found, r_k^2 = the k-th smallest d2 replaces r^2 in the membership test, the filter and the
// normaliser; otherwise the fixed radius is used.  r_k^2 is found EXACTLY by bisection on the bit
// pattern of d2 with the count-only gather (<= 63 steps).
int launch_gather_knn(ppm_ctx* c, const double* dpos, const double* dnrm, int64_t n, uint32_t k, int filter, double* drgb,
                      double* dr2k, uint32_t* dcounts) {
  if (n <= 0) return PPM_OK;
  RC(probe_sort_queries(c, dpos, n));
  RC(ens(c, c->knn_lo, (size_t)n * 8)); RC(ens(c, c->knn_hi, (size_t)n * 8)); RC(ens(c, c->knn_thr, (size_t)n * 8));
  RC(ens(c, c->knn_cnt, (size_t)n * 4)); RC(ens(c, c->knn_act, 64));
  unsigned long long* lo = c->knn_lo.as<unsigned long long>();
  unsigned long long* hi = c->knn_hi.as<unsigned long long>();
  double* thr = c->knn_thr.as<double>();
  uint32_t* cnt = c->knn_cnt.as<uint32_t>();
  unsigned int* nact = c->knn_act.as<unsigned int>();
  const double r2 = c->hps.r2;
  // photons within the fixed radius
  RC(enq_gather(c, c->stream, dpos, dnrm, (uint64_t)n, PPM_FILTER_NONE, 0, nullptr, drgb, cnt, true, -1));
  CK(c, cudaMemsetAsync(nact, 0, 4, c->stream));
  k_knn_init<<<nblk(n, 256), 256, 0, c->stream>>>(n, r2, cnt, k, lo, hi, thr, nact);
  KCHECK(c);
  for (int it = 0; it < 70; ++it) {
    unsigned int active = 0;
    CK(c, cudaMemcpyAsync(&active, nact, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    if (!active) break;
    RC(enq_gather(c, c->stream, dpos, dnrm, (uint64_t)n, PPM_FILTER_NONE, 2, thr, nullptr, cnt, true, -1));
    CK(c, cudaMemsetAsync(nact, 0, 4, c->stream));
    k_knn_step<<<nblk(n, 256), 256, 0, c->stream>>>(n, cnt, k, lo, hi, thr, nact);
    KCHECK(c);
  }
  k_knn_finish<<<nblk(n, 256), 256, 0, c->stream>>>(n, r2, thr);
  KCHECK(c);
  RC(enq_gather(c, c->stream, dpos, dnrm, (uint64_t)n, filter, 1, thr, drgb, dcounts, true, -1));
  if (dr2k) CK(c, cudaMemcpyAsync(dr2k, thr, (size_t)n * 8, cudaMemcpyDeviceToDevice, c->stream));
  return PPM_OK;
}

// probe: expand the eye paths of n rays (drays == NULL: camera rays) into the gather-node pool, growing it as needed
int probe_eye_expand(ppm_ctx* c, const double* drays, int64_t n, int64_t first_pixel, uint64_t seed, uint32_t pass, int classic, uint32_t* nn_out) {
  if (c->eye_cap < (uint64_t)n * 2) c->eye_cap = (uint64_t)n * 2;      // first guess: two gather nodes per pixel
  for (int attempt = 0;; ++attempt) {
    RC(ensure_eye(c, (uint64_t)n, c->eye_cap));
    c->hps.seed = seed; c->hps.pass = pass; c->hps.n_nodes = 0ull; c->hps.n_visited = 0ull; c->hps.status = 0u;
    RC(push_ps(c));
    if (c->scene.bvh_on) k_eye_expand<true><<<nblk(n, 128), 128, 0, c->stream>>>(c->scene, c->cam, drays, n, first_pixel, c->ps, eyenodes(c), (uint32_t)c->eye_cap,
                                                                                c->e_head.as<uint32_t>(), c->e_emit.as<double>(), classic, -1);
    else k_eye_expand<false><<<nblk(n, 128), 128, 0, c->stream>>>(c->scene, c->cam, drays, n, first_pixel, c->ps, eyenodes(c), (uint32_t)c->eye_cap,
                                                                  c->e_head.as<uint32_t>(), c->e_emit.as<double>(), classic, -1);
    KCHECK(c);
    RC(pull_ps(c));
    if (c->hps.n_nodes <= c->eye_cap) break;
    if (attempt >= 2) return fail(c, PPM_ERR_CAPACITY, "gather-node pool keeps overflowing");
    c->eye_cap = c->hps.n_nodes + c->hps.n_nodes / 4;                 // pool overflow: grow and walk again
  }
  *nn_out = (uint32_t)c->hps.n_nodes;
  return PPM_OK;
}
// probe: classic direct light at nn points in the given order (NULL = as they are)
int probe_direct_light(ppm_ctx* c, const double* dpos, const double* dnrm, int64_t nn, double* dout) {
  if (nn <= 0) return PPM_OK;
  const unsigned long long* masks = nullptr;
  c->hps.n_nodes = (unsigned long long)nn;
  RC(push_ps(c));
  if (cull_on(c)) {
    RC(ens(c, c->dl_masks, (size_t)nn * (size_t)c->scene.nlights * 8));
    RC(enq_dl_classify(c, c->stream, dpos, dnrm, (uint64_t)nn, c->dl_masks.as<unsigned long long>()));
    masks = c->dl_masks.as<unsigned long long>();
  }
  unsigned long long* dbg = nullptr;
  RC(dl_stats_begin(c, c->stream, &dbg));
  RC(enq_direct_light(c, c->stream, dpos, dnrm, (uint64_t)nn, dout, masks, nullptr, dbg, -1));
  return dl_stats_print(c, c->stream, dbg);
}
int enq_combine(ppm_ctx* c, int64_t n, int64_t first_pixel, bool with_direct, bool classic, double* dout, double* daccum, int stamp_slot) {
  const D3 amb = {c->cam.ambient[0], c->cam.ambient[1], c->cam.ambient[2]};
  k_combine<<<nblk(n, 256), 256, 0, c->stream>>>(c->ps, c->e_head.as<uint32_t>(), c->e_prev.as<uint32_t>(), c->e_w.as<double>(),
                                                (with_direct || classic) ? c->e_direct.as<double>() : nullptr,
                                                classic ? nullptr : c->e_photon.as<double>(), c->e_emit.as<double>(), n, dout, daccum,
                                                first_pixel, amb, stamp_slot);
  KCHECK(c);
  return PPM_OK;
}
// serial eye path on the main stream (ppm_trace_rays / ppm_trace_rays_classic)
int do_trace_rays(ppm_ctx* c, const double* drays, int64_t n, int64_t first_pixel, uint64_t seed, uint32_t pass, int uc, double* dout, int classic) {
  if (n <= 0) return PPM_OK;
  uint32_t nn = 0;
  RC(probe_eye_expand(c, drays, n, first_pixel, seed, pass, classic, &nn));
  const bool direct = classic || uc;
  if (direct && nn) RC(probe_direct_light(c, c->e_pos.as<double>(), c->e_nrm.as<double>(), nn, c->e_direct.as<double>()));
  if (!classic && nn) {
    if (c->cam.pfilter < PPM_FILTER_NONE || c->cam.pfilter > PPM_FILTER_GAUSS) return fail(c, PPM_ERR_ARG, "bad filter");
    RC(launch_gather(c, c->e_pos.as<double>(), c->e_nrm.as<double>(), nn, c->cam.pfilter, c->e_photon.as<double>(), nullptr));
  }
  c->hps.status = 0u;
  RC(push_ps(c));
  return enq_combine(c, n, first_pixel, uc != 0, classic != 0, dout, nullptr, -1);
}

// ---- a rendered pass ------------------------------------------------------------------------------------------------
int ensure_accum(ppm_ctx* c) {
  uint64_t npix = (uint64_t)c->cam.xreso * (uint64_t)c->cam.yreso;
  if (c->accum_pixels == npix && c->accum.p) return PPM_OK;
  // sum image and pass counter live in ONE allocation so a single reduce covers both
  RC(ens(c, c->accum, (size_t)(npix * 3 + 1) * 8));
  CK(c, cudaMemsetAsync(c->accum.p, 0, (size_t)(npix * 3 + 1) * 8, c->stream));
  c->accum_pixels = npix;
  return PPM_OK;
}

// Per-scene calibration (once per scene / camera / seed / photon budget): the grid region and the buffer capacities of
// a pass come from a small calibration pass with its own Philox pass id, so they do not depend on which passes a
// context renders -- every rank of a multi-GPU frame derives the same grid and therefore the same summation order.
int calibrate(ppm_ctx* c, uint64_t seed, int64_t nphoton, int uc) {
  CalKey k;
  k.scene_ver = c->scene_ver; k.cam_ver = c->cam_ver; k.seed = seed; k.nphoton = nphoton; k.uc = uc;
  if (c->cal_valid && k == c->cal_key) return PPM_OK;
  c->cal_valid = false;
  double pw = 0.0, pw_cal = 0.0;
  int64_t ns[PPM_MAX_LIGHTS], ns_cal[PPM_MAX_LIGHTS];
  int rc;
  if ((rc = ppm_photon_budget(c->scene.lights, c->scene.nlights, nphoton, &pw, ns))) return fail(c, rc, "photon budget");
  const int64_t ncal = std::min<int64_t>(nphoton, PPM_CAL_PHOTONS);
  if ((rc = ppm_photon_budget(c->scene.lights, c->scene.nlights, ncal, &pw_cal, ns_cal))) return fail(c, rc, "photon budget");
  int64_t total = 0, total_cal = 0;
  for (int i = 0; i < c->scene.nlights; ++i) { total += ns[i]; total_cal += ns_cal[i]; }
  RC(do_trace_photons(c, seed, PPM_CAL_PASS, uc, ns_cal, pw_cal));
  Bounds b; int have = 0;
  RC(compute_bounds(c, c->n_rec, 0.01, &b, &have));
  c->cal_bounds = b; c->cal_have_bounds = have;
  const double per = total_cal > 0 ? (double)c->n_rec / (double)total_cal : 1.0;
  uint64_t rec_cap = (uint64_t)(per * (double)total * 1.25) + 65536;
  rec_cap = std::min<uint64_t>(rec_cap, (uint64_t)std::max<int64_t>(total, 1) * PPM_MAX_TRACE);
  c->rec_cap = std::max<uint64_t>(rec_cap, 1);
  const int64_t npix = (int64_t)c->cam.xreso * c->cam.yreso;
  uint32_t nn = 0;
  RC(probe_eye_expand(c, nullptr, npix, 0, seed, PPM_CAL_PASS, 0, &nn));
  c->pass_eye_cap = (uint64_t)nn + (uint64_t)nn / 12 + 16384;
  c->have_map = false; c->n_rec = 0;
  c->cal_key = k; c->cal_valid = true;
  return PPM_OK;
}
// buffers of a rendered pass at the calibrated capacities
int ensure_pass(ppm_ctx* c, int64_t total) {
  const uint64_t npix = (uint64_t)c->cam.xreso * (uint64_t)c->cam.yreso;
  RC(ensure_records(c, c->rec_cap, (uint64_t)total));
  RC(ensure_build(c, c->rec_cap, (uint64_t)total));
  RC(ensure_eye(c, npix, c->pass_eye_cap));
  RC(ensure_query(c, c->pass_eye_cap));
  RC(ens(c, c->pass_img, (size_t)npix * 24));
  RC(ensure_accum(c));
  return PPM_OK;
}

// Stream priorities are not carried into captured kernel nodes, and without them the few-CTA kernels of the photon
// branch queue behind the 16 k CTAs of the eye-path expansion.  While capturing, the nodes recorded from the main
// stream are collected segment by segment (set difference of the graph's node list) and get the high priority as a
// node attribute afterwards.
std::vector<cudaGraphNode_t> capture_nodes(ppm_ctx* c) {
  std::vector<cudaGraphNode_t> v;
  cudaStreamCaptureStatus stt = cudaStreamCaptureStatusNone;
  cudaGraph_t g = nullptr;
  if (cudaStreamGetCaptureInfo(c->stream, &stt, nullptr, &g, nullptr, nullptr) != cudaSuccess || stt != cudaStreamCaptureStatusActive || !g) {
    cudaGetLastError();
    return v;
  }
  size_t n = 0;
  if (cudaGraphGetNodes(g, nullptr, &n) != cudaSuccess || n == 0) { cudaGetLastError(); return v; }
  v.resize(n);
  if (cudaGraphGetNodes(g, v.data(), &n) != cudaSuccess) { cudaGetLastError(); v.clear(); return v; }
  v.resize(n);
  std::sort(v.begin(), v.end());
  return v;
}
void seg_begin(ppm_ctx* c) { if (c->capturing) c->seg_before = capture_nodes(c); }
void seg_end(ppm_ctx* c) {
  if (!c->capturing) return;
  std::vector<cudaGraphNode_t> now = capture_nodes(c);
  for (cudaGraphNode_t n : now)
    if (!std::binary_search(c->seg_before.begin(), c->seg_before.end(), n)) c->hi_nodes.push_back(n);
}

// The launch sequence of ONE pass: photon branch on `stream`, eye branch on `stream2`, joined before the gather (needs
// map + sorted queries) and before the combine (needs direct light + estimates).  Pure enqueue: runs as is in stream
// mode and is what the graph capture records.
int enq_pass(ppm_ctx* c, const LightSplit& ls, int64_t total, int uc, bool accumulate, unsigned long long* dl_dbg) {
  typedef ppm_ctx E;
  cudaStream_t sa = c->stream, sb = c->stream2;
  const int64_t npix = (int64_t)c->cam.xreso * c->cam.yreso;
  const uint64_t ecap = c->pass_eye_cap;
  const bool cull = uc && cull_on(c);
  seg_begin(c);
  k_pass_begin<<<1, 32, 0, sa>>>(c->ps, c->bt);
  KCHECK(c);
  CK(c, cudaEventRecord(c->ev[E::EV_FORK], sa));
  CK(c, cudaStreamWaitEvent(sb, c->ev[E::EV_FORK], 0));
  // photon branch
  RC(enq_trace(c, sa, ls, total, uc, c->rec_cap));
  RC(enq_stamp(c, sa, ST_TRACE_END));
  RC(enq_build(c, sa, true, (uint64_t)total, c->rec_cap));
  RC(enq_stamp(c, sa, ST_BUILD_END));
  seg_end(c);
  // eye branch
  if (c->scene.bvh_on) k_eye_expand<true><<<nblk(npix, 128), 128, 0, sb>>>(c->scene, c->cam, nullptr, npix, 0, c->ps, eyenodes(c), (uint32_t)ecap, c->e_head.as<uint32_t>(),
                                                                          c->e_emit.as<double>(), 0, ST_EXPAND_BEGIN);
  else k_eye_expand<false><<<nblk(npix, 128), 128, 0, sb>>>(c->scene, c->cam, nullptr, npix, 0, c->ps, eyenodes(c), (uint32_t)ecap, c->e_head.as<uint32_t>(),
                                                            c->e_emit.as<double>(), 0, ST_EXPAND_BEGIN);
  KCHECK(c);
  RC(enq_stamp(c, sb, ST_EXPAND_END));
  if (cull) RC(enq_dl_classify(c, sb, c->e_pos.as<double>(), c->e_nrm.as<double>(), ecap, c->dl_masks.as<unsigned long long>()));
  RC(enq_stamp(c, sb, ST_CLASSIFY_END));
  RC(enq_query_sort(c, sb, c->e_pos.as<double>(), ecap));
  RC(enq_stamp(c, sb, ST_QSORT_END));
  CK(c, cudaEventRecord(c->ev[E::EV_NODES], sb));
  if (uc) {
    // direct light over the nodes in the cell-sorted order of the gather (coherent culling masks), beside k_gather
    RC(enq_direct_light(c, sb, c->e_pos.as<double>(), c->e_nrm.as<double>(), ecap, c->e_direct.as<double>(),
                        cull ? c->dl_masks.as<unsigned long long>() : nullptr, c->sidx.as<uint32_t>(), dl_dbg, ST_DL_BEGIN));
    RC(enq_stamp(c, sb, ST_DL_END));
  }
  CK(c, cudaEventRecord(c->ev[E::EV_DL], sb));
  // gather
  seg_begin(c);
  CK(c, cudaStreamWaitEvent(sa, c->ev[E::EV_NODES], 0));
  if (!c->opt_graph) CK(c, cudaEventRecord(c->ev[E::EV_G0], sa));
  RC(enq_gather(c, sa, c->e_pos.as<double>(), c->e_nrm.as<double>(), ecap, c->cam.pfilter, 0, nullptr, c->e_photon.as<double>(), nullptr, false,
                ST_GATHER_BEGIN));
  RC(enq_stamp(c, sa, ST_GATHER_END));
  if (!c->opt_graph) CK(c, cudaEventRecord(c->ev[E::EV_G1], sa));
  // combine + accumulate
  CK(c, cudaStreamWaitEvent(sa, c->ev[E::EV_DL], 0));
  RC(enq_combine(c, npix, 0, uc != 0, false, c->pass_img.as<double>(), accumulate ? c->accum.as<double>() : nullptr, ST_COMBINE_BEGIN));
  k_pass_end<<<1, 32, 0, sa>>>(c->ps, c->out, accumulate ? c->accum.as<double>() + (size_t)npix * 3 : nullptr);
  KCHECK(c);
  seg_end(c);
  return PPM_OK;
}

int build_graph(ppm_ctx* c, const LightSplit& ls, int64_t total, int uc) {
  GraphKey k;
  k.scene_ver = c->scene_ver; k.cam_ver = c->cam_ver; k.buf_gen = c->buf_gen; k.rec_cap = c->rec_cap; k.eye_cap = c->pass_eye_cap;
  k.nphoton = total; k.uc = uc; k.cull = c->opt_dl_cull; k.heavy = c->opt_gather_heavy; k.accumulate = 1;
  if (c->gexec && k == c->gkey) return PPM_OK;
  if (c->gexec) { cudaGraphExecDestroy(c->gexec); c->gexec = nullptr; }
  CK(c, cudaStreamSynchronize(c->stream));
  CK(c, cudaStreamSynchronize(c->stream2));
  const uint64_t l0 = c->launches;
  CK(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
  c->capturing = true; c->hi_nodes.clear();
  int rc = enq_pass(c, ls, total, uc, true, nullptr);
  c->capturing = false;
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(c->stream, &g);
  if (rc) { if (g) cudaGraphDestroy(g); cudaGetLastError(); return rc; }
  if (e != cudaSuccess) { c->err = std::string("graph capture: ") + cudaGetErrorString(e); cudaGetLastError(); return PPM_ERR_CUDA; }
  {
    // node priorities: main-stream nodes high, everything else low
    size_t n = 0;
    std::vector<cudaGraphNode_t> all;
    if (cudaGraphGetNodes(g, nullptr, &n) == cudaSuccess && n) { all.resize(n); cudaGraphGetNodes(g, all.data(), &n); all.resize(n); }
    std::sort(c->hi_nodes.begin(), c->hi_nodes.end());
    for (cudaGraphNode_t nd : all) {
      cudaGraphNodeType ty;
      if (cudaGraphNodeGetType(nd, &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
      cudaLaunchAttributeValue v;
      std::memset(&v, 0, sizeof v);
      v.priority = std::binary_search(c->hi_nodes.begin(), c->hi_nodes.end(), nd) ? c->prio_hi : c->prio_lo;
      cudaGraphKernelNodeSetAttribute(nd, cudaLaunchAttributePriority, &v);
    }
    cudaGetLastError();
  }
  e = cudaGraphInstantiateWithFlags(&c->gexec, g, cudaGraphInstantiateFlagUseNodePriority);   // else the launch stream's priority applies to every node
  cudaGraphDestroy(g);
  if (e != cudaSuccess) { c->gexec = nullptr; c->err = std::string("graph instantiate: ") + cudaGetErrorString(e); return PPM_ERR_CUDA; }
  c->graph_kernels = c->launches - l0;
  c->launches = l0;
  c->gkey = k;
  return PPM_OK;
}

struct BatchStats {
  double ms[8] = {0};
  uint64_t ct[8] = {0};
};
void add_pass_stats(BatchStats& s, const PassOut& o, int64_t emitted, uint64_t launches) {
  auto d = [&](int a, int b) { return (o.stamp[a] && o.stamp[b] && o.stamp[b] >= o.stamp[a]) ? (double)(o.stamp[b] - o.stamp[a]) * 1e-6 : 0.0; };
  s.ms[0] += d(ST_BEGIN, ST_TRACE_END);
  s.ms[1] += d(ST_TRACE_END, ST_BUILD_END);
  s.ms[2] += d(ST_EXPAND_BEGIN, ST_EXPAND_END);
  s.ms[3] += d(ST_DL_BEGIN, ST_DL_END);
  s.ms[4] += d(ST_CLASSIFY_END, ST_QSORT_END) + d(ST_GATHER_BEGIN, ST_GATHER_END);
  s.ms[5] += d(ST_COMBINE_BEGIN, ST_END);
  s.ms[6] += d(ST_BEGIN, ST_END);
  s.ms[7] += d(ST_GATHER_BEGIN, ST_GATHER_END);
  s.ct[0] += (uint64_t)emitted; s.ct[1] += o.n_rec; s.ct[2] += o.n_visited; s.ct[3] += o.n_nodes; s.ct[4] += o.sum_k;
  s.ct[5] += launches; s.ct[6] += o.cand;
}

// the lanes of a context: itself + twins, all with the same scene, camera, switches, calibration
int sync_twin(ppm_ctx* c, ppm_ctx* t) {
  if (t->scene_ver != c->scene_ver || !t->have_scene) {
    t->scene = c->scene; t->scene_ver = c->scene_ver; t->have_scene = true;
    int rc = upload_cull(t);
    if (rc) return fail(c, rc, "lane: " + t->err);
  }
  t->cam = c->cam; t->cam_ver = c->cam_ver; t->have_camera = true;
  t->opt_dl_cull = c->opt_dl_cull; t->opt_gather_heavy = c->opt_gather_heavy; t->opt_dl_stats = c->opt_dl_stats; t->opt_graph = c->opt_graph;
  t->cal_bounds = c->cal_bounds; t->cal_have_bounds = c->cal_have_bounds;
  t->rec_cap = c->rec_cap; t->pass_eye_cap = c->pass_eye_cap;
  return PPM_OK;
}

// Renders the passes idx[0..n) of a batch (pass id = first + idx*stride, radius2[idx]) on `lanes` lanes; returns the
// reports in outs[idx].  One host thread: a pass is one graph launch (or, in stream mode, one enqueue sequence).
int run_passes(ppm_ctx* c, std::vector<ppm_ctx*>& lane, const std::vector<int32_t>& idx, uint64_t seed, uint32_t first_pass, uint32_t stride,
               const double* radius2, double power, const LightSplit& ls, int64_t total, int uc, std::vector<PassOut>& outs) {
  const int lanes = (int)lane.size();
  size_t done = 0;
  while (done < idx.size()) {
    const size_t chunk = std::min(idx.size() - done, (size_t)PPM_BATCH_MAX * (size_t)lanes);
    // batch tables
    std::vector<int> cnt(lanes, 0);
    for (size_t k = 0; k < chunk; ++k) {
      const int j = (int)(k % (size_t)lanes);
      ppm_ctx* x = lane[j];
      const int32_t i = idx[done + k];
      x->bt_h->pass[cnt[j]] = first_pass + (uint32_t)i * stride;
      x->bt_h->r2[cnt[j]] = radius2[i];
      ++cnt[j];
    }
    for (int j = 0; j < lanes; ++j) {
      ppm_ctx* x = lane[j];
      if (!cnt[j]) continue;
      x->bt_h->seed = seed; x->bt_h->power = power;
      CK(c, cudaMemcpyAsync(x->bt, x->bt_h, sizeof(BatchDev), cudaMemcpyHostToDevice, x->stream));
      // bounds + cursor: the grid region of the pass state and the table position
      x->hps.cursor = 0u;
      x->hps.bounds = x->cal_bounds; x->hps.have_bounds = x->cal_have_bounds;
      CK(c, cudaMemcpyAsync(x->ps, &x->hps, sizeof(PassDev), cudaMemcpyHostToDevice, x->stream));
    }
    // the passes, alternating over the lanes
    for (size_t k = 0; k < chunk; ++k) {
      ppm_ctx* x = lane[k % (size_t)lanes];
      if (x->opt_graph && x->gexec) {
        cudaError_t e = cudaGraphLaunch(x->gexec, x->stream);
        if (e != cudaSuccess) return fail(c, PPM_ERR_CUDA, std::string("graph launch: ") + cudaGetErrorString(e));
      } else {
        unsigned long long* dbg = nullptr;
        int rc = dl_stats_begin(x, x->stream2, &dbg);
        if (!rc) rc = enq_pass(x, ls, total, uc, true, dbg);
        if (!rc && dbg) { cudaStreamSynchronize(x->stream); rc = dl_stats_print(x, x->stream2, dbg); }
        if (rc) return x == c ? rc : fail(c, rc, "lane: " + x->err);
        if (!x->opt_graph) {                               // stream mode: CUDA events around k_gather (+ heavy parts) as a cross-check of the stamps
          cudaStreamSynchronize(x->stream);
          float f = 0.f;
          if (cudaEventElapsedTime(&f, x->ev[ppm_ctx::EV_G0], x->ev[ppm_ctx::EV_G1]) == cudaSuccess) x->ms[7] += (double)f; else cudaGetLastError();
        }
      }
    }
    // reports
    for (int j = 0; j < lanes; ++j) {
      ppm_ctx* x = lane[j];
      if (!cnt[j]) continue;
      CK(c, cudaMemcpyAsync(x->out_h, x->out, sizeof(PassOut) * (size_t)cnt[j], cudaMemcpyDeviceToHost, x->stream));
    }
    for (int j = 0; j < lanes; ++j) {
      if (!cnt[j]) continue;
      cudaError_t e = wait_stream(lane[j], lane[j]->stream);
      if (e != cudaSuccess) return fail(c, PPM_ERR_CUDA, std::string("pass batch: ") + cudaGetErrorString(e));
    }
    std::vector<int> pos(lanes, 0);
    for (size_t k = 0; k < chunk; ++k) {
      const int j = (int)(k % (size_t)lanes);
      outs[(size_t)idx[done + k]] = lane[j]->out_h[pos[j]++];
    }
    done += chunk;
  }
  return PPM_OK;
}

int render_batch(ppm_ctx* c, uint64_t seed, uint32_t first_pass, uint32_t stride, int32_t npass, int64_t nphoton, const double* radius2,
                 int uc, int lanes_wanted) {
  if (!c->have_scene || c->scene.nlights == 0) return fail(c, PPM_ERR_STATE, "scene with lights not set");
  if (!c->have_camera) return fail(c, PPM_ERR_STATE, "camera not set");
  if (nphoton <= 0) return fail(c, PPM_ERR_ARG, "nphoton must be positive");
  if (c->cam.pfilter < PPM_FILTER_NONE || c->cam.pfilter > PPM_FILTER_GAUSS) return fail(c, PPM_ERR_ARG, "bad filter");
  for (int32_t i = 0; i < npass; ++i)
    if (!(radius2[i] > 0.0)) return fail(c, PPM_ERR_ARG, "radius2 must be positive");
  CK(c, cudaSetDevice(c->device));
  int rc;
  double power;
  int64_t ns[PPM_MAX_LIGHTS];
  if ((rc = ppm_photon_budget(c->scene.lights, c->scene.nlights, nphoton, &power, ns))) return fail(c, rc, "photon budget");
  LightSplit ls;
  int64_t total = 0;
  RC(light_split(c, ns, &ls, &total));
  if ((uint64_t)total >= (1ull << 32) - 16) return fail(c, PPM_ERR_CAPACITY, "at most 2^32-16 photons per pass");
  RC(calibrate(c, seed, nphoton, uc));
  int lanes = std::max(1, std::min(lanes_wanted, (int)PPM_MAX_LANES));
  lanes = std::min<int>(lanes, npass);
  std::vector<ppm_ctx*> lane(lanes, nullptr);
  lane[0] = c;
  for (int j = 1; j < lanes; ++j) {
    if ((int)c->twins.size() < j) {
      ppm_ctx* t = nullptr;
      rc = ppm_create(c->device, &t);
      if (rc) return fail(c, rc, "cannot create another lane");
      c->twins.push_back(t);
    }
    lane[j] = c->twins[j - 1];
  }
  std::vector<int32_t> todo(npass);
  for (int32_t i = 0; i < npass; ++i) todo[i] = i;
  std::vector<PassOut> outs((size_t)npass);
  BatchStats st;
  uint64_t retried = 0, stream_launches = 0;
  int last_lane = 0;
  for (int attempt = 0; !todo.empty(); ++attempt) {
    for (size_t k = 0; k < todo.size(); ++k)
      if (todo[k] == npass - 1) last_lane = (int)(k % (size_t)lanes);
    uint64_t l0 = 0;
    for (int j = 0; j < lanes; ++j) l0 += lane[j]->launches;
    for (int j = 0; j < lanes; ++j) {
      ppm_ctx* x = lane[j];
      if (j > 0) RC(sync_twin(c, x));
      if ((rc = ensure_pass(x, total))) return x == c ? rc : fail(c, rc, "lane: " + x->err);
      x->ms[7] = 0.0;
      if (x->opt_graph && !x->opt_dl_stats) {
        if ((rc = build_graph(x, ls, total, uc))) return x == c ? rc : fail(c, rc, "lane: " + x->err);
      } else if (x->gexec) { cudaGraphExecDestroy(x->gexec); x->gexec = nullptr; }
    }
    RC(run_passes(c, lane, todo, seed, first_pass, stride, radius2, power, ls, total, uc, outs));
    for (int j = 0; j < lanes; ++j) stream_launches += lane[j]->launches;
    stream_launches -= l0;
    // passes that overflowed a buffer did not count: grow and render them again
    std::vector<int32_t> again;
    uint64_t need_rec = 0, need_nodes = 0;
    for (int32_t i : todo) {
      const PassOut& o = outs[(size_t)i];
      if (o.status) {
        again.push_back(i);
        if (o.status & PPM_ST_REC_OVERFLOW) need_rec = std::max<uint64_t>(need_rec, o.n_rec);
        if (o.status & PPM_ST_NODE_OVERFLOW) need_nodes = std::max<uint64_t>(need_nodes, o.n_nodes);
      } else {
        add_pass_stats(st, o, total, c->opt_graph && !c->opt_dl_stats ? c->graph_kernels : 0);
      }
    }
    if (!again.empty()) {
      if (attempt >= 3) return fail(c, PPM_ERR_CAPACITY, "pass buffers keep overflowing");
      if (need_rec) c->rec_cap = std::min<uint64_t>(need_rec + need_rec / 4 + 65536, (uint64_t)total * PPM_MAX_TRACE);
      if (need_nodes) c->pass_eye_cap = need_nodes + need_nodes / 8 + 16384;
      if (!need_rec && !need_nodes) return fail(c, PPM_ERR_STATE, "a pass reported an inconsistent state");
      retried += again.size();
    }
    todo.swap(again);
  }
  // merge the twins' accumulators (and pass counters) into ours; the last pass image follows the last pass
  const int64_t nacc = (int64_t)c->accum_pixels * 3 + 1;
  for (int j = 1; j < lanes; ++j) {
    k_accum_merge<<<nblk(nacc, 256), 256, 0, c->stream>>>(c->accum.as<double>(), lane[j]->accum.as<double>(), nacc);
    KCHECK(c);
  }
  if (last_lane != 0)
    CK(c, cudaMemcpyAsync(c->pass_img.p, lane[last_lane]->pass_img.p, (size_t)c->accum_pixels * 24, cudaMemcpyDeviceToDevice, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  if (!c->opt_graph) {                                    // stream mode: k_gather by CUDA events
    double g = 0.0;
    for (int j = 0; j < lanes; ++j) g += lane[j]->ms[7];
    st.ms[7] = g;
  }
  st.ct[5] += stream_launches;                            // (graph mode: the graph's kernel nodes were counted per pass above)
  st.ct[7] = retried;
  {
    const PassOut& o = outs[(size_t)npass - 1];
    for (int k = 0; k < PPM_NSTAMP; ++k)
      c->timeline[k] = (o.stamp[k] && o.stamp[k] >= o.stamp[ST_BEGIN]) ? (double)(o.stamp[k] - o.stamp[ST_BEGIN]) * 1e-6 : 0.0;
  }
  std::memcpy(c->ms, st.ms, sizeof st.ms);
  std::memcpy(c->counters, st.ct, sizeof st.ct);
  // the photon set and the map of this lane's last pass stay current for the probe entry points
  RC(pull_ps(c));
  c->n_rec = c->hps.n_map; c->power = power; c->rec_traced = true; c->rec_nphoton = (uint64_t)total;
  c->have_map = c->hps.status == 0u;
  return PPM_OK;
}

// ---- NCCL, loaded at run time -------------------------------------------------------------------------------------------
struct NcclUid { char internal[128]; };
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(NcclUid*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUid, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*Reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string why;
};
NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return &api;
  tried = true;
  // the copy already mapped into the process (e.g. PyTorch's) is found first by its soname
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.h) break;
  }
  if (!api.h) { api.why = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?"); return &api; }
  api.GetUniqueId = (int (*)(NcclUid*))dlsym(api.h, "ncclGetUniqueId");
  api.CommInitRank = (int (*)(void**, int, NcclUid, int))dlsym(api.h, "ncclCommInitRank");
  api.CommDestroy = (int (*)(void*))dlsym(api.h, "ncclCommDestroy");
  api.Reduce = (int (*)(const void*, void*, size_t, int, int, int, void*, cudaStream_t))dlsym(api.h, "ncclReduce");
  api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(api.h, "ncclAllReduce");
  api.GetErrorString = (const char* (*)(int))dlsym(api.h, "ncclGetErrorString");
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.Reduce || !api.AllReduce) {
    api.why = "libnccl.so.2 lacks an expected symbol";
    api.h = nullptr;
  }
  return &api;
}
std::string nccl_err(NcclApi* a, int r) { return a->GetErrorString ? std::string(a->GetErrorString(r)) : ("nccl error " + std::to_string(r)); }
const int kNcclFloat64 = 8, kNcclSum = 0;

int opt_from_env(const char* name, int dflt) {
  const char* e = std::getenv(name);
  if (!e || !e[0]) return dflt;
  return std::atoi(e);
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
  // The main stream gets the highest priority: its small photon-branch kernels must be dispatched ahead of the
  // remaining blocks of the long eye-branch kernels on stream2.
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  c->prio_hi = prio_hi; c->prio_lo = prio_lo;
  bool ok = cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi) == cudaSuccess;
  ok = ok && cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, prio_lo) == cudaSuccess;
  for (int i = 0; ok && i < ppm_ctx::EV_COUNT; ++i)
    ok = cudaEventCreateWithFlags(&c->ev[i], (i == ppm_ctx::EV_G0 || i == ppm_ctx::EV_G1) ? cudaEventDefault : cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&c->ps, sizeof(PassDev)) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&c->bt, sizeof(BatchDev)) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&c->out, sizeof(PassOut) * PPM_BATCH_MAX) == cudaSuccess;
  ok = ok && cudaMallocHost((void**)&c->bt_h, sizeof(BatchDev)) == cudaSuccess;
  ok = ok && cudaMallocHost((void**)&c->out_h, sizeof(PassOut) * PPM_BATCH_MAX) == cudaSuccess;
  if (!ok) { cudaGetLastError(); ppm_destroy(c); return PPM_ERR_CUDA; }
  std::memset(&c->hps, 0, sizeof c->hps);
  c->hps.grid.nx = c->hps.grid.ny = c->hps.grid.nz = 1; c->hps.grid.ncells = 1; c->hps.grid.inv_cell = 1.0;
  cudaMemcpy(c->ps, &c->hps, sizeof(PassDev), cudaMemcpyHostToDevice);
  std::memset(&c->scene, 0, sizeof c->scene);
  ppm_camera_default(&c->cam);
  // switches: the environment is read once, here
  c->opt_lanes = std::max(1, std::min(opt_from_env("PPM_LANES", PPM_DEFAULT_LANES), (int)PPM_MAX_LANES));
  c->opt_dl_cull = opt_from_env("PPM_DL_CULL", 1) != 0;
  c->opt_gather_heavy = opt_from_env("PPM_GATHER_HEAVY", 1) != 0;
  c->opt_dl_stats = opt_from_env("PPM_DL_STATS", 0) != 0;
  c->opt_graph = opt_from_env("PPM_GRAPH", 1) != 0;
  c->opt_bvh = opt_from_env("PPM_BVH", 0) != 0;
  *out = c;
  return PPM_OK;
}

void ppm_destroy(ppm_ctx* c) {
  if (!c) return;
  for (ppm_ctx* t : c->twins) ppm_destroy(t);
  c->twins.clear();
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->stream2) cudaStreamSynchronize(c->stream2);
  if (c->comm) { NcclApi* a = nccl_api(); if (a->h) a->CommDestroy(c->comm); c->comm = nullptr; }
  if (c->gexec) cudaGraphExecDestroy(c->gexec);
  DBuf* all[] = {&c->r_pos, &c->r_dir, &c->r_wl, &c->r_tag, &c->pmask, &c->pbase, &c->order0, &c->cid, &c->words_p, &c->start_p, &c->key_a, &c->key_b,
                 &c->val_a, &c->val_b, &c->rs_table, &c->tsum_p, &c->m_P, &c->m_D, &c->m_orig, &c->bbox, &c->axis_hist, &c->qcell, &c->words_q,
                 &c->qcnt, &c->qstart, &c->qrank, &c->qpic, &c->skey, &c->sidx, &c->tsum_q, &c->heavy, &c->knn_lo, &c->knn_hi, &c->knn_thr,
                 &c->knn_cnt, &c->knn_act, &c->st_in0, &c->st_in1, &c->st_out0, &c->st_out1, &c->st_out2, &c->st_out3, &c->st_out4, &c->e_head,
                 &c->e_prev, &c->e_pos, &c->e_nrm, &c->e_w, &c->e_emit, &c->e_direct, &c->e_photon, &c->dl_dbg, &c->dl_masks, &c->cull,
                 &c->pass_img, &c->accum, &c->bvh_nodes, &c->bvh_prims, &c->g_prims};
  for (DBuf* b : all) b->release();
  if (c->ps) cudaFree(c->ps);
  if (c->bt) cudaFree(c->bt);
  if (c->out) cudaFree(c->out);
  if (c->bt_h) cudaFreeHost(c->bt_h);
  if (c->out_h) cudaFreeHost(c->out_h);
  for (int i = 0; i < ppm_ctx::EV_COUNT; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char* ppm_last_error(const ppm_ctx* c) { return c ? c->err.c_str() : "null context"; }
void* ppm_stream(ppm_ctx* c) { return c ? (void*)c->stream : nullptr; }

static int* opt_slot(ppm_ctx* c, const char* name) {
  if (!name) return nullptr;
  if (!std::strcmp(name, "lanes")) return &c->opt_lanes;
  if (!std::strcmp(name, "dl_cull")) return &c->opt_dl_cull;
  if (!std::strcmp(name, "gather_heavy")) return &c->opt_gather_heavy;
  if (!std::strcmp(name, "dl_stats")) return &c->opt_dl_stats;
  if (!std::strcmp(name, "graph")) return &c->opt_graph;
  if (!std::strcmp(name, "bvh")) return &c->opt_bvh;
  return nullptr;
}
int ppm_option_set(ppm_ctx* c, const char* name, int64_t value) {
  if (!c) return PPM_ERR_ARG;
  int* s = opt_slot(c, name);
  if (!s) return fail(c, PPM_ERR_ARG, std::string("unknown option ") + (name ? name : "(null)"));
  if (s == &c->opt_lanes) {
    if (value < 1 || value > PPM_MAX_LANES) return fail(c, PPM_ERR_ARG, "lanes must be 1..8");
    *s = (int)value;
  } else {
    *s = value != 0;
  }
  return PPM_OK;
}
int ppm_option_get(ppm_ctx* c, const char* name, int64_t* value) {
  if (!c || !value) return PPM_ERR_ARG;
  int* s = opt_slot(c, name);
  if (!s) return fail(c, PPM_ERR_ARG, std::string("unknown option ") + (name ? name : "(null)"));
  *value = *s;
  return PPM_OK;
}

int ppm_scene_set(ppm_ctx* c, const ppm_prim* prims, int32_t nprims, const ppm_material* mats, int32_t nmats,
                  const ppm_light* lights, int32_t nlights) {
  if (!c) return PPM_ERR_ARG;
  if (!prims || !mats || nprims <= 0 || nmats <= 0 || nlights < 0 || (nlights > 0 && !lights)) return fail(c, PPM_ERR_ARG, "null/empty scene arrays");
  if (nmats > PPM_MAX_MATS || nlights > PPM_MAX_LIGHTS) return fail(c, PPM_ERR_CAPACITY, "scene exceeds 48 materials / 8 lights");
  for (int i = 0; i < nprims; ++i) {
    if (prims[i].material < 0 || prims[i].material >= nmats) return fail(c, PPM_ERR_ARG, "primitive material index out of range");
    if (prims[i].type < PPM_SHAPE_POINT || prims[i].type > PPM_SHAPE_PARALLELOGRAM) return fail(c, PPM_ERR_ARG, "bad shape type");
  }
  for (int i = 0; i < nlights; ++i)
    if (lights[i].type < PPM_LIGHT_POINT || lights[i].type > PPM_LIGHT_SUN) return fail(c, PPM_ERR_ARG, "bad light type");
  // Up to PPM_MAX_PRIMS primitives travel in the kernel parameters and every ray tests all of them, as the reference
  // does.  Larger scenes (or any scene with the "bvh" option set) go through the hierarchy of bvh_types.h.
  const bool bvh = nprims > PPM_MAX_PRIMS || c->opt_bvh != 0;
  // the same scene again (a host loop that hands the scene over every pass): keep the calibration and the pass graph
  if (c->have_scene && (c->scene.bvh_on != 0) == bvh && (int32_t)c->h_prims.size() == nprims && (int32_t)c->h_mats.size() == nmats &&
      (int32_t)c->h_lights.size() == nlights && std::memcmp(c->h_prims.data(), prims, sizeof(ppm_prim) * nprims) == 0 &&
      std::memcmp(c->h_mats.data(), mats, sizeof(ppm_material) * nmats) == 0 &&
      (nlights == 0 || std::memcmp(c->h_lights.data(), lights, sizeof(ppm_light) * nlights) == 0))
    return PPM_OK;
  DevScene* ns = new DevScene;
  std::unique_ptr<DevScene> ns_owner(ns);
  std::memset(ns, 0, sizeof *ns);
  ns->nmats = nmats; ns->nlights = nlights; ns->nprims_total = nprims;
  std::memcpy(ns->mats, mats, sizeof(ppm_material) * nmats);
  if (nlights) std::memcpy(ns->lights, lights, sizeof(ppm_light) * nlights);
  std::vector<BvhNode> nodes;
  std::vector<BvhPrim> bprims;
  int depth = 0;
  if (!bvh) {
    ns->nprims = nprims;
    std::memcpy(ns->prims, prims, sizeof(ppm_prim) * nprims);
  } else {
    // the constant list keeps the infinite planes (in object order), the hierarchy takes the bounded primitives
    int nu = 0;
    for (int i = 0; i < nprims; ++i) {
      if (prims[i].type != PPM_SHAPE_PLAIN) continue;
      if (nu >= PPM_MAX_PRIMS) return fail(c, PPM_ERR_CAPACITY, "scene has more than 64 infinite planes");
      ns->prims[nu] = prims[i]; ns->unb_obj[nu] = i; ++nu;
    }
    ns->nprims = nu; ns->bvh_on = 1;
    std::string why;
    if (!ppmhost::bvh_build(prims, nprims, nodes, bprims, &depth, why)) return fail(c, PPM_ERR_CAPACITY, "BVH: " + why);
    // class (b) of the shadow-ray culling: primitives lying in the plane of an area light (its own emitter geometry)
    for (int li = 0; li < nlights; ++li) {
      const ppm_light& l = lights[li];
      if (l.type != PPM_LIGHT_PARALLELOGRAM) continue;
      const double cx[3] = {l.dir1[1] * l.dir2[2] - l.dir2[1] * l.dir1[2], l.dir1[2] * l.dir2[0] - l.dir2[2] * l.dir1[0],
                            l.dir1[0] * l.dir2[1] - l.dir2[0] * l.dir1[1]};
      const double cn = std::sqrt(cx[0] * cx[0] + cx[1] * cx[1] + cx[2] * cx[2]);
      if (!(cn > 0.0 && cn < 1e150)) continue;
      const double nl[3] = {cx[0] / cn, cx[1] / cn, cx[2] / cn};
      for (BvhPrim& q : bprims)
        if (cull_in_light_plane(l, nl, prims[q.obj])) q.type |= 1 << (8 + li);
    }
  }
  ns->types.nwords = ns->nprims > 32 ? 2 : 1;
  for (int o = 0; o < ns->nprims; ++o) {
    const unsigned long long bit = 1ull << o;
    if (ns->prims[o].type == PPM_SHAPE_PLAIN) ns->types.plain |= bit;
    else if (ns->prims[o].type == PPM_SHAPE_SPHERE) ns->types.sphere |= bit;
    else if (ns->prims[o].type == PPM_SHAPE_POLYGON) ns->types.poly |= bit;
    else if (ns->prims[o].type == PPM_SHAPE_PARALLELOGRAM) ns->types.para |= bit;
  }
  CK(c, cudaSetDevice(c->device));
  CK(c, cudaStreamSynchronize(c->stream));
  CK(c, cudaStreamSynchronize(c->stream2));
  for (ppm_ctx* t : c->twins) { CK(c, cudaStreamSynchronize(t->stream)); CK(c, cudaStreamSynchronize(t->stream2)); }
  c->have_scene = false;
  if (bvh) {
    RC(ens(c, c->g_prims, sizeof(ppm_prim) * (size_t)nprims));
    CK(c, cudaMemcpy(c->g_prims.p, prims, sizeof(ppm_prim) * (size_t)nprims, cudaMemcpyHostToDevice));
    ns->gprims = c->g_prims.as<ppm_prim>();
    if (!nodes.empty()) {
      RC(ens(c, c->bvh_nodes, sizeof(BvhNode) * nodes.size()));
      RC(ens(c, c->bvh_prims, sizeof(BvhPrim) * bprims.size()));
      CK(c, cudaMemcpy(c->bvh_nodes.p, nodes.data(), sizeof(BvhNode) * nodes.size(), cudaMemcpyHostToDevice));
      CK(c, cudaMemcpy(c->bvh_prims.p, bprims.data(), sizeof(BvhPrim) * bprims.size(), cudaMemcpyHostToDevice));
      ns->bvh = c->bvh_nodes.as<BvhNode>();
      ns->bprims = c->bvh_prims.as<BvhPrim>();
    }
  }
  c->bvh_nnodes = nodes.size(); c->bvh_nprims = bprims.size(); c->bvh_depth = depth;
  c->scene = *ns;
  c->h_prims.assign(prims, prims + nprims);
  c->h_mats.assign(mats, mats + nmats);
  c->h_lights.assign(lights, lights + nlights);
  RC(upload_cull(c));
  c->have_scene = true;
  c->scene_ver++;
  return PPM_OK;
}

int ppm_camera_set(ppm_ctx* c, const ppm_camera* cam) {
  if (!c || !cam) return PPM_ERR_ARG;
  if (cam->xreso <= 0 || cam->yreso <= 0) return fail(c, PPM_ERR_ARG, "bad resolution");
  if (cam->pfilter < PPM_FILTER_NONE || cam->pfilter > PPM_FILTER_GAUSS) return fail(c, PPM_ERR_ARG, "bad photon filter");
  if (c->have_camera && std::memcmp(&c->cam, cam, sizeof *cam) == 0) return PPM_OK;   // unchanged: keep calibration and graph
  c->cam = *cam;
  c->have_camera = true;
  c->cam_ver++;
  return PPM_OK;
}

int ppm_intersect(ppm_ctx* c, const double* rays6, int64_t n, int32_t* hit_idx, double* t, double* pos3, double* nrm3, int32_t* io) {
  if (!c) return PPM_ERR_ARG;
  if (!c->have_scene) return fail(c, PPM_ERR_STATE, "scene not set");
  if (n < 0 || (n > 0 && (!rays6 || !hit_idx))) return fail(c, PPM_ERR_ARG, "null rays / hit_idx");
  if (n == 0) return PPM_OK;
  CK(c, cudaSetDevice(c->device));
  const void* drays; void *dh, *dt, *dp, *dn, *di;
  RC(stage_in(c, rays6, (size_t)n * 48, c->st_in0, &drays));
  RC(stage_out(c, hit_idx, (size_t)n * 4, c->st_out0, &dh));
  RC(stage_out(c, t, (size_t)n * 8, c->st_out1, &dt));
  RC(stage_out(c, pos3, (size_t)n * 24, c->st_out2, &dp));
  RC(stage_out(c, nrm3, (size_t)n * 24, c->st_out3, &dn));
  RC(stage_out(c, io, (size_t)n * 4, c->st_out4, &di));
  if (c->scene.bvh_on) k_intersect<true><<<nblk(n, 128), 128, 0, c->stream>>>(c->scene, (const double*)drays, n, (int32_t*)dh, (double*)dt, (double*)dp,
                                                                             (double*)dn, (int32_t*)di);
  else k_intersect<false><<<nblk(n, 128), 128, 0, c->stream>>>(c->scene, (const double*)drays, n, (int32_t*)dh, (double*)dt, (double*)dp,
                                                               (double*)dn, (int32_t*)di);
  KCHECK(c);
  RC(finish_out(c, hit_idx, (size_t)n * 4, dh));
  RC(finish_out(c, t, (size_t)n * 8, dt));
  RC(finish_out(c, pos3, (size_t)n * 24, dp));
  RC(finish_out(c, nrm3, (size_t)n * 24, dn));
  RC(finish_out(c, io, (size_t)n * 4, di));
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

int ppm_emit_photons(ppm_ctx* c, uint64_t seed, uint32_t pass, const int64_t* n_per_light, ppm_photon* out) {
  if (!c) return PPM_ERR_ARG;
  if (!c->have_scene || c->scene.nlights == 0) return fail(c, PPM_ERR_STATE, "scene with lights not set");
  if (!n_per_light || !out) return fail(c, PPM_ERR_ARG, "null argument");
  CK(c, cudaSetDevice(c->device));
  LightSplit ls; int64_t total;
  RC(light_split(c, n_per_light, &ls, &total));
  if (total == 0) return PPM_OK;
  void* d;
  RC(stage_out(c, out, (size_t)total * sizeof(ppm_photon), c->st_out0, &d));
  k_emit<<<nblk(total, 128), 128, 0, c->stream>>>(c->scene, ls, seed, pass, total, (ppm_photon*)d);
  KCHECK(c);
  RC(finish_out(c, out, (size_t)total * sizeof(ppm_photon), d));
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

int ppm_trace_photons(ppm_ctx* c, uint64_t seed, uint32_t pass, int uc, const int64_t* n_per_light, double power, uint64_t* n_stored) {
  if (!c) return PPM_ERR_ARG;
  if (!c->have_scene || c->scene.nlights == 0) return fail(c, PPM_ERR_STATE, "scene with lights not set");
  if (!n_per_light) return fail(c, PPM_ERR_ARG, "null n_per_light");
  CK(c, cudaSetDevice(c->device));
  RC(do_trace_photons(c, seed, pass, uc, n_per_light, power));
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
  void* d;
  size_t bytes = (size_t)c->n_rec * sizeof(ppm_photon);
  RC(stage_out(c, out, bytes, c->st_out0, &d));
  k_export<<<nblk((int64_t)c->n_rec, 256), 256, 0, c->stream>>>(recbuf(c), c->n_rec, (ppm_photon*)d);
  KCHECK(c);
  RC(finish_out(c, out, bytes, d));
  if (tags) CK(c, cudaMemcpyAsync(tags, c->r_tag.p, (size_t)c->n_rec * 8, is_device_ptr(tags) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

int ppm_photons_import(ppm_ctx* c, const ppm_photon* in, uint64_t n, double power) {
  if (!c) return PPM_ERR_ARG;
  if (n > 0 && !in) return fail(c, PPM_ERR_ARG, "null input");
  if (n >= (1ull << 32) - 16) return fail(c, PPM_ERR_CAPACITY, "at most 2^32-16 photon records");
  CK(c, cudaSetDevice(c->device));
  RC(ensure_records(c, n ? n : 1, 0));
  if (n) {
    const void* d;
    RC(stage_in(c, in, (size_t)n * sizeof(ppm_photon), c->st_in0, &d));
    k_import<<<nblk((int64_t)n, 256), 256, 0, c->stream>>>((const ppm_photon*)d, n, recbuf(c));
    KCHECK(c);
    CK(c, cudaStreamSynchronize(c->stream));
  }
  c->n_rec = n; c->power = power; c->have_map = false;
  c->rec_traced = false; c->rec_nphoton = 0;
  return PPM_OK;
}

int ppm_map_build(ppm_ctx* c, double radius2) {
  if (!c) return PPM_ERR_ARG;
  CK(c, cudaSetDevice(c->device));
  return do_map_build(c, radius2, true);
}

int ppm_within(ppm_ctx* c, const double* q3, int64_t nq, uint32_t* idx, uint32_t* count, uint32_t cap) {
  if (!c) return PPM_ERR_ARG;
  if (!c->have_map) return fail(c, PPM_ERR_STATE, "photon map not built");
  if (nq < 0 || (nq > 0 && (!q3 || !count || (cap > 0 && !idx)))) return fail(c, PPM_ERR_ARG, "null argument");
  if (nq == 0) return PPM_OK;
  CK(c, cudaSetDevice(c->device));
  const void* dq; void *di, *dc;
  size_t ib = (size_t)nq * (cap ? cap : 1) * 4;
  RC(stage_in(c, q3, (size_t)nq * 24, c->st_in0, &dq));
  RC(ens(c, c->st_out0, ib));
  di = c->st_out0.p;
  CK(c, cudaMemsetAsync(di, 0, ib, c->stream));      // slots beyond a query's count stay 0
  RC(stage_out(c, count, (size_t)nq * 4, c->st_out1, &dc));
  k_within<<<nblk(nq, 128), 128, 0, c->stream>>>(c->ps, cellindex(c), mapsoa(c), (const double*)dq, nq, (uint32_t*)di, (uint32_t*)dc, cap);
  KCHECK(c);
  RC(finish_out(c, count, (size_t)nq * 4, dc));
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
  const void *dp, *dn; void *dr, *dc;
  RC(stage_in(c, pos3, (size_t)n * 24, c->st_in0, &dp));
  RC(stage_in(c, nrm3, (size_t)n * 24, c->st_in1, &dn));
  RC(stage_out(c, rgb3, (size_t)n * 24, c->st_out0, &dr));
  RC(stage_out(c, counts, (size_t)n * 4, c->st_out1, &dc));
  RC(launch_gather(c, (const double*)dp, (const double*)dn, n, filter, (double*)dr, (uint32_t*)dc));
  RC(finish_out(c, rgb3, (size_t)n * 24, dr));
  RC(finish_out(c, counts, (size_t)n * 4, dc));
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
  const void *dp, *dn; void *dr, *dk, *dc;
  RC(stage_in(c, pos3, (size_t)n * 24, c->st_in0, &dp));
  RC(stage_in(c, nrm3, (size_t)n * 24, c->st_in1, &dn));
  RC(stage_out(c, rgb3, (size_t)n * 24, c->st_out0, &dr));
  RC(stage_out(c, r2k, (size_t)n * 8, c->st_out1, &dk));
  RC(stage_out(c, counts, (size_t)n * 4, c->st_out2, &dc));
  RC(launch_gather_knn(c, (const double*)dp, (const double*)dn, n, k, filter, (double*)dr, (double*)dk, (uint32_t*)dc));
  RC(finish_out(c, rgb3, (size_t)n * 24, dr));
  RC(finish_out(c, r2k, (size_t)n * 8, dk));
  RC(finish_out(c, counts, (size_t)n * 4, dc));
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

int ppm_direct_light(ppm_ctx* c, const double* pos3, const double* nrm3, int64_t n, double* rgb3) {
  if (!c) return PPM_ERR_ARG;
  if (!c->have_scene) return fail(c, PPM_ERR_STATE, "scene not set");
  if (n < 0 || (n > 0 && (!pos3 || !nrm3 || !rgb3))) return fail(c, PPM_ERR_ARG, "null argument");
  if (n == 0) return PPM_OK;
  CK(c, cudaSetDevice(c->device));
  const void *dp, *dn; void* dr;
  RC(stage_in(c, pos3, (size_t)n * 24, c->st_in0, &dp));
  RC(stage_in(c, nrm3, (size_t)n * 24, c->st_in1, &dn));
  RC(stage_out(c, rgb3, (size_t)n * 24, c->st_out0, &dr));
  RC(probe_direct_light(c, (const double*)dp, (const double*)dn, n, (double*)dr));
  RC(finish_out(c, rgb3, (size_t)n * 24, dr));
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

int ppm_generate_rays(ppm_ctx* c, uint64_t seed, uint32_t pass, double* rays6) {
  if (!c) return PPM_ERR_ARG;
  if (!c->have_camera) return fail(c, PPM_ERR_STATE, "camera not set");
  if (!rays6) return fail(c, PPM_ERR_ARG, "null output");
  CK(c, cudaSetDevice(c->device));
  int64_t n = (int64_t)c->cam.xreso * c->cam.yreso;
  void* d;
  RC(stage_out(c, rays6, (size_t)n * 48, c->st_out0, &d));
  k_gen_rays<<<nblk(n, 128), 128, 0, c->stream>>>(c->cam, seed, pass, n, (double*)d);
  KCHECK(c);
  RC(finish_out(c, rays6, (size_t)n * 48, d));
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

int ppm_trace_rays_classic(ppm_ctx* c, const double* rays6, int64_t n, int64_t first_pixel, uint64_t seed, uint32_t pass, double* rgb3) {
  if (!c) return PPM_ERR_ARG;
  if (!c->have_scene) return fail(c, PPM_ERR_STATE, "scene not set");
  if (n < 0 || first_pixel < 0 || (n > 0 && (!rays6 || !rgb3))) return fail(c, PPM_ERR_ARG, "null argument");
  if (n == 0) return PPM_OK;
  CK(c, cudaSetDevice(c->device));
  const void* dr; void* dout;
  RC(stage_in(c, rays6, (size_t)n * 48, c->st_in0, &dr));
  RC(stage_out(c, rgb3, (size_t)n * 24, c->st_out0, &dout));
  RC(do_trace_rays(c, (const double*)dr, n, first_pixel, seed, pass, 1, (double*)dout, 1));
  RC(finish_out(c, rgb3, (size_t)n * 24, dout));
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
  const void* dr; void* dout;
  // the eye-path buffers must not alias the staging of the rays / the output
  RC(stage_in(c, rays6, (size_t)n * 48, c->st_in1, &dr));
  RC(stage_out(c, rgb3, (size_t)n * 24, c->st_out4, &dout));
  RC(do_trace_rays(c, (const double*)dr, n, first_pixel, seed, pass, uc, (double*)dout, 0));
  RC(finish_out(c, rgb3, (size_t)n * 24, dout));
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

int ppm_render_pass(ppm_ctx* c, uint64_t seed, uint32_t pass, int64_t nphoton, double radius2, int uc) {
  if (!c) return PPM_ERR_ARG;
  return render_batch(c, seed, pass, 1, 1, nphoton, &radius2, uc, 1);
}

// A batch of passes with cross-pass overlap.  Passes are independent, and within one pass the FP64-bound kernels
// (direct light, gather) and the latency-bound ones (photon tracing, eye-path expansion, map build) cannot fill the
// GPU together.  Lanes -- this context and internal twins on the same GPU, each with its own streams, buffers and
// pass graph -- render alternating passes so that different phases of different passes co-schedule.  One host thread
// launches everything; nothing waits for the host until the batch ends.  The twins' accumulators are merged
// afterwards, so ppm_accum_read / ppm_image_mean see every pass.
int ppm_render_passes(ppm_ctx* c, uint64_t seed, uint32_t first_pass, uint32_t pass_stride, int32_t npass, int64_t nphoton,
                      const double* radius2, int uc) {
  if (!c) return PPM_ERR_ARG;
  if (npass < 0 || (npass > 0 && !radius2)) return fail(c, PPM_ERR_ARG, "bad pass batch");
  if (npass == 0) return PPM_OK;
  if (!c->have_scene || !c->have_camera) return fail(c, PPM_ERR_STATE, "scene and camera must be set");
  int rc = render_batch(c, seed, first_pass, pass_stride, npass, nphoton, radius2, uc, c->opt_lanes);
  if (rc) {
    // a failed batch leaves no half-summed twin behind: fold whatever the lanes accumulated into the parent
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    cudaGetLastError();
    const int64_t nacc = (int64_t)c->accum_pixels * 3 + 1;
    for (ppm_ctx* t : c->twins)
      if (t->accum.p && t->accum_pixels == c->accum_pixels && c->accum.p)
        k_accum_merge<<<nblk(nacc, 256), 256, 0, c->stream>>>(c->accum.as<double>(), t->accum.as<double>(), nacc);
    cudaStreamSynchronize(c->stream);
    cudaGetLastError();
  }
  return rc;
}

int ppm_last_pass_stats(ppm_ctx* c, double ms[8], uint64_t counters[8]) {
  if (!c) return PPM_ERR_ARG;
  if (ms) std::memcpy(ms, c->ms, sizeof c->ms);
  if (counters) std::memcpy(counters, c->counters, sizeof c->counters);
  return PPM_OK;
}

int ppm_last_pass_timeline(ppm_ctx* c, double ms_since_begin[16]) {
  if (!c || !ms_since_begin) return PPM_ERR_ARG;
  static_assert(PPM_NSTAMP == 16, "timeline slots");
  std::memcpy(ms_since_begin, c->timeline, sizeof c->timeline);
  return PPM_OK;
}

static int copy_out(ppm_ctx* c, void* user, const void* dev, size_t bytes) {
  CK(c, cudaMemcpyAsync(user, dev, bytes, is_device_ptr(user) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
  CK(c, wait_stream(c, c->stream));
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
  for (ppm_ctx* t : c->twins)                           // the lanes' partial sums go too
    if (t->accum.p && t->accum_pixels) CK(c, cudaMemsetAsync(t->accum.p, 0, (size_t)(t->accum_pixels * 3 + 1) * 8, c->stream));
  if (c->accum.p && c->accum_pixels) CK(c, cudaMemsetAsync(c->accum.p, 0, (size_t)(c->accum_pixels * 3 + 1) * 8, c->stream));
  if (c->have_camera) RC(ensure_accum(c));
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

// Checkpoint / resume.  The reference's pass images are files (util/iterator.rb:96-117), so a killed job keeps what it has
// rendered and averager2.rb:49-62 sums whatever is there; here the sums live on the device, so they can be taken out
// (ppm_accum_read / ppm_accum_save) and put back (ppm_accum_add / ppm_accum_load) -- also how pass sums rendered
// elsewhere are merged into a frame.
int ppm_accum_add(ppm_ctx* c, const double* rgb3, uint32_t n_pass) {
  if (!c || !rgb3) return PPM_ERR_ARG;
  if (!c->have_camera) return fail(c, PPM_ERR_STATE, "camera not set");
  CK(c, cudaSetDevice(c->device));
  RC(ensure_accum(c));
  const void* dsrc;
  RC(stage_in(c, rgb3, (size_t)c->accum_pixels * 24, c->st_in0, &dsrc));
  const int64_t n = (int64_t)c->accum_pixels * 3;
  k_accum_add<<<nblk(n, 256), 256, 0, c->stream>>>(c->accum.as<double>(), (const double*)dsrc, n, (double)n_pass);
  KCHECK(c);
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}
namespace {
struct AccFileHeader { char magic[8]; uint32_t xreso, yreso, n_pass, _pad; };
const char kAccMagic[8] = {'P', 'P', 'M', 'A', 'C', 'C', '1', '\n'};
}
int ppm_accum_save(ppm_ctx* c, const char* path) {
  if (!c || !path) return PPM_ERR_ARG;
  if (!c->accum.p || !c->accum_pixels) return fail(c, PPM_ERR_STATE, "no accumulator yet");
  std::vector<double> sum((size_t)c->accum_pixels * 3);
  uint32_t n = 0;
  RC(ppm_accum_read(c, sum.data(), &n));
  AccFileHeader h;
  std::memcpy(h.magic, kAccMagic, 8);
  h.xreso = (uint32_t)c->cam.xreso; h.yreso = (uint32_t)c->cam.yreso; h.n_pass = n; h._pad = 0;
  // written beside the target and renamed: a job killed while saving leaves the previous checkpoint intact
  const std::string tmp = std::string(path) + ".tmp";
  FILE* f = std::fopen(tmp.c_str(), "wb");
  if (!f) return fail(c, PPM_ERR_IO, std::string("cannot write ") + tmp);
  bool ok = std::fwrite(&h, sizeof h, 1, f) == 1 && std::fwrite(sum.data(), 8, sum.size(), f) == sum.size();
  ok = (std::fclose(f) == 0) && ok;
  if (!ok || std::rename(tmp.c_str(), path) != 0) { std::remove(tmp.c_str()); return fail(c, PPM_ERR_IO, std::string("cannot write ") + path); }
  return PPM_OK;
}
int ppm_accum_load(ppm_ctx* c, const char* path, uint32_t* n_pass) {
  if (!c || !path) return PPM_ERR_ARG;
  if (!c->have_camera) return fail(c, PPM_ERR_STATE, "camera not set");
  FILE* f = std::fopen(path, "rb");
  if (!f) return fail(c, PPM_ERR_IO, std::string("cannot read ") + path);
  AccFileHeader h;
  std::vector<double> sum;
  bool ok = std::fread(&h, sizeof h, 1, f) == 1 && std::memcmp(h.magic, kAccMagic, 8) == 0;
  if (ok && (h.xreso != (uint32_t)c->cam.xreso || h.yreso != (uint32_t)c->cam.yreso)) {
    std::fclose(f);
    return fail(c, PPM_ERR_ARG, "checkpoint resolution differs from the camera's");
  }
  if (ok) {
    sum.resize((size_t)h.xreso * h.yreso * 3);
    ok = std::fread(sum.data(), 8, sum.size(), f) == sum.size();
  }
  std::fclose(f);
  if (!ok) return fail(c, PPM_ERR_PARSE, std::string("not an accumulator checkpoint: ") + path);
  RC(ppm_accum_add(c, sum.data(), h.n_pass));
  if (n_pass) *n_pass = h.n_pass;
  return PPM_OK;
}

int ppm_accum_read(ppm_ctx* c, double* rgb3, uint32_t* n_pass) {
  if (!c) return PPM_ERR_ARG;
  if (!c->accum.p || !c->accum_pixels) return fail(c, PPM_ERR_STATE, "no accumulator yet");
  CK(c, cudaSetDevice(c->device));
  if (rgb3) RC(copy_out(c, rgb3, c->accum.p, (size_t)c->accum_pixels * 24));
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
  RC(ensure_accum(c));
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
  double np = 0.0;
  CK(c, cudaMemcpyAsync(&np, c->accum.as<double>() + n, 8, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  if (!(np > 0.0)) return fail(c, PPM_ERR_STATE, "no pass has been accumulated");
  void* d;
  RC(stage_out(c, rgb3, (size_t)n * 8, c->st_out0, &d));
  k_scale<<<nblk(n, 256), 256, 0, c->stream>>>(c->accum.as<double>(), c->accum.as<double>() + n, n, (double*)d);
  KCHECK(c);
  RC(finish_out(c, rgb3, (size_t)n * 8, d));
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

// ---- multi-GPU frame: one NCCL sum-reduce of [3*W*H sums | pass count] (util/averager2.rb:49-62,86) --------------------
int ppm_comm_unique_id(void* id128) {
  if (!id128) return PPM_ERR_ARG;
  NcclApi* a = nccl_api();
  if (!a->h) return PPM_ERR_STATE;
  NcclUid id;
  std::memset(&id, 0, sizeof id);
  if (a->GetUniqueId(&id) != 0) return PPM_ERR_CUDA;
  std::memcpy(id128, &id, sizeof id);
  return PPM_OK;
}
int ppm_comm_init(ppm_ctx* c, int32_t nranks, int32_t rank, const void* id128) {
  if (!c || !id128 || nranks <= 0 || rank < 0 || rank >= nranks) return c ? fail(c, PPM_ERR_ARG, "bad communicator arguments") : PPM_ERR_ARG;
  NcclApi* a = nccl_api();
  if (!a->h) return fail(c, PPM_ERR_STATE, a->why);
  CK(c, cudaSetDevice(c->device));
  if (c->comm) { a->CommDestroy(c->comm); c->comm = nullptr; }
  NcclUid id;
  std::memcpy(&id, id128, sizeof id);
  int r = a->CommInitRank(&c->comm, nranks, id, rank);
  if (r != 0) { c->comm = nullptr; return fail(c, PPM_ERR_CUDA, "ncclCommInitRank: " + nccl_err(a, r)); }
  return PPM_OK;
}
int ppm_comm_destroy(ppm_ctx* c) {
  if (!c) return PPM_ERR_ARG;
  if (c->comm) {
    NcclApi* a = nccl_api();
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (a->h) a->CommDestroy(c->comm);
    c->comm = nullptr;
  }
  return PPM_OK;
}
int ppm_accum_reduce(ppm_ctx* c, void* nccl_comm, int32_t root) {
  if (!c) return PPM_ERR_ARG;
  NcclApi* a = nccl_api();
  if (!a->h) return fail(c, PPM_ERR_STATE, a->why);
  void* comm = nccl_comm ? nccl_comm : c->comm;
  if (!comm) return fail(c, PPM_ERR_STATE, "no communicator: call ppm_comm_init or pass an ncclComm_t");
  if (!c->have_camera) return fail(c, PPM_ERR_STATE, "camera not set");
  CK(c, cudaSetDevice(c->device));
  RC(ensure_accum(c));
  const size_t n = (size_t)c->accum_pixels * 3 + 1;
  int r = root < 0 ? a->AllReduce(c->accum.p, c->accum.p, n, kNcclFloat64, kNcclSum, comm, c->stream)
                   : a->Reduce(c->accum.p, c->accum.p, n, kNcclFloat64, kNcclSum, root, comm, c->stream);
  if (r != 0) return fail(c, PPM_ERR_CUDA, "nccl reduce: " + nccl_err(a, r));
  CK(c, cudaStreamSynchronize(c->stream));
  return PPM_OK;
}

}  // extern "C"
