// kernels_photon.cuh -- calc_intersection probe, photon emission and tracing, record import/export
// Part of the single translation unit engine.cu (compiled -fmad=false, sm_100a); see DESIGN.md section 6.
#ifndef PPM_KERNELS_PHOTON_CUH_
#define PPM_KERNELS_PHOTON_CUH_

#include "dev_core.cuh"
#include "pass_state.cuh"


template <bool BVH>
__global__ void k_intersect(const __grid_constant__ DevScene sc, const double* __restrict__ rays6, int64_t n,
                            int32_t* __restrict__ hit, double* __restrict__ t, double* __restrict__ pos3,
                            double* __restrict__ nrm3, int32_t* __restrict__ io) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  D3 p = ld3(rays6 + i * 6), d = ld3(rays6 + i * 6 + 3);
  Isect is;
  bool ok = nearest_hit<BVH>(sc, p, d, is);
  hit[i] = ok ? is.obj : -1;
  if (t) t[i] = ok ? is.t : 0.0;
  if (pos3) st3(pos3 + i * 3, ok ? is.pos : mk3(0, 0, 0));
  if (nrm3) st3(nrm3 + i * 3, ok ? is.nvec : mk3(0, 0, 0));
  if (io) io[i] = ok ? is.io : 0;
}

struct LightSplit {
  int64_t first[PPM_MAX_LIGHTS + 1];   // first[l] = global index of light l's first photon
};
__device__ __forceinline__ int light_of(const LightSplit& ls, int nlights, int64_t i) {
  int l = 0;
  while (l + 1 < nlights && i >= ls.first[l + 1]) ++l;
  return l;
}

__global__ void k_emit(const __grid_constant__ DevScene sc, const __grid_constant__ LightSplit ls, uint64_t seed,
                       uint32_t pass, int64_t n, ppm_photon* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Philox rng(seed, pass, PPM_DOMAIN_PHOTON, (uint64_t)i, 0);
  int wl; D3 pos, dir;
  generate_photon(sc.lights[light_of(ls, sc.nlights, i)], rng, wl, pos, dir);
  st3(out[i].pos, pos); st3(out[i].dir, dir);
  out[i].wl = wl; out[i]._pad = 0;
}

// Unsorted photon records as produced by tracing / import.
struct RecBuf {
  double* pos3;     // [cap][3]
  double* dir3;     // [cap][3]
  uint8_t* wl;      // [cap]
  uint64_t* tag;    // [cap]  (photon index << 4) | depth
};

// Persistent photon tracer with path regeneration.  Every photon path is still its own
// counter-based Philox stream (seed, pass, photon index), so results do not depend on which
// lane traces it: a lane whose photon is absorbed immediately claims the next photon index
// from a global ticket counter (one atomic per warp per refill) instead of idling until the
// longest path of its warp ends.  One loop iteration = (optional) emission + one bounce.
// Records are appended with one atomic per warp (warp-aggregated compaction).
// seed / pass come from the device-resident pass state, the record counter and the ticket live there too: the
// kernel is launched with the same arguments every pass (CUDA graph).  Every photon leaves the set of depths at which
// it stored a record in pmask[photon] (kernels_map.cuh: the tag order of the records then costs one scan).
#ifndef PPM_TRACE_MINB
#define PPM_TRACE_MINB 6
#endif
template <bool BVH>
__global__ void __launch_bounds__(128, PPM_TRACE_MINB)
k_trace_photons(const __grid_constant__ DevScene sc, const __grid_constant__ LightSplit ls, PassDev* ps,
                int uc, int64_t n, RecBuf rec, unsigned long long cap, uint32_t* __restrict__ pmask) {
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const uint64_t seed = ps->seed;
  const uint32_t pass = ps->pass;
  unsigned long long* const counter = &ps->n_rec;
  unsigned long long* const ticket = &ps->ticket;
  bool alive = false, exhausted = false, need_dir = false;
  int64_t idx = 0;
  int wl = 0, medium = -1, depth = 0;
  uint32_t stored = 0;                                   // depths at which this photon stored a record
  D3 pos = mk3(0, 0, 0), dir = mk3(1, 0, 0);
  Philox rng(seed, pass, PPM_DOMAIN_PHOTON, 0, 0);
  for (;;) {
    // ---- refill dead lanes --------------------------------------------------------------
    const unsigned need = __ballot_sync(FULL, !alive && !exhausted);
    if (need) {
      unsigned long long base = 0;
      const int leader = __ffs(need) - 1;
      if ((int)lane == leader) base = atomicAdd(ticket, (unsigned long long)__popc(need));
      base = __shfl_sync(FULL, base, leader);
      if (!alive && !exhausted) {
        const int64_t i = (int64_t)(base + __popc(need & lt_mask));
        if (i < n) {
          idx = i;
          rng = Philox(seed, pass, PPM_DOMAIN_PHOTON, (uint64_t)i, 0);
          need_dir = generate_photon_t<true>(sc.lights[light_of(ls, sc.nlights, i)], rng, wl, pos, dir);   // dir = normal if deferred
          medium = -1; depth = 0; alive = true; stored = 0u;
        } else {
          exhausted = true;
        }
      }
    }
    if (!__any_sync(FULL, alive)) break;
    // ---- deferred diffuse directions: emission of the lanes regenerated above and the diffuse bounces of
    //      the previous iteration (`dir` holds the normal to sample about) --------------------------------------
    if (need_dir) { dir = diffuse_reflection(dir, rng); need_dir = false; }
    // ---- one bounce ----------------------------------------------------------------------
    Isect is;
    bool store = false;
    const bool was_alive = alive;
    const D3 in_dir = dir;
    const int l = depth;
    if (alive) {
      if (!nearest_hit<BVH>(sc, pos, dir, is)) {
        alive = false;
      } else {
        store = (uc == 0 || l > 0) && surf_store_photon(sc.mats[is.mat]);
        D3 nd;
        bool diffuse;
        const bool go = photon_bounce(sc, is, wl, dir, rng, medium, nd, diffuse);
        pos = is.pos;
        ++depth;
        if (go && depth < PPM_MAX_TRACE) { dir = diffuse ? is.nvec : nd; need_dir = diffuse; }
        else alive = false;                                              // `if l >= MAX_TRACE { return vec![] }`
      }
    }
    const unsigned m = __ballot_sync(FULL, store);
    if (m) {
      unsigned long long base = 0;
      const int leader = __ffs(m) - 1;
      if ((int)lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
      base = __shfl_sync(FULL, base, leader);
      if (store) {
        const unsigned long long slot = base + __popc(m & lt_mask);
        if (slot < cap) {
          st3(rec.pos3 + slot * 3, is.pos);
          st3(rec.dir3 + slot * 3, in_dir);
          rec.wl[slot] = (uint8_t)wl;
          rec.tag[slot] = ((uint64_t)idx << 4) | (uint64_t)l;
        }
        stored |= 1u << l;
      }
    }
    if (was_alive && !alive && pmask) pmask[idx] = stored;                // the path has ended: publish its depth set
  }
}

__global__ void k_import(const ppm_photon* __restrict__ in, uint64_t n, RecBuf rec) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < 3; ++k) { rec.pos3[i * 3 + k] = in[i].pos[k]; rec.dir3[i * 3 + k] = in[i].dir[k]; }
  rec.wl[i] = (uint8_t)in[i].wl;
  rec.tag[i] = i << 4;
}
__global__ void k_export(RecBuf rec, uint64_t n, ppm_photon* __restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < 3; ++k) { out[i].pos[k] = rec.pos3[i * 3 + k]; out[i].dir[k] = rec.dir3[i * 3 + k]; }
  out[i].wl = rec.wl[i]; out[i]._pad = 0;
}

#endif
