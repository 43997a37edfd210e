// kernels_map.cuh -- photon map build: bounding box, trimmed region histograms, cell keys, scatter
// Part of the single translation unit engine.cu (compiled -fmad=false, sm_100a); see DESIGN.md section 6.
#ifndef PPM_KERNELS_MAP_CUH_
#define PPM_KERNELS_MAP_CUH_

#include "dev_core.cuh"
#include "kernels_photon.cuh"   // RecBuf

#include <cstring>

// ---- photon map: uniform grid, cell edge >= r, cells linearised x-fastest ----
struct Grid {
  double org[3];
  double inv_cell;
  int32_t nx, ny, nz;
  uint32_t ncells;
};
__device__ __forceinline__ unsigned long long enc_ord(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
static inline double dec_ord(unsigned long long e) {
  unsigned long long b = (e & 0x8000000000000000ull) ? (e & 0x7fffffffffffffffull) : ~e;
  double v;
  std::memcpy(&v, &b, 8);
  return v;
}
// mm[0..2] = min xyz, mm[3..5] = max xyz (order-preserving encoding)
__global__ void k_bbox(const double* __restrict__ pos3, uint64_t n, unsigned long long* __restrict__ mm) {
  unsigned long long lo[3] = {~0ull, ~0ull, ~0ull}, hi[3] = {0ull, 0ull, 0ull};
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    for (int k = 0; k < 3; ++k) {
      unsigned long long e = enc_ord(pos3[i * 3 + k]);
      lo[k] = e < lo[k] ? e : lo[k];
      hi[k] = e > hi[k] ? e : hi[k];
    }
  // warp shuffle reduction, then one shared-memory step per block: 6 global atomics per block instead of per warp
  __shared__ unsigned long long slo[3][8], shi[3][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = 0; k < 3; ++k) {
    for (int o = 16; o > 0; o >>= 1) {
      unsigned long long a = __shfl_xor_sync(0xffffffffu, lo[k], o), b = __shfl_xor_sync(0xffffffffu, hi[k], o);
      lo[k] = a < lo[k] ? a : lo[k];
      hi[k] = b > hi[k] ? b : hi[k];
    }
    if (lane == 0) { slo[k][warp] = lo[k]; shi[k][warp] = hi[k]; }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    const int k = threadIdx.x % 3;
    const bool is_hi = threadIdx.x >= 3;
    const int nw = (blockDim.x + 31) >> 5;
    unsigned long long v = is_hi ? shi[k][0] : slo[k][0];
    for (int w = 1; w < nw; ++w) {
      const unsigned long long x = is_hi ? shi[k][w] : slo[k][w];
      v = is_hi ? (x > v ? x : v) : (x < v ? x : v);
    }
    if (is_hi) atomicMax(&mm[3 + k], v); else atomicMin(&mm[k], v);
  }
}
// Per-axis histograms (AXIS_BINS bins over [lo, lo + AXIS_BINS*w)) for the trimmed grid region.
#define AXIS_BINS 1024
struct AxisRange { double lo[3], inv_w[3]; };
__global__ void __launch_bounds__(256)
k_axis_hist(const double* __restrict__ pos3, uint64_t n, AxisRange ar, uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[3 * AXIS_BINS];
  for (int i = threadIdx.x; i < 3 * AXIS_BINS; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    for (int k = 0; k < 3; ++k) {
      double b = floor((pos3[i * 3 + k] - ar.lo[k]) * ar.inv_w[k]);
      int bi = b < 0.0 ? 0 : (b > (double)(AXIS_BINS - 1) ? AXIS_BINS - 1 : (int)b);
      atomicAdd(&sh[k * AXIS_BINS + bi], 1u);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * AXIS_BINS; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
// cell coordinate, clamped into the grid (see the note on the trimmed region in do_map_build)
__device__ __forceinline__ int cell_coord(const Grid& g, double p, int ax) {
  const int nmax = (ax == 0 ? g.nx : (ax == 1 ? g.ny : g.nz)) - 1;
  double f = floor((p - g.org[ax]) * g.inv_cell);
  return f < 0.0 ? 0 : (f > (double)nmax ? nmax : (int)f);   // NaN -> 0
}
// sort key = (cell << tag_bits) | tag: photons ordered by cell, then by (photon index, depth) -> the map
// is bit-reproducible whatever order the tracing atomics produced.  tag_bits = bits of the largest tag
// of this photon set (24 for 1 M photons), so the radix sort runs over tag_bits + cell bits only.
__global__ void k_cell_key(Grid g, const double* __restrict__ pos3, const uint64_t* __restrict__ tag, uint64_t n, int tag_bits,
                           uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ hist) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int cx = cell_coord(g, pos3[i * 3], 0), cy = cell_coord(g, pos3[i * 3 + 1], 1), cz = cell_coord(g, pos3[i * 3 + 2], 2);
  uint32_t c = ((uint32_t)cz * (uint32_t)g.ny + (uint32_t)cy) * (uint32_t)g.nx + (uint32_t)cx;
  keys[i] = ((uint64_t)c << tag_bits) | (tag[i] & ((1ull << tag_bits) - 1ull));
  vals[i] = (uint32_t)i;
  atomicAdd(&hist[c], 1u);
}
// Sorted photon map, laid out for 16-byte vector loads: per photon two double2
// for (px, py | pz, wavelength-bits) and two for (dx, dy | dz, 0).  64 B physical
// per photon (49 B of information).
struct MapSoA {
  double2* P;       // [n][2]
  double2* D;       // [n][2]
  uint32_t* orig;   // index in the unsorted (import/export) order
};
__global__ void k_scatter(RecBuf rec, const uint32_t* __restrict__ vals, uint64_t n, MapSoA m) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s = vals[i];
  const double* p = rec.pos3 + (uint64_t)s * 3;
  const double* d = rec.dir3 + (uint64_t)s * 3;
  m.P[i * 2] = make_double2(p[0], p[1]);
  m.P[i * 2 + 1] = make_double2(p[2], __longlong_as_double((long long)rec.wl[s]));
  m.D[i * 2] = make_double2(d[0], d[1]);
  m.D[i * 2 + 1] = make_double2(d[2], 0.0);
  m.orig[i] = s;
}

#endif
