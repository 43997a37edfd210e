// kernels_map.cuh -- photon map and query ordering without host round trips:
//   (sizes come from PassDev, the device-resident state of the pass: pass_state.cuh)
//   * compact cell index: one occupancy bit per grid cell + a prefix count per 64-cell word (16 B per word,
//     L2 resident at every radius) + start offsets of the OCCUPIED cells only -- replaces the dense
//     cell_start table (105 MB at r = 0.019) that had to be cleared and scanned every pass
//   * device-length exclusive scans (3 phases, no library)
//   * photon sort = tag order for free (per-photon depth masks + one scan) followed by a hand-written
//     stable LSD radix sort over the compact cell rank (11-bit digits): the map is ordered by
//     (cell, photon index, depth) whatever order the tracing atomics produced -> bit-reproducible
//   * query sort = counting sort by cell with one atomic per query (the order inside a cell is free:
//     every query's sum runs in map order whatever lane it sits in)
// Replaces build_photonmap (photonmap.rs:23-29: N kd-tree inserts).  Part of the single translation unit
// engine.cu (compiled -fmad=false, sm_100a); see DESIGN.md sections 5-7.
#ifndef PPM_KERNELS_MAP_CUH_
#define PPM_KERNELS_MAP_CUH_

#include "dev_core.cuh"
#include "pass_state.cuh"
#include "kernels_photon.cuh"   // RecBuf

#include <cstring>

// ---- bounding box / trimmed region of a photon set (probe entry points and the per-scene calibration) -----------
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

// ---- compact cell index ---------------------------------------------------------------------------------------
// cell coordinate, clamped into the grid
__device__ __forceinline__ int cell_coord(const Grid& g, double p, int ax) {
  const int nmax = (ax == 0 ? g.nx : (ax == 1 ? g.ny : g.nz)) - 1;
  double f = floor((p - g.org[ax]) * g.inv_cell);
  return f < 0.0 ? 0 : (f > (double)nmax ? nmax : (int)f);   // NaN -> 0
}
__device__ __forceinline__ uint32_t cell_of(const Grid& g, double x, double y, double z) {
  const int cx = cell_coord(g, x, 0), cy = cell_coord(g, y, 1), cz = cell_coord(g, z, 2);
  return ((uint32_t)cz * (uint32_t)g.ny + (uint32_t)cy) * (uint32_t)g.nx + (uint32_t)cx;
}
// One word per 64 consecutive cells: occupancy bits + number of occupied cells in all earlier words.  16 bytes so
// that ONE vector load answers "how many occupied cells lie before cell c" (the rank of c).
struct __align__(16) IdxWord {
  unsigned long long bits;
  uint32_t prefix;
  uint32_t _pad;
};
struct CellIndex {
  IdxWord* words;      // [(ncells >> 6) + 1] (the last word is a sentinel whose prefix = occupied cells)
  uint32_t* start;     // [occupied + 1]: first element of the k-th occupied cell; start[occupied] = n
};
__device__ __forceinline__ uint32_t cell_rank(const IdxWord* __restrict__ words, uint32_t c) {
  const uint4 w = __ldg(reinterpret_cast<const uint4*>(words + (c >> 6)));
  const unsigned long long bits = ((unsigned long long)w.y << 32) | (unsigned long long)w.x;
  return w.z + (uint32_t)__popcll(bits & ((1ull << (c & 63u)) - 1ull));
}
// number of elements in cells < c  (c in [0, ncells])
__device__ __forceinline__ uint32_t cell_begin(const CellIndex& ix, uint32_t c) { return __ldg(ix.start + cell_rank(ix.words, c)); }

// occupancy bit of cell c.  Millions of points share a few thousand words: the bit is almost always set already, and a
// plain (L2) load answers that without the serialised same-address atomic (k_query_mark: 0.26 -> 0.03 ms at 1080p).
__device__ __forceinline__ void mark_cell(IdxWord* __restrict__ words, uint32_t c) {
  const unsigned long long bit = 1ull << (c & 63u);
  unsigned long long* w = &words[c >> 6].bits;
  if (!(__ldcg(w) & bit)) atomicOr(w, bit);
}
__device__ __forceinline__ uint32_t idx_nwords(const PassDev* ps) { return (ps->grid.ncells >> 6) + 1u; }
__global__ void k_index_clear(const PassDev* __restrict__ ps, IdxWord* __restrict__ words) {
  const uint32_t nw = idx_nwords(ps);
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += gridDim.x * blockDim.x) reinterpret_cast<uint4*>(words)[i] = z;
}

// ---- exclusive scans whose length lives on the device (3 phases: tile totals, scan of the totals, apply) -------
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total) {   // 256 threads
  __shared__ uint32_t wsum[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < SCAN_THREADS / 32; ++w) {
    const uint32_t s = wsum[w];
    if (w < warp) base += s;
    tot += s;
  }
  __syncthreads();
  *total = tot;
  return base + inc - v;
}
template <class Op>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(Op op, uint32_t* __restrict__ tile_sums) {
  const uint32_t n = op.n();
  const uint32_t t0 = blockIdx.x * SCAN_TILE;
  if (t0 >= n) return;
  uint32_t s = 0;
  const uint32_t i0 = t0 + threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k)
    if (i0 + k < n) s += op.load(i0 + k);
  uint32_t tot;
  block_excl_scan(s, &tot);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}
template <class Op>
__global__ void __launch_bounds__(1024) k_scan_sums(Op op, uint32_t* __restrict__ tile_sums) {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t carry_s;
  const uint32_t n = op.n();
  const uint32_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < ntiles; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < ntiles ? tile_sums[i] : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t wb = 0, tot = 0;
    for (int w = 0; w < 32; ++w) {
      const uint32_t s = wsum[w];
      if (w < warp) wb += s;
      tot += s;
    }
    const uint32_t carry = carry_s;
    if (i < ntiles) tile_sums[i] = carry + wb + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) op.finish(carry_s);
}
template <class Op>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(Op op, const uint32_t* __restrict__ tile_sums) {
  const uint32_t n = op.n();
  const uint32_t t0 = blockIdx.x * SCAN_TILE;
  if (t0 >= n) return;
  uint32_t v[SCAN_ITEMS];
  uint32_t s = 0;
  const uint32_t i0 = t0 + threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = i0 + k < n ? op.load(i0 + k) : 0u;
    s += v[k];
  }
  uint32_t tot;
  uint32_t run = tile_sums[blockIdx.x] + block_excl_scan(s, &tot);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (i0 + k < n) op.store(i0 + k, run, v[k]);
    run += v[k];
  }
}

// scan of the index words: prefix = occupied cells before the word; the occupied total goes to *n_occ.
// `zero` (optional): the per-occupied-cell counters of the word's cells are cleared on the way (query counting sort).
struct ScanWords {
  const PassDev* ps;
  IdxWord* words;
  uint32_t* n_occ;
  uint32_t* zero;
  __device__ uint32_t n() const { return idx_nwords(ps); }
  __device__ uint32_t load(uint32_t i) const { return (uint32_t)__popcll(words[i].bits); }
  __device__ void store(uint32_t i, uint32_t excl, uint32_t cnt) const {
    words[i].prefix = excl;
    if (zero) for (uint32_t k = 0; k < cnt; ++k) zero[excl + k] = 0u;
  }
  __device__ void finish(uint32_t total) const { *n_occ = total; }
};
// plain u32 array with a device-side length (+1 slot: out[n] = total)
struct ScanU32 {
  const uint32_t* len;
  const uint32_t* in;
  uint32_t* out;
  __device__ uint32_t n() const { return *len; }
  __device__ uint32_t load(uint32_t i) const { return in[i]; }
  __device__ void store(uint32_t i, uint32_t excl, uint32_t) const { out[i] = excl; }
  __device__ void finish(uint32_t total) const { out[*len] = total; }
};
// fixed-length in-place u32 scan (radix-sort offset tables)
struct ScanFixed {
  uint32_t* a;
  uint32_t len;
  __device__ uint32_t n() const { return len; }
  __device__ uint32_t load(uint32_t i) const { return a[i]; }
  __device__ void store(uint32_t i, uint32_t excl, uint32_t) const { a[i] = excl; }
  __device__ void finish(uint32_t) const {}
};
// Per-photon depth masks -> first tag-ordered slot of every photon.  The tracer leaves, per emitted photon, the set
// of depths at which it stored a record; record (photon, depth) then sits at base[photon] + popc(mask below depth)
// in (photon, depth) order: the sort by tag costs one scan.  finish() validates the pass' record count.
struct ScanMasks {
  PassDev* ps;
  const uint32_t* mask;
  uint32_t* base;
  uint32_t nphoton;
  uint32_t cap;
  __device__ uint32_t n() const { return nphoton; }
  __device__ uint32_t load(uint32_t i) const { return (uint32_t)__popc(mask[i]); }
  __device__ void store(uint32_t i, uint32_t excl, uint32_t) const { base[i] = excl; }
  __device__ void finish(uint32_t total) const {
    if (ps->n_rec > (unsigned long long)cap || (unsigned long long)total != ps->n_rec) { atomicOr(&ps->status, PPM_ST_REC_OVERFLOW); ps->n_map = 0u; }
    else ps->n_map = total;
  }
};

// ---- photon side ------------------------------------------------------------------------------------------------
// Sorted photon map, laid out for 16-byte vector loads: per photon two double2 for (px, py | pz, wavelength-bits)
// and two for (dx, dy | dz, 0).  64 B physical per photon (49 B of information).
struct MapSoA {
  double2* P;       // [n][2]
  double2* D;       // [n][2]
  uint32_t* orig;   // index in the unsorted (import/export) order
};
// per record: tag-ordered position -> order0, grid cell -> cid, occupancy bit.  IDENTITY: imported photons are
// already in tag order (tag = index << 4).
template <bool IDENTITY>
__global__ void k_photon_place(PassDev* ps, RecBuf rec, const uint32_t* __restrict__ mask, const uint32_t* __restrict__ base,
                               uint32_t* __restrict__ order0, uint32_t* __restrict__ cid, IdxWord* __restrict__ words, int stamp_slot) {
  stamp(ps, stamp_slot);
  const uint32_t n = ps->n_map;
  const Grid g = ps->grid;
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    uint32_t tp = s;
    if (!IDENTITY) {
      const uint64_t tag = rec.tag[s];
      const uint32_t idx = (uint32_t)(tag >> 4), depth = (uint32_t)(tag & 15u);
      tp = base[idx] + (uint32_t)__popc(mask[idx] & ((1u << depth) - 1u));
    }
    if (tp < n) order0[tp] = s;
    const uint32_t c = cell_of(g, rec.pos3[(uint64_t)s * 3], rec.pos3[(uint64_t)s * 3 + 1], rec.pos3[(uint64_t)s * 3 + 2]);
    cid[s] = c;
    mark_cell(words, c);
  }
}

// Stable LSD radix sort, RS_BITS-bit digits.  RS_BLOCKS blocks each own one contiguous tile of the input; a pass is
// (digit histogram per tile) -> (scan of the digit-major [digit][tile] table) -> (stable scatter).
#define RS_BITS 11
#define RS_DIGITS (1 << RS_BITS)
#define RS_THREADS 256
__device__ __forceinline__ void rs_tile(uint32_t n, uint32_t nblocks, uint32_t b, uint32_t& t0, uint32_t& t1) {
  uint32_t tile = (n + nblocks - 1) / nblocks;
  tile = (tile + RS_THREADS - 1) / RS_THREADS * RS_THREADS;
  const unsigned long long a = (unsigned long long)tile * b, e = a + tile;
  t0 = a < n ? (uint32_t)a : n;
  t1 = e < n ? (uint32_t)e : n;
}
// FIRST: the keys are made here: key = compact rank of the photon's cell, value = its record slot, in tag order
template <bool FIRST>
__global__ void __launch_bounds__(RS_THREADS)
k_rs_hist(const PassDev* __restrict__ ps, const uint32_t* __restrict__ keys_in, int shift, uint32_t* __restrict__ table,
          const uint32_t* __restrict__ order0, const uint32_t* __restrict__ cid, const IdxWord* __restrict__ words,
          uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
  __shared__ uint32_t hist[RS_DIGITS];
  for (int d = threadIdx.x; d < RS_DIGITS; d += RS_THREADS) hist[d] = 0u;
  __syncthreads();
  const uint32_t n = ps->n_map;
  uint32_t t0, t1;
  rs_tile(n, gridDim.x, blockIdx.x, t0, t1);
  for (uint32_t i = t0 + threadIdx.x; i < t1; i += RS_THREADS) {
    uint32_t key;
    if (FIRST) {
      const uint32_t s = order0[i];
      key = cell_rank(words, cid[s]);
      keys_out[i] = key; vals_out[i] = s;
    } else {
      key = keys_in[i];
    }
    atomicAdd(&hist[(key >> shift) & (RS_DIGITS - 1)], 1u);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < RS_DIGITS; d += RS_THREADS) table[(uint32_t)d * gridDim.x + blockIdx.x] = hist[d];
}
// table = exclusive scan of the histogram table: table[d][b] = first output slot of digit d of tile b
__global__ void __launch_bounds__(RS_THREADS)
k_rs_scatter(const PassDev* __restrict__ ps, const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, int shift,
             const uint32_t* __restrict__ table, uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
  __shared__ uint32_t cnt[RS_DIGITS];
  for (int d = threadIdx.x; d < RS_DIGITS; d += RS_THREADS) cnt[d] = table[(uint32_t)d * gridDim.x + blockIdx.x];
  __syncthreads();
  const uint32_t n = ps->n_map;
  uint32_t t0, t1;
  rs_tile(n, gridDim.x, blockIdx.x, t0, t1);
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (uint32_t r0 = t0; r0 < t1; r0 += RS_THREADS) {                 // uniform trip count per block
    const uint32_t i = r0 + threadIdx.x;
    const bool valid = i < t1;
    const uint32_t key = valid ? keys_in[i] : 0u;
    const uint32_t val = valid ? vals_in[i] : 0u;
    const uint32_t d = valid ? ((key >> shift) & (RS_DIGITS - 1)) : 0xFFFFFFFFu;
    uint32_t dst = 0;
    // the warps of the block take their turns in order: element order inside the tile is preserved (stable)
    for (unsigned w = 0; w < RS_THREADS / 32; ++w) {
      if (warp == w) {
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned before = __popc(peers & lt_mask);
        uint32_t base = 0;
        if (valid) base = cnt[d];
        __syncwarp();
        if (valid) {
          dst = base + before;
          if (before == 0u) cnt[d] = base + __popc(peers);
        }
      }
      __syncthreads();
    }
    if (valid) { keys_out[dst] = key; vals_out[dst] = val; }
  }
}
// final order -> SoA map + start offsets of the occupied cells (a cell starts where the rank changes)
__global__ void k_map_scatter(PassDev* ps, RecBuf rec, const uint32_t* __restrict__ rank, const uint32_t* __restrict__ slot, MapSoA m,
                              uint32_t* __restrict__ start) {
  const uint32_t n = ps->n_map;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t s = slot[i];
    const double* p = rec.pos3 + (uint64_t)s * 3;
    const double* d = rec.dir3 + (uint64_t)s * 3;
    m.P[(uint64_t)i * 2] = make_double2(p[0], p[1]);
    m.P[(uint64_t)i * 2 + 1] = make_double2(p[2], __longlong_as_double((long long)rec.wl[s]));
    m.D[(uint64_t)i * 2] = make_double2(d[0], d[1]);
    m.D[(uint64_t)i * 2 + 1] = make_double2(d[2], 0.0);
    m.orig[i] = s;
    const uint32_t r = rank[i];
    if (i == 0u || rank[i - 1] != r) start[r] = i;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) start[ps->n_occ_p] = n;
}

// ---- query side: counting sort by cell ---------------------------------------------------------------------------
#define MAP_UNROLL 4
__global__ void k_query_mark(PassDev* ps, const double* __restrict__ qpos3, uint32_t cap, uint32_t* __restrict__ qcell,
                             IdxWord* __restrict__ words, int stamp_slot) {
  stamp(ps, stamp_slot);
  const unsigned long long made = ps->n_nodes;
  const uint32_t n = made > (unsigned long long)cap ? cap : (uint32_t)made;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ps->n_query = n;
    if (made > (unsigned long long)cap) atomicOr(&ps->status, PPM_ST_NODE_OVERFLOW);
  }
  const Grid g = ps->grid;
  // the kernel is a chain of memory round trips with a handful of instructions in between: four queries per thread
  // and trip, all loads issued before the first use
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += MAP_UNROLL * stride) {
    double x[MAP_UNROLL], y[MAP_UNROLL], z[MAP_UNROLL];
#pragma unroll
    for (int k = 0; k < MAP_UNROLL; ++k) {
      const uint32_t i = i0 + (uint32_t)k * stride;
      if (i < n) { x[k] = qpos3[(uint64_t)i * 3]; y[k] = qpos3[(uint64_t)i * 3 + 1]; z[k] = qpos3[(uint64_t)i * 3 + 2]; }
    }
#pragma unroll
    for (int k = 0; k < MAP_UNROLL; ++k) {
      const uint32_t i = i0 + (uint32_t)k * stride;
      if (i < n) {
        const uint32_t c = cell_of(g, x[k], y[k], z[k]);
        qcell[i] = c;
        mark_cell(words, c);
      }
    }
  }
}
// one atomic per query: its position inside its cell (any order: a query's sum runs in map order wherever it sits)
__global__ void k_query_count(const PassDev* __restrict__ ps, const uint32_t* __restrict__ qcell, const IdxWord* __restrict__ words,
                              uint32_t* __restrict__ cnt, uint32_t* __restrict__ qrank, uint32_t* __restrict__ qpos_in_cell) {
  const uint32_t n = ps->n_query;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += MAP_UNROLL * stride) {
   uint32_t rk[MAP_UNROLL];
#pragma unroll
   for (int k = 0; k < MAP_UNROLL; ++k) {               // the look-ups of the trip's queries first (independent round trips)
     const uint32_t i = i0 + (uint32_t)k * stride;
     rk[k] = i < n ? cell_rank(words, qcell[i]) : 0u;
   }
#pragma unroll
   for (int k = 0; k < MAP_UNROLL; ++k) {
    const uint32_t i = i0 + (uint32_t)k * stride;
    if (i >= n) continue;
    const uint32_t r = rk[k];
    qrank[i] = r;
    // neighbouring nodes come from neighbouring pixels and mostly share a cell: one atomic per distinct cell of the warp
    const unsigned act = __activemask();
    const unsigned peers = __match_any_sync(act, r);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if ((int)(threadIdx.x & 31u) == leader) base = atomicAdd(&cnt[r], (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    qpos_in_cell[i] = base + (uint32_t)__popc(peers & ((1u << (threadIdx.x & 31u)) - 1u));
   }
  }
}
__global__ void k_query_scatter(const PassDev* __restrict__ ps, const uint32_t* __restrict__ qcell, const uint32_t* __restrict__ qrank,
                                const uint32_t* __restrict__ qpos_in_cell, const uint32_t* __restrict__ start,
                                uint32_t* __restrict__ skey, uint32_t* __restrict__ sidx) {
  const uint32_t n = ps->n_query;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += MAP_UNROLL * stride) {
    uint32_t r[MAP_UNROLL], pic[MAP_UNROLL], cell[MAP_UNROLL], st[MAP_UNROLL];
#pragma unroll
    for (int k = 0; k < MAP_UNROLL; ++k) {
      const uint32_t i = i0 + (uint32_t)k * stride;
      if (i < n) { r[k] = qrank[i]; pic[k] = qpos_in_cell[i]; cell[k] = qcell[i]; }
    }
#pragma unroll
    for (int k = 0; k < MAP_UNROLL; ++k) {
      const uint32_t i = i0 + (uint32_t)k * stride;
      if (i < n) st[k] = __ldg(start + r[k]);
    }
#pragma unroll
    for (int k = 0; k < MAP_UNROLL; ++k) {
      const uint32_t i = i0 + (uint32_t)k * stride;
      if (i < n) { const uint32_t s = st[k] + pic[k]; skey[s] = cell[k]; sidx[s] = i; }
    }
  }
}

#endif
