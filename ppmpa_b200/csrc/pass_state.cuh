// pass_state.cuh -- the device-resident state of one pass: grid, radius, RNG keys, counters, phase stamps.
// Every kernel on the pass path reads its inputs and sizes from here, so a whole pass is a fixed sequence of launches
// with fixed arguments (one CUDA graph per lane) and never waits for the host (ppmpa.rs:74-84 + iterator.rb:96-106).
// Part of the single translation unit engine.cu (compiled -fmad=false, sm_100a); see DESIGN.md section 6.
#ifndef PPM_PASS_STATE_CUH_
#define PPM_PASS_STATE_CUH_

#include "dev_core.cuh"

// ---- uniform grid, cell edge >= r, cells linearised x-fastest -------------------------------------------------
struct Grid {
  double org[3];
  double inv_cell;
  int32_t nx, ny, nz;
  uint32_t ncells;
  uint32_t mx, sx1, sx2, my, sy1, sy2;   // exact division of a cell id by nx / ny without a divide (grid_div)
};
// Division of a 32-bit value by an invariant divisor (Granlund-Montgomery, branch free): q = (t + ((n - t) >> s1)) >> s2
// with t = mulhi(m, n).  k_gather splits every query's cell id into coordinates; two of these replace three hardware
// integer divisions (~60 instructions) per query.
__host__ __device__ inline void grid_div_init(uint32_t d, uint32_t& m, uint32_t& s1, uint32_t& s2) {
  uint32_t l = 0;
  while ((1ull << l) < (unsigned long long)d) ++l;
  m = (uint32_t)((((1ull << l) - (unsigned long long)d) << 32) / (unsigned long long)d + 1ull);
  s1 = l < 1u ? l : 1u;
  s2 = l > 0u ? l - 1u : 0u;
}
__device__ __forceinline__ uint32_t grid_div(uint32_t n, uint32_t m, uint32_t s1, uint32_t s2) {
  const uint32_t t = __umulhi(m, n);
  return (t + ((n - t) >> s1)) >> s2;
}
struct Bounds { double lo[3], hi[3]; };
#define PPM_CELL_CAP 67108864.0          // 2^26 cells: 2^20 index words = 16 MB
#define PPM_CELL_CAP_WORDS ((1u << 20) + 2u)

// The grid of a pass: a function of the region and the radius only, evaluated identically on the host (probe entry
// points) and on the device (k_pass_begin).  Cell edge slightly > r: the 27-cell walk can never miss.  The origin is
// padded by half a cell so that surfaces bounding the photon cloud (the room's walls) sit mid-cell: the +-1 ulp noise
// of hit points on such a plane cannot straddle a cell boundary.  Photons and queries outside the region are CLAMPED
// into the boundary cells; clamping is monotone and non-expansive, so two points within r still differ by at most one
// cell per axis and the walk stays exact for ANY region.
__host__ __device__ inline Grid make_grid(const Bounds& b, int have_bounds, double radius2) {
  Grid g;
  double cell = sqrt(radius2) * (1.0 + 1.0 / 1024.0);
  g.org[0] = g.org[1] = g.org[2] = 0.0;
  g.nx = g.ny = g.nz = 1; g.inv_cell = 1.0 / cell; g.ncells = 1;
  grid_div_init(1u, g.mx, g.sx1, g.sx2); grid_div_init(1u, g.my, g.sy1, g.sy2);
  if (!have_bounds) return g;
  for (int it = 0; it < 64; ++it) {
    double dims[3];
    for (int k = 0; k < 3; ++k) dims[k] = floor((b.hi[k] - (b.lo[k] - 0.5 * cell)) / cell) + 2.0;
    if (dims[0] * dims[1] * dims[2] <= PPM_CELL_CAP && dims[0] < 2e9 && dims[1] < 2e9 && dims[2] < 2e9 &&
        dims[0] >= 1.0 && dims[1] >= 1.0 && dims[2] >= 1.0) {
      g.nx = (int32_t)dims[0]; g.ny = (int32_t)dims[1]; g.nz = (int32_t)dims[2];
      break;
    }
    cell *= 2.0;
  }
  for (int k = 0; k < 3; ++k) g.org[k] = b.lo[k] - 0.5 * cell;
  g.inv_cell = 1.0 / cell;
  g.ncells = (uint32_t)g.nx * (uint32_t)g.ny * (uint32_t)g.nz;
  grid_div_init((uint32_t)g.nx, g.mx, g.sx1, g.sx2); grid_div_init((uint32_t)g.ny, g.my, g.sy1, g.sy2);
  return g;
}

// ---- device-resident pass state -----------------------------------------------------------------------------
// One per lane.  k_pass_begin fills the inputs from the batch table and zeroes the counters; every later kernel of
// the pass reads its sizes from here, so nothing on the pass path waits for the host.
#define PPM_ST_REC_OVERFLOW 1u           // more photon records than the record buffers hold
#define PPM_ST_NODE_OVERFLOW 2u          // more gather nodes than the node pool holds
#define PPM_NSTAMP 16
#define PPM_NSLOT 64
struct PassDev {
  Grid grid;
  Bounds bounds;
  int32_t have_bounds;
  uint32_t pass;                         // Philox pass id of this pass
  uint64_t seed;
  double r2, power;
  double inv_pi_r2;                      // (1 / pi) / r2: the normaliser of the fixed-radius estimate (tracer.rs:193), once per pass
  unsigned long long n_rec;              // photon records appended by the tracer (may exceed the capacity)
  unsigned long long ticket;             // photon ticket of k_trace_photons
  unsigned long long n_nodes;            // gather nodes made by k_eye_expand (may exceed the capacity)
  unsigned long long n_visited;          // eye-path nodes visited
  unsigned long long sum_k;              // sum over the queries of the photons within r   (totals of the slots below,
  unsigned long long cand;               // candidate distance tests of k_gather            made by k_pass_end / the probes)
  unsigned long long sum_k_s[PPM_NSLOT]; // per-CTA-group slots: 170 k warps adding to ONE address serialise in L2
  unsigned long long cand_s[PPM_NSLOT];
  uint32_t n_map;                        // photons in the map of this pass
  uint32_t n_query;                      // gather queries of this pass (= min(n_nodes, capacity))
  uint32_t n_occ_p, n_occ_q;             // occupied cells: photons, queries
  uint32_t heavy[4];                     // [0] reservations, [1] ticket of k_gather_heavy, [2] groups, [3] parts published
  uint32_t status;                       // PPM_ST_* flags of this pass
  uint32_t cursor;                       // next entry of the batch table
  unsigned long long stamp[PPM_NSTAMP];  // %globaltimer at phase boundaries (ns)
};
// what a lane reports per pass (read by the host once per batch)
struct PassOut {
  unsigned long long n_rec, n_nodes, n_visited, sum_k, cand;
  uint32_t status, n_occ_p, n_occ_q, heavy_parts;
  unsigned long long stamp[PPM_NSTAMP];
};
// batch table of a lane: the passes it renders, in order
#define PPM_BATCH_MAX 1024
struct BatchDev {
  uint64_t seed;
  double power;
  uint32_t pass[PPM_BATCH_MAX];
  double r2[PPM_BATCH_MAX];
};
enum { ST_BEGIN = 0, ST_TRACE_END, ST_BUILD_END, ST_EXPAND_BEGIN, ST_EXPAND_END, ST_CLASSIFY_END, ST_QSORT_END,
       ST_DL_BEGIN, ST_DL_END, ST_GATHER_BEGIN, ST_GATHER_END, ST_COMBINE_BEGIN, ST_END };
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// first thread of a kernel: phase boundary stamp (block 0 is dispatched first)
__device__ __forceinline__ void stamp(PassDev* ps, int slot) {
  if (slot >= 0 && blockIdx.x == 0 && threadIdx.x == 0) ps->stamp[slot] = globaltimer_ns();
}
__global__ void k_stamp(PassDev* ps, int slot) { ps->stamp[slot] = globaltimer_ns(); }

__global__ void k_pass_begin(PassDev* ps, const BatchDev* __restrict__ bt) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const uint32_t i = ps->cursor % PPM_BATCH_MAX;
  ps->seed = bt->seed; ps->power = bt->power; ps->pass = bt->pass[i]; ps->r2 = bt->r2[i];
  ps->grid = make_grid(ps->bounds, ps->have_bounds, bt->r2[i]);
  ps->inv_pi_r2 = (1.0 / PPM_PI) / bt->r2[i];
  ps->n_rec = 0ull; ps->ticket = 0ull; ps->n_nodes = 0ull; ps->n_visited = 0ull; ps->sum_k = 0ull; ps->cand = 0ull;
  for (int k = 0; k < PPM_NSLOT; ++k) { ps->sum_k_s[k] = 0ull; ps->cand_s[k] = 0ull; }
  ps->n_map = 0u; ps->n_query = 0u; ps->n_occ_p = 0u; ps->n_occ_q = 0u;
  ps->heavy[0] = ps->heavy[1] = ps->heavy[2] = ps->heavy[3] = 0u;
  ps->status = 0u;
  for (int k = 0; k < PPM_NSTAMP; ++k) ps->stamp[k] = 0ull;
  ps->stamp[ST_BEGIN] = globaltimer_ns();
}
// end of a pass: report, bump the pass counter of the accumulator if the pass counted, advance the batch cursor
__global__ void k_pass_end(PassDev* ps, PassOut* __restrict__ out, double* __restrict__ npass_acc) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  ps->stamp[ST_END] = globaltimer_ns();
  for (int k = 0; k < PPM_NSLOT; ++k) { ps->sum_k += ps->sum_k_s[k]; ps->cand += ps->cand_s[k]; }
  PassOut& o = out[ps->cursor % PPM_BATCH_MAX];
  o.n_rec = ps->n_rec; o.n_nodes = ps->n_nodes; o.n_visited = ps->n_visited; o.sum_k = ps->sum_k; o.cand = ps->cand;
  o.status = ps->status; o.n_occ_p = ps->n_occ_p; o.n_occ_q = ps->n_occ_q; o.heavy_parts = ps->heavy[3];
  for (int k = 0; k < PPM_NSTAMP; ++k) o.stamp[k] = ps->stamp[k];
  if (npass_acc && ps->status == 0u) npass_acc[0] += 1.0;
  ps->cursor = ps->cursor + 1u;
}

#endif
