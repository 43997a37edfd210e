// kernels_gather.cuh -- radiance estimate: query keys, warp-cooperative gather, k-NN radius search, within probe
// Part of the single translation unit engine.cu (compiled -fmad=false, sm_100a); see DESIGN.md section 6.
#ifndef PPM_KERNELS_GATHER_CUH_
#define PPM_KERNELS_GATHER_CUH_

#include "dev_core.cuh"
#include "kernels_map.cuh"      // Grid, MapSoA, cell_coord

// ---- gather -------------------------------------------------------------------
// tracer.rs:198-216
__device__ __forceinline__ double filter_cone(double d, double rmax) {
  const double K_CONE = 1.1;
  const double FAC_K = 1.0 - 2.0 / (3.0 * K_CONE);
  double d2 = sqrt(d / rmax) / K_CONE;
  return d2 > 1.0 ? 0.0 : (1.0 - d2) / FAC_K;
}
__device__ __forceinline__ double filter_gauss(double d, double rmax) {
  const double ALPHA = 0.918, BETA = 1.953, E_BETA = 1.0 - 0.14184788965323, CORR = 0.5;
  double e_r = 1.0 - exp(-BETA * d / (rmax * 2.0));
  return e_r > E_BETA ? 0.0 : ALPHA * (1.0 - e_r / E_BETA) + CORR;
}

// Queries are keyed by their (padded) grid cell and radix sorted, so that the 32
// lanes of a warp hold queries of the same cell (or of a few cells).
__global__ void k_query_key(Grid g, const double* __restrict__ qpos3, int64_t n, uint32_t* __restrict__ keys,
                            uint32_t* __restrict__ vals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int cx = cell_coord(g, qpos3[i * 3], 0), cy = cell_coord(g, qpos3[i * 3 + 1], 1), cz = cell_coord(g, qpos3[i * 3 + 2], 2);
  uint32_t key = ((uint32_t)cz * (uint32_t)g.ny + (uint32_t)cy) * (uint32_t)g.nx + (uint32_t)cx;
  keys[i] = key;
  vals[i] = (uint32_t)i;
}

// v2: warp-cooperative gather.  A warp owns 32 cell-sorted queries (one per lane).
// For each distinct cell among them, the photons of the (2R+1)^3 neighbourhood -- (2R+1)^2
// x-contiguous runs of the sorted map, R = 1 (cell edge r: 9 runs of 3 cells) or R = 2 (cell
// edge r/2: 25 runs of 5 cells, 30 % fewer candidates) -- form one virtual candidate stream;
// 32 candidates at a time are fetched with coalesced 16-byte loads, staged in shared memory,
// and every lane tests the SAME photon (broadcast LDS.128) against its own query -- no
// per-lane loop lengths, no scattered global loads.
#define GATHER_WARPS 4
#define GATHER_SPAN 3        // a group may span cells cx .. cx+3 of one row
// MODE 0: fixed radius r2 (estimate_radiance).  MODE 1: per-query squared radius r2q[] (k-NN estimate:
// membership, filter rmax and normaliser all use the query's own radius).  MODE 2: count only,
// members are d2 <= r2q[] (the bisection steps of the k-NN radius search).
template <int FILTER, int MODE>
__global__ void __launch_bounds__(GATHER_WARPS * 32)
k_gather(Grid g, const uint32_t* __restrict__ cell_start, MapSoA m, const uint32_t* __restrict__ qkey,
         const uint32_t* __restrict__ qidx, const double* __restrict__ qpos3, const double* __restrict__ qnrm3, int64_t n,
         double power, double r2_fixed, const double* __restrict__ r2q, double* __restrict__ rgb3, uint32_t* __restrict__ counts,
         unsigned long long* __restrict__ sum_k) {
  constexpr int REACH = 1;
  __shared__ double2 sP[GATHER_WARPS][32][2];
  __shared__ double2 sD[GATHER_WARPS][32][2];
  __shared__ uint32_t sEnd[GATHER_WARPS][32], sOff[GATHER_WARPS][32];   // per run: cumulative end, start - exclusive prefix
  constexpr int W = 2 * REACH + 1, ROWS = W * W;
  static_assert(ROWS <= 32, "one lane per run");
  const unsigned FULL = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t s = ((int64_t)blockIdx.x * GATHER_WARPS + warp) * 32 + lane;
  const bool valid = s < n;
  const uint32_t key = valid ? qkey[s] : 0xFFFFFFFFu;
  const uint32_t qi = valid ? qidx[s] : 0u;
  double qx = 0.0, qy = 0.0, qz = 0.0;
  D3 nv = mk3(0.0, 0.0, 0.0);
  double r2 = r2_fixed;
  if (valid) {
    qx = qpos3[(uint64_t)qi * 3]; qy = qpos3[(uint64_t)qi * 3 + 1]; qz = qpos3[(uint64_t)qi * 3 + 2];
    if (MODE != 2) nv = ld3(qnrm3 + (uint64_t)qi * 3);
    if (MODE != 0) r2 = r2q[qi];
  }
  double rr = 0.0, rg = 0.0, rb = 0.0;
  uint32_t cnt = 0;
  const uint32_t nxp = (uint32_t)g.nx, nyp = (uint32_t)g.ny;
  unsigned pending = __ballot_sync(FULL, valid);
  while (pending) {
    const int leader = __ffs(pending) - 1;
    const uint32_t ck = __shfl_sync(FULL, key, leader);
    // Group = the pending lanes whose cell lies in the leader's row (same cy, cz) at most
    // GATHER_SPAN cells to the right of the leader's cell (keys are sorted, x fastest).  They
    // share ONE candidate stream covering [cx_leader - R, cx_last + R]: a superset of every
    // lane's own neighbourhood, so the extra candidates simply fail the distance test.
    const bool act = valid && key >= ck && key - ck <= (uint32_t)GATHER_SPAN && key / nxp == ck / nxp;
    const unsigned grp = __ballot_sync(FULL, act);
    pending &= ~grp;
    const uint32_t klast = __shfl_sync(FULL, key, 31 - __clz((int)grp));
    const int cx = (int)(ck % nxp), cy = (int)((ck / nxp) % nyp), cz = (int)(ck / (nxp * nyp));
    const int x0 = max(cx - REACH, 0), x1 = min((int)(klast % nxp) + REACH, g.nx - 1);
    // lane l < ROWS looks up run l = (dz, dy) of the neighbourhood
    uint32_t rbeg = 0, rlen = 0;
    if (lane < ROWS && x0 <= x1) {
      const int z = cz + lane / W - REACH, y = cy + lane % W - REACH;
      if (z >= 0 && z < g.nz && y >= 0 && y < g.ny) {
        const uint32_t row = ((uint32_t)z * (uint32_t)g.ny + (uint32_t)y) * (uint32_t)g.nx;
        rbeg = cell_start[row + x0];
        rlen = cell_start[row + x1 + 1] - rbeg;
      }
    }
    uint32_t pre = rlen;                               // inclusive prefix of the run lengths
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(FULL, pre, o);
      if (lane >= o) pre += t;
    }
    const uint32_t total = __shfl_sync(FULL, pre, 31);
    __syncwarp();
    sEnd[warp][lane] = pre;
    sOff[warp][lane] = rbeg - (pre - rlen);
    __syncwarp();
    for (uint32_t base = 0; base < total; base += 32) {
      const uint32_t v = base + lane;
      if (v < total) {
        int run = 0;                                   // number of runs that end at or before v (binary search)
#pragma unroll
        for (int step = 16; step > 0; step >>= 1)
          if (sEnd[warp][run + step - 1] <= v) run += step;
        const uint32_t o = sOff[warp][run];
        const uint64_t j = (uint64_t)(v + o) * 2;
        sP[warp][lane][0] = m.P[j]; sP[warp][lane][1] = m.P[j + 1];
        if (MODE != 2) { sD[warp][lane][0] = m.D[j]; sD[warp][lane][1] = m.D[j + 1]; }
      }
      __syncwarp();
      const int mcount = (int)min(32u, total - base);
      if (act) {
        for (int t = 0; t < mcount; ++t) {
          const double2 a = sP[warp][t][0], b = sP[warp][t][1];
          // squared_euclidean: ((qx-px)^2 + (qy-py)^2) + (qz-pz)^2, member iff d2 <= r2
          const double ax = qx - a.x, ay = qy - a.y, az = qz - b.x;
          const double d2 = (ax * ax + ay * ay) + az * az;
          if (d2 <= r2) {
            ++cnt;
            if (MODE == 2) continue;
            const double wt = FILTER == PPM_FILTER_NONE ? 1.0 : (FILTER == PPM_FILTER_CONE ? filter_cone(d2, r2) : filter_gauss(d2, r2));
            const double2 c = sD[warp][t][0], d = sD[warp][t][1];
            // photon_to_radiance, optics.rs:224-233
            const double cos0 = (nv.x * c.x + nv.y * c.y) + nv.z * d.x;
            const double pw2 = cos0 < 0.0 ? (wt * power) * -cos0 : 0.0;
            const int w = (int)__double_as_longlong(b.y);
            if (w == 0) rr = rr + pw2; else if (w == 1) rg = rg + pw2; else rb = rb + pw2;
          }
        }
      }
      __syncwarp();
    }
  }
  if (valid) {
    if (MODE != 2) {
      const double sc = (1.0 / PPM_PI) / r2;          // rad * (ONE_PI / radius), tracer.rs:193
      rgb3[(uint64_t)qi * 3] = rr * sc; rgb3[(uint64_t)qi * 3 + 1] = rg * sc; rgb3[(uint64_t)qi * 3 + 2] = rb * sc;
    }
    if (counts) counts[qi] = cnt;
  }
  if (sum_k) {
    unsigned long long c = cnt;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
    if (lane == 0 && c) atomicAdd(sum_k, c);
  }
}

// ---- k-NN radius search: exact k-th smallest d2 per query by bisection on the bit pattern of d2 --------
// (non-negative doubles order like their bit patterns).  State per query: [lo, hi] as uint64 bits,
// hi always satisfies count(d2 <= hi) >= k.  done[] = 1 when fewer than k photons lie within r (fixed radius).
__global__ void k_knn_init(int64_t n, double r2, const uint32_t* __restrict__ cnt, uint32_t k, unsigned long long* __restrict__ lo,
                           unsigned long long* __restrict__ hi, double* __restrict__ thr, unsigned int* __restrict__ n_active) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long rb = (unsigned long long)__double_as_longlong(r2);
  if (cnt[i] < k) { lo[i] = hi[i] = rb; thr[i] = r2; return; }        // fewer than k within r: fixed radius
  lo[i] = 0ull; hi[i] = rb;
  thr[i] = __longlong_as_double((long long)(rb >> 1));
  atomicAdd(n_active, 1u);
}
__global__ void k_knn_step(int64_t n, const uint32_t* __restrict__ cnt, uint32_t k, unsigned long long* __restrict__ lo,
                           unsigned long long* __restrict__ hi, double* __restrict__ thr, unsigned int* __restrict__ n_active) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long l = lo[i], h = hi[i];
  if (l >= h) return;
  const unsigned long long mid = l + ((h - l) >> 1);                   // thr[i] was asdouble(mid)
  if (cnt[i] >= k) h = mid; else l = mid + 1;
  lo[i] = l; hi[i] = h;
  if (l < h) { thr[i] = __longlong_as_double((long long)(l + ((h - l) >> 1))); atomicAdd(n_active, 1u); }
  else thr[i] = __longlong_as_double((long long)h);
}
// a k-th distance of exactly zero (k coincident photons at the query) cannot normalise: fall back to r2
__global__ void k_knn_finish(int64_t n, double r2, double* __restrict__ thr) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !(thr[i] > 0.0)) thr[i] = r2;
}

__global__ void k_within(Grid g, const uint32_t* __restrict__ cell_start, MapSoA m, const double* __restrict__ qpos3,
                         int64_t n, double r2, uint32_t* __restrict__ idx, uint32_t* __restrict__ counts, uint32_t cap) {
  int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const double qx = qpos3[q * 3], qy = qpos3[q * 3 + 1], qz = qpos3[q * 3 + 2];
  uint32_t cnt = 0;
  int cx = cell_coord(g, qx, 0), cy = cell_coord(g, qy, 1), cz = cell_coord(g, qz, 2);
  const int R = 1;
  int x0 = max(cx - R, 0), x1 = min(cx + R, g.nx - 1);
  if (x0 <= x1)
    for (int z = max(cz - R, 0); z <= min(cz + R, g.nz - 1); ++z)
      for (int y = max(cy - R, 0); y <= min(cy + R, g.ny - 1); ++y) {
        uint32_t row = ((uint32_t)z * (uint32_t)g.ny + (uint32_t)y) * (uint32_t)g.nx;
        uint32_t b = cell_start[row + x0], e = cell_start[row + x1 + 1];
        for (uint32_t j = b; j < e; ++j) {
          const double2 a = m.P[(uint64_t)j * 2], b = m.P[(uint64_t)j * 2 + 1];
          double ax = qx - a.x, ay = qy - a.y, az = qz - b.x;
          double d2 = (ax * ax + ay * ay) + az * az;
          if (d2 <= r2) {
            if (cnt < cap) idx[(uint64_t)q * cap + cnt] = m.orig[j];
            ++cnt;
          }
        }
      }
  counts[q] = cnt;
}

#endif
