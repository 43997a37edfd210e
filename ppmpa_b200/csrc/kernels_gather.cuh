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

// Queries are ordered by grid cell (counting sort, kernels_map.cuh), so that the 32 lanes of a warp hold
// queries of the same cell (or of a few cells).
// v2: warp-cooperative gather.  A warp owns 32 cell-sorted queries (one per lane).
// For each distinct cell among them, the photons of the 3x3x3 neighbourhood -- nine
// x-contiguous runs of the sorted map (cell edge >= r) -- form one virtual candidate stream;
// 32 candidates at a time are fetched with coalesced 16-byte loads, staged in shared memory,
// and every lane tests the SAME photon (broadcast LDS.128) against its own query -- no
// per-lane loop lengths, no scattered global loads.
#ifndef GATHER_WARPS
#define GATHER_WARPS 4
#endif
#ifndef GATHER_YSPAN
#define GATHER_YSPAN 2       // ... and rows cy .. cy+2 of one z layer
#endif
#ifndef GATHER_SPAN
#define GATHER_SPAN 4        // a group may span cells cx .. cx+4 of one row (3: logical 0.586 -> 0.596 over the north-star schedule, 5: 0.596)
#endif

// A group whose candidate stream is longer than GATHER_HEAVY_MIN is not processed by its warp alone: the warp
// publishes it as S = ceil(total / GATHER_HEAVY_MIN) (<= 64) independent PARTS in a device-side list; part k takes the
// 32-candidate chunks k, k + S, k + 2S, ... of the stream.  k_gather_heavy (launched only when the list is not empty)
// lets every warp of the GPU claim parts one at a time, write the 32 partial sums of its part to a pool, and the warp
// that completes a group adds the S partials in part order and writes the result: deterministic whoever did which part.
// (Letting the light kernel's warps help after their own queries was tried: helpers leave whenever the list is
// momentarily empty, so the tail ran on a handful of warps -- 36 ms.)
// Work concentrated in few queries (a sunlit patch that holds most photons: 10^4..10^5 neighbours for ~1 % of the
// queries, BASELINE config 3) otherwise leaves ~1000 long-running warps for 592 schedulers (config 3 at 1024^2:
// k_gather 8.2 -> 3 ms).  Config 2 has no heavy groups from the second pass on.
#define GATHER_HEAVY_MIN 2048u
#define GATHER_HEAVY_MAXPARTS 64u
struct HeavyGroup { uint32_t s_base, grp, ck, klast, nparts, part0, done, _pad; };  // warp's first sorted query, lane mask, leader / last cell key, parts, completed parts
struct HeavyPart { uint32_t group; };
struct HeavyPartial { double rgb[3][32]; uint32_t cnt[32]; };                      // one part's partial sums, one column per lane
struct HeavyList {
  unsigned int* ctr;        // = PassDev::heavy: [0] reservation counter, [1] ticket of k_gather_heavy, [2] groups, [3] parts published
  HeavyGroup* groups;       // [cap_parts / 2]
  HeavyPart* parts;         // [cap_parts]
  HeavyPartial* partials;   // [cap_parts]
  uint32_t cap_parts;
};

// lanes 0..8 look up the nine runs of the group's neighbourhood; returns the length of the candidate stream and
// leaves, per run, its cumulative end (sEnd) and start minus exclusive prefix (sOff) in the warp's shared arrays
// (cx, cy, cz) = cell coordinates of the group's leader, xlast = x coordinate of its last cell: the callers keep the
// coordinates of their own query's cell in registers (three integer divisions per QUERY instead of five per GROUP: at
// small radius a warp forms several groups and the divisions were 11 % of the kernel's instructions)
// ylast = y coordinate of the group's last row (cy for a single-row group): the stream then covers the rows
// cy - 1 .. ylast + 1 of the three z layers, z-major, y next, x fastest -- the order of the sorted map, so every query
// still meets the photons of its own 27 cells in the same order whatever group it was served in.
__device__ __forceinline__ uint32_t gather_group_runs(const Grid& g, const CellIndex& ix, int cx, int cy, int cz, int xlast, int ylast,
                                                      int lane, uint32_t* sEnd, uint32_t* sOff) {
  constexpr int REACH = 1;
  const int W = ylast - cy + 2 * REACH + 1, ROWS = (2 * REACH + 1) * W;      // <= 32 (GATHER_YSPAN <= 7)
  const unsigned FULL = 0xffffffffu;
  const int x0 = max(cx - REACH, 0), x1 = min(xlast + REACH, g.nx - 1);
  uint32_t rbeg = 0, rlen = 0;
  if (lane < ROWS && x0 <= x1) {
    const int z = cz + lane / W - REACH, y = cy + lane % W - REACH;
    if (z >= 0 && z < g.nz && y >= 0 && y < g.ny) {
      const uint32_t row = ((uint32_t)z * (uint32_t)g.ny + (uint32_t)y) * (uint32_t)g.nx;
      rbeg = cell_begin(ix, row + (uint32_t)x0);
      rlen = cell_begin(ix, row + (uint32_t)x1 + 1u) - rbeg;
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
  sEnd[lane] = pre;
  sOff[lane] = rbeg - (pre - rlen);
  __syncwarp();
  return total;
}
// one member photon's contribution: filter weight x photon_to_radiance (optics.rs:224-233), into its wavelength's channel
template <int FILTER>
__device__ __forceinline__ void gather_accept(double d2, double r2, double2 c, double2 d, int w, D3 nv, double power,
                                              double& rr, double& rg, double& rb) {
  const double wt = FILTER == PPM_FILTER_NONE ? 1.0 : (FILTER == PPM_FILTER_CONE ? filter_cone(d2, r2) : filter_gauss(d2, r2));
  const double cos0 = (nv.x * c.x + nv.y * c.y) + nv.z * d.x;
  const double pw2 = cos0 < 0.0 ? (wt * power) * -cos0 : 0.0;
  if (w == 0) rr = rr + pw2; else if (w == 1) rg = rg + pw2; else rb = rb + pw2;
}
#ifndef GATHER_LAZYWL
#define GATHER_LAZYWL 1      // the wavelength word of an accepted candidate is read again from shared memory instead of kept
#endif
#ifndef GATHER_UNROLL
#define GATHER_UNROLL 6      // candidates per trip of the test loop (1: 0.560, 2: 0.550 ms at r = 0.0189; profiles/r2_gather_unroll.txt)
#endif
// chunks base = first, first + stride, ... of the candidate stream: stage 32 candidates, test them against the lane's query
// returns the number of candidates staged (warp-uniform)
template <int FILTER, int MODE>
__device__ __forceinline__ uint32_t gather_chunks(const MapSoA& m, uint32_t total, uint32_t first, uint32_t stride, int lane, bool act,
                                              const uint32_t* sEnd, const uint32_t* sOff, double2 (*sP)[2], double2 (*sD)[2],
                                              double qx, double qy, double qz, D3 nv, double r2, double power,
                                              double& rr, double& rg, double& rb, uint32_t& cnt) {
  uint32_t staged = 0;
  for (uint32_t base = first; base < total; base += stride) {
    const uint32_t v = base + lane;
    if (v < total) {
      int run = 0;                                   // number of runs that end at or before v (binary search)
#pragma unroll
      for (int step = 16; step > 0; step >>= 1)
        if (sEnd[run + step - 1] <= v) run += step;
      const uint32_t o = sOff[run];
      const uint64_t j = (uint64_t)(v + o) * 2;
      sP[lane][0] = m.P[j]; sP[lane][1] = m.P[j + 1];
      if (MODE != 2) { sD[lane][0] = m.D[j]; sD[lane][1] = m.D[j + 1]; }
    }
    __syncwarp();
    const int mcount = (int)min(32u, total - base);
    staged += (uint32_t)mcount;
    if (act) {
      int t = 0;
#if GATHER_UNROLL > 1
      // GATHER_UNROLL candidates per trip: their distance chains are independent (the accept blocks still run in stream
      // order, so every query adds its photons in the same order)
      for (; t + GATHER_UNROLL <= mcount; t += GATHER_UNROLL) {
        double d2u[GATHER_UNROLL];
#if !GATHER_LAZYWL
        double wlu[GATHER_UNROLL];
#endif
#pragma unroll
        for (int u = 0; u < GATHER_UNROLL; ++u) {
          const double2 a = sP[t + u][0], b = sP[t + u][1];
          const double ax = qx - a.x, ay = qy - a.y, az = qz - b.x;
          d2u[u] = (ax * ax + ay * ay) + az * az;
#if !GATHER_LAZYWL
          wlu[u] = b.y;
#endif
        }
#pragma unroll
        for (int u = 0; u < GATHER_UNROLL; ++u) {
          if (d2u[u] <= r2) {
            ++cnt;
#if GATHER_LAZYWL
            const double wl = sP[t + u][1].y;
#else
            const double wl = wlu[u];
#endif
            if (MODE != 2) gather_accept<FILTER>(d2u[u], r2, sD[t + u][0], sD[t + u][1], (int)__double_as_longlong(wl), nv, power, rr, rg, rb);
          }
        }
      }
#endif
      for (; t < mcount; ++t) {
        const double2 a = sP[t][0], b = sP[t][1];
        // squared_euclidean: ((qx-px)^2 + (qy-py)^2) + (qz-pz)^2, member iff d2 <= r2
        const double ax = qx - a.x, ay = qy - a.y, az = qz - b.x;
        const double d2 = (ax * ax + ay * ay) + az * az;
        if (d2 <= r2) {
          ++cnt;
          if (MODE == 2) continue;
          gather_accept<FILTER>(d2, r2, sD[t][0], sD[t][1], (int)__double_as_longlong(b.y), nv, power, rr, rg, rb);
        }
      }
    }
    __syncwarp();
  }
  return staged;
}
// MODE 0: fixed radius r2 (estimate_radiance).  MODE 1: per-query squared radius r2q[] (k-NN estimate:
// membership, filter rmax and normaliser all use the query's own radius).  MODE 2: count only,
// members are d2 <= r2q[] (the bisection steps of the k-NN radius search).
// The number of queries, the grid, the radius and the photon power come from the pass state.
// 56 registers without a min-blocks hint (9 CTAs per SM).  Measured: 48 registers / 10 CTAs: k_gather 0.594 -> 0.639 ms
// at r = 0.019 and 2.05 -> 2.33 ms at r = 0.1; 40 / 12: 0.670 and 2.66 ms; a hint of 1 makes ptxas take 80 registers
// (0.635 ms) -- gpurun_out/r2r_gather_minb.txt
template <int FILTER, int MODE>
#ifdef GATHER_MINB
__global__ void __launch_bounds__(GATHER_WARPS * 32, GATHER_MINB)
#else
__global__ void __launch_bounds__(GATHER_WARPS * 32)
#endif
k_gather(PassDev* ps, CellIndex ix, MapSoA m, const uint32_t* __restrict__ qkey,
         const uint32_t* __restrict__ qidx, const double* __restrict__ qpos3, const double* __restrict__ qnrm3,
         const double* __restrict__ r2q, double* __restrict__ rgb3, uint32_t* __restrict__ counts, HeavyList hl, int stamp_slot) {
  __shared__ double2 sP[GATHER_WARPS][32][2];
  __shared__ double2 sD[GATHER_WARPS][32][2];
  __shared__ uint32_t sEnd[GATHER_WARPS][32], sOff[GATHER_WARPS][32];   // per run: cumulative end, start - exclusive prefix
  stamp(ps, stamp_slot);
  const unsigned FULL = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = (int64_t)ps->n_query;
  const int64_t s_base = ((int64_t)blockIdx.x * GATHER_WARPS + warp) * 32;
  if (s_base >= n) return;                             // whole warp beyond the query list (grids are sized for the capacity)
  const Grid g = ps->grid;
  const double power = ps->power, r2_fixed = ps->r2;
  const int64_t s = s_base + lane;
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
  bool deferred = false;                               // this lane's query was handed to k_gather_heavy
  unsigned long long tests = 0;                        // (candidates staged) x (lanes of the group): distance tests done
  const uint32_t nxp = (uint32_t)g.nx, nyp = (uint32_t)g.ny;
  const uint32_t qrow = grid_div(key, g.mx, g.sx1, g.sx2);   // this query's cell: row id and coordinates, computed once
  const uint32_t qzc = grid_div(qrow, g.my, g.sy1, g.sy2);
  const int qcx = (int)(key - qrow * nxp), qcy = (int)(qrow - qzc * nyp), qcz = (int)qzc;
  unsigned pending = __ballot_sync(FULL, valid);
  while (pending) {
    const int leader = __ffs(pending) - 1;
    const uint32_t ck = __shfl_sync(FULL, key, leader);
    const int cx = __shfl_sync(FULL, qcx, leader), cy = __shfl_sync(FULL, qcy, leader), cz = __shfl_sync(FULL, qcz, leader);
    // Group = the pending lanes whose cell lies in the leader's row (same cy, cz) at most
    // GATHER_SPAN cells to the right of the leader's cell (keys are sorted, x fastest).  They
    // share ONE candidate stream covering [cx_leader - R, cx_last + R]: a superset of every
    // lane's own neighbourhood, so the extra candidates simply fail the distance test.
    // ... and, since a surface that is one cell thick in x (a wall facing x) leaves a single cell per row, the
    // cells of up to GATHER_YSPAN further rows of the same z layer in the same x range.
    // (groups are no longer contiguous key ranges: a lane served with an earlier leader must not be served again)
    bool act = ((pending >> lane) & 1u) != 0u && qcz == cz && qcy >= cy && qcy - cy <= GATHER_YSPAN && qcx >= cx && qcx - cx <= GATHER_SPAN;
    unsigned grp = __ballot_sync(FULL, act);
    const int last = 31 - __clz((int)grp);
    uint32_t klast = __shfl_sync(FULL, key, last);        // keys are sorted: != ck iff the group has more than one cell
    const int xlast = __reduce_max_sync(FULL, act ? qcx : cx), ylast = __reduce_max_sync(FULL, act ? qcy : cy);
    uint32_t total = gather_group_runs(g, ix, cx, cy, cz, xlast, ylast, lane, sEnd[warp], sOff[warp]);
    if (hl.ctr && total > GATHER_HEAVY_MIN && klast != ck) {
      // A heavy stream is split into parts by its length, and the parts fix the order in which a query's photons are
      // summed.  Narrow the group to the leader's cell alone: the stream, its split and therefore every sum then
      // depend on the query's cell only, not on which other queries happen to share the warp (bit-reproducible
      // whatever order the counting sort left inside a cell).  The other lanes stay pending.
      act = valid && key == ck;
      grp = __ballot_sync(FULL, act);
      klast = ck;
      total = gather_group_runs(g, ix, cx, cy, cz, cx, cy, lane, sEnd[warp], sOff[warp]);
    }
    pending &= ~grp;
    if (hl.ctr && total > GATHER_HEAVY_MIN) {
      const uint32_t nparts = min(GATHER_HEAVY_MAXPARTS, (total + GATHER_HEAVY_MIN - 1u) / GATHER_HEAVY_MIN);
      uint32_t part0 = 0xFFFFFFFFu;
      if (lane == 0) {
        // Reserve nparts consecutive parts with ONE atomicAdd (a CAS loop collapses when every warp of a uniformly
        // dense map publishes at once).  The counter only grows, so the successful reservations are exactly a prefix
        // [0, ctr[3]) of the pool; once it is exhausted the warps do the work themselves.
        if (*(volatile unsigned int*)hl.ctr < hl.cap_parts) {
          const unsigned int old = atomicAdd(hl.ctr, nparts);
          if (old + nparts <= hl.cap_parts) { part0 = old; atomicMax(hl.ctr + 3, old + nparts); }
        }
        if (part0 != 0xFFFFFFFFu) {
          const unsigned int gs = atomicAdd(hl.ctr + 2, 1u);   // groups <= parts / 2: cannot overflow
          HeavyGroup* h = hl.groups + gs;
          h->s_base = (uint32_t)s_base; h->grp = grp; h->ck = ck; h->klast = klast; h->nparts = nparts; h->part0 = part0; h->done = 0u;
          for (uint32_t k = 0; k < nparts; ++k) hl.parts[part0 + k].group = gs;
        }
      }
      part0 = __shfl_sync(FULL, part0, 0);
      if (part0 != 0xFFFFFFFFu) {
        if (act) deferred = true;
        continue;
      }
    }
    const uint32_t staged = gather_chunks<FILTER, MODE>(m, total, 0u, 32u, lane, act, sEnd[warp], sOff[warp], sP[warp], sD[warp], qx, qy, qz,
                                                        nv, r2, power, rr, rg, rb, cnt);
    tests += (unsigned long long)staged * (unsigned)__popc(grp);
  }
  if (valid && !deferred) {
    if (MODE != 2) {
      const double sc = MODE == 0 ? ps->inv_pi_r2 : (1.0 / PPM_PI) / r2;   // rad * (ONE_PI / radius), tracer.rs:193
      rgb3[(uint64_t)qi * 3] = rr * sc; rgb3[(uint64_t)qi * 3 + 1] = rg * sc; rgb3[(uint64_t)qi * 3 + 2] = rb * sc;
    }
    if (counts) counts[qi] = cnt;
  }
  {
    const unsigned c = __reduce_add_sync(FULL, cnt);
    if (lane == 0 && c) atomicAdd(&ps->sum_k_s[blockIdx.x & (PPM_NSLOT - 1)], (unsigned long long)c);
    if (lane == 0 && tests) atomicAdd(&ps->cand_s[blockIdx.x & (PPM_NSLOT - 1)], tests);
  }
}

// Heavy parts: every warp claims parts by ticket (ctr[1]) until the list (ctr[3] parts, all published before this
// kernel starts) is exhausted.  Launched after every k_gather with a fixed grid; without parts it returns at once.
template <int FILTER, int MODE>
__global__ void __launch_bounds__(GATHER_WARPS * 32)
k_gather_heavy(PassDev* ps, CellIndex ix, MapSoA m, const uint32_t* __restrict__ qidx,
               const double* __restrict__ qpos3, const double* __restrict__ qnrm3,
               const double* __restrict__ r2q, double* __restrict__ rgb3, uint32_t* __restrict__ counts, HeavyList hl) {
  __shared__ double2 sP[GATHER_WARPS][32][2];
  __shared__ double2 sD[GATHER_WARPS][32][2];
  __shared__ uint32_t sEnd[GATHER_WARPS][32], sOff[GATHER_WARPS][32];
  const unsigned int nparts_total = hl.ctr[3];
  if (nparts_total == 0u) return;
  const unsigned FULL = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Grid g = ps->grid;
  const int64_t n = (int64_t)ps->n_query;
  const double power = ps->power, r2_fixed = ps->r2;
  for (;;) {
    unsigned int t = 0xFFFFFFFFu;
    if (lane == 0) { t = atomicAdd(hl.ctr + 1, 1u); if (t >= nparts_total) t = 0xFFFFFFFFu; }
    t = __shfl_sync(FULL, t, 0);
    if (t == 0xFFFFFFFFu) break;
    const uint32_t gs = __ldcg(&hl.parts[t].group);
    HeavyGroup* hg = hl.groups + gs;
    const uint32_t h_base = __ldcg(&hg->s_base), h_grp = __ldcg(&hg->grp), h_ck = __ldcg(&hg->ck), h_klast = __ldcg(&hg->klast),
                   h_np = __ldcg(&hg->nparts), h_p0 = __ldcg(&hg->part0);
    const int64_t hs = (int64_t)h_base + lane;
    const bool hact = ((h_grp >> lane) & 1u) != 0u && hs < n;
    uint32_t hqi = 0;
    double hx = 0.0, hy = 0.0, hz = 0.0, hr2 = r2_fixed;
    D3 hnv = mk3(0.0, 0.0, 0.0);
    if (hact) {
      hqi = qidx[hs];
      hx = qpos3[(uint64_t)hqi * 3]; hy = qpos3[(uint64_t)hqi * 3 + 1]; hz = qpos3[(uint64_t)hqi * 3 + 2];
      if (MODE != 2) hnv = ld3(qnrm3 + (uint64_t)hqi * 3);
      if (MODE != 0) hr2 = r2q[hqi];
    }
    const uint32_t hnx = (uint32_t)g.nx, hny = (uint32_t)g.ny;
    const uint32_t total = gather_group_runs(g, ix, (int)(h_ck % hnx), (int)((h_ck / hnx) % hny), (int)(h_ck / (hnx * hny)), (int)(h_klast % hnx),
                                             (int)((h_klast / hnx) % hny), lane, sEnd[warp], sOff[warp]);
    double ar = 0.0, ag = 0.0, ab = 0.0;
    uint32_t ac = 0;
    const uint32_t staged = gather_chunks<FILTER, MODE>(m, total, (t - h_p0) * 32u, h_np * 32u, lane, hact, sEnd[warp], sOff[warp], sP[warp],
                                                        sD[warp], hx, hy, hz, hnv, hr2, power, ar, ag, ab, ac);
    if (lane == 0 && staged) atomicAdd(&ps->cand_s[blockIdx.x & (PPM_NSLOT - 1)], (unsigned long long)staged * (unsigned)__popc(h_grp));
    HeavyPartial* hp = hl.partials + t;
    __stcg(&hp->rgb[0][lane], ar); __stcg(&hp->rgb[1][lane], ag); __stcg(&hp->rgb[2][lane], ab); __stcg(&hp->cnt[lane], ac);
    __threadfence();
    __syncwarp();
    unsigned int fin = 0;
    if (lane == 0) fin = atomicAdd(&hg->done, 1u) + 1u == h_np ? 1u : 0u;
    fin = __shfl_sync(FULL, fin, 0);
    if (!fin) continue;
    __threadfence();                                     // this warp completed the group: add the parts in order
    double tr = 0.0, tg = 0.0, tb = 0.0;
    uint32_t tc = 0;
    for (uint32_t k = 0; k < h_np; ++k) {
      const HeavyPartial* q = hl.partials + h_p0 + k;
      tr = tr + __ldcg(&q->rgb[0][lane]); tg = tg + __ldcg(&q->rgb[1][lane]); tb = tb + __ldcg(&q->rgb[2][lane]); tc += __ldcg(&q->cnt[lane]);
    }
    if (hact) {
      if (MODE != 2) {
        const double sc = (1.0 / PPM_PI) / hr2;
        rgb3[(uint64_t)hqi * 3] = tr * sc; rgb3[(uint64_t)hqi * 3 + 1] = tg * sc; rgb3[(uint64_t)hqi * 3 + 2] = tb * sc;
      }
      if (counts) counts[hqi] = tc;
    }
    {
      const unsigned c = __reduce_add_sync(FULL, hact ? tc : 0u);
      if (lane == 0 && c) atomicAdd(&ps->sum_k_s[blockIdx.x & (PPM_NSLOT - 1)], (unsigned long long)c);
    }
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

__global__ void k_within(const PassDev* __restrict__ ps, CellIndex ix, MapSoA m, const double* __restrict__ qpos3,
                         int64_t n, uint32_t* __restrict__ idx, uint32_t* __restrict__ counts, uint32_t cap) {
  int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const Grid g = ps->grid;
  const double r2 = ps->r2;
  const double qx = qpos3[q * 3], qy = qpos3[q * 3 + 1], qz = qpos3[q * 3 + 2];
  uint32_t cnt = 0;
  int cx = cell_coord(g, qx, 0), cy = cell_coord(g, qy, 1), cz = cell_coord(g, qz, 2);
  const int R = 1;
  int x0 = max(cx - R, 0), x1 = min(cx + R, g.nx - 1);
  if (x0 <= x1)
    for (int z = max(cz - R, 0); z <= min(cz + R, g.nz - 1); ++z)
      for (int y = max(cy - R, 0); y <= min(cy + R, g.ny - 1); ++y) {
        uint32_t row = ((uint32_t)z * (uint32_t)g.ny + (uint32_t)y) * (uint32_t)g.nx;
        uint32_t b = cell_begin(ix, row + (uint32_t)x0), e = cell_begin(ix, row + (uint32_t)x1 + 1u);
        for (uint32_t j = b; j < e; ++j) {
          const double2 a = m.P[(uint64_t)j * 2], b2 = m.P[(uint64_t)j * 2 + 1];
          double ax = qx - a.x, ay = qy - a.y, az = qz - b2.x;
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
