// xnb_sweep_cl.cuh -- pair sweep over "compiled lists" (sm_100a).
//
// The neighbour lists stay in the reference's GridChunkNeighbors format (chunk_neighbors.h:42-186); that is what
// xnb_chunk_neighbors builds and what xnb_view_chunk_neighbors exports.  For its own pair sweep the library also keeps a
// second, derived copy of the same lists, laid out for the way the sweep kernel reads them:
//
//   * the swept cells are cut into tiles of ti x tj x tk cells; one block sweeps one tile and stages the positions of
//     the tile's halo box in shared memory.  A candidate is stored as the u16 index of the neighbour in that staged
//     array (times 8: the byte offset of its z, half the byte offset of its {x,y} pair), so the sweep needs no (cell
//     code, count) headers, no cell-base lookups and one shift-add per address;
//   * the particles of a tile are numbered 0..n_tile-1 (tile cell after tile cell); 32 consecutive particles form a
//     GROUP = one warp of the sweep.  The lists of a group are stored row by row, row k holding words 4k..4k+3 of all 32
//     lists (8 bytes per lane, 256 bytes per row): every warp load is one fully coalesced 256-byte line, streamed once;
//   * lists of a group are padded to the group's longest list with the particle's OWN staged index: d2 = 0 is rejected
//     by the reference's own test (d2 > 0, impl_default.h:186), so pads need no flag and no branch.
//
// k_cl_compile derives this copy from the reference-format streams after every rebuild; k_lj_sweep_cl is the sweep.
// Results are identical to the reference-format sweep (k_lj_sweep): same candidates in the same order, same arithmetic.
#pragma once
#include "xnb_kernels.cuh"

namespace xnb {

struct ClTileP
{
  int ti, tj, tk;                       // cells per tile
  int gap;                              // neighbour cell layers
  int lo[3], hi[3];                     // swept cell range [lo,hi)
  int tiles_i, tiles_j, tiles_k;
  int cap;                              // staging capacity in particles (<= 8191: a u16 list word is 8 x the staged index)
  int nh_max, tc_max;                   // nominal halo cells / tile cells: sizes of the shared-memory tables
  int gmax;                             // group-table entries per tile
};

struct ClTile { int ci0, cj0, ck0, tci, tcj, tck, tcells, bx0, by0, bz0, HX, HY, HZ, NH; };

XNB_DEVINL ClTile cl_tile(const GridP& g, const ClTileP& tp, int b)
{
  ClTile T;
  const int t_i = b % tp.tiles_i; b /= tp.tiles_i;
  const int t_j = b % tp.tiles_j; const int t_k = b / tp.tiles_j;
  T.ci0 = tp.lo[0] + t_i * tp.ti; T.cj0 = tp.lo[1] + t_j * tp.tj; T.ck0 = tp.lo[2] + t_k * tp.tk;
  T.tci = min(tp.ti, tp.hi[0] - T.ci0); T.tcj = min(tp.tj, tp.hi[1] - T.cj0); T.tck = min(tp.tk, tp.hi[2] - T.ck0);
  T.tcells = T.tci * T.tcj * T.tck;
  T.bx0 = max(T.ci0 - tp.gap, 0); const int bx1 = min(T.ci0 + T.tci - 1 + tp.gap, g.dims[0] - 1);
  T.by0 = max(T.cj0 - tp.gap, 0); const int by1 = min(T.cj0 + T.tcj - 1 + tp.gap, g.dims[1] - 1);
  T.bz0 = max(T.ck0 - tp.gap, 0); const int bz1 = min(T.ck0 + T.tck - 1 + tp.gap, g.dims[2] - 1);
  T.HX = bx1 - T.bx0 + 1; T.HY = by1 - T.by0 + 1; T.HZ = bz1 - T.bz0 + 1;
  T.NH = T.HX * T.HY * T.HZ;
  return T;
}

// shared-memory tables of a tile (both kernels): particle prefix of the halo cells (= staged index of each cell's first
// particle), global index of each halo cell's first particle, particle prefix of the tile cells, halo index of each tile cell
struct ClTables { uint32_t *hstart, *hfirst, *tstart, *thalo; };

XNB_DEVINL size_t cl_tables_bytes(int nh_max, int tc_max) { return (((size_t)(2 * nh_max + 2 * tc_max + 2) * 4 + 15) & ~(size_t)15); }
// dynamic shared memory of the sweep: tables | {x,y}[cap] | z[cap]
__host__ __device__ inline size_t cl_sweep_smem_bytes(int nh_max, int tc_max, int cap, int /*warps*/)
{
  return (((size_t)(2 * nh_max + 2 * tc_max + 2) * 4 + 15) & ~(size_t)15) + (size_t)cap * 24;
}

XNB_DEVINL ClTables cl_tables(unsigned char* smem, const ClTileP& tp)
{
  ClTables t;
  t.hstart = reinterpret_cast<uint32_t*>(smem);
  t.hfirst = t.hstart + tp.nh_max + 1;
  t.tstart = t.hfirst + tp.nh_max;
  t.thalo = t.tstart + tp.tc_max + 1;
  return t;
}

// block-cooperative; ends with a barrier.  Tile cells are numbered q = (kk * tcj + jj) * tci + ii.
// sel_mode 1: only tile cells outside the inner range take part (their particles are the tile particles); the others are still
// staged as candidates
XNB_DEVINL void cl_setup(const GridP& g, const ClTile& T, const ClTables& tb, const uint32_t* __restrict__ cell_start,
                         const uint32_t* __restrict__ cell_count, uint32_t* s_scan, int sel_mode = 0)
{
  const int HXY = T.HX * T.HY;
  uint32_t carry = 0;
  for (int base = 0; base < T.NH; base += blockDim.x)
  {
    const int h = base + threadIdx.x;
    uint32_t cnt = 0;
    if (h < T.NH)
    {
      const int hxq = h % T.HX, hyq = (h / T.HX) % T.HY, hzq = h / HXY;
      const int c = ijk_to_index(g.dims, T.bx0 + hxq, T.by0 + hyq, T.bz0 + hzq);
      cnt = cell_count[c]; tb.hfirst[h] = cell_start[c];
    }
    uint32_t total;
    const uint32_t off = block_exclusive_scan<uint32_t>(cnt, &total, s_scan);
    if (h < T.NH) tb.hstart[h] = carry + off;
    carry += total;
  }
  if (threadIdx.x == 0) tb.hstart[T.NH] = carry;
  __syncthreads();
  if (threadIdx.x < 32)
  {
    const int q = threadIdx.x;      // tcells <= 32
    uint32_t cnt = 0; int h = 0;
    if (q < T.tcells)
    {
      const int ii = q % T.tci, jj = (q / T.tci) % T.tcj, kk = q / (T.tci * T.tcj);
      h = ((T.ck0 + kk - T.bz0) * T.HY + (T.cj0 + jj - T.by0)) * T.HX + (T.ci0 + ii - T.bx0);
      cnt = tb.hstart[h + 1] - tb.hstart[h];
      if (sel_mode == 1)
      {
        const int ci = T.ci0 + ii, cj = T.cj0 + jj, ck = T.ck0 + kk;
        if (ci >= g.gl && ci < g.dims[0] - g.gl && cj >= g.gl && cj < g.dims[1] - g.gl && ck >= g.gl && ck < g.dims[2] - g.gl) cnt = 0;
      }
    }
    uint32_t x = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (q >= o) x += y; }
    if (q < T.tcells) { tb.tstart[q] = x - cnt; tb.thalo[q] = (uint32_t)h; }
    if (q == T.tcells - 1) tb.tstart[T.tcells] = x;
  }
  __syncthreads();
}

// tile cell of tile particle t (binary search in the prefix table; tcells <= 32)
XNB_DEVINL int cl_find_cell(const uint32_t* tstart, int tcells, uint32_t t)
{
  int lo = 0, hi = tcells - 1;
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (t >= tstart[mid]) lo = mid; else hi = mid - 1; }
  return lo;
}

// ------------------------------------------------------------------------------------------------------------------
// k_cl_compile: reference-format streams -> compiled lists.  One block per tile, one warp per group of 32 tile
// particles, one thread per list.  A thread reads its list with aligned 16-byte loads and walks it with the state
// machine of the format (chunknbh_stream_info / the nested loops of impl_default.h:143-179): cell code -> count ->
// `count` candidates (a flat walk diverges less than the nested loops: measured 0.46 against 0.60 ms at C2).  A candidate's staged index is hstart[halo cell of its group] + p_b; four of them make one 8-byte
// word of the lane's column in the group's rows.
// counters: [0] rows used (bump allocator) [1] max groups of a tile [2] max staged particles of a tile; *n_candidates += list entries
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_cl_compile(GridP g, ClTileP tp, const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_count,
             const uint16_t* const* __restrict__ cell_stream, uint2* __restrict__ groups, uint2* __restrict__ rows,
             uint32_t cap_rows, uint32_t* __restrict__ counters, unsigned long long* __restrict__ n_candidates)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t s_scan[32];
  const ClTile T = cl_tile(g, tp, (int)blockIdx.x);
  const ClTables tb = cl_tables(smem_raw, tp);
  cl_setup(g, T, tb, cell_start, cell_count, s_scan);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const uint32_t n_tile = tb.tstart[T.tcells], n_halo = tb.hstart[T.NH];
  const uint32_t ngroups = (n_tile + 31u) >> 5;
  uint2* const gt = groups + (size_t)blockIdx.x * (size_t)tp.gmax;
  if (threadIdx.x == 0) { atomicMax(&counters[1], ngroups); atomicMax(&counters[2], n_halo); }
  for (uint32_t q = min(ngroups, (uint32_t)tp.gmax) + threadIdx.x; q < (uint32_t)tp.gmax; q += blockDim.x) gt[q] = make_uint2(0u, 0u);
  if (ngroups > (uint32_t)tp.gmax || n_halo > (uint32_t)tp.cap) return;      // the host reads the counters and re-runs with more room
  const int HXY = T.HX * T.HY;

  for (uint32_t grp = warp; grp < ngroups; grp += nwarp)
  {
    // ---- this lane's particle: list location, candidate count, own staged index
    const uint32_t t = grp * 32u + lane;
    const bool active = t < n_tile;
    const uint16_t* lst = nullptr; uint32_t len = 0, ncand = 0, ngrp = 0; int hb2 = 0;
    // (idle lanes of the last group stand on tile particle 0, exactly as in the sweep, so that their pads are "self" too)
    const int q = cl_find_cell(tb.tstart, T.tcells, active ? t : 0u);
    const uint32_t pa = (active ? t : 0u) - tb.tstart[q], na = tb.tstart[q + 1] - tb.tstart[q];
    const int hA = (int)tb.thalo[q];
    const uint32_t self = tb.hstart[hA] + pa;
    if (active)
    {
      const int ii = q % T.tci, jj = (q / T.tci) % T.tcj, kk = q / (T.tci * T.tcj);
      const int c = ijk_to_index(g.dims, T.ci0 + ii, T.cj0 + jj, T.ck0 + kk);
      const uint16_t* cs = cell_stream[c];
      const uint32_t off0 = reinterpret_cast<const uint32_t*>(cs)[pa], off1 = reinterpret_cast<const uint32_t*>(cs)[pa + 1];
      lst = cs + 2u * (na + 1u) + off0;               // first word behind the group counter (offsets are biased by the number of tables = 1)
      len = off1 - off0 - 1u;
      ngrp = (uint32_t)lst[-1];
      ncand = len - 2u * ngrp;
      hb2 = hA - 16 * (HXY + T.HX + 1);               // halo index of the cell with raw 5-bit fields (0,0,0)
    }
    uint32_t trips = (ncand + 3u) >> 2, csum = ncand;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { trips = max(trips, __shfl_xor_sync(0xffffffffu, trips, o)); csum += __shfl_xor_sync(0xffffffffu, csum, o); }
    uint32_t row0 = 0;
    if (lane == 0) { row0 = atomicAdd(&counters[0], trips); atomicAdd(n_candidates, (unsigned long long)csum); }
    row0 = __shfl_sync(0xffffffffu, row0, 0);
    const bool fits = row0 + trips <= cap_rows;        // else: keep counting, write nothing
    if (lane == 0) gt[grp] = make_uint2(row0, fits ? trips : 0u);
    if (!fits) continue;

    uint2* col = rows + ((size_t)row0 * 32u + (uint32_t)lane);     // this lane's column: word k of the list lives in col[(k >> 2) * 32]
    unsigned long long buf = 0ull;
    uint32_t r = 0;                                   // candidates emitted
    if (len)
    {
      const uint4* chunk = reinterpret_cast<const uint4*>(reinterpret_cast<uintptr_t>(lst) & ~(uintptr_t)15);
      uint32_t skip = (uint32_t)((reinterpret_cast<uintptr_t>(lst) & 15) >> 1);     // words of the first chunk in front of the list
      uint32_t left = len;                            // stream words still to walk
      uint32_t cnt = 0, base = 0; bool have_code = false;
      uint4 cur = __ldg(chunk);
      while (left)
      {
        const uint4 nxt = (left + skip > 8u) ? __ldg(chunk + 1) : cur;
        const uint32_t w2[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
        for (int v = 0; v < 8; v++)
        {
          const uint32_t w = (v & 1) ? (w2[v >> 1] >> 16) : (w2[v >> 1] & 0xffffu);
          if (skip) { skip--; continue; }
          if (!left) continue;
          left--;
          if (cnt)
          {
            buf = (buf >> 16) | ((unsigned long long)((base + w) << 3) << 48);     // a list word = 8 x staged index (<= 65528)
            cnt--; r++;
            if ((r & 3u) == 0u) col[(size_t)((r >> 2) - 1u) * 32u] = make_uint2((uint32_t)buf, (uint32_t)(buf >> 32));
          }
          else if (have_code) { cnt = w; have_code = false; }
          else
          {
            // halo cell (cia + ri, cja + rj, cka + rk), (ri,rj,rk) = 5-bit fields of the code minus 16 (chunk_neighbors.h:137-162)
            base = tb.hstart[hb2 + (int)(w >> 10) * HXY + (int)((w >> 5) & 31u) * T.HX + (int)(w & 31u)];
            have_code = true;
          }
        }
        cur = nxt; chunk++;
      }
    }
    // pads: the particle's own staged index, up to the group's trip count
    while (r < 4u * trips)
    {
      buf = (buf >> 16) | ((unsigned long long)(self << 3) << 48);
      r++;
      if ((r & 3u) == 0u) col[(size_t)((r >> 2) - 1u) * 32u] = make_uint2((uint32_t)buf, (uint32_t)(buf >> 32));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// k_lj_sweep_cl: the pair sweep (compute_cell_particle_pairs_impl_default.h:87-239 with the Lennard-Jones functor,
// lennard_jones.cu:46-56,106-124) over compiled lists.  MODE / EV as k_lj_sweep.
// One block = one tile; one thread = one tile particle; one warp = one group.  Staged in shared memory (cp.async): the
// positions of the tile's halo box as {x,y} + z.  Per trip a lane reads its four next candidates with one 8-byte load
// (the warp: one 256-byte row, two trips ahead), fetches the four positions and evaluates the four pairs with
// independent FP64 chains.  Arithmetic identical to k_lj_sweep.
// VAR 0: <= 576 threads, two blocks per SM (<= 56 registers); VAR 1: <= 1024 threads.
// ------------------------------------------------------------------------------------------------------------------
#ifndef XNB_CL_PREFETCH_ROWS
#define XNB_CL_PREFETCH_ROWS 8
#endif
constexpr uint32_t CL_PREFETCH_ROWS = XNB_CL_PREFETCH_ROWS;      // rows pulled towards L2 ahead of the trip that needs them (4 and 16 measured: no better)

XNB_DEVINL uint2 ld_stream8(const uint2* p)
{
  uint2 v;
  asm volatile("ld.global.cs.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
XNB_DEVINL void lds_f64x2(uint32_t a, double& x, double& y) { asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a)); }
XNB_DEVINL void lds_f64(uint32_t a, double& x) { asm("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(a)); }
// 0 < d2 <= rc2 as ONE predicate: d2 != 0 read off the high word feeds the FP64 compare (DSETP.LE.AND); written in C the two
// tests become two predicates and twice the selects
XNB_DEVINL bool in_cut(double d2, double rc2)
{
  uint32_t r;
  asm("{ .reg .pred p, q; .reg .b32 l, h; mov.b64 {l, h}, %1; setp.ne.s32 q, h, 0; setp.le.and.f64 p, %1, %2, q; selp.u32 %0, 1, 0, p; }"
      : "=r"(r) : "d"(d2), "d"(rc2));
  return r != 0u;
}
// MODE 2 = MODE 1 + the FIRST HALF OF THE NEXT STEP in the same epilogue (xnb_run_steps): with a and the kicked v in registers,
//   r' = r + (v dt + a dt^2/2), v' = v + a dt/2, displacement test of r' against the backup (k_verlet_first_half's arithmetic, same
// operations in the same order), r' written to the OTHER position buffer (other blocks still stage r), the count added to `counter`.
struct NextHalfP
{
  double dt, dt2, thr2;
  double *nrx, *nry, *nrz;
  const uint32_t* atom_cell; const uint32_t* backup;
  unsigned long long* counter;
};
template <class F, int MODE, bool EV, int VAR>
__global__ void __launch_bounds__(VAR == 0 ? 576 : VAR == 1 ? 1024 : VAR == 2 ? 288 : 576, VAR == 0 ? 2 : VAR == 1 ? 1 : VAR == 2 ? 3 : 1)
k_lj_sweep_cl(GridP g, ClTileP tp, int n_inner, int n_total, F lj, double dth, NextHalfP nh,
              const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz,
              double* __restrict__ vx, double* __restrict__ vy, double* __restrict__ vz,
              double* __restrict__ fx, double* __restrict__ fy, double* __restrict__ fz,
              const uint8_t* __restrict__ type, const double* __restrict__ mass,
              const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_count,
              const uint2* __restrict__ groups, const uint2* __restrict__ rows,
              double* __restrict__ ev_partials /* [gridDim.x][7] */, uint32_t* __restrict__ err,
              const unsigned long long* __restrict__ skip_if_nonzero, const uint32_t* __restrict__ tile_list)
{
  // speculative launch (xnb_run_steps): the displacement counter says a rebuild is due -> this launch is void
  if (skip_if_nonzero && *skip_if_nonzero) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t s_scan[32];
  // tile_list: this launch sweeps a subset of the tiles (interior tiles while the halo is in flight, boundary tiles after)
  const uint32_t tile = tile_list ? tile_list[blockIdx.x] : blockIdx.x;
  const ClTile T = cl_tile(g, tp, (int)tile);
  const ClTables tb = cl_tables(smem_raw, tp);
  double2* XY = reinterpret_cast<double2*>(smem_raw + cl_tables_bytes(tp.nh_max, tp.tc_max));   // [cap]
  double* Z = reinterpret_cast<double*>(XY + tp.cap);                                             // [cap]
  cl_setup(g, T, tb, cell_start, cell_count, s_scan);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const uint32_t n_tile = tb.tstart[T.tcells], n_halo = tb.hstart[T.NH];
  const uint32_t ngroups = (n_tile + 31u) >> 5;
  const bool bad = ngroups > (uint32_t)tp.gmax || n_halo > (uint32_t)tp.cap;   // cannot happen after a successful compile
  if (bad && threadIdx.x == 0) atomicOr(err, DERR_TILE_CAPACITY);

  const uint2* const gt = groups + (size_t)tile * (size_t)tp.gmax;
  if (!bad && threadIdx.x < ngroups * 32u)
  {
    // start pulling this warp's first list rows towards L2 while the positions are staged
    const uint2 ge = gt[threadIdx.x >> 5];
    const uint2* R = rows + ((size_t)ge.x * 32u + (uint32_t)lane);
    for (uint32_t k = 0; k < min(ge.y, CL_PREFETCH_ROWS); k++) asm volatile("prefetch.global.L2 [%0];" :: "l"(R + (size_t)k * 32u));
  }
  if (!bad && n_tile > 0)
  {
    // halo positions: one warp per halo cell, lanes over its particles (coalesced reads)
    for (int h = warp; h < T.NH; h += nwarp)
    {
      const uint32_t d0 = tb.hstart[h], cnt = tb.hstart[h + 1] - d0, s0 = tb.hfirst[h];
      for (uint32_t p = lane; p < cnt; p += 32)
      {
        cp_async8(&XY[d0 + p].x, rx + s0 + p); cp_async8(&XY[d0 + p].y, ry + s0 + p); cp_async8(Z + d0 + p, rz + s0 + p);
      }
    }
    cp_async_wait_all();
  }
  __syncthreads();

  PairAcc acc;
  acc.e = acc.wxx = acc.wyy = acc.wzz = acc.wxy = acc.wxz = acc.wyz = 0.;
  const double rc2 = lj.rcut2();
  const uint32_t xyb = (uint32_t)__cvta_generic_to_shared(XY), zb = (uint32_t)__cvta_generic_to_shared(Z);

  if (!bad)
  for (uint32_t t = threadIdx.x; t < ngroups * 32u; t += blockDim.x)
  {
    const bool active = t < n_tile;
    const int q = cl_find_cell(tb.tstart, T.tcells, active ? t : 0u);
    const uint32_t pa = (active ? t : 0u) - tb.tstart[q];
    const uint32_t self = tb.hstart[tb.thalo[q]] + pa;
    const uint32_t i = tb.hfirst[tb.thalo[q]] + pa;
    const double2 ra = XY[self];
    const double xa = ra.x, ya = ra.y, za = Z[self];
    // epilogue operands fetched now so that their latency hides behind the pair loop
    double m = 1.0, ux = 0., uy = 0., uz = 0.;
    if (MODE == 1 && active) { m = mass[type[i]]; if (dth != 0.0) { ux = vx[i]; uy = vy[i]; uz = vz[i]; } }
    acc.ax = acc.ay = acc.az = 0.;
    const uint2 ge = gt[t >> 5];
    const uint32_t trips = ge.y;
    const uint2* R = rows + ((size_t)ge.x * 32u + (uint32_t)lane);
    // register prefetch ONE row ahead, issued at the very top of a trip and pinned there by a warp barrier.  (Written as a
    // two-rows-ahead rotation, ptxas sank the load to the bottom of the loop body; the six scoreboards of a warp are all needed by
    // the eight position gathers of the body, so the loop head then waited for that load every trip: 15 % of all stall samples,
    // profiles/r1m_sweep_cl.md.)  The rows are streamed from HBM exactly once, so they are pulled into L2 well ahead of that.
    const uint32_t selfw = (self << 3) | (self << 19);
    uint2 w0 = make_uint2(selfw, selfw), w1 = w0;
    if (trips > 0u) w0 = ld_stream8(R);
    for (uint32_t k = 0; k < trips; k++)
    {
      if (k + 1u < trips) w1 = ld_stream8(R + (size_t)(k + 1u) * 32u);
      if (k + CL_PREFETCH_ROWS < trips) asm volatile("prefetch.global.L2 [%0];" :: "l"(R + (size_t)(k + CL_PREFETCH_ROWS) * 32u));
      __syncwarp();
      // a list word is 8 x the staged index of the candidate: byte offset of its z, half the byte offset of its {x,y}
      // (PRMT for the low halves: written as `& 0xffff` the compiler folds the mask into the scaled address and needs two more instructions)
      const uint32_t j[4] = {__byte_perm(w0.x, 0u, 0x4410), w0.x >> 16, __byte_perm(w0.y, 0u, 0x4410), w0.y >> 16};
      double dx[4], dy[4], dz[4], d2[4]; bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; u++)
      {
        double px, py, pz;
        lds_f64x2(xyb + j[u] + j[u], px, py); lds_f64(zb + j[u], pz);
        dx[u] = __dadd_rn(px, -xa); dy[u] = __dadd_rn(py, -ya); dz[u] = __dadd_rn(pz, -za);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) d2[u] = norm2_exact(dx[u], dy[u], dz[u]);
      // accept iff d2 > 0 && d2 <= rcut2 (impl_default.h:186).  d2 > 0 is read off the high word (a sum of squares is never
      // -0; below 2^-1022 it counts as zero: DESIGN.md 4), so the test costs one integer and one FP64 compare
#pragma unroll
      for (int u = 0; u < 4; u++) ok[u] = in_cut(d2[u], rc2);
      pair_apply4<EV>(lj, dx, dy, dz, d2, ok, j, acc);
      w0 = w1;
    }
    bool over = false;
    if (active)
    {
      double ax = acc.ax, ay = acc.ay, az = acc.az;
      if (MODE == 0)
      {
        fx[i] += ax; fy[i] += ay; fz[i] += az;
      }
      else
      {
        if (MODE == 2) { m = mass[type[i]]; ux = vx[i]; uy = vy[i]; uz = vz[i]; }      // (MODE 2 has no register to park them in during the loop)
        ax = __ddiv_rn(ax, m); ay = __ddiv_rn(ay, m); az = __ddiv_rn(az, m);
        fx[i] = ax; fy[i] = ay; fz[i] = az;
        if (MODE == 2)
        {
          // second half of this step, then the first half of the next (push_f_v{0.5}; push_f_v_r{1.0}; push_f_v{0.5})
          ux = __dadd_rn(ux, __dmul_rn(ax, dth)); uy = __dadd_rn(uy, __dmul_rn(ay, dth)); uz = __dadd_rn(uz, __dmul_rn(az, dth));
          const double x = __dadd_rn(xa, __dadd_rn(__dmul_rn(ux, nh.dt), __dmul_rn(ax, nh.dt2)));
          const double y = __dadd_rn(ya, __dadd_rn(__dmul_rn(uy, nh.dt), __dmul_rn(ay, nh.dt2)));
          const double z = __dadd_rn(za, __dadd_rn(__dmul_rn(uz, nh.dt), __dmul_rn(az, nh.dt2)));
          nh.nrx[i] = x; nh.nry[i] = y; nh.nrz[i] = z;
          vx[i] = __dadd_rn(ux, __dmul_rn(ax, dth));
          vy[i] = __dadd_rn(uy, __dmul_rn(ay, dth));
          vz[i] = __dadd_rn(uz, __dmul_rn(az, dth));
          over = displ_over_test(g, nh.atom_cell[i], nh.backup + 3 * (size_t)i, x, y, z, nh.thr2);
        }
        else if (dth != 0.0)
        {
          vx[i] = __dadd_rn(ux, __dmul_rn(ax, dth));
          vy[i] = __dadd_rn(uy, __dmul_rn(ay, dth));
          vz[i] = __dadd_rn(uz, __dmul_rn(az, dth));
        }
      }
    }
    if (MODE == 2) block_count_add(over, nh.counter);
  }
  if (MODE != 0)
  {
    // zero_particle_force{ghost:true}: ghost particles keep f = 0 (each block clears its slice of the ghost range)
    const int ng = n_total - n_inner;
    const int per = (ng + (int)gridDim.x - 1) / (int)gridDim.x;
    const int g0 = n_inner + (int)blockIdx.x * per, g1 = min(g0 + per, n_total);
    for (int i = g0 + (int)threadIdx.x; i < g1; i += blockDim.x) { fx[i] = 0.; fy[i] = 0.; fz[i] = 0.; }
  }
  if (EV)
  {
    __shared__ double red[7][32];
    double vals[7] = {acc.e, acc.wxx, acc.wyy, acc.wzz, acc.wxy, acc.wxz, acc.wyz};
#pragma unroll
    for (int q = 0; q < 7; q++)
    {
      double v = vals[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[q][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 7)
    {
      double v = 0.;
      for (int wq = 0; wq < nwarp; wq++) v += red[threadIdx.x][wq];
      ev_partials[(size_t)blockIdx.x * 7 + threadIdx.x] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// k_pair_sweep_sym: Newton-3 pair sweep over half_symmetric GridChunkNeighbors lists --
// compute_cell_particle_pairs<Symmetric = true> (impl_default.h:143-239) with ComputePairOptionalLocks<true>
// (compute_pair_optional_args.h:152-161): each listed pair is evaluated once, f_a += de*dr and f_b -= de*dr, b may be a
// ghost (update_force_from_ghost returns that part to its owner).  Where the reference serialises the concurrent updates
// of a cell's particles with per-cell spin locks, this kernel uses FP64 atomic adds (RED.ADD.F64): the sums are the same
// up to the order of the additions.  One thread per inner particle, reading the reference-format stream directly.
// Forces must have been zeroed (zero_particle_force{ghost: true}).
// ------------------------------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128)
k_pair_sweep_sym(GridP g, int n_inner, F func, const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz,
                 double* __restrict__ fx, double* __restrict__ fy, double* __restrict__ fz,
                 const uint32_t* __restrict__ atom_cell, const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_count,
                 const uint16_t* const* __restrict__ cell_stream)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_inner) return;
  const uint32_t ca = atom_cell[i];
  const uint32_t na = cell_count[ca], pa = (uint32_t)i - cell_start[ca];
  const uint16_t* cs = cell_stream[ca];
  const uint32_t off0 = reinterpret_cast<const uint32_t*>(cs)[pa];
  const uint16_t* lst = cs + 2u * (na + 1u) + off0;          // first word behind the group counter (offsets are biased by the number of tables = 1)
  uint32_t ngrp = (uint32_t)lst[-1];
  const double xa = rx[i], ya = ry[i], za = rz[i];
  const double rc2 = func.rcut2();
  double ax = 0., ay = 0., az = 0.;
  const int dxy = g.dims[0] * g.dims[1];
  bool stop = false;
  for (; ngrp > 0u && !stop; ngrp--)
  {
    const uint32_t code = *lst++; uint32_t n = *lst++;
    // chunk_neighbors.h:137-162: 5-bit fields of the code minus 16 = cell of b relative to the cell of a
    const int cb = (int)ca + ((int)(code >> 10) - 16) * dxy + ((int)((code >> 5) & 31u) - 16) * g.dims[0] + ((int)(code & 31u) - 16);
    const uint32_t sb = cell_start[cb];
    for (; n > 0u; n--)
    {
      const uint32_t pb = *lst++;
      if ((uint32_t)cb > ca || ((uint32_t)cb == ca && pb > pa)) { stop = true; break; }        // impl_default.h:181
      const uint32_t j = sb + pb;
      const double dx = __dadd_rn(rx[j], -xa), dy = __dadd_rn(ry[j], -ya), dz = __dadd_rn(rz[j], -za);
      const double d2 = norm2_exact(dx, dy, dz);
      if (d2 > 0.0 && d2 <= rc2)
      {
        double px = 0., py = 0., pz = 0.;
        func(make_double3(dx, dy, dz), d2, px, py, pz, PairNbh{j}, 1.0);
        ax += px; ay += py; az += pz;
        atomicAdd(fx + j, -px); atomicAdd(fy + j, -py); atomicAdd(fz + j, -pz);
      }
    }
  }
  atomicAdd(fx + i, ax); atomicAdd(fy + i, ay); atomicAdd(fz + i, az);
}

// update_force_from_ghost (UpdateFromGhosts<fx,fy,fz,UpdateValueAdd>, mpi/update_force_from_ghost.cu:44): unpack side.  Send-list
// entry q created one ghost of particle send_src[q]; that ghost's force (received in `stage`, or read in place for the periodic
// self images) is added to the particle.  A particle can have several images, hence the atomic adds.
__global__ void k_ghost_unpack_add(int n_send, const uint32_t* __restrict__ send_src, int self_first, int self_end, uint32_t self_dst,
                                   const double* __restrict__ stage /* planes of n_send doubles: fx fy fz */,
                                   double* __restrict__ fx, double* __restrict__ fy, double* __restrict__ fz)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_send) return;
  const uint32_t s = send_src[q];
  double x, y, z;
  if (q >= self_first && q < self_end) { const uint32_t d = self_dst + (uint32_t)(q - self_first); x = fx[d]; y = fy[d]; z = fz[d]; }
  else { const size_t n = (size_t)n_send; x = stage[q]; y = stage[n + q]; z = stage[2 * n + q]; }
  atomicAdd(fx + s, x); atomicAdd(fy + s, y); atomicAdd(fz + s, z);
}

} // namespace xnb
