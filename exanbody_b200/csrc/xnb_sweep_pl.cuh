// xnb_sweep_pl.cuh -- pair sweep over compiled lists for LARGE cells (C1, C4, the Ni deck: 256 particles per cell): k_lj_sweep_pl.
//
// reference: compute/include/exanb/compute/compute_cell_particle_pairs_impl_default.h:87-239 with the functor of
// contribs/md/lennard_jones/lennard_jones.cu:46-56,106-124 -- same contract, same arithmetic as k_lj_sweep_cl (xnb_sweep_cl.cuh).
//
// For cells of hundreds of particles the sweep's tile is ONE cell and its 27-cell halo is 166 KB of staged positions: one block of
// eight or nine warps per SM, and the kernel waits on latency (C1 0.19 ms, C4 0.82 ms).  A particle's compiled list is ordered by
// staged index, i.e. z-plane of cells after z-plane, so k_nbh_big (xnb_nbh_big.cuh) emits it in PLANE SEGMENTS -- plane p of a group
// = trips_p rows, indices relative to the first particle of that plane, padded with the plane's sentinel slot -- and this kernel stages
// one plane at a time (a third of the halo: three blocks per SM), sweeps the segment, and moves on:
//   for p in planes: barrier; stage the 9 cells of plane p (+ a far-away sentinel at slot cap_pl); barrier; trips_p rows of 4 candidates.
// One warp = one group (blockDim = 32 x groups of the fullest cell), one thread = one particle; its own position comes from the flat
// arrays.  A pad is the sentinel (1e30, 1e30, 1e30): d2 = 3e60 fails the cut test like any far candidate, no flag, no branch.
// groups[tile * gmax + g] = (first row, trips_0 | trips_1 << 10 | trips_2 << 20).
#pragma once
#include "xnb_sweep_cl.cuh"

namespace xnb {

constexpr int SWEEP_PL_MAX_THREADS = 512;
// dynamic shared memory: tables | {x,y}[cap_pl + 1] | z[cap_pl + 1]
__host__ __device__ inline size_t pl_sweep_smem_bytes(int nh_max, int tc_max, int cap_pl)
{
  return (((size_t)(2 * nh_max + 2 * tc_max + 2) * 4 + 15) & ~(size_t)15) + (size_t)(cap_pl + 2) * 24;
}

template <class F, int MODE, bool EV>
__global__ void __launch_bounds__(SWEEP_PL_MAX_THREADS, 2)
k_lj_sweep_pl(GridP g, ClTileP tp, int cap_pl, int n_inner, int n_total, F lj, double dth, NextHalfP nh,
              const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz,
              double* __restrict__ vx, double* __restrict__ vy, double* __restrict__ vz,
              double* __restrict__ fx, double* __restrict__ fy, double* __restrict__ fz,
              const uint8_t* __restrict__ type, const double* __restrict__ mass,
              const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_count,
              const uint2* __restrict__ groups, const uint2* __restrict__ rows,
              double* __restrict__ ev_partials /* [gridDim.x][7] */, uint32_t* __restrict__ err,
              const unsigned long long* __restrict__ skip_if_nonzero, const uint32_t* __restrict__ tile_list)
{
  if (skip_if_nonzero && *skip_if_nonzero) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t s_scan[32];
  const uint32_t tile = tile_list ? tile_list[blockIdx.x] : blockIdx.x;
  const ClTile T = cl_tile(g, tp, (int)tile);
  const ClTables tb = cl_tables(smem_raw, tp);
  double2* XY = reinterpret_cast<double2*>(smem_raw + cl_tables_bytes(tp.nh_max, tp.tc_max));   // [cap_pl + 1]
  double* Z = reinterpret_cast<double*>(XY + (cap_pl + 1));                                       // [cap_pl + 1]
  cl_setup(g, T, tb, cell_start, cell_count, s_scan);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const uint32_t n_tile = tb.tstart[T.tcells];
  const uint32_t ngroups = (n_tile + 31u) >> 5;
  const int HXY = T.HX * T.HY;
  uint32_t plane_max = 0;
  for (int z = 0; z < T.HZ; z++) plane_max = max(plane_max, tb.hstart[(z + 1) * HXY] - tb.hstart[z * HXY]);
  const bool bad = ngroups > (uint32_t)nwarp || ngroups > (uint32_t)tp.gmax || plane_max > (uint32_t)cap_pl || T.tcells != 1;   // cannot happen after a successful build
  if (bad && threadIdx.x == 0) atomicOr(err, DERR_TILE_CAPACITY);

  const uint32_t t = threadIdx.x;                               // tile particle: group = warp
  const bool have_group = !bad && (uint32_t)warp < ngroups;
  const bool active = have_group && t < n_tile;
  const int hA = (int)tb.thalo[0];
  const uint32_t i = tb.hfirst[hA] + (active ? t : 0u);
  double xa = 0., ya = 0., za = 0.;
  double m = 1.0, ux = 0., uy = 0., uz = 0.;
  if (!bad && n_tile > 0) { xa = rx[i]; ya = ry[i]; za = rz[i]; }
  if (MODE == 1 && active) { m = mass[type[i]]; if (dth != 0.0) { ux = vx[i]; uy = vy[i]; uz = vz[i]; } }
  PairAcc acc;
  acc.e = acc.wxx = acc.wyy = acc.wzz = acc.wxy = acc.wxz = acc.wyz = 0.;
  acc.ax = acc.ay = acc.az = 0.;
  const double rc2 = lj.rcut2();
  const uint32_t xyb = (uint32_t)__cvta_generic_to_shared(XY), zb = (uint32_t)__cvta_generic_to_shared(Z);
  uint2 ge = make_uint2(0u, 0u);
  if (have_group) ge = groups[(size_t)tile * (size_t)tp.gmax + (size_t)warp];
  const uint2* R = rows + ((size_t)ge.x * 32u + (uint32_t)lane);
  const uint32_t padw = ((uint32_t)cap_pl << 3) | ((uint32_t)cap_pl << 19);        // four sentinel candidates

  if (!bad && n_tile > 0)
  for (int p = 0; p < T.HZ; p++)
  {
    const uint32_t trips = (ge.y >> (10 * p)) & 1023u;
    // first row of the segment travels while the plane is staged
    uint2 w0 = make_uint2(padw, padw), w1 = w0;
    if (trips > 0u) w0 = ld_stream8(R);
    for (uint32_t k = 1; k < min(trips, CL_PREFETCH_ROWS); k++) asm volatile("prefetch.global.L2 [%0];" :: "l"(R + (size_t)k * 32u));
    __syncthreads();                                            // everybody is done with the previous plane
    const uint32_t pl0 = tb.hstart[p * HXY];
    for (int h = p * HXY + warp; h < (p + 1) * HXY; h += nwarp)
    {
      const uint32_t d0 = tb.hstart[h] - pl0, cnt = tb.hstart[h + 1] - tb.hstart[h], s0 = tb.hfirst[h];
      for (uint32_t q = lane; q < cnt; q += 32)
      {
        cp_async8(&XY[d0 + q].x, rx + s0 + q); cp_async8(&XY[d0 + q].y, ry + s0 + q); cp_async8(Z + d0 + q, rz + s0 + q);
      }
    }
    if (threadIdx.x == 0) { XY[cap_pl] = make_double2(1e30, 1e30); Z[cap_pl] = 1e30; }      // the sentinel every pad points at
    cp_async_wait_all();
    __syncthreads();
    for (uint32_t k = 0; k < trips; k++)
    {
      if (k + 1u < trips) w1 = ld_stream8(R + (size_t)(k + 1u) * 32u);
      if (k + CL_PREFETCH_ROWS < trips) asm volatile("prefetch.global.L2 [%0];" :: "l"(R + (size_t)(k + CL_PREFETCH_ROWS) * 32u));
      __syncwarp();
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
#pragma unroll
      for (int u = 0; u < 4; u++) ok[u] = in_cut(d2[u], rc2);
      pair_apply4<EV>(lj, dx, dy, dz, d2, ok, j, acc);
      w0 = w1;
    }
    R += (size_t)trips * 32u;
  }
  bool over = false;
  if (active)
  {
    double ax = acc.ax, ay = acc.ay, az = acc.az;
    if (MODE == 0) { fx[i] += ax; fy[i] += ay; fz[i] += az; }
    else
    {
      if (MODE == 2) { m = mass[type[i]]; ux = vx[i]; uy = vy[i]; uz = vz[i]; }
      ax = __ddiv_rn(ax, m); ay = __ddiv_rn(ay, m); az = __ddiv_rn(az, m);
      fx[i] = ax; fy[i] = ay; fz[i] = az;
      if (MODE == 2)
      {
        // second half of this step, then the first half of the next (as k_lj_sweep_cl MODE 2)
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
  if (MODE != 0)
  {
    // zero_particle_force{ghost:true}: ghost particles keep f = 0 (each block clears its slice of the ghost range)
    const int ng = n_total - n_inner;
    const int per = (ng + (int)gridDim.x - 1) / (int)gridDim.x;
    const int g0 = n_inner + (int)blockIdx.x * per, g1 = min(g0 + per, n_total);
    for (int q = g0 + (int)threadIdx.x; q < g1; q += blockDim.x) { fx[q] = 0.; fy[q] = 0.; fz[q] = 0.; }
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

} // namespace xnb
