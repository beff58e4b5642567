// xnb_kernels.cuh -- hand-written sm_100a kernels of the exaNBody LJ hot path.
//
//   K1  binning          : k_bin_locate -> scan -> k_bin_scatter -> k_cell_sort -> k_gather      (move_particles)
//   K1c in-cell sub-grid : k_amr_sizes -> scan -> k_cell_sort<AMR> -> k_gather                    (rebuild_amr)
//   K1e backup           : k_backup_r                                                             (backup_r)
//   K6/7 ghosts          : k_ghost_count -> scan -> k_ghost_fill -> k_ghost_cells, k_ghost_pack   (ghost_comm_scheme, ghost_update_*)
//   K2  neighbour build  : k_nbh_bits (xnb_nbh_bits.cuh); here the per-particle two-pass fallback
//                          k_nbh_build<COUNT> -> k_nbh_cell_sizes -> scan -> k_nbh_build<FILL>    (chunk_neighbors)
//   K3  pair sweep       : k_lj_force<...>                                                         (lennard_jones_force [+ fused epilogue])
//   K4/5 integrate       : k_verlet_first_half, k_push_f_v_r, k_push_f_v, k_displ_over, ...
//
// Data layout (HBM): flat SoA over particles; inner particles [0,n_inner) sorted by local cell index, ghost particles
// [n_inner,n_total) grouped by ghost cell; cell_start[c]/cell_count[c] give each cell's slice (the per-cell SoA view
// of the reference's CellParticles).  Neighbour streams live in one u16 pool, one 16-byte aligned slice per cell in
// the exact GridChunkNeighbors format.
#pragma once
#include "xnb_common.cuh"
#include "xnb_pair_functor.cuh"
#include <type_traits>

namespace xnb {

// ------------------------------------------------------------------------------------------------------------------
// exclusive scan (three-kernel, any length)
// ------------------------------------------------------------------------------------------------------------------
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

template <class T>
XNB_DEVINL T block_exclusive_scan(T v, T* total, T* smem /* 32 entries */)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  T x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { T y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) smem[warp] = x;
  __syncthreads();
  if (warp == 0)
  {
    T w = (lane < nwarp) ? smem[lane] : T(0);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { T y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
    smem[lane] = w;   // inclusive over warps
  }
  __syncthreads();
  const T warp_off = (warp > 0) ? smem[warp - 1] : T(0);
  *total = smem[nwarp - 1];
  __syncthreads();
  return warp_off + x - v;
}

template <class TIn, class T>
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_tiles(const TIn* __restrict__ in, T* __restrict__ out, T* __restrict__ tile_sums, size_t n)
{
  __shared__ T sm[32];
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  T v[SCAN_ITEMS]; T s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < n) ? (T)in[base + k] : T(0); s += v[k]; }
  T total;
  T off = block_exclusive_scan<T>(s, &total, sm);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = off; off += v[k]; }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

template <class T>
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_sums(T* __restrict__ tile_sums, size_t n_tiles, T* __restrict__ grand_total)
{
  __shared__ T sm[32];
  T carry = 0;
  for (size_t base = 0; base < n_tiles; base += SCAN_TILE)
  {
    const size_t b = base + (size_t)threadIdx.x * SCAN_ITEMS;
    T v[SCAN_ITEMS]; T s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (b + k < n_tiles) ? tile_sums[b + k] : T(0); s += v[k]; }
    T total;
    T off = carry + block_exclusive_scan<T>(s, &total, sm);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (b + k < n_tiles) tile_sums[b + k] = off; off += v[k]; }
    carry += total;
  }
  if (threadIdx.x == 0 && grand_total) *grand_total = carry;
}

template <class T>
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_add(T* __restrict__ out, const T* __restrict__ tile_sums, size_t n)
{
  const T add = tile_sums[blockIdx.x];
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) out[base + k] += add;
}

// ------------------------------------------------------------------------------------------------------------------
// K1 binning.  reference: move_particles_across_cells.h:104-156 (wrap + locate, domain.h:160-191)
// ------------------------------------------------------------------------------------------------------------------
constexpr uint32_t KEY_LEAVING = 0xFFFFFFFFu;

// periodic wrap of one coordinate and domain cell location, bit-identical to the oracle's domain_periodic_location
XNB_DEVINL int locate_axis(double& r, double dmin, double cs, int ddim, int periodic)
{
  const double rel = __dadd_rn(r, -dmin);
  long long loc = (long long)floor(__ddiv_rn(rel, cs));
  if ((loc < 0 || loc >= ddim) && periodic)
  {
    const long long o = loc;
    loc = ((loc % ddim) + ddim) % ddim;
    r = __dadd_rn(r, __dmul_rn((double)(loc - o), cs));
  }
  return (int)max(min(loc, (long long)INT32_MAX), (long long)INT32_MIN);
}

// one thread per inner particle: wrap position in place, compute destination local cell, count per cell.
// rank[i] = arrival order inside the destination cell (arbitrary; fixed afterwards by k_cell_sort).
// Particles whose destination is outside this rank's inner block get KEY_LEAVING and are appended to leave_list.
__global__ void k_bin_locate(GridP g, int n, double* __restrict__ rx, double* __restrict__ ry, double* __restrict__ rz,
                             uint32_t* __restrict__ key, uint32_t* __restrict__ rank, uint32_t* __restrict__ cell_count,
                             uint32_t* __restrict__ leave_list, uint32_t* __restrict__ leave_count, uint32_t* __restrict__ err)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = rx[i], y = ry[i], z = rz[i];
  const int li = locate_axis(x, g.dmin[0], g.cs, g.ddims[0], g.periodic[0]);
  const int lj = locate_axis(y, g.dmin[1], g.cs, g.ddims[1], g.periodic[1]);
  const int lk = locate_axis(z, g.dmin[2], g.cs, g.ddims[2], g.periodic[2]);
  rx[i] = x; ry[i] = y; rz[i] = z;
  if (li < 0 || li >= g.ddims[0] || lj < 0 || lj >= g.ddims[1] || lk < 0 || lk >= g.ddims[2])
  {
    atomicOr(err, DERR_LOST_PARTICLE);   // left a non periodic domain: reference drops it into otb_particles for good
    key[i] = KEY_LEAVING; rank[i] = 0;
    return;
  }
  if (li < g.bstart[0] || li >= g.bend[0] || lj < g.bstart[1] || lj >= g.bend[1] || lk < g.bstart[2] || lk >= g.bend[2])
  {
    // otb_particles (move_particles_across_cells.h:124-138): handed to migrate (multi-GPU)
    key[i] = KEY_LEAVING;
    rank[i] = atomicAdd(leave_count, 1u);
    if (leave_list) leave_list[rank[i]] = (uint32_t)i;
    return;
  }
  const int c = ijk_to_index(g.dims, li - g.off[0], lj - g.off[1], lk - g.off[2]);
  key[i] = (uint32_t)c;
  rank[i] = atomicAdd(&cell_count[c], 1u);
}

// cell_start for INNER cells from the exclusive scan of counts over all local cells (ghost cells hold 0 here)
__global__ void k_bin_scatter(int n, const uint32_t* __restrict__ key, const uint32_t* __restrict__ rank,
                              const uint32_t* __restrict__ cell_start, uint32_t* __restrict__ perm)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t c = key[i];
  if (c == KEY_LEAVING) return;
  perm[cell_start[c] + rank[i]] = (uint32_t)i;
}

XNB_DEVINL void cell_origin(const GridP& g, uint32_t c, double& ox, double& oy, double& oz)
{
  const int ci = c % g.dims[0], cj = (c / g.dims[0]) % g.dims[1], ck = c / (g.dims[0] * g.dims[1]);
  ox = __dadd_rn(g.org[0], __dmul_rn((double)(g.off[0] + ci), g.cs));
  oy = __dadd_rn(g.org[1], __dmul_rn((double)(g.off[1] + cj), g.cs));
  oz = __dadd_rn(g.org[2], __dmul_rn((double)(g.off[2] + ck), g.cs));
}

// ------------------------------------------------------------------------------------------------------------------
// in-cell ordering.  One block per cell.  perm[cell_start[c] .. +n) holds source indices; they are re-ordered by
//   AMR=false : particle id  (the reference's in-cell order after move_particles depends on OpenMP scheduling, i.e. is
//               unspecified; ordering by id makes ours deterministic AND independent of the domain decomposition)
//   AMR=true  : (sub-cell index, particle id): the sub-cell grouping of project_particles_in_sub_grids
//               (amr_grid_algorithm.h:188-299); sub_grid_cells[] gets the cumulative end offsets (:283-292)
// ------------------------------------------------------------------------------------------------------------------
constexpr int CELLSORT_MAX = 2048;   // particles per cell sorted out of shared memory; larger cells (the format allows 65535) rank through global scratch
constexpr int CELLSORT_THREADS = 128;

// big_keys / big_srcs: global scratch of n_src entries for cells beyond CELLSORT_MAX (slices [cell_start, cell_start + n) are disjoint)
// k_cell_sort for cells of at most 64 particles (C2, C3, C5: the usual case), one WARP per cell: two keys per lane in registers,
// rank by shuffles.  A fuller cell (the occupancy may have changed since the host last looked) is still sorted correctly, by the same warp
// through the global scratch (slow, rare).  In-cell order = ascending particle id, as k_cell_sort<false>.
__global__ void __launch_bounds__(256)
k_cell_sort_warp(int n_cells, const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_count,
                 const uint32_t* __restrict__ perm_in, uint32_t* __restrict__ perm_out, const unsigned long long* __restrict__ id,
                 unsigned long long* __restrict__ big_keys, uint32_t* __restrict__ big_srcs, uint32_t* __restrict__ err)
{
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int c = blockIdx.x * wpb + (threadIdx.x >> 5); c < n_cells; c += gridDim.x * wpb)
  {
    const int n = (int)cell_count[c];
    if (n == 0) continue;
    if (n > 65535) { if (lane == 0) atomicOr(err, DERR_CELL_OVERFLOW); continue; }
    const uint32_t s0 = cell_start[c];
    if (n <= 64)
    {
      // two keys per lane (particles lane and lane + 32); rank = number of smaller keys among the n
      uint32_t src0 = 0, src1 = 0; unsigned long long key0 = ~0ull, key1 = ~0ull;
      if (lane < n)
      {
        src0 = perm_in[s0 + lane];
        const unsigned long long pid = id[src0];
        if (pid >> 52) atomicOr(err, DERR_ID_RANGE);
        key0 = pid & ((1ull << 52) - 1ull);
      }
      if (lane + 32 < n)
      {
        src1 = perm_in[s0 + lane + 32];
        const unsigned long long pid = id[src1];
        if (pid >> 52) atomicOr(err, DERR_ID_RANGE);
        key1 = pid & ((1ull << 52) - 1ull);
      }
      int r0 = 0, r1 = 0;
      const int na = min(n, 32);
      for (int u = 0; u < na; u++) { const unsigned long long k = __shfl_sync(0xffffffffu, key0, u); r0 += (k < key0) ? 1 : 0; r1 += (k < key1) ? 1 : 0; }
      for (int u = 32; u < n; u++) { const unsigned long long k = __shfl_sync(0xffffffffu, key1, u - 32); r0 += (k < key0) ? 1 : 0; r1 += (k < key1) ? 1 : 0; }
      if (lane < n) perm_out[s0 + r0] = src0;
      if (lane + 32 < n) perm_out[s0 + r1] = src1;
      continue;
    }
    for (int t = lane; t < n; t += 32)
    {
      const uint32_t src = perm_in[s0 + t];
      const unsigned long long pid = id[src];
      if (pid >> 52) atomicOr(err, DERR_ID_RANGE);
      big_keys[s0 + t] = pid & ((1ull << 52) - 1ull); big_srcs[s0 + t] = src;
    }
    __syncwarp();
    for (int t = lane; t < n; t += 32)
    {
      const unsigned long long my = big_keys[s0 + t];
      int r = 0;
      for (int u = 0; u < n; u++) r += (big_keys[s0 + u] < my) ? 1 : 0;
      perm_out[s0 + r] = big_srcs[s0 + t];
    }
  }
}

template <bool AMR>
__global__ void __launch_bounds__(CELLSORT_THREADS)
k_cell_sort(GridP g, const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_count,
            const uint32_t* __restrict__ perm_in, uint32_t* __restrict__ perm_out,
            const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz,
            const unsigned long long* __restrict__ id,
            const uint8_t* __restrict__ side_lut, const unsigned long long* __restrict__ sub_grid_start,
            uint32_t* __restrict__ sub_grid_cells, unsigned long long* __restrict__ big_keys, uint32_t* __restrict__ big_srcs, uint32_t* __restrict__ err)
{
  __shared__ unsigned long long s_keys[CELLSORT_MAX];
  __shared__ uint32_t s_srcs[CELLSORT_MAX];
  __shared__ uint32_t hist[AMR ? 4096 : 1];
  const int c = blockIdx.x;
  const int n = (int)cell_count[c];
  if (n == 0) return;
  if (n > 65535) { if (threadIdx.x == 0) atomicOr(err, DERR_CELL_OVERFLOW); return; }
  const uint32_t s0 = cell_start[c];
  const bool big = n > CELLSORT_MAX;
  if (big && !big_keys) { if (threadIdx.x == 0) atomicOr(err, DERR_SORT_CAPACITY); return; }
  unsigned long long* const keys = big ? big_keys + s0 : s_keys;
  uint32_t* const srcs = big ? big_srcs + s0 : s_srcs;
  int side = 1;
  if (AMR)
  {
    side = side_lut[min(n, 65535)];
    if (side < 1) side = 1;
  }
  const int nsub = side * side * side;
  if (AMR) { for (int q = threadIdx.x; q < nsub; q += blockDim.x) hist[q] = 0; __syncthreads(); }
  // cell low corner: origin + (offset+loc)*cell_size   (grid.h:113-116, oracle Grid::cell_position)
  double lx, ly, lz;
  cell_origin(g, (uint32_t)c, lx, ly, lz);
  for (int t = threadIdx.x; t < n; t += blockDim.x)
  {
    const uint32_t src = perm_in[s0 + t];
    unsigned long long sub = 0;
    if (AMR && side > 1)
    {
      // particle_pcoord (grid.h:184-190) then trunc(p*side) clamped (amr_grid_algorithm.h:188-205)
      const double px = __ddiv_rn(__dadd_rn(rx[src], -lx), g.cs), py = __ddiv_rn(__dadd_rn(ry[src], -ly), g.cs), pz = __ddiv_rn(__dadd_rn(rz[src], -lz), g.cs);
      long long si = (long long)__dmul_rn(px, (double)side), sj = (long long)__dmul_rn(py, (double)side), sk = (long long)__dmul_rn(pz, (double)side);
      si = max(0ll, min(si, (long long)side - 1)); sj = max(0ll, min(sj, (long long)side - 1)); sk = max(0ll, min(sk, (long long)side - 1));
      sub = (unsigned long long)((sk * side + sj) * side + si);
      atomicAdd(&hist[(int)sub], 1u);
    }
    const unsigned long long pid = id[src];
    if (pid >> 52) atomicOr(err, DERR_ID_RANGE);
    keys[t] = (sub << 52) | (pid & ((1ull << 52) - 1ull));
    srcs[t] = src;
  }
  __syncthreads();
  // rank sort (32..256 particles per cell in the benchmark configurations); ids are unique so the ranks are a permutation
  for (int t = threadIdx.x; t < n; t += blockDim.x)
  {
    const unsigned long long my = keys[t];
    int r = 0;
    for (int u = 0; u < n; u++) r += (keys[u] < my) ? 1 : 0;
    perm_out[s0 + r] = srcs[t];
  }
  if (AMR && side > 1)
  {
    // cumulative END offsets of sub-cells 0..nsub-2
    __syncthreads();
    if (threadIdx.x == 0)
    {
      const unsigned long long sg0 = sub_grid_start[c];
      uint32_t acc = 0;
      for (int q = 0; q < nsub - 1; q++) { acc += hist[q]; sub_grid_cells[sg0 + q] = acc; }
    }
  }
}

// number of sub_grid_cells entries per cell: max(side^3-1,0), inner cells only (ghost cells are empty when rebuild_amr
// runs in the reference: update-particles.msp:48-52), reference amr_grid_algorithm.h:400-417
__global__ void k_amr_sizes(GridP g, const uint32_t* __restrict__ cell_count, const uint8_t* __restrict__ side_lut,
                            uint32_t* __restrict__ sg_size, uint32_t* __restrict__ max_side)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= g.n_cells) return;
  const int ci = c % g.dims[0], cj = (c / g.dims[0]) % g.dims[1], ck = c / (g.dims[0] * g.dims[1]);
  const bool inner = ci >= g.gl && ci < g.dims[0] - g.gl && cj >= g.gl && cj < g.dims[1] - g.gl && ck >= g.gl && ck < g.dims[2] - g.gl;
  int side = 0;
  if (inner) side = side_lut[min(cell_count[c], 65535u)];
  sg_size[c] = (side > 1) ? (uint32_t)(side * side * side - 1) : 0u;
  if (side > 1) atomicMax(max_side, (uint32_t)side);
}

// gather all particle fields through a permutation (dest j takes source perm[j]); also writes atom_cell
// (n_src: entries of the source arrays; a slot whose perm entry was never written -- error paths: lost particles,
// in-cell sort capacity -- holds the 0xFFFFFFFF sentinel and is skipped, the error word tells the host)
__global__ void k_gather(int n, uint32_t n_src, const uint32_t* __restrict__ perm, ParticlesP src, ParticlesP dst,
                         const uint32_t* __restrict__ key_src, uint32_t* __restrict__ atom_cell)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t s = perm[j];
  if (s >= n_src) return;
  dst.rx[j] = src.rx[s]; dst.ry[j] = src.ry[s]; dst.rz[j] = src.rz[s];
  dst.vx[j] = src.vx[s]; dst.vy[j] = src.vy[s]; dst.vz[j] = src.vz[s];
  dst.fx[j] = src.fx[s]; dst.fy[j] = src.fy[s]; dst.fz[j] = src.fz[s];
  dst.id[j] = src.id[s]; dst.type[j] = src.type[s];
  atom_cell[j] = key_src[s];
}

// identity permutation restricted to cell slices (used by rebuild_amr, where particles are already cell sorted)
__global__ void k_iota(int n, uint32_t* __restrict__ p) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = (uint32_t)i; }

__global__ void k_fill_u32(size_t n, uint32_t* __restrict__ p, uint32_t v) { const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

// ------------------------------------------------------------------------------------------------------------------
// backup_r.  reference: io/backup_r.cpp:56-78, core/backup_r.h:36-51
// ------------------------------------------------------------------------------------------------------------------
XNB_DEVINL uint32_t encode_double_u32(double x, double o, double r)
{
  const double xo = __dadd_rn(x, -o);
  long long q = (long long)__ddiv_rn(__dmul_rn(xo, 4294967296.0), r);
  q = max(0ll, min(q, 4294967295ll));
  const uint32_t a = (uint32_t)q, b = a - 1u, c = a + 1u;
  const double ea = fabs(__dadd_rn(restore_u32_double(a, o, r), -x));
  const double eb = fabs(__dadd_rn(restore_u32_double(b, o, r), -x));
  const double ec = fabs(__dadd_rn(restore_u32_double(c, o, r), -x));
  if (eb < ea) return b;
  if (ec < ea) return c;
  return a;
}

__global__ void k_backup_r(GridP g, int n_inner, const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz,
                           const uint32_t* __restrict__ atom_cell, uint32_t* __restrict__ backup)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_inner) return;
  double ox, oy, oz;
  cell_origin(g, atom_cell[i], ox, oy, oz);
  backup[3 * (size_t)i + 0] = encode_double_u32(rx[i], ox, g.cs);
  backup[3 * (size_t)i + 1] = encode_double_u32(ry[i], oy, g.cs);
  backup[3 * (size_t)i + 2] = encode_double_u32(rz[i], oz, g.cs);
}

// displacement test of one particle (particle_displ_over.cu:47-65): |r - restore(backup)|^2 >= threshold^2
XNB_DEVINL bool displ_over_test(const GridP& g, uint32_t cell, const uint32_t* __restrict__ rb, double x, double y, double z, double thr2)
{
  double ox, oy, oz;
  cell_origin(g, cell, ox, oy, oz);
  const double dx = __dadd_rn(x, -restore_u32_double(rb[0], ox, g.cs));
  const double dy = __dadd_rn(y, -restore_u32_double(rb[1], oy, g.cs));
  const double dz = __dadd_rn(z, -restore_u32_double(rb[2], oz, g.cs));
  return norm2_exact(dx, dy, dz) >= thr2;
}

XNB_DEVINL void block_count_add(bool flag, unsigned long long* counter)
{
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(counter, (unsigned long long)__popc(m));
}

__global__ void k_displ_over(GridP g, int n_inner, const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz,
                             const uint32_t* __restrict__ atom_cell, const uint32_t* __restrict__ backup, double thr2,
                             unsigned long long* __restrict__ counter)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool over = false;
  if (i < n_inner) over = displ_over_test(g, atom_cell[i], backup + 3 * (size_t)i, rx[i], ry[i], rz[i], thr2);
  block_count_add(over, counter);
}

// ------------------------------------------------------------------------------------------------------------------
// integrators.  reference: push_vec3_2nd_order.h:29-39, push_vec3_1st_order.h:29-38 (identity xform).
// Arithmetic is written with explicit round-to-nearest ops in the reference's order so that it is bit-identical to
// the oracle (which is compiled with -ffp-contract=off).
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_push_f_v_r(int n, double dt, double dt2, double* __restrict__ rx, double* __restrict__ ry, double* __restrict__ rz,
                             const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ vz,
                             const double* __restrict__ fx, const double* __restrict__ fy, const double* __restrict__ fz)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  rx[i] = __dadd_rn(rx[i], __dadd_rn(__dmul_rn(vx[i], dt), __dmul_rn(fx[i], dt2)));
  ry[i] = __dadd_rn(ry[i], __dadd_rn(__dmul_rn(vy[i], dt), __dmul_rn(fy[i], dt2)));
  rz[i] = __dadd_rn(rz[i], __dadd_rn(__dmul_rn(vz[i], dt), __dmul_rn(fz[i], dt2)));
}

__global__ void k_push_f_v(int n, double dt, double* __restrict__ vx, double* __restrict__ vy, double* __restrict__ vz,
                           const double* __restrict__ fx, const double* __restrict__ fy, const double* __restrict__ fz)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  vx[i] = __dadd_rn(vx[i], __dmul_rn(fx[i], dt));
  vy[i] = __dadd_rn(vy[i], __dmul_rn(fy[i], dt));
  vz[i] = __dadd_rn(vz[i], __dmul_rn(fz[i], dt));
}

// K4: verlet_first_half (push_f_v_r{1.0} + push_f_v{0.5}) fused with the particle_displ_over reduction.
__global__ void k_verlet_first_half(GridP g, int n, double dt, double dt2, double dth,
                                    double* __restrict__ rx, double* __restrict__ ry, double* __restrict__ rz,
                                    double* __restrict__ vx, double* __restrict__ vy, double* __restrict__ vz,
                                    const double* __restrict__ fx, const double* __restrict__ fy, const double* __restrict__ fz,
                                    const uint32_t* __restrict__ atom_cell, const uint32_t* __restrict__ backup, double thr2,
                                    unsigned long long* __restrict__ counter)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool over = false;
  if (i < n)
  {
    const double ax = fx[i], ay = fy[i], az = fz[i];
    const double ux = vx[i], uy = vy[i], uz = vz[i];
    const double x = __dadd_rn(rx[i], __dadd_rn(__dmul_rn(ux, dt), __dmul_rn(ax, dt2)));
    const double y = __dadd_rn(ry[i], __dadd_rn(__dmul_rn(uy, dt), __dmul_rn(ay, dt2)));
    const double z = __dadd_rn(rz[i], __dadd_rn(__dmul_rn(uz, dt), __dmul_rn(az, dt2)));
    rx[i] = x; ry[i] = y; rz[i] = z;
    vx[i] = __dadd_rn(ux, __dmul_rn(ax, dth));
    vy[i] = __dadd_rn(uy, __dmul_rn(ay, dth));
    vz[i] = __dadd_rn(uz, __dmul_rn(az, dth));
    over = displ_over_test(g, atom_cell[i], backup + 3 * (size_t)i, x, y, z, thr2);
  }
  block_count_add(over, counter);
}

__global__ void k_zero_force(int n, double* __restrict__ fx, double* __restrict__ fy, double* __restrict__ fz)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { fx[i] = 0.; fy[i] = 0.; fz[i] = 0.; }
}

// divide_force_by_type_scalar: mass (vec3_typescalar_op.cu:118-122, math_functors.h:136-141): a true division per component
__global__ void k_divide_force_by_mass(int n, double* __restrict__ fx, double* __restrict__ fy, double* __restrict__ fz,
                                       const uint8_t* __restrict__ type, const double* __restrict__ mass)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double m = mass[type[i]];
  fx[i] = __ddiv_rn(fx[i], m); fy[i] = __ddiv_rn(fy[i], m); fz[i] = __ddiv_rn(fz[i], m);
}

// ------------------------------------------------------------------------------------------------------------------
// ghosts.  reference: update_ghosts_comm_scheme.cpp:403-481 (particle selection), update_ghost_functors.h:172-254 (pack)
// A "send item" = (my inner cell, partner, partner ghost cell, boundary flags); the item list is static for a static
// decomposition and is built on the host (xnb_ctx.cu); per rebuild the device selects particles.
// ------------------------------------------------------------------------------------------------------------------
struct GhostItemsP
{
  const uint32_t* src_cell;   // my cell
  const uint32_t* flags;      // GhostBoundaryModifier flags
  const uint32_t* partner;    // partner slot -> outer bounds
  const double* outer;        // [n_partner_slots][6] partner inner bounds enlarged by ghost_dist (lo xyz, hi xyz)
  int n_items;
};

XNB_DEVINL bool ghost_selected(const GridP& g, const double* __restrict__ ob, uint32_t fl, double x, double y, double z)
{
  const double gx = coord_shift(x, g.dmin[0], g.dmax[0], fl >> 0);
  const double gy = coord_shift(y, g.dmin[1], g.dmax[1], fl >> 3);
  const double gz = coord_shift(z, g.dmin[2], g.dmax[2], fl >> 6);
  return gx >= ob[0] && gx <= ob[3] && gy >= ob[1] && gy <= ob[4] && gz >= ob[2] && gz <= ob[5];
}

// one warp per item: count selected particles
__global__ void k_ghost_count(GridP g, GhostItemsP it, const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_count,
                              const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz,
                              uint32_t* __restrict__ item_count)
{
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= it.n_items) return;
  const uint32_t c = it.src_cell[w], fl = it.flags[w];
  const double* ob = it.outer + 6 * it.partner[w];
  const uint32_t s0 = cell_start[c], n = cell_count[c];
  uint32_t cnt = 0;
  for (uint32_t p = lane; p < ((n + 31u) & ~31u); p += 32)
  {
    const bool sel = (p < n) && ghost_selected(g, ob, fl, rx[s0 + p], ry[s0 + p], rz[s0 + p]);
    cnt += __popc(__ballot_sync(0xffffffffu, sel));
  }
  if (lane == 0) item_count[w] = cnt;
}

// one warp per item: write the selected particle indices (in-cell order preserved, update_ghosts_comm_scheme.cpp:462-470)
__global__ void k_ghost_fill(GridP g, GhostItemsP it, const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_count,
                             const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz,
                             const uint32_t* __restrict__ item_offset, uint32_t* __restrict__ send_src, uint16_t* __restrict__ send_flags)
{
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= it.n_items) return;
  const uint32_t c = it.src_cell[w], fl = it.flags[w];
  const double* ob = it.outer + 6 * it.partner[w];
  const uint32_t s0 = cell_start[c], n = cell_count[c];
  uint32_t base = item_offset[w];
  for (uint32_t p = lane; p < ((n + 31u) & ~31u); p += 32)
  {
    const bool sel = (p < n) && ghost_selected(g, ob, fl, rx[s0 + p], ry[s0 + p], rz[s0 + p]);
    const unsigned m = __ballot_sync(0xffffffffu, sel);
    if (sel) { const uint32_t k = base + __popc(m & ((1u << lane) - 1u)); send_src[k] = s0 + p; send_flags[k] = (uint16_t)fl; }
    base += __popc(m);
  }
}

// receive side: one warp per received item. Sets the ghost cell slice and the cell id of its particles.
__global__ void k_ghost_cells(int n_items, const uint32_t* __restrict__ dst_cell, const uint32_t* __restrict__ recv_count,
                              const uint32_t* __restrict__ recv_offset, uint32_t n_inner,
                              uint32_t* __restrict__ cell_start, uint32_t* __restrict__ cell_count, uint32_t* __restrict__ atom_cell)
{
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_items) return;
  const uint32_t c = dst_cell[w], n = recv_count[w], s0 = n_inner + recv_offset[w];
  if (lane == 0 && n > 0) { cell_start[c] = s0; cell_count[c] = n; }
  for (uint32_t p = lane; p < n; p += 32) atom_cell[s0 + p] = c;
}

// reset ghost cells to empty (migrate_cell_particles leaves them empty: mpi/migrate_cell_particles.cpp:101-110)
__global__ void k_ghost_cells_clear(GridP g, uint32_t n_inner, uint32_t* __restrict__ cell_start, uint32_t* __restrict__ cell_count)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= g.n_cells) return;
  const int ci = c % g.dims[0], cj = (c / g.dims[0]) % g.dims[1], ck = c / (g.dims[0] * g.dims[1]);
  const bool inner = ci >= g.gl && ci < g.dims[0] - g.gl && cj >= g.gl && cj < g.dims[1] - g.gl && ck >= g.gl && ck < g.dims[2] - g.gl;
  if (!inner) { cell_start[c] = n_inner; cell_count[c] = 0; }
}

// pack: ghost g takes particle send_src[g] with the periodic shift applied (update_ghost_functors.h:41-67,172-217).
// Sends to myself (periodic images, grid_update_ghosts.h:176-187) go straight into the ghost slots [self_dst ...);
// the others go to the staging buffer, ONE CONTIGUOUS SLAB PER PARTNER (one message per partner, as the reference's
// UpdateGhostsCommManager ships one buffer per partner: update_ghosts_comm_manager.h:63-75,267-281): partner p's send-list
// entries [s0, s0 + sn) occupy 8-byte words [NW s0, NW (s0 + sn)), field-major inside the slab:
//   x[sn] y[sn] z[sn]                                     (NW = 3,  ghost_update_r)
//   x y z vx vy vz fx fy fz id type(one word each)        (NW = 11, ghost_update_all)
constexpr int GHOST_WORDS_R = 3, GHOST_WORDS_ALL = 11;
// partner whose range [base[p], base[p+1]) holds entry q (base has nranks + 1 entries)
XNB_DEVINL int ghost_partner_of(const uint32_t* __restrict__ base, int nranks, uint32_t q)
{
  int lo = 0, hi = nranks - 1;
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (q >= base[mid]) lo = mid; else hi = mid - 1; }
  return lo;
}
template <bool ALL_FIELDS>
__global__ void k_ghost_pack(GridP g, int n_send, const uint32_t* __restrict__ send_src, const uint16_t* __restrict__ send_flags,
                             ParticlesP p, int self_first, int self_end, uint32_t self_dst,
                             const uint32_t* __restrict__ send_base, int nranks, double* __restrict__ stage)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_send) return;
  const uint32_t s = send_src[q], fl = send_flags[q];
  const double x = coord_shift(p.rx[s], g.dmin[0], g.dmax[0], fl >> 0);
  const double y = coord_shift(p.ry[s], g.dmin[1], g.dmax[1], fl >> 3);
  const double z = coord_shift(p.rz[s], g.dmin[2], g.dmax[2], fl >> 6);
  if (q >= self_first && q < self_end)
  {
    const uint32_t d = self_dst + (uint32_t)(q - self_first);
    p.rx[d] = x; p.ry[d] = y; p.rz[d] = z;
    if (ALL_FIELDS)
    {
      p.vx[d] = p.vx[s]; p.vy[d] = p.vy[s]; p.vz[d] = p.vz[s];
      p.fx[d] = p.fx[s]; p.fy[d] = p.fy[s]; p.fz[d] = p.fz[s];
      p.id[d] = p.id[s]; p.type[d] = p.type[s];
    }
  }
  else
  {
    constexpr size_t NW = ALL_FIELDS ? GHOST_WORDS_ALL : GHOST_WORDS_R;
    const int pr = ghost_partner_of(send_base, nranks, (uint32_t)q);
    const size_t s0 = send_base[pr], n = send_base[pr + 1] - s0;
    double* slab = stage + NW * s0 + ((size_t)q - s0);
    slab[0] = x; slab[n] = y; slab[2 * n] = z;
    if (ALL_FIELDS)
    {
      slab[3 * n] = p.vx[s]; slab[4 * n] = p.vy[s]; slab[5 * n] = p.vz[s];
      slab[6 * n] = p.fx[s]; slab[7 * n] = p.fy[s]; slab[8 * n] = p.fz[s];
      reinterpret_cast<unsigned long long*>(slab)[9 * n] = p.id[s];
      reinterpret_cast<unsigned long long*>(slab)[10 * n] = (unsigned long long)p.type[s];
    }
  }
}
// unpack (GhostReceiveUnpackFunctor, update_ghost_functors.h:369-452): ghost i of partner p's slab -> ghost slot n_inner + i
template <bool ALL_FIELDS>
__global__ void k_ghost_unpack(int n_ghost, uint32_t n_inner, ParticlesP p, const uint32_t* __restrict__ recv_base, int nranks, int self_rank,
                               const double* __restrict__ rstage)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_ghost) return;
  const int pr = ghost_partner_of(recv_base, nranks, (uint32_t)i);
  if (pr == self_rank) return;                       // my own periodic images were written in place by the pack kernel
  constexpr size_t NW = ALL_FIELDS ? GHOST_WORDS_ALL : GHOST_WORDS_R;
  const size_t r0 = recv_base[pr], n = recv_base[pr + 1] - r0;
  const double* slab = rstage + NW * r0 + ((size_t)i - r0);
  const uint32_t d = n_inner + (uint32_t)i;
  p.rx[d] = slab[0]; p.ry[d] = slab[n]; p.rz[d] = slab[2 * n];
  if (ALL_FIELDS)
  {
    p.vx[d] = slab[3 * n]; p.vy[d] = slab[4 * n]; p.vz[d] = slab[5 * n];
    p.fx[d] = slab[6 * n]; p.fy[d] = slab[7 * n]; p.fz[d] = slab[8 * n];
    p.id[d] = reinterpret_cast<const unsigned long long*>(slab)[9 * n];
    p.type[d] = (uint8_t)reinterpret_cast<const unsigned long long*>(slab)[10 * n];
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K2 chunk neighbour build.  reference: chunk_neighbors_execute.h:110-411.
// One thread per particle (ALL cells, ghost cells included, :110).  Neighbour cells are visited in ascending (k,j,i)
// = ascending encoded cell id, particles in ascending p_b, so each particle's list comes out already sorted and unique:
// the reference's per-particle std::sort + dedup (:308-324) is a no-op for this traversal.  The sub-cell pair cache of
// the reference (amr_grid_pairs) only prunes candidates; the distance filter below decides membership.
//   COUNT pass: nb_len[i] = 1 + 2*G + N (u16 words of the particle's list), nb_cnt[i] = N
//   FILL  pass: writes the list and the particle's u32 offset-table entry
// ------------------------------------------------------------------------------------------------------------------
struct NbhOut
{
  uint32_t* nb_len;                    // per particle list length (u16 words)
  uint32_t* nb_cnt;                    // per particle neighbour count
  const uint32_t* nb_off;              // per particle offset of its list inside the cell's list area (u16 words)
  uint16_t* const* cell_stream;        // per cell stream base pointer (device)
};

template <bool FILL>
__global__ void __launch_bounds__(128)
k_nbh_build(GridP g, int n_total, int gap, double max_dist2, int half_symmetric, int skip_ghosts,
            const unsigned long long* __restrict__ sub_grid_start, const uint32_t* __restrict__ sub_grid_cells,
            const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz,
            const uint32_t* __restrict__ atom_cell, const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_count,
            NbhOut out, uint32_t* __restrict__ err)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_total) return;
  const uint32_t ca = atom_cell[i];
  const int ai = ca % g.dims[0], aj = (ca / g.dims[0]) % g.dims[1], ak = ca / (g.dims[0] * g.dims[1]);
  const double xa = rx[i], ya = ry[i], za = rz[i];
  const double nbh_reach = sub_grid_start ? sqrt(max_dist2) : 0.0;
  uint16_t* w = nullptr;
  if (FILL)
  {
    const uint32_t na = cell_count[ca], pa = (uint32_t)i - cell_start[ca];
    uint16_t* base = out.cell_stream[ca];
    const uint32_t off = out.nb_off[i];
    // offset table entry: u16 index of the list relative to the first list, + number of tables (1)  (:279-283)
    reinterpret_cast<uint32_t*>(base)[pa] = off + 1u;
    if (pa == na - 1u) reinterpret_cast<uint32_t*>(base)[na] = off + out.nb_len[i] + 1u;   // closing entry (:390-398)
    w = base + 2u * (na + 1u) + off;
  }
  uint32_t groups = 0, total = 0;
  uint16_t* wg = w;            // position of the group counter
  if (FILL) w++;
  const int k0 = max(ak - gap, 0), k1 = min(ak + gap, g.dims[2] - 1);
  const int j0 = max(aj - gap, 0), j1 = min(aj + gap, g.dims[1] - 1);
  const int i0 = max(ai - gap, 0), i1 = min(ai + gap, g.dims[0] - 1);
  for (int bk = k0; bk <= k1; bk++)
    for (int bj = j0; bj <= j1; bj++)
      for (int bi = i0; bi <= i1; bi++)
      {
        const int cb = ijk_to_index(g.dims, bi, bj, bk);
        uint32_t nb = cell_count[cb];
        // NeighborFilterHalfSymGhost (neighbor_filter_func.h:36-52): half_symmetric keeps b "before" a (cell_b < cell_a, or the
        // same cell and p_b < p_a); skip_ghosts drops every b of a ghost cell
        if (half_symmetric) { if ((uint32_t)cb > ca) continue; if ((uint32_t)cb == ca) nb = (uint32_t)i - cell_start[ca]; }
        if (skip_ghosts && (bi < g.gl || bi >= g.dims[0] - g.gl || bj < g.gl || bj >= g.dims[1] - g.gl || bk < g.gl || bk >= g.dims[2] - g.gl)) continue;
        if (nb == 0) continue;
        const uint32_t sb = cell_start[cb];
        uint32_t cnt = 0;
        uint16_t* wc = w;      // group header position (enc, n)
        // sub-cells of cell b (AmrGrid, amr_grid.h:31-61): the particles of a cell are stored sub-cell after sub-cell, so a
        // sub-cell whose box lies farther than the list radius from r_a is skipped as a whole.  This only prunes (the role of the
        // reference's AmrSubCellPairCache, amr_grid_algorithm.cpp:102-218): membership is still decided by the exact test below,
        // and sub-cells are visited in ascending order, i.e. ascending p_b.
        uint32_t nsub = 1u, side = 1u; unsigned long long sg0 = 0ull;
        if (sub_grid_start) { sg0 = sub_grid_start[cb]; nsub = (uint32_t)(sub_grid_start[cb + 1] - sg0) + 1u; while (side * side * side < nsub) side++; }
        // per axis, the sub-cell indices of cell b whose slab lies within the list radius of r_a (a box around the sphere; integer
        // tests per sub-cell, the sub-cells stay in lockstep across the warp so that the candidate loads remain broadcasts)
        int slo[3] = {0, 0, 0}, shi[3] = {0, 0, 0};
        if (nsub > 1u)
        {
          double box[3]; cell_origin(g, (uint32_t)cb, box[0], box[1], box[2]);
          const double inv_h = (double)side / g.cs, reach = nbh_reach + 1e-9 * g.cs;
          const double ra[3] = {xa, ya, za};
#pragma unroll
          for (int d = 0; d < 3; d++)
          {
            const double rel = ra[d] - box[d];
            slo[d] = max((int)floor((rel - reach) * inv_h), 0); shi[d] = min((int)floor((rel + reach) * inv_h), (int)side - 1);
          }
        }
        uint32_t si = 0u, sj = 0u, sk = 0u;
        for (uint32_t sc = 0; sc < nsub; sc++)
        {
          uint32_t p0 = 0u, p1 = nb;
          if (nsub > 1u)
          {
            const bool in_box = (int)si >= slo[0] && (int)si <= shi[0] && (int)sj >= slo[1] && (int)sj <= shi[1] && (int)sk >= slo[2] && (int)sk <= shi[2];
            if (++si == side) { si = 0u; if (++sj == side) { sj = 0u; ++sk; } }
            if (!in_box) continue;
            p0 = sc > 0u ? sub_grid_cells[sg0 + sc - 1u] : 0u;
            p1 = min(sc + 1u < nsub ? sub_grid_cells[sg0 + sc] : nb, nb);
          }
          for (uint32_t pb = p0; pb < p1; pb++)
          {
            const uint32_t j = sb + pb;
            // :225-227  dr = r_a - r_b ; d2 = |dr|^2 ; keep if not self, d2 > 0, d2 <= max_dist^2
            const double d2 = norm2_exact(__dadd_rn(xa, -rx[j]), __dadd_rn(ya, -ry[j]), __dadd_rn(za, -rz[j]));
            if (j != (uint32_t)i && d2 > 0.0 && d2 <= max_dist2)
            {
              if (FILL) wc[2 + cnt] = (uint16_t)pb;
              cnt++;
            }
          }
        }
        if (cnt > 0)
        {
          if (FILL)
          {
            // encode_cell_index (chunk_neighbors.h:137-150)
            wc[0] = (uint16_t)(((((bk - ak) + 16) << 5) + ((bj - aj) + 16)) << 5) + (uint16_t)((bi - ai) + 16);
            wc[1] = (uint16_t)cnt;
            w = wc + 2 + cnt;
          }
          if (cnt >= 65535u) atomicOr(err, DERR_GROUP_OVERFLOW);
          groups++; total += cnt;
        }
      }
  if (FILL) *wg = (uint16_t)groups;
  else
  {
    if (groups >= 65535u) atomicOr(err, DERR_GROUP_OVERFLOW);
    out.nb_len[i] = 1u + 2u * groups + total;
    out.nb_cnt[i] = total;
  }
}

// one warp per cell: in-cell exclusive offsets of the particle lists, the cell's stream size (u16 words, unpadded and
// padded to 16 bytes) and the running maximum neighbour count (m_max_neighbors, chunk_neighbors_execute.h:403-406)
__global__ void k_nbh_cell_sizes(GridP g, int n_cells, const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_count,
                                 const uint32_t* __restrict__ nb_len, const uint32_t* __restrict__ nb_cnt, uint32_t* __restrict__ nb_off,
                                 uint32_t* __restrict__ stream_size, uint32_t* __restrict__ stream_size_padded, uint32_t* __restrict__ max_nbh,
                                 uint32_t* __restrict__ max_cell_count, uint32_t* __restrict__ max_stream, uint32_t* __restrict__ max_chunk_words,
                                 unsigned long long* __restrict__ inner_words, uint32_t* __restrict__ err)
{
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= n_cells) return;
  const uint32_t n = cell_count[c], s0 = cell_start[c];
  if (n == 0) { if (lane == 0) { stream_size[c] = 0; stream_size_padded[c] = 0; } return; }
  if (n > 65535u && lane == 0) atomicOr(err, DERR_CELL_OVERFLOW);
  uint32_t run = 0, mx = 0, mxc = 0;
  for (uint32_t p0 = 0; p0 < n; p0 += 32)
  {
    const uint32_t p = p0 + lane;
    const uint32_t v = (p < n) ? nb_len[s0 + p] : 0u;
    if (p < n) mx = max(mx, nb_cnt[s0 + p]);
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (p < n) nb_off[s0 + p] = run + x - v;
    const uint32_t chunk = __shfl_sync(0xffffffffu, x, 31);
    mxc = max(mxc, chunk);      // list words of one 32-particle chunk (k_nbh_emit image size)
    run += chunk;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0)
  {
    const uint32_t sz = 2u * (n + 1u) + run;
    stream_size[c] = sz;
    stream_size_padded[c] = (sz + 7u) & ~7u;
    atomicMax(max_nbh, mx);
    atomicMax(max_cell_count, n);
    atomicMax(max_stream, (sz + 7u) & ~7u);
    atomicMax(max_chunk_words, mxc);
    const int ci = c % g.dims[0], cj = (c / g.dims[0]) % g.dims[1], ck = c / (g.dims[0] * g.dims[1]);
    const bool inner = ci >= g.gl && ci < g.dims[0] - g.gl && cj >= g.gl && cj < g.dims[1] - g.gl && ck >= g.gl && ck < g.dims[2] - g.gl;
    if (inner) { atomicAdd(inner_words, (unsigned long long)((sz + 7u) & ~7u)); atomicAdd(max_chunk_words + 1, 1u); }   // [+1]: non-empty inner cells
  }
}

// per cell: stream pointer (nullptr for empty cells, chunk_neighbors_host_write_accessor.h:47-52) and size in BYTES
__global__ void k_nbh_pointers(int n_cells, uint16_t* __restrict__ pool, const unsigned long long* __restrict__ stream_off,
                               const uint32_t* __restrict__ stream_size, uint16_t** __restrict__ cell_stream, uint32_t* __restrict__ cell_stream_bytes)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const uint32_t sz = stream_size[c];
  cell_stream[c] = sz ? pool + stream_off[c] : nullptr;
  cell_stream_bytes[c] = sz * 2u;
  if (sz)
  {
    // zero the alignment padding so the pool content is deterministic
    const uint32_t pad = ((sz + 7u) & ~7u) - sz;
    for (uint32_t q = 0; q < pad; q++) pool[stream_off[c] + sz + q] = 0;
  }
}

__global__ void k_clear_bits_u32(uint32_t* p, uint32_t bits) { *p &= ~bits; }

// occupancy of the local grid: out[0] = largest cell count (all cells), out[1] = non-empty inner cells
__global__ void k_cell_stats(GridP g, const uint32_t* __restrict__ cell_count, uint32_t* __restrict__ out)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t v = 0; bool ne = false;
  if (c < g.n_cells)
  {
    v = cell_count[c];
    const int ci = c % g.dims[0], cj = (c / g.dims[0]) % g.dims[1], ck = c / (g.dims[0] * g.dims[1]);
    ne = v > 0 && ci >= g.gl && ci < g.dims[0] - g.gl && cj >= g.gl && cj < g.dims[1] - g.gl && ck >= g.gl && ck < g.dims[2] - g.gl;
  }
  const unsigned m = __ballot_sync(0xffffffffu, ne);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) { if (v) atomicMax(out, v); if (m) atomicAdd(out + 1, (uint32_t)__popc(m)); }
}

// maximum of a u32 array (cell counts) -> *out (atomicMax)
__global__ void k_max_u32(int n, const uint32_t* __restrict__ a, uint32_t* __restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t v = (i < n) ? a[i] : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0 && v) atomicMax(out, v);
}

// ------------------------------------------------------------------------------------------------------------------
// K3 pair sweep with the Lennard-Jones functor (the hot-path kernel).
// reference: compute_cell_particle_pairs_impl_default.h:87-239 (stream decoding, d2 re-test against rcut^2),
//            lennard_jones.cu:46-56,106-124 (functor, buffer-less call form).
// dr = r_b - r_a (:183); accept iff d2 > 0 && d2 <= rcut2 (:186) with d2 evaluated exactly like the oracle
// (((x*x)+(y*y))+(z*z), every operation rounded).  The functor is restated with one reciprocal instead of sqrt + two
// divisions:  de/r = -24 eps (2 s12 - s6) / d2 = (24 eps - 48 eps s6) (s6 / d2),  s6 = (sigma^2/d2)^3.
// MODE 0: f += sum            (op lennard_jones_force, accumulating like the reference functor)
// MODE 1: f  = sum / m[type]; v += f*dth   (zero_particle_force + lennard_jones_force + divide_force_by_type_scalar
//                                           + verlet_second_half fused; dth = 0 leaves v untouched bit-for-bit)
// EV: additionally accumulate per-block partial sums of energy and virial (oracle-defined observables)
//
// One block = one TILE of ti x tj cells (same k), one thread per particle of the tile.  The block stages into shared
// memory (cp.async)
//   (a) the positions of every particle of the tile's halo box ((ti+2gap) x (tj+2gap) x (2gap+1) cells, clamped to the
//       grid) as {x,y} pairs + z: each neighbour cell is fetched from HBM/L2 once per tile with coalesced reads, and a
//       candidate costs one LDS.128 + one LDS.64;
//   (b) the u16 neighbour streams of the tile's cells, verbatim (each cell stream is one 16-byte aligned block).
// A short pre-pass (one thread per particle, hopping over its own group headers) rewrites every non-candidate word of
// the staged copy -- group counter, (cell code, count) headers, alignment padding -- as 0x8000 | base, base = index of
// the neighbour cell's first particle in the staged halo.  After that a list is a flat sequence in which a word >= 0x8000
// sets the current base and anything else is a candidate p_b, so the main loop has no nested structure and no
// divergent branch: every trip reads four words with one aligned LDS.64 and evaluates up to four candidates with
// independent FP64 dependency chains.
// ------------------------------------------------------------------------------------------------------------------
// (LJP, the Lennard-Jones functors and the functor concept: xnb_pair_functor.cuh)

struct TileP
{
  int ti, tj;        // cells per tile along i and j
  int gap;           // neighbour cell layers = ceil(nbh_dist / cell_size)
  int lo[3], hi[3];  // swept cell range [lo,hi) per axis (inner cells, or all cells when ghost=true)
  int tiles_i, tiles_j;
  int cap;           // staging capacity in particles (< 32768)
  int cap_s;         // staging capacity in u16 stream words (multiple of 8)
  int hx, hy, hz;    // nominal halo box dims
};

// asynchronous global -> shared copies (LDGSTS): issued back to back, completed by cp_async_wait_all()
XNB_DEVINL void cp_async8(void* smem_dst, const void* gsrc)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
XNB_DEVINL void cp_async8_sh(uint32_t smem_addr, const void* gsrc)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
XNB_DEVINL void cp_async16(void* smem_dst, const void* gsrc)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
XNB_DEVINL void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

constexpr uint32_t SW_MARK = 0x8000u;      // staged stream word >= SW_MARK: "set base", not a candidate
constexpr int SWEEP_MAX_THREADS = 384;

// decode one aligned chunk of four staged stream words: a word >= SW_MARK sets the base, anything else is a candidate
// (if `live`); FIRST: the chunk's first `skip` words belong to the previous list
template <bool FIRST>
XNB_DEVINL void sweep_decode(const uint2 c, uint32_t skip, bool live, uint32_t& sb, uint32_t (&j)[4], bool (&ok)[4])
{
  const uint32_t w[4] = {c.x & 0xffffu, c.x >> 16, c.y & 0xffffu, c.y >> 16};
#pragma unroll
  for (int u = 0; u < 4; u++)
  {
    const bool mark = w[u] >= SW_MARK;
    if (mark) sb = w[u] & 0x7fffu;
    ok[u] = !mark && live && (!FIRST || (uint32_t)u >= skip);
    j[u] = sb + w[u];
  }
}

XNB_DEVINL void sweep_load(const double2* __restrict__ XY, const double* __restrict__ Z, const uint32_t (&j)[4], const bool (&ok)[4],
                           double2 (&pxy)[4], double (&pz)[4])
{
  // branch free: a slot that is not a candidate reads particle 0 (all such lanes hit the same address: a broadcast)
#pragma unroll
  for (int u = 0; u < 4; u++) { const uint32_t jj = ok[u] ? j[u] : 0u; pxy[u] = XY[jj]; pz[u] = Z[jj]; }
}

// one trip of the sweep: evaluate the four candidates whose positions were loaded last trip (pxy, pz, okc) while the
// positions of the next trip (decoded from chunk cn) are fetched into (nxy, nz, okn)
template <bool EV, class F>
XNB_DEVINL void sweep_trip(const F& lj, unsigned long long rc2b, const double2* __restrict__ XY, const double* __restrict__ Z, double xa, double ya, double za,
                           const uint2 cn, bool live_next, uint32_t& sb,
                           const double2 (&pxy)[4], const double (&pz)[4], const bool (&okc)[4], const uint32_t (&jc)[4],
                           double2 (&nxy)[4], double (&nz)[4], bool (&okn)[4], uint32_t (&jn)[4], PairAcc& acc)
{
  sweep_decode<false>(cn, 0u, live_next, sb, jn, okn);
  sweep_load(XY, Z, jn, okn, nxy, nz);
  double dx[4], dy[4], dz[4], d2[4]; bool ok[4];
#pragma unroll
  for (int u = 0; u < 4; u++)
  {
    dx[u] = __dadd_rn(pxy[u].x, -xa); dy[u] = __dadd_rn(pxy[u].y, -ya); dz[u] = __dadd_rn(pz[u], -za);
    d2[u] = norm2_exact(dx[u], dy[u], dz[u]);
  }
  // accept iff d2 > 0 && d2 <= rcut2 (impl_default.h:186):  bits(d2) - 1 < bits(rcut2) as unsigned integers
  // (d2 >= +0, never NaN for finite positions) -- one predicate, off the FP64 pipe
#pragma unroll
  for (int u = 0; u < 4; u++) ok[u] = okc[u] && (unsigned long long)(__double_as_longlong(d2[u]) - 1ll) < rc2b;
  pair_apply4<EV>(lj, dx, dy, dz, d2, ok, jc, acc);
}

template <class F, int MODE, bool EV>
__global__ void __launch_bounds__(SWEEP_MAX_THREADS)
k_lj_sweep(GridP g, TileP tp, int n_inner, int n_total, F lj, double dth,
           const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz,
           double* __restrict__ vx, double* __restrict__ vy, double* __restrict__ vz,
           double* __restrict__ fx, double* __restrict__ fy, double* __restrict__ fz,
           const uint8_t* __restrict__ type, const double* __restrict__ mass,
           const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_count,
           const uint16_t* const* __restrict__ cell_stream, const uint32_t* __restrict__ stream_size,
           double* __restrict__ ev_partials /* [gridDim.x][7] */, const unsigned long long* __restrict__ skip_if_nonzero)
{
  // speculative launch (xnb_run_steps): the displacement counter says a rebuild is due -> this launch is void
  if (skip_if_nonzero && *skip_if_nonzero) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nh_max = tp.hx * tp.hy * tp.hz;
  const int tc_max = tp.ti * tp.tj;
  uint32_t* hstart = reinterpret_cast<uint32_t*>(smem_raw);                       // [nh_max + 1] particle prefix of the halo cells
  uint32_t* hfirst = hstart + nh_max + 1;                                           // [nh_max] flat index of each halo cell's first particle
  uint32_t* tstart = hfirst + nh_max;                                               // [tc_max + 1] particle prefix of the tile cells
  uint32_t* tfirst = tstart + tc_max + 1;                                           // [tc_max] flat index of each tile cell's first particle
  uint32_t* sstart = tfirst + tc_max;                                               // [tc_max + 1] u16 offset of each staged tile cell stream in S
  uint32_t* ssize = sstart + tc_max + 1;                                            // [tc_max] stream size of each tile cell (u16 words, unpadded)
  const size_t head = (((size_t)(2 * nh_max + 1 + 4 * tc_max + 2) * 4 + 15) & ~(size_t)15);
  uint16_t* S = reinterpret_cast<uint16_t*>(smem_raw + head);                       // [cap_s + 16] (the loop reads up to 4 chunks ahead)
  double2* XY = reinterpret_cast<double2*>(smem_raw + head + (size_t)(tp.cap_s + 16) * 2);  // [cap]
  double* Z = reinterpret_cast<double*>(XY + tp.cap);                               // [cap]
  __shared__ uint32_t s_scan[32];
  __shared__ int s_fit;

  // ---- tile geometry
  int b = blockIdx.x;
  const int t_i = b % tp.tiles_i; b /= tp.tiles_i;
  const int t_j = b % tp.tiles_j; const int ck = tp.lo[2] + b / tp.tiles_j;
  const int ci0 = tp.lo[0] + t_i * tp.ti, cj0 = tp.lo[1] + t_j * tp.tj;
  const int tci = min(tp.ti, tp.hi[0] - ci0), tcj = min(tp.tj, tp.hi[1] - cj0);
  const int tcells = tci * tcj;
  const int bx0 = max(ci0 - tp.gap, 0), bx1 = min(ci0 + tci - 1 + tp.gap, g.dims[0] - 1);
  const int by0 = max(cj0 - tp.gap, 0), by1 = min(cj0 + tcj - 1 + tp.gap, g.dims[1] - 1);
  const int bz0 = max(ck - tp.gap, 0), bz1 = min(ck + tp.gap, g.dims[2] - 1);
  const int HX = bx1 - bx0 + 1, HY = by1 - by0 + 1, HZ = bz1 - bz0 + 1;
  const int NH = HX * HY * HZ, HXY = HX * HY;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;

  // ---- prefix tables: halo cells (block scan) and tile cells (one warp)
  {
    uint32_t carry = 0;
    for (int base = 0; base < NH; base += blockDim.x)
    {
      const int h = base + threadIdx.x;
      uint32_t cnt = 0;
      if (h < NH)
      {
        const int hxq = h % HX, hyq = (h / HX) % HY, hzq = h / HXY;
        const int c = ijk_to_index(g.dims, bx0 + hxq, by0 + hyq, bz0 + hzq);
        cnt = cell_count[c]; hfirst[h] = cell_start[c];
      }
      uint32_t total;
      const uint32_t off = block_exclusive_scan<uint32_t>(cnt, &total, s_scan);
      if (h < NH) hstart[h] = carry + off;
      carry += total;
    }
    if (threadIdx.x == 0) hstart[NH] = carry;
    if (warp == nwarp - 1)
    {
      const int q = lane;
      uint32_t cnt = 0, sw = 0;
      if (q < tcells)
      {
        const int c = ijk_to_index(g.dims, ci0 + q % tci, cj0 + q / tci, ck);
        cnt = cell_count[c]; tfirst[q] = cell_start[c]; ssize[q] = stream_size[c]; sw = (stream_size[c] + 7u) & ~7u;
      }
      uint32_t x = cnt, y = sw;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
      {
        const uint32_t x2 = __shfl_up_sync(0xffffffffu, x, o), y2 = __shfl_up_sync(0xffffffffu, y, o);
        if (q >= o) { x += x2; y += y2; }
      }
      if (q < tcells) { tstart[q] = x - cnt; sstart[q] = y - sw; }
      if (q == tcells - 1) { tstart[tcells] = x; sstart[tcells] = y; }
    }
  }
  __syncthreads();
  const int n_tile = (int)tstart[tcells];
  if (threadIdx.x == 0) s_fit = (hstart[NH] <= (uint32_t)tp.cap && sstart[tcells] <= (uint32_t)tp.cap_s) ? 1 : 0;
  __syncthreads();
  const bool staged = s_fit != 0;

  if (staged && n_tile > 0)
  {
    // ---- (a) halo positions: one warp per halo cell, lanes over its particles (coalesced reads)
    for (int h = warp; h < NH; h += nwarp)
    {
      const uint32_t d0 = hstart[h], cnt = hstart[h + 1] - d0, s0 = hfirst[h];
      for (uint32_t p = lane; p < cnt; p += 32)
      {
        cp_async8(&XY[d0 + p].x, rx + s0 + p); cp_async8(&XY[d0 + p].y, ry + s0 + p); cp_async8(Z + d0 + p, rz + s0 + p);
      }
    }
    // ---- (b) streams of the tile cells, 16 bytes per copy
    for (int q = 0; q < tcells; q++)
    {
      const int c = ijk_to_index(g.dims, ci0 + q % tci, cj0 + q / tci, ck);
      const uint4* src = reinterpret_cast<const uint4*>(cell_stream[c]);
      uint4* dst = reinterpret_cast<uint4*>(S + sstart[q]);
      const int nv = (int)((sstart[q + 1] - sstart[q]) >> 3);
      for (int v = threadIdx.x; v < nv; v += blockDim.x) cp_async16(dst + v, src + v);
    }
    cp_async_wait_all();
    __syncthreads();
    // ---- (c) pre-pass: every thread rewrites the non-candidate words of its own list
    for (int t = (int)threadIdx.x; t < n_tile; t += blockDim.x)
    {
      int q = 0;
      while (q + 1 < tcells && (uint32_t)t >= tstart[q + 1]) q++;
      const uint32_t pa = (uint32_t)t - tstart[q], na = tstart[q + 1] - tstart[q];
      uint16_t* cs = S + sstart[q];
      const uint32_t off0 = reinterpret_cast<const uint32_t*>(cs)[pa], off1 = reinterpret_cast<const uint32_t*>(cs)[pa + 1];
      uint16_t* lists = cs + 2u * (na + 1u);
      uint16_t* p = lists + (off0 - 1u);              // group counter (offsets are biased by the number of tables = 1)
      uint16_t* const end = lists + (off1 - 1u);
      const int cia = ci0 + q % tci, cja = cj0 + q / tci;
      const int hb2 = ((ck - bz0) * HY + (cja - by0)) * HX + (cia - bx0) - 16 * (HXY + HX + 1);
      *p++ = (uint16_t)SW_MARK;
      while (p < end)
      {
        const uint32_t enc = p[0], n = p[1];
        // halo index of cell (cia + ri, cja + rj, ck + rk), (ri,rj,rk) = 5-bit fields of enc minus 16 (chunk_neighbors.h:137-162)
        const uint32_t sb = hstart[hb2 + (int)(enc >> 10) * HXY + (int)((enc >> 5) & 31u) * HX + (int)(enc & 31u)];
        p[0] = (uint16_t)(SW_MARK | sb); p[1] = (uint16_t)(SW_MARK | sb);
        p += 2u + n;
      }
      if (pa == na - 1u)
      {
        // alignment padding behind the last list of the cell
        uint16_t* const pend = cs + (sstart[q + 1] - sstart[q]);
        for (uint16_t* z = cs + ssize[q]; z < pend; z++) *z = (uint16_t)SW_MARK;
      }
    }
    __syncthreads();
  }

  PairAcc acc;
  acc.e = acc.wxx = acc.wyy = acc.wzz = acc.wxy = acc.wxz = acc.wyz = 0.;
  const unsigned long long rc2b = (unsigned long long)__double_as_longlong(lj.rcut2());   // bits(d2) - 1 < bits(rcut2) <=> d2 in (0, rcut2]

  for (int t = (int)threadIdx.x; t < n_tile; t += blockDim.x)
  {
    int q = 0;
    while (q + 1 < tcells && (uint32_t)t >= tstart[q + 1]) q++;
    const uint32_t pa = (uint32_t)t - tstart[q];
    const uint32_t i = tfirst[q] + pa;
    const uint32_t na = tstart[q + 1] - tstart[q];
    const double xa = rx[i], ya = ry[i], za = rz[i];
    // epilogue operands fetched now so that their latency hides behind the pair loop
    double m = 1.0, ux = 0., uy = 0., uz = 0.;
    if (MODE == 1) { m = mass[type[i]]; if (dth != 0.0) { ux = vx[i]; uy = vy[i]; uz = vz[i]; } }
    acc.ax = acc.ay = acc.az = 0.;
    if (staged)
    {
      const uint32_t cbase = sstart[q];
      const uint32_t off0 = reinterpret_cast<const uint32_t*>(S + cbase)[pa], off1 = reinterpret_cast<const uint32_t*>(S + cbase)[pa + 1];
      const uint32_t kb = cbase + 2u * (na + 1u) + off0;            // first word after the group counter
      const uint32_t ke = cbase + 2u * (na + 1u) + (off1 - 1u);     // end of the list
      const uint2* S64 = reinterpret_cast<const uint2*>(S);
      uint32_t k = kb & ~3u;
      uint32_t sb = 0;
      // software pipeline: chunk t+2 is being read and the positions of trip t+1 are in flight while trip t is evaluated;
      // two trips per iteration so that the two register sets (A, B) swap roles without copies
      uint32_t jA[4], jB[4]; bool okA[4], okB[4]; double2 xyA[4], xyB[4]; double zA[4], zB[4];
      sweep_decode<true>(S64[k >> 2], kb & 3u, k < ke, sb, jA, okA);
      sweep_load(XY, Z, jA, okA, xyA, zA);
      uint2 c1 = S64[(k >> 2) + 1];
      for (; k < ke; k += 8u)
      {
        const uint2 c2 = S64[(k >> 2) + 2], c3 = S64[(k >> 2) + 3];
        sweep_trip<EV>(lj, rc2b, XY, Z, xa, ya, za, c1, k + 4u < ke, sb, xyA, zA, okA, jA, xyB, zB, okB, jB, acc);
        sweep_trip<EV>(lj, rc2b, XY, Z, xa, ya, za, c2, k + 8u < ke, sb, xyB, zB, okB, jB, xyA, zA, okA, jA, acc);
        c1 = c3;
      }
    }
    else
    {
      // tile too large for the staging buffers: walk the stream in global memory (nested form of impl_default.h:143-179)
      const int ca = ijk_to_index(g.dims, ci0 + q % tci, cj0 + q / tci, ck);
      const uint16_t* base = cell_stream[ca];
      const uint32_t off = reinterpret_cast<const uint32_t*>(base)[pa];
      const uint16_t* s = base + 2u * (na + 1u) + (off - 1u);
      int groups = *s++;
      for (; groups > 0; --groups)
      {
        const uint32_t enc = *s++;
        int cnt = *s++;
        const int ri = (int)(enc & 31u) - 16, rj = (int)((enc >> 5) & 31u) - 16, rk = (int)((enc >> 10) & 31u) - 16;
        const uint32_t sbg = cell_start[ca + (rk * g.dims[1] + rj) * g.dims[0] + ri];
        for (; cnt > 0; --cnt)
        {
          const uint32_t j = sbg + *s++;
          const double dx = __dadd_rn(rx[j], -xa), dy = __dadd_rn(ry[j], -ya), dz = __dadd_rn(rz[j], -za);
          const double d2 = norm2_exact(dx, dy, dz);
          if ((unsigned long long)(__double_as_longlong(d2) - 1ll) < rc2b) pair_apply1<EV>(lj, dx, dy, dz, d2, j, acc);
        }
      }
    }
    double ax = acc.ax, ay = acc.ay, az = acc.az;
    if (MODE == 0)
    {
      fx[i] += ax; fy[i] += ay; fz[i] += az;
    }
    else
    {
      ax = __ddiv_rn(ax, m); ay = __ddiv_rn(ay, m); az = __ddiv_rn(az, m);
      fx[i] = ax; fy[i] = ay; fz[i] = az;
      if (dth != 0.0)
      {
        vx[i] = __dadd_rn(ux, __dmul_rn(ax, dth));
        vy[i] = __dadd_rn(uy, __dmul_rn(ay, dth));
        vz[i] = __dadd_rn(uz, __dmul_rn(az, dth));
      }
    }
  }
  if (MODE == 1)
  {
    // zero_particle_force{ghost:true}: ghost particles keep f = 0 (each block clears its slice of the ghost range)
    const int ng = n_total - n_inner;
    const int per = (ng + (int)gridDim.x - 1) / (int)gridDim.x;
    const int g0 = n_inner + (int)blockIdx.x * per, g1 = min(g0 + per, n_total);
    for (int i = g0 + (int)threadIdx.x; i < g1; i += blockDim.x) { fx[i] = 0.; fy[i] = 0.; fz[i] = 0.; }
  }
  if (EV)
  {
    __shared__ double red[7][SWEEP_MAX_THREADS / 32];
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

// kinetic energy partial sums: sum 1/2 m v^2 per block
__global__ void __launch_bounds__(256)
k_ekin(int n, const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ vz,
       const uint8_t* __restrict__ type, const double* __restrict__ mass, double* __restrict__ partials)
{
  __shared__ double red[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.;
  if (i < n) v = 0.5 * mass[type[i]] * (vx[i] * vx[i] + vy[i] * vy[i] + vz[i] * vz[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) { double s = 0.; for (int w = 0; w < 8; w++) s += red[w]; partials[blockIdx.x] = s; }
}

// FP64 FMA peak probe: 8 independent dependent-chains per thread
__global__ void __launch_bounds__(256) k_dfma_probe(double* out, int iters, double a, double b)
{
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int k = 0; k < iters; k++)
  {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

} // namespace xnb

// ------------------------------------------------------------------------------------------------------------------
// migrate_cell_particles (multi-GPU hand-off of particles that left this rank's block).
// reference: otb_particles of move_particles_across_cells.h:124-138 consumed by mpi/migrate_cell_particles.cpp:101-143.
// ------------------------------------------------------------------------------------------------------------------
namespace xnb {

__global__ void k_migrate_dest(GridP g, int n_leave, const uint32_t* __restrict__ leave_list,
                               const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz,
                               const int* __restrict__ blocks /* nranks x 6 : start ijk, end ijk */, int nranks,
                               uint32_t* __restrict__ dest_rank, uint32_t* __restrict__ dest_pos, uint32_t* __restrict__ dest_count,
                               uint32_t* __restrict__ err)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_leave) return;
  const uint32_t i = leave_list[q];
  const int li = (int)floor(__ddiv_rn(__dadd_rn(rx[i], -g.dmin[0]), g.cs));
  const int lj = (int)floor(__ddiv_rn(__dadd_rn(ry[i], -g.dmin[1]), g.cs));
  const int lk = (int)floor(__ddiv_rn(__dadd_rn(rz[i], -g.dmin[2]), g.cs));
  int owner = -1;
  for (int r = 0; r < nranks; r++)
  {
    const int* b = blocks + 6 * r;
    if (li >= b[0] && li < b[3] && lj >= b[1] && lj < b[4] && lk >= b[2] && lk < b[5]) { owner = r; break; }
  }
  if (owner < 0) { atomicOr(err, DERR_LOST_PARTICLE); dest_rank[q] = 0xFFFFFFFFu; dest_pos[q] = 0; return; }
  dest_rank[q] = (uint32_t)owner;
  dest_pos[q] = atomicAdd(&dest_count[owner], 1u);
}

// migrating particles travel like ghosts with all fields: ONE contiguous slab per destination rank, field-major, 8-byte words
// (x y z vx vy vz fx fy fz id type); destination r's entries [base[r], base[r+1]) occupy words [11 base[r], 11 base[r+1]) and the
// receiver scatters them with k_ghost_unpack<true> (same layout on the wire).
__global__ void k_migrate_pack(int n_leave, const uint32_t* __restrict__ leave_list, const uint32_t* __restrict__ dest_rank,
                               const uint32_t* __restrict__ dest_pos, const uint32_t* __restrict__ dest_base /* nranks + 1 */, ParticlesP p,
                               double* __restrict__ stage)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_leave) return;
  const uint32_t r = dest_rank[q];
  if (r == 0xFFFFFFFFu) return;
  const uint32_t i = leave_list[q];
  const size_t s0 = dest_base[r], n = dest_base[r + 1] - s0;
  double* slab = stage + (size_t)GHOST_WORDS_ALL * s0 + dest_pos[q];
  slab[0] = p.rx[i]; slab[n] = p.ry[i]; slab[2 * n] = p.rz[i];
  slab[3 * n] = p.vx[i]; slab[4 * n] = p.vy[i]; slab[5 * n] = p.vz[i];
  slab[6 * n] = p.fx[i]; slab[7 * n] = p.fy[i]; slab[8 * n] = p.fz[i];
  reinterpret_cast<unsigned long long*>(slab)[9 * n] = p.id[i];
  reinterpret_cast<unsigned long long*>(slab)[10 * n] = (unsigned long long)p.type[i];
}

// simple_cost_model (src/mpi/include/exanb/mpi/simple_cost_model.h:67-146) on the device: cost of every INNER cell of this rank's block
// from its particle count, p = N / cell volume, cost = p d1 + p^2 d2 + p^3 d3 + c (ghost cells carry no cost, :103), written into the
// cost array of the whole domain grid at the cell's domain index (the array all ranks then sum: load_balance_rcb.cpp:270)
__global__ void k_cell_costs(GridP g, const uint32_t* __restrict__ cell_count, double d3, double d2, double d1, double cc, double* __restrict__ domain_costs)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= g.n_cells) return;
  const int i = c % g.dims[0], j = (c / g.dims[0]) % g.dims[1], k = c / (g.dims[0] * g.dims[1]);
  if (i < g.gl || i >= g.dims[0] - g.gl || j < g.gl || j >= g.dims[1] - g.gl || k < g.gl || k >= g.dims[2] - g.gl) return;
  const double cell_volume = __dmul_rn(__dmul_rn(g.cs, g.cs), g.cs);
  const double pvol = __ddiv_rn((double)cell_count[c], cell_volume);
  const double cost = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(pvol, d1), __dmul_rn(__dmul_rn(pvol, pvol), d2)), __dmul_rn(__dmul_rn(__dmul_rn(pvol, pvol), pvol), d3)), cc);
  const size_t di = (size_t)(g.off[0] + i), dj = (size_t)(g.off[1] + j), dk = (size_t)(g.off[2] + k);
  domain_costs[(dk * (size_t)g.ddims[1] + dj) * (size_t)g.ddims[0] + di] = cost;
}

} // namespace xnb
