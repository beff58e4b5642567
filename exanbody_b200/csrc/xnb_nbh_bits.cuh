// xnb_nbh_bits.cuh -- K2, the chunk neighbour build (ops amr_grid_pairs + chunk_neighbors) as ONE tiled kernel: k_nbh_bits.
//
// reference: particle_neighbors/include/exanb/particle_neighbors/chunk_neighbors_execute.h:110-411 (what a list holds and in
// which order), chunk_neighbors.h:42-186 (the GridChunkNeighbors stream format), neighbor_filter_func.h (no filter here).
//
// One block = one tile of the pair sweep (ti x tj x tk cells, xnb_sweep_cl.cuh): the build produces, for the same tile
// particles in the same groups of 32, BOTH outputs the library keeps:
//   * the reference-format streams of the tile's cells (byte-identical to the oracle's), and
//   * the compiled rows the sweep reads (u16 word = 8 x index of the neighbour in the tile's staged halo box),
// so nothing re-reads the streams to derive the second copy.
//
// How a list is built.  The halo box of the tile is staged in shared memory as fp32 PAIRS of candidates,
// {x0,x1,y0,y1 | z0,z1,w0,w1} with (x,y,z) = -2 (r - O), w = |r - O|^2, O the centre of the box.  x-adjacent cells are
// contiguous in the staged order, so the (2 gap + 1)^2 cell rows of a particle's neighbourhood are contiguous candidate
// ranges.  A lane owns one particle a (lane-per-particle form); all lanes sweep the same candidates (broadcast LDS.128),
//     t = |a|^2 - (max_dist^2 + band) + x_a q.x + y_a q.y + z_a q.z + q.w  =  d2_fp32 - (max_dist^2 + band)
// two candidates per instruction (FFMA2 / FADD2), and the SIGN BIT of t is shifted into a register: accept bits in
// registers, one 32-bit mask per (particle, 32 staged candidates), no per-candidate shared-memory traffic.  The integer
// minimum of the raw bits (VIMNMX3) is the accepted value closest to the threshold: only if it lies inside the error band of
// the fp32 evaluation are the accepted candidates of that mask re-examined and the ambiguous ones decided by the
// reference's exact fp64 test on the original coordinates (d2 > 0 && d2 <= max_dist^2, chunk_neighbors_execute.h:225-227),
// so the lists are bit-identical to an all-fp64 build.  Segments of a group that hold only a few particles of a cell are
// built candidate-per-lane instead (one ballot per particle and 32 candidates).
// One warp owns one tile cell, 32 of its particles at a time.  The accept masks of a neighbourhood row never leave the registers:
// they are expanded at once into the lane's list in shared memory (byte entries when cells hold < 256 particles), in staged
// order = ascending cell code, ascending p_b = the order the reference obtains with its sort (:308-324).  From its finished
// list a lane writes both outputs: the reference-format words (8-byte stores into the cell's stream) and the compiled-row
// words.  A group of the sweep (32 consecutive tile particles, possibly of two cells) owns a fixed block of rows, so its lanes
// are written independently; a short pass after a block barrier pads every group to its longest list.
#pragma once
#include "xnb_sweep_cl.cuh"

namespace xnb {

struct NbhBitsP
{
  int cap_l;             // list capacity per particle (elements of the list areas): 1 + 2 groups + entries
  int emit_rows;         // write the compiled rows + group table (sweep tiles); 0: streams only
  int sel_mode;          // 0: every tile cell is built; 1: only cells outside the inner range (ghost-cell lists, built lazily)
  int slot_words;        // stream capacity per cell (u16 words, multiple of 8)
  int cap_trips;         // rows per group in the compiled-list buffer (a group owns a fixed block of rows: no allocator)
  int planes;            // k_nbh_big: 1 = the compiled rows of a group come in one segment per z-plane of halo cells (k_lj_sweep_pl)
  int cap_pl;            // planes: the sentinel slot pads point at (= staging capacity of a plane in the sweep)
  double max_dist2;
};

// counters written by k_nbh_bits (u32): what it used and what it would have needed
enum { NB_ROWS = 0, NB_GMAX = 1, NB_CAP = 2, NB_SLOTS = 3, NB_SLOT_WORDS = 4, NB_MAX_NBH = 5, NB_NONEMPTY = 6, NB_MAX_CELL = 7, NB_MAX_STREAM = 8,
       NB_OVERFLOW = 9, NB_AMBIGUOUS = 10, NB_TRIPS = 11, NB_SURV = 12, NB_U32_COUNT = 13 };      // NB_SURV: k_nbh_big, most scratch rows a group needed
// u64 totals: [0] list entries of the built particles [1] padded stream words of all built cells [2] of the inner cells among them

XNB_DEVINL bool cl_cell_is_inner(const GridP& g, int ci, int cj, int ck)
{
  return ci >= g.gl && ci < g.dims[0] - g.gl && cj >= g.gl && cj < g.dims[1] - g.gl && ck >= g.gl && ck < g.dims[2] - g.gl;
}

// staged candidate j of the halo box: pair j >> 1, component j & 1
struct NbPair { float4 xy, zw; };       // {x0, x1, y0, y1}, {z0, z1, w0, w1}
XNB_DEVINL float4 nb_candidate(const NbPair* __restrict__ S, uint32_t j)
{
  const NbPair p = S[j >> 1];
  return (j & 1u) ? make_float4(p.xy.y, p.xy.w, p.zw.y, p.zw.w) : make_float4(p.xy.x, p.xy.z, p.zw.x, p.zw.z);
}

XNB_DEVINL unsigned long long nb_fma2(float a, unsigned long long b, unsigned long long c)
{
  unsigned long long r, a2;
  asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a2), "l"(b), "l"(c));
  return r;
}
XNB_DEVINL unsigned long long nb_add2(unsigned long long a, unsigned long long b)
{
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
XNB_DEVINL unsigned long long nb_pack2(float a) { unsigned long long r; asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(a)); return r; }

// per-particle constants of the classification: position relative to O (fp32) and c = |a|^2 - (max_dist^2 + band)
struct NbSelf { float x, y, z, c; };

// the value the hot loop computes for (a, candidate j), same operations in the same order (used by the slow paths)
XNB_DEVINL float nb_value(const NbSelf& a, const float4 q) { return fmaf(a.z, q.z, fmaf(a.y, q.y, fmaf(a.x, q.x, a.c))) + q.w; }

// accept bits of one particle against the n4 (multiple of 4, <= 32) staged candidates that start at j0 (j0 % 4 == 0): bit k of the
// result = candidate j0 + k has t < 0.  Cells are staged padded to multiples of four with never-accepted entries, so no candidate
// needs masking.  trk = integer minimum of the raw bits of every t computed (the negative value closest to zero, if any is
// negative).  OWN (the particle's own cell): zbits = candidates with t + zc < 0, i.e. fp32 d2 below the zero band.
template <bool OWN>
XNB_DEVINL uint32_t nb_block_bits(const NbPair* __restrict__ S, uint32_t j0, uint32_t n4, const NbSelf& a, float zc, int& trk, uint32_t& zbits)
{
  const unsigned long long c2 = nb_pack2(a.c);
  const unsigned long long zc2 = nb_pack2(zc);
  uint32_t acc = 0, zacc = 0;
  int t_min = trk;
  const NbPair* __restrict__ P = S + (j0 >> 1);
#pragma unroll 2
  for (uint32_t j = 0; j < n4; j += 4u)
  {
    const NbPair p0 = P[j >> 1], p1 = P[(j >> 1) + 1u];
    const unsigned long long* u0 = reinterpret_cast<const unsigned long long*>(&p0);
    const unsigned long long* u1 = reinterpret_cast<const unsigned long long*>(&p1);
    const unsigned long long d0 = nb_add2(nb_fma2(a.z, u0[2], nb_fma2(a.y, u0[1], nb_fma2(a.x, u0[0], c2))), u0[3]);
    const unsigned long long d1 = nb_add2(nb_fma2(a.z, u1[2], nb_fma2(a.y, u1[1], nb_fma2(a.x, u1[0], c2))), u1[3]);
    const uint32_t b0 = (uint32_t)d0, b1 = (uint32_t)(d0 >> 32), b2 = (uint32_t)d1, b3 = (uint32_t)(d1 >> 32);
    acc = __funnelshift_l(b0, acc, 1); acc = __funnelshift_l(b1, acc, 1); acc = __funnelshift_l(b2, acc, 1); acc = __funnelshift_l(b3, acc, 1);
    t_min = min(t_min, min(min((int)b0, (int)b1), min((int)b2, (int)b3)));
    if (OWN)
    {
      const unsigned long long e0 = nb_add2(d0, zc2), e1 = nb_add2(d1, zc2);
      zacc = __funnelshift_l((uint32_t)e0, zacc, 1); zacc = __funnelshift_l((uint32_t)(e0 >> 32), zacc, 1);
      zacc = __funnelshift_l((uint32_t)e1, zacc, 1); zacc = __funnelshift_l((uint32_t)(e1 >> 32), zacc, 1);
    }
  }
  trk = t_min;
  // candidate j0 + k sits at bit (n4 - 1 - k) of acc
  if (OWN) zbits = __brev(zacc) >> (32u - n4);
  return __brev(acc) >> (32u - n4);
}

// exact decision of the reference for one pair of flat particle indices (rare)
__device__ __noinline__ bool nb_exact(uint32_t self, uint32_t j, double max_dist2,
                                      const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz)
{
  const double d2 = norm2_exact(__dadd_rn(rx[self], -rx[j]), __dadd_rn(ry[self], -ry[j]), __dadd_rn(rz[self], -rz[j]));
  return j != self && d2 > 0.0 && d2 <= max_dist2;
}

// slow path of one (particle, block) mask: every accepted candidate whose fp32 value lies inside the band (or, OWN, inside the
// zero band) is decided exactly
// (gfirst: flat index of the neighbour cell's first particle; bb: staged index of bit 0, i.e. of the cell's particle pb0)
__device__ __noinline__ uint32_t nb_block_exact(const NbPair* __restrict__ S, uint32_t gfirst, uint32_t pb0, uint32_t bb, uint32_t m, uint32_t zm,
                                                const NbSelf a, float band2, uint32_t gself, double max_dist2,
                                                const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz)
{
  uint32_t out = m, w = m;
  while (w)
  {
    const uint32_t b = (uint32_t)__ffs((int)w) - 1u; w &= w - 1u;
    const uint32_t idx = bb + b;
    const float t = nb_value(a, nb_candidate(S, idx));
    if (t > -band2 || ((zm >> b) & 1u))
      if (!nb_exact(gself, gfirst + pb0 + b, max_dist2, rx, ry, rz)) out &= ~(1u << b);
  }
  return out;
}

struct NbhBitsOut
{
  // reference-format streams (GridChunkNeighbors)
  uint16_t* pool; uint16_t** cell_stream; uint32_t* stream_size; uint32_t* cell_stream_bytes; unsigned long long* stream_off;
  // compiled rows: group g of tile b owns rows [(b * gmax + g) * cap_trips, ... + cap_trips)
  uint2* groups; uint2* rows;
  uint32_t* counters; unsigned long long* totals;
};

constexpr int NBH_BITS_MAX_THREADS = 288;   // nine warps: the 17-18 groups of a 4x2x2 tile of 32-particle cells in two even rounds
constexpr int NBH_CELL_BLOCKS = 3;      // accept masks (32 candidates each) per neighbour cell: cells of up to 96 staged (padded) particles
constexpr uint32_t NBH_SLACK_PAIRS = 32u * NBH_CELL_BLOCKS / 2u + 4u;     // a lane of a smaller cell reads on to the warp's longest cell
// staged candidates: every halo cell padded to a multiple of four, + slack
__host__ __device__ inline uint32_t nb_cap_pairs(int cap, int nh_max) { return ((((uint32_t)cap + 3u * (uint32_t)nh_max + 31u) & ~31u) >> 1) + NBH_SLACK_PAIRS; }
// dynamic shared memory of k_nbh_bits: tables | plen[gmax * 32] u16 | hpad[nh_max + 1] | staged pairs | per-warp list areas
__host__ __device__ inline size_t nb_list_area_bytes(int cap_l, int elem) { return (((size_t)32 * cap_l + 64) * (size_t)elem + 15) & ~(size_t)15; }
__host__ __device__ inline size_t nb_smem_bytes(int nh_max, int tc_max, int gmax, int cap, int cap_l, int elem, int nwarp)
{
  return (((size_t)(2 * nh_max + 2 * tc_max + 2) * 4 + 15) & ~(size_t)15) + ((((size_t)gmax * 64) + 15) & ~(size_t)15) + ((((size_t)nh_max + 1) * 4 + 15) & ~(size_t)15) +
         (size_t)nb_cap_pairs(cap, nh_max) * 32 + (size_t)nwarp * nb_list_area_bytes(cap_l, elem);
}

template <class LT> XNB_DEVINL void nb_sts(uint32_t a, uint32_t v)
{
  if (sizeof(LT) == 1) asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory");
  else asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}

// List areas: one per lane, cap_l elements of LT (u8 when cells hold < 256 particles and the neighbourhood has < 128 cells:
// p_b and counts fit a byte, a group header is (0x80 | neighbour slot, n); else u16).  A list is kept in the reference's layout:
// [groups][(cell slot, n, p_b x n) x groups].
__global__ void __launch_bounds__(NBH_BITS_MAX_THREADS, 2)
k_nbh_bits(GridP g, ClTileP tp, NbhBitsP bp,
           const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz,
           const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_count,
           NbhBitsOut out, uint32_t* __restrict__ err)
{
  typedef uint8_t LT;               // byte lists: p_b, counts (cells of < 96 staged particles) and slots (< 128 neighbour cells) fit a byte
  constexpr bool U8 = true;
  constexpr uint32_t ES = (uint32_t)sizeof(LT);
  constexpr uint32_t FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t s_scan[32];
  __shared__ uint32_t s_next, s_ovf;
  __shared__ unsigned s_rmax;
  __shared__ uint32_t s_stat[NB_U32_COUNT];
  __shared__ unsigned long long s_tot[3];
  __shared__ uint16_t s_enc[128];        // neighbour slot -> encoded cell index (chunk_neighbors.h:137-150)
  __shared__ uint32_t s_flag[32];        // group g: list lengths published
  const ClTile T = cl_tile(g, tp, (int)blockIdx.x);
  const ClTables tb = cl_tables(smem_raw, tp);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int gap = tp.gap, n1 = 2 * gap + 1;
  const size_t tbytes = cl_tables_bytes(tp.nh_max, tp.tc_max);
  uint16_t* const plen = reinterpret_cast<uint16_t*>(smem_raw + tbytes);              // list length (stream words) of every tile particle
  const size_t plen_bytes = (((size_t)tp.gmax * 64) + 15) & ~(size_t)15;
  const uint32_t cap_pairs = nb_cap_pairs(tp.cap, tp.nh_max);
  uint32_t* const hpad = reinterpret_cast<uint32_t*>(smem_raw + tbytes + plen_bytes);   // staged (padded) index of every halo cell's first particle
  const size_t hpad_bytes = (((size_t)tp.nh_max + 1) * 4 + 15) & ~(size_t)15;
  NbPair* const S = reinterpret_cast<NbPair*>(smem_raw + tbytes + plen_bytes + hpad_bytes);
  const uint32_t cap_l = (uint32_t)bp.cap_l;
  LT* const Lw = reinterpret_cast<LT*>(smem_raw + tbytes + plen_bytes + hpad_bytes + (size_t)cap_pairs * sizeof(NbPair) + (size_t)warp * nb_list_area_bytes(bp.cap_l, (int)ES));
  LT* const L = Lw + (size_t)lane * cap_l;                                             // this lane's list

  cl_setup(g, T, tb, cell_start, cell_count, s_scan, bp.sel_mode);
  if (threadIdx.x == 0) { s_next = 0u; s_ovf = 0u; s_rmax = 0u; s_tot[0] = s_tot[1] = s_tot[2] = 0ull; }
  if (threadIdx.x < NB_U32_COUNT) s_stat[threadIdx.x] = 0u;
  if (threadIdx.x < 32) s_flag[threadIdx.x] = 0u;
  if (U8)
    for (int sl = threadIdx.x; sl < n1 * n1 * n1; sl += blockDim.x)
    {
      const int ri = sl % n1 - gap, rj = (sl / n1) % n1 - gap, rk = sl / (n1 * n1) - gap;
      s_enc[sl] = (uint16_t)((((rk + 16) << 5) + (rj + 16)) << 5) + (uint16_t)(ri + 16);
    }
  const uint32_t n_tile = tb.tstart[T.tcells], n_halo = tb.hstart[T.NH];
  const uint32_t ngroups = (n_tile + 31u) >> 5;
  uint2* const gt = out.groups + (size_t)blockIdx.x * (size_t)tp.gmax;
  if (bp.emit_rows) for (uint32_t q = min(ngroups, (uint32_t)tp.gmax) + threadIdx.x; q < (uint32_t)tp.gmax; q += blockDim.x) gt[q] = make_uint2(0u, 0u);
  if (threadIdx.x == 0) { atomicMax(&out.counters[NB_GMAX], ngroups); atomicMax(&out.counters[NB_CAP], n_halo); }
  if (ngroups > (uint32_t)tp.gmax || n_halo > (uint32_t)tp.cap)
  {
    if (threadIdx.x == 0) atomicOr(&out.counters[NB_OVERFLOW], 1u);
    return;                                                                      // the host reads the counters and re-runs with more room
  }
  // empty tile cells: no stream
  if ((int)threadIdx.x < T.tcells)
  {
    const int q = (int)threadIdx.x;
    const int ii = q % T.tci, jj = (q / T.tci) % T.tcj, kk = q / (T.tci * T.tcj);
    const int ci = T.ci0 + ii, cj = T.cj0 + jj, ck = T.ck0 + kk;
    const int hq = (int)tb.thalo[q];
    if (tb.hstart[hq + 1] == tb.hstart[hq] && !(bp.sel_mode == 1 && cl_cell_is_inner(g, ci, cj, ck)))
    {
      const int ca = ijk_to_index(g.dims, ci, cj, ck);
      out.cell_stream[ca] = nullptr; out.stream_size[ca] = 0u; out.cell_stream_bytes[ca] = 0u;
      out.stream_off[ca] = (unsigned long long)ca * (unsigned long long)bp.slot_words;
    }
  }
  __syncthreads();

  // ---- padded staging index of every halo cell (cells start at multiples of four), then the staged pairs themselves, relative to
  // the centre O of the box
  {
    uint32_t carry = 0;
    for (int base = 0; base < T.NH; base += blockDim.x)
    {
      const int h = base + threadIdx.x;
      const uint32_t cnt = h < T.NH ? ((tb.hstart[h + 1] - tb.hstart[h] + 3u) & ~3u) : 0u;
      uint32_t total;
      const uint32_t off = block_exclusive_scan<uint32_t>(cnt, &total, s_scan);
      if (h < T.NH) hpad[h] = carry + off;
      carry += total;
    }
    if (threadIdx.x == 0) hpad[T.NH] = carry;
  }
  __syncthreads();
  if (n_tile > 0u)
  {
    const double ox = __dadd_rn(g.org[0], __dmul_rn((double)(g.off[0] + T.bx0) + 0.5 * (double)T.HX, g.cs));
    const double oy = __dadd_rn(g.org[1], __dmul_rn((double)(g.off[1] + T.by0) + 0.5 * (double)T.HY, g.cs));
    const double oz = __dadd_rn(g.org[2], __dmul_rn((double)(g.off[2] + T.bz0) + 0.5 * (double)T.HZ, g.cs));
    float* const Sf = reinterpret_cast<float*>(S);
    float rmax = 0.f;
    for (int h = warp; h < T.NH; h += nwarp)
    {
      const uint32_t d0 = hpad[h], cnt = tb.hstart[h + 1] - tb.hstart[h], cnt4 = hpad[h + 1] - d0, s0 = tb.hfirst[h];
      for (uint32_t p = lane; p < cnt4; p += 32)
      {
        const uint32_t j = d0 + p;
        float* e = Sf + (size_t)(j >> 1) * 8u + (j & 1u);
        if (p < cnt)
        {
          const float x = (float)(rx[s0 + p] - ox), y = (float)(ry[s0 + p] - oy), z = (float)(rz[s0 + p] - oz);
          rmax = fmaxf(rmax, fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z))));
          e[0] = -2.f * x; e[2] = -2.f * y; e[4] = -2.f * z; e[6] = (float)((double)x * x + (double)y * y + (double)z * z);
        }
        else { e[0] = 0.f; e[2] = 0.f; e[4] = 0.f; e[6] = INFINITY; }      // pad: never within any distance
      }
    }
    // slack behind the last candidate: never accepted
    for (uint32_t j = hpad[T.NH] + threadIdx.x; j < cap_pairs * 2u; j += blockDim.x)
    {
      float* e = Sf + (size_t)(j >> 1) * 8u + (j & 1u);
      e[0] = 0.f; e[2] = 0.f; e[4] = 0.f; e[6] = INFINITY;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(FULL, rmax, o));
    if (lane == 0) atomicMax(&s_rmax, __float_as_uint(rmax));
  }
  __syncthreads();
  // classification band: |fp32 value - exact| <= 2^-24 (81 R^2 + 5 max_dist2); the band used is more than twice that
  const double Rm = (double)__uint_as_float(s_rmax);
  const double band = 5.9604644775390625e-08 * (128.0 * Rm * Rm + 4.0 * bp.max_dist2);
  const float band2 = (float)(2.0 * band);                 // accepted and t > -2 band: ambiguous
  const float zc = (float)(bp.max_dist2);                  // t + zc < 0  <=>  fp32 d2 < band: inside the zero band (own cell)
  const uint32_t Lsh = (uint32_t)__cvta_generic_to_shared(L);
  const int q_first = cl_find_cell(tb.tstart, T.tcells, 0u);                   // cell of tile particle 0 (idle lanes of the last group stand on it)
  volatile uint32_t* const vflag = s_flag;
  // every neighbour cell of every tile cell lies inside the grid unless the halo box was clipped
  const bool all_exist = T.HX == T.tci + 2 * gap && T.HY == T.tcj + 2 * gap && T.HZ == T.tck + 2 * gap;

  // ============ one warp per GROUP of the sweep (32 consecutive tile particles, lane = particle; the lanes of a group belong to one,
  // two or -- small cells -- a few tile cells): lists in shared memory, then both outputs from them ======
  for (;;)
  {
    uint32_t ug = 0;
    if (lane == 0) ug = atomicAdd(&s_next, 1u);
    const uint32_t grp = __shfl_sync(FULL, ug, 0);
    if (grp >= ngroups) break;
    const uint32_t t = grp * 32u + (uint32_t)lane;                             // tile particle: group t >> 5, lane t & 31 of the sweep
    const bool active = t < n_tile;
    const int q = active ? cl_find_cell(tb.tstart, T.tcells, t) : q_first;
    const int ii = q % T.tci, jj = (q / T.tci) % T.tcj, kk = q / (T.tci * T.tcj);
    const int cia = T.ci0 + ii, cja = T.cj0 + jj, cka = T.ck0 + kk;
    const int ca = ijk_to_index(g.dims, cia, cja, cka);
    const uint32_t tq = tb.tstart[q];
    const uint32_t nA = tb.tstart[q + 1] - tq;
    const uint32_t pa = active ? t - tq : 0u;
    const int hA = (int)tb.thalo[q];
    const uint32_t self = tb.hstart[hA] + pa, gself = tb.hfirst[hA] + pa;          // index in the sweep's staged halo / in the flat arrays
    NbSelf me;
    {
      const float4 qa = nb_candidate(S, hpad[hA] + pa);
      me.x = -0.5f * qa.x; me.y = -0.5f * qa.y; me.z = -0.5f * qa.z;
      me.c = active ? (float)((double)qa.w - (bp.max_dist2 + band)) : INFINITY;      // idle lanes accept nothing
    }
    // compiled rows of this group: lane's column, four entries per 8-byte word (word k of the column = entries 4k .. 4k+3)
    const uint32_t row0 = (uint32_t)(((size_t)blockIdx.x * (size_t)tp.gmax + grp) * (size_t)bp.cap_trips);
    uint2* const col = out.rows + (size_t)row0 * 32u + (uint32_t)lane;
    uint2* colp = col;                                                             // where the lane's pending word goes
    const uint32_t cap_e = bp.emit_rows ? 4u * (uint32_t)bp.cap_trips : 0xffffffffu;       // entries the lane's column can take
    const uint32_t rows_on = bp.emit_rows ? 0u : 4u;                              // (r & 3) == 4 never holds: no rows
    uint32_t buf_lo = 0u, buf_hi = 0u, r = 0u;
    bool room = true;                                                              // the lane's list and column still take everything

    uint32_t w = 1u, ngrp = 0u, ncand = 0u;          // L[0] = group counter; w = elements of the list so far
    // ---- every lane sweeps the neighbourhood of ITS cell (lanes of one cell read the same candidates: broadcast loads); the accept
    // masks of a neighbour cell stay in registers and are expanded into the lane's list and its column of the rows right away
    for (int rk = -gap; rk <= gap; rk++)
      for (int rj = -gap; rj <= gap; rj++)
        for (int ri = -gap; ri <= gap; ri++)
        {
          const int bi = cia + ri, bj = cja + rj, bk = cka + rk;
          const bool exist = all_exist || (bi >= 0 && bi < g.dims[0] && bj >= 0 && bj < g.dims[1] && bk >= 0 && bk < g.dims[2]);
          const int hB = exist ? hA + (rk * T.HY + rj) * T.HX + ri : 0;
          const uint32_t hp = hpad[hB], nB4 = exist ? hpad[hB + 1] - hp : 0u;          // padded staged range of the cell
          const uint32_t nmax = __reduce_max_sync(FULL, nB4);
          if (nmax == 0u) continue;
          if (nmax > 32u * (uint32_t)NBH_CELL_BLOCKS) { if (lane == 0) { s_ovf = 1u; atomicOr(&out.counters[NB_OVERFLOW], 4u); } continue; }    // host: other kernels
          const bool own = rk == 0 && rj == 0 && ri == 0;
          // bit b of cw[i] = p_b 32 i + b of this cell accepted
          uint32_t cw[NBH_CELL_BLOCKS], cnt = 0;
#pragma unroll
          for (int k = 0; k < NBH_CELL_BLOCKS; k++)
          {
            cw[k] = 0u;
            if ((uint32_t)(32 * k) < nmax)
            {
              const uint32_t bb = hp + (uint32_t)(32 * k), n4 = min(32u, nmax - (uint32_t)(32 * k));
              // candidates of this lane's cell among the n4 (the others belong to the cells behind it in the staged order)
              const uint32_t nv = nB4 > (uint32_t)(32 * k) ? min(32u, nB4 - (uint32_t)(32 * k)) : 0u;
              const uint32_t valid = nv >= 32u ? FULL : ((1u << nv) - 1u);
              int trk = 0x7fffffff; uint32_t zb = 0u, mk;
              if (own)
              {
                mk = nb_block_bits<true>(S, bb, n4, me, zc, trk, zb) & valid;
                const uint32_t sb = pa - (uint32_t)(32 * k);
                if (sb < 32u) { mk &= ~(1u << sb); zb &= ~(1u << sb); }      // never a neighbour of itself
                zb &= mk;
              }
              else mk = nb_block_bits<false>(S, bb, n4, me, zc, trk, zb) & valid;
              const bool amb = mk != 0u && ((trk < 0 && __int_as_float(trk) > -band2) || zb != 0u);
              if (__any_sync(FULL, amb))
              {
                if (amb) { mk = nb_block_exact(S, tb.hfirst[hB], (uint32_t)(32 * k), bb, mk, zb, me, band2, gself, bp.max_dist2, rx, ry, rz); atomicAdd(&s_stat[NB_AMBIGUOUS], 1u); }
              }
              cw[k] = mk; cnt += (uint32_t)__popc(mk);
            }
          }
          if (cnt == 0u) continue;
          room = room && w + 2u + cnt <= cap_l && ncand + cnt <= cap_e;
          if (room)
          {
            // group header: neighbour slot (byte lists) or the encoded cell index (chunk_neighbors.h:137-150), then the count
            nb_sts<LT>(Lsh + w * ES, U8 ? (0x80u | (uint32_t)(((rk + gap) * n1 + (rj + gap)) * n1 + (ri + gap))) : (uint32_t)((((rk + 16) << 5) + (rj + 16)) << 5) + (uint32_t)(ri + 16));
            nb_sts<LT>(Lsh + (w + 1u) * ES, cnt);
            uint32_t wa = Lsh + (w + 2u) * ES;
            const uint32_t hs8 = tb.hstart[hB] << 3;                     // 8 x staged index (sweep) of the cell's first particle
#pragma unroll
            for (int i = 0; i < NBH_CELL_BLOCKS; i++)
            {
              uint32_t x = __brev(cw[i]);
              const uint32_t hs8i = hs8 + (uint32_t)(256 * i);
              while (x)
              {
                const uint32_t b = (uint32_t)__clz((int)x);
                x ^= 0x80000000u >> b;
                nb_sts<LT>(wa, (uint32_t)(32 * i) + b); wa += ES;
                // compiled-row entry: shifted into the lane's pending word, stored every fourth entry
                buf_lo = __funnelshift_r(buf_lo, buf_hi, 16);
                buf_hi = __byte_perm(buf_hi, hs8i + (b << 3), 0x5432);
                r++;
                if ((r & 3u) == rows_on) { *colp = make_uint2(buf_lo, buf_hi); colp += 32; }
              }
            }
          }
          w += 2u + cnt; ngrp++; ncand += cnt;
        }
    const uint32_t len = active ? w : 0u;                 // = 1 + 2 groups + entries
    // ---- capacities
    const uint32_t mxl = __reduce_max_sync(FULL, len);
    const uint32_t my_trips = (ncand + 3u) >> 2;
    const uint32_t trips = __reduce_max_sync(FULL, my_trips);
    const uint32_t mxc = __reduce_max_sync(FULL, ncand), csum = __reduce_add_sync(FULL, ncand);
    if (lane == 0)
    {
      atomicMax(&s_stat[NB_SLOTS], mxl); atomicMax(&s_stat[NB_TRIPS], trips); atomicMax(&s_stat[NB_MAX_NBH], mxc);
      atomicAdd(&s_tot[0], (unsigned long long)csum);
    }
    if (active && (ngrp >= 65535u || ncand >= 65535u || len >= 65535u)) atomicOr(err, DERR_GROUP_OVERFLOW);
    const bool lists_ok = __all_sync(FULL, room);        // else: lists incomplete, the host re-runs with more room
    if (!lists_ok && lane == 0) s_ovf = 1u;
    // ---- compiled rows: the pending word, then pads (the particle's own staged index: d2 = 0 is never inside the cut) up to the
    // group's longest list
    if (bp.emit_rows)
    {
      const uint32_t pad = self << 3, pw = pad | (pad << 16);
      if (lists_ok)
      {
        while (r & 3u) { buf_lo = __funnelshift_r(buf_lo, buf_hi, 16); buf_hi = __byte_perm(buf_hi, pad, 0x5432); r++; if ((r & 3u) == 0u) *colp = make_uint2(buf_lo, buf_hi); }
        for (uint32_t k = my_trips; k < trips; k++) col[(size_t)k * 32u] = make_uint2(pw, pw);
        if (lane == 0) { gt[grp] = make_uint2(row0, trips); atomicAdd(&s_stat[NB_ROWS], trips); }
      }
      else if (lane == 0) { gt[grp] = make_uint2(row0, 0u); s_ovf = 1u; }
    }
    // ---- position of every list in its cell's stream: lengths of the cell's particles in front of it.  Those inside this group come
    // from a scan; the part of the group's first cell that lies in earlier groups is read from the lengths those groups published
    if (active) { plen[t] = (uint16_t)len; if (lists_ok) nb_sts<LT>(Lsh, ngrp); }
    __syncwarp();
    if (lane == 0) { __threadfence_block(); vflag[grp] = 1u; }
    uint32_t x = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(FULL, x, o); if (lane >= o) x += y; }
    const uint32_t same = __match_any_sync(FULL, q);
    const int first = __ffs((int)same) - 1;
    const uint32_t ex = x - len;
    uint32_t off = ex - __shfl_sync(FULL, ex, first);
    const int q0 = __shfl_sync(FULL, q, 0);
    const uint32_t t0 = __shfl_sync(FULL, tq, 0);
    if (t0 < grp * 32u)
    {
      // wait for the groups that hold the front part of the cell (they were claimed earlier and never wait for this one)
      const uint32_t g0 = t0 >> 5;
      for (;;)
      {
        const uint32_t gi = g0 + (uint32_t)lane;
        const bool ready = gi >= grp || vflag[gi] != 0u;
        if (__all_sync(FULL, ready)) break;
        __nanosleep(64);
      }
      __threadfence_block();
      uint32_t carry = 0;
      for (uint32_t u = t0 + (uint32_t)lane; u < grp * 32u; u += 32u) carry += (uint32_t)reinterpret_cast<volatile uint16_t*>(plen)[u];
      carry = __reduce_add_sync(FULL, carry);
      if (q == q0) off += carry;
    }
    const unsigned long long slot_off = (unsigned long long)ca * (unsigned long long)bp.slot_words;
    uint16_t* const base = out.pool + slot_off;
    uint16_t* const lists = base + 2u * (nA + 1u);
    const bool fits = lists_ok && 2u * (nA + 1u) + off + len <= (uint32_t)bp.slot_words;
    if (fits && active)
    {
      // offset table entry (chunk_neighbors_execute.h:279-283) and closing entry (:390-398)
      reinterpret_cast<uint32_t*>(base)[pa] = off + 1u;
      if (pa == nA - 1u) reinterpret_cast<uint32_t*>(base)[nA] = off + len + 1u;
      // ---- the reference-format list: every lane copies its own list, 8-byte stores once the destination is aligned (the lists of a
      // group are adjacent in the stream, partial sectors merge in L2)
      uint16_t* const dst = lists + off;
      const uint32_t head = min(len, (uint32_t)(((8u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 7u)) & 7u) >> 1));
      uint32_t v = 0;
      for (; v < head; v++) dst[v] = (uint16_t)L[v];
      for (; v + 4u <= len; v += 4u)
      {
        const uint32_t w0 = L[v], w1 = L[v + 1], w2 = L[v + 2], w3 = L[v + 3];
        *reinterpret_cast<uint2*>(dst + v) = make_uint2(w0 | (w1 << 16), w2 | (w3 << 16));
      }
      for (; v < len; v++) dst[v] = (uint16_t)L[v];
      if (U8)
      {
        // byte lists: the group headers hold 0x80 | slot; hop over them and write the cell codes
        uint32_t pos = 1;
        for (uint32_t gq = 0; gq < ngrp; gq++) { const uint32_t code = L[pos], n = L[pos + 1]; dst[pos] = s_enc[code & 0x7fu]; pos += 2u + n; }
      }
    }
    // ---- per cell bookkeeping, by the lane of the cell's last particle
    if (active && pa == nA - 1u)
    {
      const uint32_t sz = 2u * (nA + 1u) + off + len;
      const uint32_t szp = (sz + 7u) & ~7u;
      const bool cfits = lists_ok && szp <= (uint32_t)bp.slot_words;
      out.cell_stream[ca] = cfits ? base : nullptr;
      out.stream_size[ca] = sz; out.cell_stream_bytes[ca] = sz * 2u; out.stream_off[ca] = slot_off;
      if (cfits) for (uint32_t p = sz; p < szp; p++) base[p] = 0;       // deterministic padding
      if (!cfits) s_ovf = 1u;
      atomicMax(&s_stat[NB_MAX_CELL], nA); atomicMax(&s_stat[NB_MAX_STREAM], szp); atomicMax(&s_stat[NB_SLOT_WORDS], szp);
      atomicAdd(&s_tot[1], (unsigned long long)szp);
      if (cl_cell_is_inner(g, cia, cja, cka)) { atomicAdd(&s_tot[2], (unsigned long long)szp); atomicAdd(&s_stat[NB_NONEMPTY], 1u); }
    }
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    if (s_ovf) atomicOr(&out.counters[NB_OVERFLOW], 2u);
    atomicMax(&out.counters[NB_SLOTS], s_stat[NB_SLOTS]); atomicMax(&out.counters[NB_SLOT_WORDS], s_stat[NB_SLOT_WORDS]);
    atomicMax(&out.counters[NB_MAX_NBH], s_stat[NB_MAX_NBH]); atomicAdd(&out.counters[NB_NONEMPTY], s_stat[NB_NONEMPTY]);
    atomicMax(&out.counters[NB_MAX_CELL], s_stat[NB_MAX_CELL]); atomicMax(&out.counters[NB_MAX_STREAM], s_stat[NB_MAX_STREAM]);
    atomicMax(&out.counters[NB_TRIPS], s_stat[NB_TRIPS]); atomicAdd(&out.counters[NB_ROWS], s_stat[NB_ROWS]);
    if (s_stat[NB_AMBIGUOUS]) atomicAdd(&out.counters[NB_AMBIGUOUS], s_stat[NB_AMBIGUOUS]);
    for (int q = 0; q < 3; q++) if (s_tot[q]) atomicAdd(&out.totals[q], s_tot[q]);
  }
}

} // namespace xnb
