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
// Masks are walked in staged order = ascending cell code, ascending p_b = the order the reference obtains with its sort
// (:308-324): phase 1 writes the compiled rows (lock-step, coalesced 256-byte rows) and parks the masks in global
// memory (L2); phase 2, after a block barrier that frees the staging memory, expands them into the streams through
// shared memory with coalesced 16-byte stores.
#pragma once
#include "xnb_sweep_cl.cuh"

namespace xnb {

struct NbhBitsP
{
  int cap_slots;         // mask words per lane a warp's area holds
  int emit_rows;         // write the compiled rows + group table (sweep tiles); 0: streams only
  int sel_mode;          // 0: every tile cell is built; 1: only cells outside the inner range (ghost-cell lists, built lazily)
  int slot_words;        // stream capacity per cell (u16 words, multiple of 8)
  uint32_t cap_rows;     // rows the compiled-list buffer holds
  int lane_min;          // segments of a group with at least this many particles of one cell are built lane-per-particle
  double max_dist2;
};

// counters written by k_nbh_bits (u32): what it used and what it would have needed
enum { NB_ROWS = 0, NB_GMAX = 1, NB_CAP = 2, NB_SLOTS = 3, NB_SLOT_WORDS = 4, NB_MAX_NBH = 5, NB_NONEMPTY = 6, NB_MAX_CELL = 7, NB_MAX_STREAM = 8,
       NB_OVERFLOW = 9, NB_AMBIGUOUS = 10, NB_U32_COUNT = 12 };
// u64 totals: [0] list entries of the built particles [1] padded stream words of all built cells [2] of the inner cells among them

XNB_DEVINL bool cl_cell_is_inner(const GridP& g, int ci, int cj, int ck)
{
  return ci >= g.gl && ci < g.dims[0] - g.gl && cj >= g.gl && cj < g.dims[1] - g.gl && ck >= g.gl && ck < g.dims[2] - g.gl;
}

// a row of a particle's neighbourhood: the cells (cia-gap..cia+gap, cja+rj, cka+rk) clamped to the grid; their particles are
// the contiguous staged range [a0, a1)
struct NbRow { uint32_t a0, a1; int h0, ncell, ri0, rj, rk; };
XNB_DEVINL bool nb_row(const GridP& g, const ClTile& T, const uint32_t* __restrict__ hstart, int gap, int cia, int cja, int cka, int r, NbRow& R)
{
  const int n1 = 2 * gap + 1;
  R.rk = r / n1 - gap; R.rj = r - (r / n1) * n1 - gap;
  const int bk = cka + R.rk, bj = cja + R.rj;
  if (bk < 0 || bk >= g.dims[2] || bj < 0 || bj >= g.dims[1]) return false;
  const int bi0 = max(cia - gap, 0), bi1 = min(cia + gap, g.dims[0] - 1);
  R.h0 = ((bk - T.bz0) * T.HY + (bj - T.by0)) * T.HX + (bi0 - T.bx0);
  R.ncell = bi1 - bi0 + 1; R.ri0 = bi0 - cia;
  R.a0 = hstart[R.h0]; R.a1 = hstart[R.h0 + R.ncell];
  return R.a1 > R.a0;
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

// accept bits of one particle against the staged candidates [j0, j1) of one 32-aligned block (j0 >> 5 == (j1 - 1) >> 5):
// bit (j & 31) of the result = candidate j has t < 0.  trk = integer minimum of the raw bits of every t computed (the negative
// value closest to zero, if any is negative).  OWN: zbits = candidates with t + zc < 0, i.e. fp32 d2 below the zero band.
template <bool OWN>
XNB_DEVINL uint32_t nb_block_bits(const NbPair* __restrict__ S, uint32_t j0, uint32_t j1, const NbSelf& a, float zc, int& trk, uint32_t& zbits)
{
  const uint32_t js = j0 & ~3u, je = (j1 + 3u) & ~3u;
  const unsigned long long c2 = nb_pack2(a.c);
  const unsigned long long zc2 = nb_pack2(zc);
  uint32_t acc = 0, zacc = 0;
  int t_min = trk;
#pragma unroll 2
  for (uint32_t j = js; j < je; j += 4u)
  {
    const NbPair p0 = S[j >> 1], p1 = S[(j >> 1) + 1u];
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
  // candidate j sits at bit (je - 1 - j) of acc
  const uint32_t p4 = je - js;
  const uint32_t lo = j0 & 31u, n = j1 - j0;
  const uint32_t valid = (n >= 32u ? 0xffffffffu : ((1u << n) - 1u)) << lo;
  const uint32_t m = ((__brev(acc) >> (32u - p4)) << (js & 31u)) & valid;
  if (OWN) zbits = ((__brev(zacc) >> (32u - p4)) << (js & 31u)) & valid;
  return m;
}

// global index of the staged candidate idx of row R
XNB_DEVINL uint32_t nb_global_index(const ClTables& tb, const NbRow& R, uint32_t idx, uint32_t* p_b = nullptr, int* cell_in_row = nullptr)
{
  int c = 0;
  while (c + 1 < R.ncell && idx >= tb.hstart[R.h0 + c + 1]) c++;
  const uint32_t pb = idx - tb.hstart[R.h0 + c];
  if (p_b) *p_b = pb;
  if (cell_in_row) *cell_in_row = c;
  return tb.hfirst[R.h0 + c] + pb;
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
__device__ __noinline__ uint32_t nb_block_exact(const NbPair* __restrict__ S, const ClTables tb, const NbRow R, uint32_t ab, uint32_t m, uint32_t zm,
                                                const NbSelf a, float band2, uint32_t gself, double max_dist2,
                                                const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz)
{
  uint32_t out = m, w = m;
  while (w)
  {
    const uint32_t b = (uint32_t)__ffs((int)w) - 1u; w &= w - 1u;
    const uint32_t idx = (ab << 5) + b;
    const float t = nb_value(a, nb_candidate(S, idx));
    if (t > -band2 || ((zm >> b) & 1u))
      if (!nb_exact(gself, nb_global_index(tb, R, idx), max_dist2, rx, ry, rz)) out &= ~(1u << b);
  }
  return out;
}

// walks the accept masks of one particle in staged order (= stream order)
struct NbWalk
{
  int cia, cja, cka, r, nrows;
  uint32_t ab, ab_end, s, m;
  XNB_DEVINL void open(int ci, int cj, int ck, int gap) { cia = ci; cja = cj; cka = ck; r = -1; nrows = (2 * gap + 1) * (2 * gap + 1); ab = 1u; ab_end = 0u; s = 0xffffffffu; m = 0u; }
  // next staged index, 0xFFFFFFFF when the list is exhausted.  M: this lane's mask words, stride 32 u32
  XNB_DEVINL uint32_t next(const GridP& g, const ClTile& T, const uint32_t* __restrict__ hstart, int gap, const uint32_t* __restrict__ M)
  {
    while (m == 0u)
    {
      if (ab < ab_end) { ab++; }
      else
      {
        NbRow R; bool ok = false;
        while (++r < nrows) { if (nb_row(g, T, hstart, gap, cia, cja, cka, r, R)) { ok = true; break; } }
        if (!ok) { r = nrows; return 0xffffffffu; }
        ab = R.a0 >> 5; ab_end = (R.a1 - 1u) >> 5;
      }
      s++;
      m = M[s * 32u];
    }
    const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
    m &= m - 1u;
    return (ab << 5) + b;
  }
};

struct NbhBitsOut
{
  // reference-format streams (GridChunkNeighbors)
  uint16_t* pool; uint16_t** cell_stream; uint32_t* stream_size; uint32_t* cell_stream_bytes; unsigned long long* stream_off;
  // compiled rows
  uint2* groups; uint2* rows;
  // masks parked between the two phases: [(tile * gmax + group) * cap_slots + slot][32]
  uint32_t* gmasks;
  uint32_t* counters; unsigned long long* totals;
};

constexpr int NBH_BITS_THREADS = 256;
__host__ __device__ inline uint32_t nb_cap_pairs(int cap) { return ((((uint32_t)cap + 31u) & ~31u) >> 1) + 2u; }
// dynamic shared memory of k_nbh_bits: tables | plen | staged pairs | per-warp mask areas
__host__ __device__ inline size_t nb_smem_bytes(int nh_max, int tc_max, int gmax, int cap, int cap_slots, int nwarp)
{
  return (((size_t)(2 * nh_max + 2 * tc_max + 2) * 4 + 15) & ~(size_t)15) + ((((size_t)gmax * 64) + 15) & ~(size_t)15) + (size_t)nb_cap_pairs(cap) * 32 + (size_t)nwarp * (size_t)cap_slots * 128;
}

__global__ void __launch_bounds__(NBH_BITS_THREADS, 2)
k_nbh_bits(GridP g, ClTileP tp, NbhBitsP bp,
           const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz,
           const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_count,
           NbhBitsOut out, uint32_t* __restrict__ err)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t s_scan[32];
  __shared__ uint32_t s_next, s_ovf;
  __shared__ unsigned s_rmax;
  __shared__ uint32_t s_stat[NB_U32_COUNT];
  __shared__ unsigned long long s_tot[3];
  const ClTile T = cl_tile(g, tp, (int)blockIdx.x);
  const ClTables tb = cl_tables(smem_raw, tp);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int gap = tp.gap;
  // dynamic shared memory: tables | plen[gmax * 32] u16 | staged pairs [cap_pairs] | per-warp mask areas
  const size_t tbytes = cl_tables_bytes(tp.nh_max, tp.tc_max);
  uint16_t* const plen = reinterpret_cast<uint16_t*>(smem_raw + tbytes);
  const size_t plen_bytes = (((size_t)tp.gmax * 64) + 15) & ~(size_t)15;
  const uint32_t cap_pairs = nb_cap_pairs(tp.cap);                                 // staged candidates rounded up to 32, + slack
  NbPair* const S = reinterpret_cast<NbPair*>(smem_raw + tbytes + plen_bytes);
  uint32_t* const Mall = reinterpret_cast<uint32_t*>(smem_raw + tbytes + plen_bytes + (size_t)cap_pairs * sizeof(NbPair));
  uint32_t* const Mw = Mall + (size_t)warp * (size_t)bp.cap_slots * 32u;          // this warp's masks [slot][32]
  const size_t arena_bytes = (size_t)cap_pairs * sizeof(NbPair) + (size_t)nwarp * (size_t)bp.cap_slots * 128u;

  cl_setup(g, T, tb, cell_start, cell_count, s_scan, bp.sel_mode);
  if (threadIdx.x == 0) { s_next = 0u; s_ovf = 0u; s_rmax = 0u; s_tot[0] = s_tot[1] = s_tot[2] = 0ull; }
  if (threadIdx.x < NB_U32_COUNT) s_stat[threadIdx.x] = 0u;
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
  __syncthreads();
  if (n_tile == 0u)
  {
    // nothing to build here; cells of the tile that are selected and empty still get their (empty) stream entries
    for (int q = threadIdx.x; q < T.tcells; q += blockDim.x)
    {
      const int ii = q % T.tci, jj = (q / T.tci) % T.tcj, kk = q / (T.tci * T.tcj);
      if (bp.sel_mode == 1 && cl_cell_is_inner(g, T.ci0 + ii, T.cj0 + jj, T.ck0 + kk)) continue;
      const int c = ijk_to_index(g.dims, T.ci0 + ii, T.cj0 + jj, T.ck0 + kk);
      out.cell_stream[c] = nullptr; out.stream_size[c] = 0u; out.cell_stream_bytes[c] = 0u; out.stream_off[c] = (unsigned long long)c * (unsigned long long)bp.slot_words;
    }
    return;
  }

  // ---- stage the halo box: pairs of candidates, relative to the centre O of the box
  const double ox = __dadd_rn(g.org[0], __dmul_rn((double)(g.off[0] + T.bx0) + 0.5 * (double)T.HX, g.cs));
  const double oy = __dadd_rn(g.org[1], __dmul_rn((double)(g.off[1] + T.by0) + 0.5 * (double)T.HY, g.cs));
  const double oz = __dadd_rn(g.org[2], __dmul_rn((double)(g.off[2] + T.bz0) + 0.5 * (double)T.HZ, g.cs));
  {
    float* const Sf = reinterpret_cast<float*>(S);
    float rmax = 0.f;
    for (int h = warp; h < T.NH; h += nwarp)
    {
      const uint32_t d0 = tb.hstart[h], cnt = tb.hstart[h + 1] - d0, s0 = tb.hfirst[h];
      for (uint32_t p = lane; p < cnt; p += 32)
      {
        const float x = (float)(rx[s0 + p] - ox), y = (float)(ry[s0 + p] - oy), z = (float)(rz[s0 + p] - oz);
        rmax = fmaxf(rmax, fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z))));
        const uint32_t j = d0 + p;
        float* e = Sf + (size_t)(j >> 1) * 8u + (j & 1u);
        e[0] = -2.f * x; e[2] = -2.f * y; e[4] = -2.f * z; e[6] = (float)((double)x * x + (double)y * y + (double)z * z);
      }
    }
    // slack behind the last candidate: never accepted
    for (uint32_t j = n_halo + threadIdx.x; j < cap_pairs * 2u; j += blockDim.x)
    {
      float* e = Sf + (size_t)(j >> 1) * 8u + (j & 1u);
      e[0] = 0.f; e[2] = 0.f; e[4] = 0.f; e[6] = INFINITY;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
    if (lane == 0) atomicMax(&s_rmax, __float_as_uint(rmax));
  }
  __syncthreads();
  // classification band: |fp32 value - exact| <= 2^-24 (81 R^2 + 5 max_dist2); the band used is more than twice that
  const double Rm = (double)__uint_as_float(s_rmax);
  const double band = 5.9604644775390625e-08 * (192.0 * Rm * Rm + 12.0 * bp.max_dist2);
  const float band2 = (float)(2.0 * band);                 // accepted and t > -2 band: ambiguous
  const float zc = (float)(bp.max_dist2);                  // t + zc < 0  <=>  fp32 d2 < band: inside the zero band (own cell row)
  const int mid_row = gap * (2 * gap + 1) + gap;

  // ================================================= phase 1: masks, compiled rows ===========================================
  for (;;)
  {
    uint32_t grp = 0;
    if (lane == 0) grp = atomicAdd(&s_next, 1u);
    grp = __shfl_sync(0xffffffffu, grp, 0);
    if (grp >= ngroups) break;
    const uint32_t t = grp * 32u + lane;
    const bool active = t < n_tile;
    const int q = cl_find_cell(tb.tstart, T.tcells, active ? t : 0u);
    const uint32_t pa = (active ? t : 0u) - tb.tstart[q];
    const int hA = (int)tb.thalo[q];
    const uint32_t self = tb.hstart[hA] + pa, gself = tb.hfirst[hA] + pa;
    const int ii = q % T.tci, jj = (q / T.tci) % T.tcj, kk = q / (T.tci * T.tcj);
    const int cia = T.ci0 + ii, cja = T.cj0 + jj, cka = T.ck0 + kk;
    NbSelf me;
    {
      const float4 qa = nb_candidate(S, self);
      me.x = -0.5f * qa.x; me.y = -0.5f * qa.y; me.z = -0.5f * qa.z;
      me.c = (float)((double)qa.w - (bp.max_dist2 + band));
    }
    uint32_t ncand = 0, nslots = 0;
    // ---- segments of the group: lanes of the same tile cell
    uint32_t todo = __ballot_sync(0xffffffffu, active);
    while (todo)
    {
      const int first = __ffs((int)todo) - 1;
      const int qs = __shfl_sync(0xffffffffu, q, first);
      const uint32_t seg = __ballot_sync(0xffffffffu, active && q == qs) & todo;
      todo &= ~seg;
      const bool mine = (seg >> lane) & 1u;
      const int sci = __shfl_sync(0xffffffffu, cia, first), scj = __shfl_sync(0xffffffffu, cja, first), sck = __shfl_sync(0xffffffffu, cka, first);
      uint32_t slot = 0;
      if (__popc(seg) >= bp.lane_min)
      {
        // ---- lane = particle; every lane sweeps the same candidates (broadcast loads)
        NbSelf a = me;
        if (!mine) { a.x = a.y = a.z = 0.f; a.c = INFINITY; }          // lanes of other segments accept nothing
        const int nrows = (2 * gap + 1) * (2 * gap + 1);
        for (int r = 0; r < nrows; r++)
        {
          NbRow R;
          if (!nb_row(g, T, tb.hstart, gap, sci, scj, sck, r, R)) continue;
          const bool own = r == mid_row;
          for (uint32_t ab = R.a0 >> 5; ab <= (R.a1 - 1u) >> 5; ab++, slot++)
          {
            const uint32_t j0 = max(R.a0, ab << 5), j1 = min(R.a1, (ab << 5) + 32u);
            int trk = 0x7fffffff; uint32_t zb = 0u;
            uint32_t m = own ? nb_block_bits<true>(S, j0, j1, a, zc, trk, zb) : nb_block_bits<false>(S, j0, j1, a, zc, trk, zb);
            if (own) { if ((self >> 5) == ab) { m &= ~(1u << (self & 31u)); zb &= ~(1u << (self & 31u)); } zb &= m; }
            const bool amb = mine && ((trk < 0 && __int_as_float(trk) > -band2) || zb != 0u);
            if (__any_sync(0xffffffffu, amb))
            {
              if (amb) { m = nb_block_exact(S, tb, R, ab, m, zb, a, band2, gself, bp.max_dist2, rx, ry, rz); atomicAdd(&s_stat[NB_AMBIGUOUS], 1u); }
            }
            if (mine) { ncand += (uint32_t)__popc(m); if (slot < (uint32_t)bp.cap_slots) Mw[slot * 32u + lane] = m; }
          }
        }
      }
      else
      {
        // ---- a few particles of this cell: lane = candidate, one ballot per particle and block
        const int nrows = (2 * gap + 1) * (2 * gap + 1);
        for (int r = 0; r < nrows; r++)
        {
          NbRow R;
          if (!nb_row(g, T, tb.hstart, gap, sci, scj, sck, r, R)) continue;
          const bool own = r == mid_row;
          for (uint32_t ab = R.a0 >> 5; ab <= (R.a1 - 1u) >> 5; ab++, slot++)
          {
            const uint32_t j = (ab << 5) + lane;
            const bool valid = j >= R.a0 && j < R.a1;
            const float4 qj = nb_candidate(S, j);
            uint32_t gj = 0xffffffffu;
            uint32_t w = seg;
            while (w)
            {
              const int al = __ffs((int)w) - 1; w &= w - 1u;
              NbSelf a;
              a.x = __shfl_sync(0xffffffffu, me.x, al); a.y = __shfl_sync(0xffffffffu, me.y, al); a.z = __shfl_sync(0xffffffffu, me.z, al); a.c = __shfl_sync(0xffffffffu, me.c, al);
              const uint32_t self_a = __shfl_sync(0xffffffffu, self, al), gself_a = __shfl_sync(0xffffffffu, gself, al);
              const float tv = nb_value(a, qj);
              bool acc = valid && tv < 0.f && !(own && j == self_a);
              const bool amb = acc && (tv > -band2 || (own && tv + zc < 0.f));
              if (__any_sync(0xffffffffu, amb))
              {
                if (amb)
                {
                  if (gj == 0xffffffffu) gj = nb_global_index(tb, R, j);
                  acc = nb_exact(gself_a, gj, bp.max_dist2, rx, ry, rz);
                  atomicAdd(&s_stat[NB_AMBIGUOUS], 1u);
                }
              }
              const uint32_t m = __ballot_sync(0xffffffffu, acc);
              if (lane == al) { ncand += (uint32_t)__popc(m); if (slot < (uint32_t)bp.cap_slots) Mw[slot * 32u + lane] = m; }
            }
          }
        }
      }
      if (mine) nslots = slot;
    }
    // ---- capacities
    uint32_t smax = nslots;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) smax = max(smax, __shfl_xor_sync(0xffffffffu, smax, o));
    if (smax > (uint32_t)bp.cap_slots)
    {
      if (lane == 0) { atomicMax(&s_stat[NB_SLOTS], smax); s_ovf = 1u; }
      continue;                                                                  // masks incomplete: the host re-runs with more slots
    }
    if (lane == 0) atomicMax(&s_stat[NB_SLOTS], smax);
    __syncwarp();
    // ---- number of non-empty neighbour cells of every particle (group headers of its stream list)
    uint32_t ngrp_cells = 0;
    if (active)
    {
      uint32_t slot = 0;
      const int nrows = (2 * gap + 1) * (2 * gap + 1);
      for (int r = 0; r < nrows; r++)
      {
        NbRow R;
        if (!nb_row(g, T, tb.hstart, gap, cia, cja, cka, r, R)) continue;
        const uint32_t ab0 = R.a0 >> 5;
        for (int cc = 0; cc < R.ncell; cc++)
        {
          const uint32_t hs = tb.hstart[R.h0 + cc], he = tb.hstart[R.h0 + cc + 1];
          if (he == hs) continue;
          uint32_t any = 0;
          for (uint32_t ab = hs >> 5; ab <= (he - 1u) >> 5; ab++)
          {
            const uint32_t lo = max(hs, ab << 5) & 31u, n = min(he, (ab << 5) + 32u) - max(hs, ab << 5);
            const uint32_t rm = (n >= 32u ? 0xffffffffu : ((1u << n) - 1u)) << lo;
            any |= Mw[(slot + ab - ab0) * 32u + lane] & rm;
          }
          if (any) ngrp_cells++;
        }
        slot += ((R.a1 - 1u) >> 5) - ab0 + 1u;
      }
    }
    const uint32_t len = active ? 1u + 2u * ngrp_cells + ncand : 0u;
    if (active && (len > 65535u || ncand >= 65535u)) atomicOr(err, DERR_GROUP_OVERFLOW);
    plen[t] = (uint16_t)len;
    uint32_t mxc = ncand, csum = ncand, trips = (ncand + 3u) >> 2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      mxc = max(mxc, __shfl_xor_sync(0xffffffffu, mxc, o)); csum += __shfl_xor_sync(0xffffffffu, csum, o);
      trips = max(trips, __shfl_xor_sync(0xffffffffu, trips, o));
    }
    if (lane == 0) { atomicMax(&s_stat[NB_MAX_NBH], mxc); atomicAdd(&s_tot[0], (unsigned long long)csum); }
    // ---- masks -> global (phase 2 reads them back once the staging memory is free)
    {
      uint32_t* gm = out.gmasks + ((size_t)blockIdx.x * (size_t)tp.gmax + grp) * (size_t)bp.cap_slots * 32u;
      for (uint32_t s = 0; s < smax; s++) gm[s * 32u + lane] = (s < nslots) ? Mw[s * 32u + lane] : 0u;
    }
    // ---- compiled rows: four candidates per lane and row, rows written in lock-step (one coalesced 256-byte line each)
    if (bp.emit_rows)
    {
      uint32_t row0 = 0;
      if (lane == 0) row0 = atomicAdd(&out.counters[NB_ROWS], trips);
      row0 = __shfl_sync(0xffffffffu, row0, 0);
      const bool fits = row0 + trips <= bp.cap_rows;
      if (lane == 0) gt[grp] = make_uint2(row0, fits ? trips : 0u);
      if (fits)
      {
        uint2* col = out.rows + ((size_t)row0 * 32u + (uint32_t)lane);
        NbWalk wk; wk.open(cia, cja, cka, gap);
        if (!active) wk.r = wk.nrows;                           // idle lanes of the tile's last group: pads only
        const uint32_t pad = self << 3;                         // the particle's own staged index: d2 = 0 is never inside the cut
        for (uint32_t k = 0; k < trips; k++)
        {
          uint32_t wv[4];
#pragma unroll
          for (int u = 0; u < 4; u++)
          {
            uint32_t idx = 0xffffffffu;
            if (wk.r < wk.nrows) idx = wk.next(g, T, tb.hstart, gap, Mw + lane);
            wv[u] = idx == 0xffffffffu ? pad : (idx << 3);
          }
          col[(size_t)k * 32u] = make_uint2(wv[0] | (wv[1] << 16), wv[2] | (wv[3] << 16));
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (s_ovf)
  {
    if (threadIdx.x == 0) { atomicOr(&out.counters[NB_OVERFLOW], 2u); atomicMax(&out.counters[NB_SLOTS], s_stat[NB_SLOTS]); }
    return;
  }

  // ================================================= phase 2: reference-format streams =======================================
  // one warp per tile cell; per chunk of 32 particles: masks back from global into this warp's arena, every lane expands its list
  // into the staging part of the arena, the warp copies the chunk's words out with aligned 16-byte stores
  {
    const size_t per_warp = (arena_bytes / (size_t)nwarp) & ~(size_t)15;
    unsigned char* const arena = reinterpret_cast<unsigned char*>(S) + (size_t)warp * per_warp;
    for (int q = warp; q < T.tcells; q += nwarp)
    {
      const int ii = q % T.tci, jj = (q / T.tci) % T.tcj, kk = q / (T.tci * T.tcj);
      const int cia = T.ci0 + ii, cja = T.cj0 + jj, cka = T.ck0 + kk;
      if (bp.sel_mode == 1 && cl_cell_is_inner(g, cia, cja, cka)) continue;
      const int ca = ijk_to_index(g.dims, cia, cja, cka);
      const uint32_t nA = tb.tstart[q + 1] - tb.tstart[q];
      const unsigned long long slot_off = (unsigned long long)ca * (unsigned long long)bp.slot_words;
      if (nA == 0u)
      {
        if (lane == 0) { out.cell_stream[ca] = nullptr; out.stream_size[ca] = 0u; out.cell_stream_bytes[ca] = 0u; out.stream_off[ca] = slot_off; }
        continue;
      }
      uint16_t* const base = out.pool + slot_off;
      uint16_t* const lists = base + 2u * (nA + 1u);
      // slots of this cell's neighbourhood (same for all its particles)
      uint32_t nslots = 0;
      const int nrows = (2 * gap + 1) * (2 * gap + 1);
      for (int r = 0; r < nrows; r++) { NbRow R; if (nb_row(g, T, tb.hstart, gap, cia, cja, cka, r, R)) nslots += ((R.a1 - 1u) >> 5) - (R.a0 >> 5) + 1u; }
      uint32_t* const Mc = reinterpret_cast<uint32_t*>(arena);                         // [nslots][32]
      uint16_t* const stag = reinterpret_cast<uint16_t*>(arena + (size_t)nslots * 128u);
      const uint32_t stag_cap = (uint32_t)((per_warp - (size_t)nslots * 128u) / 2u);
      uint32_t run = 0;
      bool cell_fits = true;
      for (uint32_t pa0 = 0; pa0 < nA; pa0 += 32u)
      {
        const uint32_t pa = pa0 + lane;
        const bool active = pa < nA;
        const uint32_t t = tb.tstart[q] + (active ? pa : 0u);
        const uint32_t len = active ? (uint32_t)plen[t] : 0u;
        uint32_t x = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        const uint32_t chunk_total = __shfl_sync(0xffffffffu, x, 31);
        const uint32_t off = run + x - len;
        const bool fits = 2u * (nA + 1u) + run + chunk_total <= (uint32_t)bp.slot_words;
        cell_fits = cell_fits && fits;
        if (fits)
        {
          if (active)
          {
            // offset table entry (chunk_neighbors_execute.h:279-283) and closing entry (:390-398)
            reinterpret_cast<uint32_t*>(base)[pa] = off + 1u;
            if (pa == nA - 1u) reinterpret_cast<uint32_t*>(base)[nA] = off + len + 1u;
          }
          // masks of this chunk's particles
          {
            const uint32_t* gm = out.gmasks + ((size_t)blockIdx.x * (size_t)tp.gmax + (t >> 5)) * (size_t)bp.cap_slots * 32u + (t & 31u);
            if (active) for (uint32_t s = 0; s < nslots; s++) Mc[s * 32u + lane] = gm[s * 32u];
          }
          __syncwarp();
          // destination of the chunk's words; the staging copy starts at the same offset modulo 16 bytes
          uint16_t* const dst0 = lists + run;
          const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(dst0) & 15u) >> 1);
          const bool staged = chunk_total + mis + 8u <= stag_cap;
          uint16_t* w = staged ? stag + mis + (off - run) : lists + off;
          if (active)
          {
            uint16_t* const wg = w++;
            uint32_t groups = 0, slot = 0;
            for (int r = 0; r < nrows; r++)
            {
              NbRow R;
              if (!nb_row(g, T, tb.hstart, gap, cia, cja, cka, r, R)) continue;
              const uint32_t ab0 = R.a0 >> 5;
              for (int cc = 0; cc < R.ncell; cc++)
              {
                const uint32_t hs = tb.hstart[R.h0 + cc], he = tb.hstart[R.h0 + cc + 1];
                if (he == hs) continue;
                uint16_t* const hdr = w;
                uint32_t cnt = 0;
                for (uint32_t ab = hs >> 5; ab <= (he - 1u) >> 5; ab++)
                {
                  const uint32_t lo = max(hs, ab << 5) & 31u, n = min(he, (ab << 5) + 32u) - max(hs, ab << 5);
                  const uint32_t rm = (n >= 32u ? 0xffffffffu : ((1u << n) - 1u)) << lo;
                  uint32_t m = Mc[(slot + ab - ab0) * 32u + lane] & rm;
                  const uint32_t b0 = (ab << 5) - hs;                    // p_b of bit 0 of this block (may wrap below zero: only set bits are used)
                  while (m) { const uint32_t b = (uint32_t)__ffs((int)m) - 1u; m &= m - 1u; hdr[2u + cnt] = (uint16_t)(b0 + b); cnt++; }
                }
                if (cnt)
                {
                  // encode_cell_index (chunk_neighbors.h:137-150)
                  const int ri = R.ri0 + cc;
                  hdr[0] = (uint16_t)((((R.rk + 16) << 5) + (R.rj + 16)) << 5) + (uint16_t)(ri + 16);
                  hdr[1] = (uint16_t)cnt;
                  w = hdr + 2u + cnt; groups++;
                }
              }
              slot += ((R.a1 - 1u) >> 5) - ab0 + 1u;
            }
            *wg = (uint16_t)groups;
          }
          __syncwarp();
          if (staged)
          {
            // head (to the next 16-byte boundary of the destination), aligned body, tail
            const uint32_t head = min(chunk_total, (8u - mis) & 7u);
            for (uint32_t v = lane; v < head; v += 32u) dst0[v] = stag[mis + v];
            const uint32_t body = (chunk_total - head) >> 3;
            const uint4* s16 = reinterpret_cast<const uint4*>(stag + mis + head);
            uint4* d16 = reinterpret_cast<uint4*>(dst0 + head);
            for (uint32_t v = lane; v < body; v += 32u) d16[v] = s16[v];
            for (uint32_t v = head + (body << 3) + lane; v < chunk_total; v += 32u) dst0[v] = stag[mis + v];
          }
          __syncwarp();
        }
        run += chunk_total;
      }
      if (lane == 0)
      {
        const uint32_t sz = 2u * (nA + 1u) + run;
        const uint32_t szp = (sz + 7u) & ~7u;
        const bool fits = cell_fits && szp <= (uint32_t)bp.slot_words;
        out.cell_stream[ca] = fits ? base : nullptr;
        out.stream_size[ca] = sz; out.cell_stream_bytes[ca] = sz * 2u; out.stream_off[ca] = slot_off;
        if (fits) for (uint32_t p = sz; p < szp; p++) base[p] = 0;       // deterministic padding
        atomicMax(&s_stat[NB_MAX_CELL], nA); atomicMax(&s_stat[NB_MAX_STREAM], szp); atomicMax(&s_stat[NB_SLOT_WORDS], szp);
        atomicAdd(&s_tot[1], (unsigned long long)szp);
        if (cl_cell_is_inner(g, cia, cja, cka)) { atomicAdd(&s_tot[2], (unsigned long long)szp); atomicAdd(&s_stat[NB_NONEMPTY], 1u); }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    atomicMax(&out.counters[NB_SLOTS], s_stat[NB_SLOTS]); atomicMax(&out.counters[NB_SLOT_WORDS], s_stat[NB_SLOT_WORDS]);
    atomicMax(&out.counters[NB_MAX_NBH], s_stat[NB_MAX_NBH]); atomicAdd(&out.counters[NB_NONEMPTY], s_stat[NB_NONEMPTY]);
    atomicMax(&out.counters[NB_MAX_CELL], s_stat[NB_MAX_CELL]); atomicMax(&out.counters[NB_MAX_STREAM], s_stat[NB_MAX_STREAM]);
    if (s_stat[NB_AMBIGUOUS]) atomicAdd(&out.counters[NB_AMBIGUOUS], s_stat[NB_AMBIGUOUS]);
    for (int q = 0; q < 3; q++) if (s_tot[q]) atomicAdd(&out.totals[q], s_tot[q]);
  }
}

} // namespace xnb
