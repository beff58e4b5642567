// xnb_nbh_big.cuh -- K2 for LARGE cells (hundreds of particles per cell: C1, C4, the Ni deck): k_nbh_big.
//
// reference: particle_neighbors/include/exanb/particle_neighbors/chunk_neighbors_execute.h:110-411 (lists), :160-182 with
// amr/lib/amr_grid_algorithm.cpp:102-218 (AmrSubCellPairCache: only sub-cell pairs closer than the list radius are visited) and
// amr/include/exanb/amr/amr_grid_algorithm.h:66-78,188-299 (the sub-grid: a cell's particles are ordered sub-cell by sub-cell).
//
// Same idea as k_nbh_bits (xnb_nbh_bits.cuh) -- fp32 classification of staged candidates into accept bits, exact fp64 decision inside
// the error band, reference-format streams AND the compiled rows of the sweep from one kernel -- but for cells whose lists do not
// fit shared memory:
//   * one block = ONE cell (the sweep's tile for such cells is 1x1x1), one warp = one group of 32 of its particles.  The halo is staged
//     as fp32 pairs ONE z-PLANE OF CELLS AT A TIME (the lists are ordered plane by plane anyway: cell code = ((rk * 32) + rj) * 32 + ri),
//     every halo cell padded to a multiple of 32, so a 32-candidate BLOCK never straddles two cells; a third of the halo in shared
//     memory lets three to four blocks share an SM;
//   * sub-cell pruning: rebuild_amr has ordered every cell's particles sub-cell by sub-cell, so 32 consecutive particles are compact in
//     space.  Every block gets its bounding box at staging; a group of 32 tile particles (one warp, lane = particle) skips the blocks
//     whose box is farther than the list radius from the group's own box -- the job the reference gives its sub-cell pair cache.  (With
//     an arbitrary in-cell order the boxes are large and nothing is skipped: correct, just slower.)
//   * two passes instead of a list in shared memory: COUNT (accept bits -> per neighbour cell counts, list length; the masks of the
//     blocks that accept anything are parked in L2-resident scratch, 128 bytes per block and group), positions of the lists in the
//     cell's stream (scan + the lengths of the groups in front), FILL (the parked masks expanded straight into the stream and into
//     the lane's column of the compiled rows: no second staging, no second classification).
#pragma once
#include "xnb_nbh_bits.cuh"

namespace xnb {

constexpr int NBH_BIG_MAX_THREADS = 384;      // one warp per group of the cell: cells of at most 384 particles
// dynamic shared memory: tables | plen[gmax * 32] u16 | hpad[nh_max + 1] | block boxes of one plane | per-warp counts | staged pairs of one plane
__host__ __device__ inline size_t nbig_smem_bytes(int nh_max, int tc_max, int gmax, int cap32, int nslots, int nwarp)
{
  return (((size_t)(2 * nh_max + 2 * tc_max + 2) * 4 + 15) & ~(size_t)15) + ((((size_t)gmax * 64) + 15) & ~(size_t)15) + ((((size_t)nh_max + 1) * 4 + 15) & ~(size_t)15) +
         (size_t)(cap32 / 32) * 32 + (size_t)nwarp * (size_t)nslots * 64 + (size_t)cap32 * 16 + 64;
}

// u16 words of one particle's list, written in order through a 64-bit register: aligned 8-byte stores for everything but the words
// that share an 8-byte line with a neighbouring list (first and last line: 2-byte stores).  A quarter of the store requests of
// word-by-word writing, which is what bounds the fill pass (234 M list words per build at C4).
struct NbStreamWriter
{
  uint2* sp; uint32_t lo, hi, sfill, a0; bool first;
  XNB_DEVINL void begin(uint16_t* dst)
  {
    const uintptr_t a = reinterpret_cast<uintptr_t>(dst);
    a0 = (uint32_t)((a >> 1) & 3u); sp = reinterpret_cast<uint2*>(a & ~(uintptr_t)7); lo = hi = 0u; sfill = a0; first = true;
  }
  XNB_DEVINL void push(uint32_t w)
  {
    lo = __funnelshift_r(lo, hi, 16); hi = __byte_perm(hi, w, 0x5432); sfill++;
    if ((sfill & 3u) == 0u)
    {
      if (first) { uint16_t* p = reinterpret_cast<uint16_t*>(sp); const unsigned long long v = ((unsigned long long)hi << 32) | lo; for (uint32_t q = a0; q < 4u; q++) p[q] = (uint16_t)(v >> (16u * q)); first = false; }
      else *sp = make_uint2(lo, hi);
      sp++;
    }
  }
  XNB_DEVINL void end()
  {
    const uint32_t rem = sfill & 3u, base = first ? a0 : 0u;
    if (rem <= base) return;
    const uint32_t npend = rem - base;
    const unsigned long long v = ((unsigned long long)hi << 32) | lo;
    uint16_t* p = reinterpret_cast<uint16_t*>(sp);
    for (uint32_t i = 0; i < npend; i++) p[base + i] = (uint16_t)(v >> (16u * (4u - npend + i)));
  }
};

// accept bits of lane's particle against the 32 staged candidates of block `bb` (staged index of its first candidate), ambiguity resolved
XNB_DEVINL uint32_t nbig_classify(const NbPair* __restrict__ S, uint32_t bb, const NbSelf& me, bool own, uint32_t pa, uint32_t pb0, float zc, float band2,
                                  uint32_t gfirst, uint32_t gself, double max_dist2, const double* __restrict__ rx, const double* __restrict__ ry,
                                  const double* __restrict__ rz, uint32_t* ambiguous)
{
  int trk = 0x7fffffff; uint32_t zb = 0u, mk;
  if (own)
  {
    mk = nb_block_bits<true>(S, bb, 32u, me, zc, trk, zb);
    const uint32_t sb = pa - pb0;
    if (sb < 32u) { mk &= ~(1u << sb); zb &= ~(1u << sb); }      // never a neighbour of itself
    zb &= mk;
  }
  else mk = nb_block_bits<false>(S, bb, 32u, me, zc, trk, zb);
  const bool amb = mk != 0u && ((trk < 0 && __int_as_float(trk) > -band2) || zb != 0u);
  if (__any_sync(0xffffffffu, amb))
  {
    if (amb) { mk = nb_block_exact(S, gfirst, pb0, bb, mk, zb, me, band2, gself, max_dist2, rx, ry, rz); if (ambiguous) atomicAdd(ambiguous, 1u); }
  }
  return mk;
}

__global__ void __launch_bounds__(NBH_BIG_MAX_THREADS, 2)
k_nbh_big(GridP g, ClTileP tp, NbhBitsP bp, int cap32, uint32_t* __restrict__ scratch, int scratch_rows,
          const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz,
          const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_count,
          NbhBitsOut out, uint32_t* __restrict__ err)
{
  constexpr uint32_t FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t s_scan[32];
  __shared__ uint32_t s_ovf;
  __shared__ uint32_t s_stat[NB_U32_COUNT];
  __shared__ unsigned long long s_tot[3];
  const ClTile T = cl_tile(g, tp, (int)blockIdx.x);
  const ClTables tb = cl_tables(smem_raw, tp);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int gap = tp.gap, n1 = 2 * gap + 1, nslots = n1 * n1 * n1;
  const size_t tbytes = cl_tables_bytes(tp.nh_max, tp.tc_max);
  uint16_t* const plen = reinterpret_cast<uint16_t*>(smem_raw + tbytes);              // list length (stream words) of every tile particle
  const size_t plen_bytes = (((size_t)tp.gmax * 64) + 15) & ~(size_t)15;
  uint32_t* const hpad = reinterpret_cast<uint32_t*>(smem_raw + tbytes + plen_bytes);   // padded (to 32) index of every halo cell's first particle
  const size_t hpad_bytes = (((size_t)tp.nh_max + 1) * 4 + 15) & ~(size_t)15;
  float* const BB = reinterpret_cast<float*>(smem_raw + tbytes + plen_bytes + hpad_bytes);   // per block of the plane: min x y z, max x y z (+ 2 unused)
  const size_t bb_bytes = (size_t)(cap32 / 32) * 32;
  uint16_t* const cnts = reinterpret_cast<uint16_t*>(smem_raw + tbytes + plen_bytes + hpad_bytes + bb_bytes) + (size_t)warp * nslots * 32 + lane;   // [slot * 32]
  NbPair* const S = reinterpret_cast<NbPair*>(smem_raw + ((tbytes + plen_bytes + hpad_bytes + bb_bytes + (size_t)nwarp * nslots * 64 + 15) & ~(size_t)15));

  cl_setup(g, T, tb, cell_start, cell_count, s_scan, bp.sel_mode);
  if (threadIdx.x == 0) { s_ovf = 0u; s_tot[0] = s_tot[1] = s_tot[2] = 0ull; }
  if (threadIdx.x < NB_U32_COUNT) s_stat[threadIdx.x] = 0u;
  const uint32_t n_tile = tb.tstart[T.tcells], n_halo = tb.hstart[T.NH];
  const uint32_t ngroups = (n_tile + 31u) >> 5;
  uint2* const gt = out.groups + (size_t)blockIdx.x * (size_t)tp.gmax;
  if (bp.emit_rows) for (uint32_t q = min(ngroups, (uint32_t)tp.gmax) + threadIdx.x; q < (uint32_t)tp.gmax; q += blockDim.x) gt[q] = make_uint2(0u, 0u);
  // padded index of every halo cell (cells start at multiples of 32); a plane's cells are contiguous
  {
    uint32_t carry = 0;
    for (int base = 0; base < T.NH; base += blockDim.x)
    {
      const int h = base + threadIdx.x;
      const uint32_t cnt = h < T.NH ? ((tb.hstart[h + 1] - tb.hstart[h] + 31u) & ~31u) : 0u;
      uint32_t total;
      const uint32_t off = block_exclusive_scan<uint32_t>(cnt, &total, s_scan);
      if (h < T.NH) hpad[h] = carry + off;
      carry += total;
    }
    if (threadIdx.x == 0) hpad[T.NH] = carry;
  }
  __syncthreads();
  const int HXY = T.HX * T.HY;
  uint32_t plane_max = 0;
  for (int z = 0; z < T.HZ; z++) plane_max = max(plane_max, hpad[(z + 1) * HXY] - hpad[z * HXY]);
  if (threadIdx.x == 0) { atomicMax(&out.counters[NB_GMAX], ngroups); atomicMax(&out.counters[NB_CAP], n_halo); atomicMax(&out.counters[NB_SLOTS], plane_max); }
  if (ngroups > (uint32_t)tp.gmax || n_halo > (uint32_t)tp.cap || plane_max > (uint32_t)cap32 || ngroups > (uint32_t)nwarp)
  {
    if (threadIdx.x == 0) atomicOr(&out.counters[NB_OVERFLOW], ngroups > (uint32_t)nwarp && ngroups <= (uint32_t)tp.gmax ? 4u : 1u);
    return;                                                                      // the host reads the counters and re-runs with more room
  }
  const int ci0 = T.ci0, cj0 = T.cj0, ck0 = T.ck0;                               // the tile IS one cell
  const int ca = ijk_to_index(g.dims, ci0, cj0, ck0);
  const unsigned long long slot_off = (unsigned long long)ca * (unsigned long long)bp.slot_words;
  if (n_tile == 0u)
  {
    if (threadIdx.x == 0 && !(bp.sel_mode == 1 && cl_cell_is_inner(g, ci0, cj0, ck0)))
    { out.cell_stream[ca] = nullptr; out.stream_size[ca] = 0u; out.cell_stream_bytes[ca] = 0u; out.stream_off[ca] = slot_off; }
    return;
  }
  const double ox = __dadd_rn(g.org[0], __dmul_rn((double)(g.off[0] + T.bx0) + 0.5 * (double)T.HX, g.cs));
  const double oy = __dadd_rn(g.org[1], __dmul_rn((double)(g.off[1] + T.by0) + 0.5 * (double)T.HY, g.cs));
  const double oz = __dadd_rn(g.org[2], __dmul_rn((double)(g.off[2] + T.bz0) + 0.5 * (double)T.HZ, g.cs));
  // classification band from a bound on |r - O|: half the box (+ the epsilon a particle may stick out of its cell)
  const double Rm = 0.5 * (double)max(T.HX, max(T.HY, T.HZ)) * g.cs * (1.0 + 1e-9);
  const double band = 5.9604644775390625e-08 * (128.0 * Rm * Rm + 4.0 * bp.max_dist2);
  const float band2 = (float)(2.0 * band);
  const float zc = (float)(bp.max_dist2);
  const float prune2 = (float)(bp.max_dist2 + 4.0 * band + 1e-6 * bp.max_dist2);          // boxes farther apart than this hold no pair
  const int hA = (int)tb.thalo[0];
  const uint32_t nA = n_tile;
  uint16_t* const base = out.pool + slot_off;
  uint16_t* const lists = base + 2u * (nA + 1u);

  // ---- this warp's group: lane = particle
  const uint32_t grp = (uint32_t)warp;
  const bool have_group = grp < ngroups;
  const uint32_t t = grp * 32u + (uint32_t)lane;
  const bool active = have_group && t < n_tile;
  const uint32_t pa = active ? t : 0u;
  const uint32_t self = tb.hstart[hA] + pa, gself = tb.hfirst[hA] + pa;
  NbSelf me;
  float glo[3], ghi[3];
  {
    const float x = (float)(rx[gself] - ox), y = (float)(ry[gself] - oy), z = (float)(rz[gself] - oz);
    const float w = (float)((double)x * x + (double)y * y + (double)z * z);
    me.x = x; me.y = y; me.z = z;
    me.c = active ? (float)((double)w - (bp.max_dist2 + band)) : INFINITY;      // idle lanes accept nothing
    glo[0] = active ? x : INFINITY; glo[1] = active ? y : INFINITY; glo[2] = active ? z : INFINITY;
    ghi[0] = active ? x : -INFINITY; ghi[1] = active ? y : -INFINITY; ghi[2] = active ? z : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int d = 0; d < 3; d++) { glo[d] = fminf(glo[d], __shfl_xor_sync(FULL, glo[d], o)); ghi[d] = fmaxf(ghi[d], __shfl_xor_sync(FULL, ghi[d], o)); }
  }
  auto far_block = [&](uint32_t blk) -> bool {
    const float* b = BB + (size_t)blk * 8u;
    const float dx = fmaxf(0.f, fmaxf(b[0] - ghi[0], glo[0] - b[3])), dy = fmaxf(0.f, fmaxf(b[1] - ghi[1], glo[1] - b[4])), dz = fmaxf(0.f, fmaxf(b[2] - ghi[2], glo[2] - b[5]));
    return dx * dx + dy * dy + dz * dz > prune2;
  };
  // block-cooperative: stage halo plane z (fp32 pairs relative to O, box of every 32-candidate block); ends with a barrier
  auto stage_plane = [&](int z) {
    __syncthreads();                                                             // everybody is done with the previous plane
    float* const Sf = reinterpret_cast<float*>(S);
    const uint32_t pl0 = hpad[z * HXY];
    for (int h = z * HXY + warp; h < (z + 1) * HXY; h += nwarp)
    {
      const uint32_t d0 = hpad[h] - pl0, cnt = tb.hstart[h + 1] - tb.hstart[h], cnt32 = hpad[h + 1] - hpad[h], s0 = tb.hfirst[h];
      for (uint32_t p0 = 0; p0 < cnt32; p0 += 32u)
      {
        const uint32_t p = p0 + (uint32_t)lane, j = d0 + p;
        float* e = Sf + (size_t)(j >> 1) * 8u + (j & 1u);
        float x = 0.f, y = 0.f, z2 = 0.f;
        const bool valid = p < cnt;
        if (valid)
        {
          x = (float)(rx[s0 + p] - ox); y = (float)(ry[s0 + p] - oy); z2 = (float)(rz[s0 + p] - oz);
          e[0] = -2.f * x; e[2] = -2.f * y; e[4] = -2.f * z2; e[6] = (float)((double)x * x + (double)y * y + (double)z2 * z2);
        }
        else { e[0] = 0.f; e[2] = 0.f; e[4] = 0.f; e[6] = INFINITY; }      // pad: never within any distance
        float lo[3] = {valid ? x : INFINITY, valid ? y : INFINITY, valid ? z2 : INFINITY}, hi[3] = {valid ? x : -INFINITY, valid ? y : -INFINITY, valid ? z2 : -INFINITY};
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int d = 0; d < 3; d++) { lo[d] = fminf(lo[d], __shfl_xor_sync(FULL, lo[d], o)); hi[d] = fmaxf(hi[d], __shfl_xor_sync(FULL, hi[d], o)); }
        if (lane < 3) { BB[(size_t)(j >> 5) * 8u + lane] = lo[lane]; BB[(size_t)(j >> 5) * 8u + 3u + lane] = hi[lane]; }
      }
    }
    __syncthreads();
  };

  // ---- COUNT pass, plane by plane.  The accept masks of the surviving blocks are parked in this group's scratch rows (row n = the 32
  // lanes' masks of the n-th surviving block, one coalesced 128-byte store; tag = slot * 16 + block index behind the rows): the FILL
  // pass reads them back from L2 instead of staging and classifying everything again
  uint32_t ngrp = 0u, ncand = 0u, n_surv = 0u;
  uint32_t ncand_pl[3] = {0u, 0u, 0u};                 // planes mode: list entries per z-plane of the halo (T.HZ <= 3: one neighbour layer)
  uint32_t* const my_rows = scratch + ((size_t)blockIdx.x * (size_t)tp.gmax + (size_t)min(warp, tp.gmax - 1)) * (size_t)scratch_rows * 33u;
  uint32_t* const my_tags = my_rows + (size_t)scratch_rows * 32u;
  for (int rk = -gap; rk <= gap; rk++)
  {
    const int bk = ck0 + rk;
    if (bk < 0 || bk >= g.dims[2]) { if (have_group) for (int sl = 0; sl < n1 * n1; sl++) cnts[((rk + gap) * n1 * n1 + sl) * 32] = 0; continue; }     // (block-uniform)
    const int z = bk - T.bz0;
    stage_plane(z);
    if (!have_group) continue;
    const uint32_t pl0 = hpad[z * HXY];
    int slot = (rk + gap) * n1 * n1;
    for (int rj = -gap; rj <= gap; rj++) for (int ri = -gap; ri <= gap; ri++, slot++)
    {
      const int bi = ci0 + ri, bj = cj0 + rj;
      uint32_t cntc = 0u;
      if (bi >= 0 && bi < g.dims[0] && bj >= 0 && bj < g.dims[1])
      {
        const int hB = hA + (rk * T.HY + rj) * T.HX + ri;
        const uint32_t hp = hpad[hB] - pl0, nblk = (hpad[hB + 1] - hpad[hB]) >> 5;
        const bool own = rk == 0 && rj == 0 && ri == 0;
        const uint32_t gfirst = tb.hfirst[hB];
        for (uint32_t k = 0; k < nblk; k++)
        {
          if (far_block((hp >> 5) + k)) continue;
          const uint32_t mk = nbig_classify(S, hp + 32u * k, me, own, pa, 32u * k, zc, band2, gfirst, gself, bp.max_dist2, rx, ry, rz, &s_stat[NB_AMBIGUOUS]);
          cntc += (uint32_t)__popc(mk);
          if (!__any_sync(FULL, mk != 0u)) continue;                 // nobody accepts anything of this block
          if (n_surv < (uint32_t)scratch_rows) { my_rows[n_surv * 32u + (uint32_t)lane] = mk; if (lane == 0) my_tags[n_surv] = (uint32_t)slot * 16u + k; }
          n_surv++;
        }
      }
      cnts[slot * 32] = (uint16_t)cntc;
      if (cntc) { ngrp++; ncand += cntc; }
      if (bp.planes) { if (z == 0) ncand_pl[0] += cntc; else if (z == 1) ncand_pl[1] += cntc; else ncand_pl[2] += cntc; }
    }
  }
  const uint32_t len = active ? 1u + 2u * ngrp + ncand : 0u;
  // rows of the group: one segment, or (planes) one segment per halo plane, each as long as its longest lane
  uint32_t trips_pl[3] = {0u, 0u, 0u};
  if (bp.planes) for (int q = 0; q < 3; q++) trips_pl[q] = __reduce_max_sync(FULL, (ncand_pl[q] + 3u) >> 2);
  const uint32_t my_trips = (ncand + 3u) >> 2;
  const uint32_t trips = bp.planes ? trips_pl[0] + trips_pl[1] + trips_pl[2] : __reduce_max_sync(FULL, my_trips);
  if (have_group)
  {
    const uint32_t mxc = __reduce_max_sync(FULL, ncand), csum = __reduce_add_sync(FULL, ncand);
    if (lane == 0) { atomicMax(&s_stat[NB_TRIPS], trips); atomicMax(&s_stat[NB_MAX_NBH], mxc); atomicAdd(&s_tot[0], (unsigned long long)csum); }
    if (active && (ngrp >= 65535u || ncand >= 65535u || len >= 65535u)) atomicOr(err, DERR_GROUP_OVERFLOW);
    plen[t] = (uint16_t)len;
  }
  __syncthreads();
  // ---- position of every list in the cell's stream: lengths of the groups in front + scan inside the group
  uint32_t off = 0;
  bool fill = false;
  const uint32_t row0 = (uint32_t)(((size_t)blockIdx.x * (size_t)tp.gmax + grp) * (size_t)bp.cap_trips);
  uint2* const col = out.rows + (size_t)row0 * 32u + (uint32_t)lane;
  if (have_group)
  {
    uint32_t x = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(FULL, x, o); if (lane >= o) x += y; }
    uint32_t carry = 0;
    for (uint32_t u = (uint32_t)lane; u < grp * 32u; u += 32u) carry += (uint32_t)plen[u];
    off = x - len + __reduce_add_sync(FULL, carry);
    const bool fits = 2u * (nA + 1u) + off + len <= (uint32_t)bp.slot_words;
    const bool rows_ok = (!bp.emit_rows || trips <= (uint32_t)bp.cap_trips) && n_surv <= (uint32_t)scratch_rows;
    if (lane == 0) atomicMax(&s_stat[NB_SURV], n_surv);
    if (n_surv > (uint32_t)scratch_rows && lane == 0) atomicOr(&out.counters[NB_OVERFLOW], 8u);
    fill = __all_sync(FULL, fits || !active) && rows_ok;
    if (!fill && lane == 0) { s_ovf = 1u; if (bp.emit_rows) gt[grp] = make_uint2(row0, 0u); }
  }
  // ---- FILL pass: the parked masks, block by block; bits expanded straight into the stream and the lane's column of the rows
  uint16_t* dst = lists + off;
  uint2* colp = col;
  uint32_t buf_lo = 0u, buf_hi = 0u, r = 0u;
  const uint32_t rows_on = bp.emit_rows ? 0u : 4u;
  NbStreamWriter sw; sw.begin(dst);
  if (fill && active) { reinterpret_cast<uint32_t*>(base)[pa] = off + 1u; if (pa == nA - 1u) reinterpret_cast<uint32_t*>(base)[nA] = off + len + 1u; sw.push(ngrp); }
  if (fill)
  {
    int prev_slot = -1, cur_z = 0; uint32_t hs8 = 0u, seg_row = 0u;
    const uint32_t pad_pl = (uint32_t)bp.cap_pl << 3, pw_pl = pad_pl | (pad_pl << 16);
    // planes: close the segment of plane cur_z (pending word, then sentinel pads up to the group's trip count of that plane)
    auto close_plane = [&]() {
      while (r & 3u) { buf_lo = __funnelshift_r(buf_lo, buf_hi, 16); buf_hi = __byte_perm(buf_hi, pad_pl, 0x5432); r++; if ((r & 3u) == 0u) { *colp = make_uint2(buf_lo, buf_hi); colp += 32; } }
      for (uint32_t kq = r >> 2; kq < trips_pl[cur_z]; kq++) { *colp = make_uint2(pw_pl, pw_pl); colp += 32; }
      seg_row += trips_pl[cur_z]; colp = col + (size_t)seg_row * 32u; r = 0u; cur_z++;
    };
    // tags 32 at a time (one coalesced load, then shuffles); the masks one block ahead of the expansion
    uint32_t tagv = 0u, mk_next = n_surv ? my_rows[(uint32_t)lane] : 0u;
    for (uint32_t n = 0; n < n_surv; n++)
    {
      if ((n & 31u) == 0u) tagv = n + (uint32_t)lane < n_surv ? my_tags[n + (uint32_t)lane] : 0u;
      const uint32_t tag = __shfl_sync(FULL, tagv, (int)(n & 31u));
      const uint32_t mk_cur = mk_next;
      if (n + 1u < n_surv) mk_next = my_rows[(n + 1u) * 32u + (uint32_t)lane];
      const int slot = (int)(tag >> 4); const uint32_t k = tag & 15u;
      if (slot != prev_slot)
      {
        // next neighbour cell: its header (cell code, count) for the lanes that list something of it
        prev_slot = slot;
        const int ri = slot % n1 - gap, rj = (slot / n1) % n1 - gap, rk = slot / (n1 * n1) - gap;
        const uint32_t cntc = cnts[slot * 32];
        if (cntc) { sw.push((uint32_t)((((rk + 16) << 5) + (rj + 16)) << 5) + (uint32_t)(ri + 16)); sw.push(cntc); }
        hs8 = tb.hstart[hA + (rk * T.HY + rj) * T.HX + ri] << 3;
        if (bp.planes)
        {
          const int z = ck0 + rk - T.bz0;
          if (bp.emit_rows) while (cur_z < z) close_plane();
          hs8 -= tb.hstart[z * HXY] << 3;                      // indices relative to the plane's first particle
        }
      }
      uint32_t xm = __brev(mk_cur);
      const uint32_t pb0 = 32u * k, hs8k = hs8 + 256u * k;
      while (xm)
      {
        const uint32_t b = (uint32_t)__clz((int)xm);
        xm ^= 0x80000000u >> b;
        sw.push(pb0 + b);
        buf_lo = __funnelshift_r(buf_lo, buf_hi, 16);
        buf_hi = __byte_perm(buf_hi, hs8k + (b << 3), 0x5432);
        r++;
        if ((r & 3u) == rows_on) { *colp = make_uint2(buf_lo, buf_hi); colp += 32; }
      }
    }
    if (bp.planes && bp.emit_rows)
    {
      while (cur_z < T.HZ) close_plane();                    // the open segment and the planes behind it
      if (lane == 0) { gt[grp] = make_uint2(row0, trips_pl[0] | (trips_pl[1] << 10) | (trips_pl[2] << 20)); atomicAdd(&s_stat[NB_ROWS], trips); }
    }
  }
  if (fill && active) sw.end();
  if (fill && bp.emit_rows && !bp.planes)
  {
    const uint32_t pad = self << 3, pw = pad | (pad << 16);
    while (r & 3u) { buf_lo = __funnelshift_r(buf_lo, buf_hi, 16); buf_hi = __byte_perm(buf_hi, pad, 0x5432); r++; if ((r & 3u) == 0u) *colp = make_uint2(buf_lo, buf_hi); }
    for (uint32_t k = my_trips; k < trips; k++) col[(size_t)k * 32u] = make_uint2(pw, pw);
    if (lane == 0) { gt[grp] = make_uint2(row0, trips); atomicAdd(&s_stat[NB_ROWS], trips); }
  }
  // ---- per cell bookkeeping, by the lane of the cell's last particle
  if (active && pa == nA - 1u)
  {
    const uint32_t sz = 2u * (nA + 1u) + off + len;
    const uint32_t szp = (sz + 7u) & ~7u;
    const bool cfits = szp <= (uint32_t)bp.slot_words;
    out.cell_stream[ca] = cfits ? base : nullptr;
    out.stream_size[ca] = sz; out.cell_stream_bytes[ca] = sz * 2u; out.stream_off[ca] = slot_off;
    if (cfits && fill) for (uint32_t p = sz; p < szp; p++) base[p] = 0;       // deterministic padding
    if (!cfits) s_ovf = 1u;
    atomicMax(&s_stat[NB_MAX_CELL], nA); atomicMax(&s_stat[NB_MAX_STREAM], szp); atomicMax(&s_stat[NB_SLOT_WORDS], szp);
    atomicAdd(&s_tot[1], (unsigned long long)szp);
    if (cl_cell_is_inner(g, ci0, cj0, ck0)) { atomicAdd(&s_tot[2], (unsigned long long)szp); atomicAdd(&s_stat[NB_NONEMPTY], 1u); }
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    if (s_ovf) atomicOr(&out.counters[NB_OVERFLOW], 2u);
    atomicMax(&out.counters[NB_SLOT_WORDS], s_stat[NB_SLOT_WORDS]);
    atomicMax(&out.counters[NB_MAX_NBH], s_stat[NB_MAX_NBH]); atomicAdd(&out.counters[NB_NONEMPTY], s_stat[NB_NONEMPTY]);
    atomicMax(&out.counters[NB_MAX_CELL], s_stat[NB_MAX_CELL]); atomicMax(&out.counters[NB_MAX_STREAM], s_stat[NB_MAX_STREAM]);
    atomicMax(&out.counters[NB_TRIPS], s_stat[NB_TRIPS]); atomicAdd(&out.counters[NB_ROWS], s_stat[NB_ROWS]);
    atomicMax(&out.counters[NB_SURV], s_stat[NB_SURV]);
    if (s_stat[NB_AMBIGUOUS]) atomicAdd(&out.counters[NB_AMBIGUOUS], s_stat[NB_AMBIGUOUS]);
    for (int q = 0; q < 3; q++) if (s_tot[q]) atomicAdd(&out.totals[q], s_tot[q]);
  }
}

} // namespace xnb
