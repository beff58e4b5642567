// xnb_host_decomp.cpp -- static spatial decomposition and ghost item lists (host only, no CUDA).
//   reference: src/core/lib/simple_block_rcb.cpp:27-59, src/grid_cell_particles/init_rcb_grid.cpp:65-77,
//              src/mpi/update_ghosts_comm_scheme.cpp:130-196,429-443, src/mpi/include/exanb/mpi/ghosts_comm_scheme.h:46-81
#include "../../include/xnb_hotpath.h"
#include "xnb_host_decomp.hpp"
#include <algorithm>
#include <vector>
#include <cmath>

namespace xnb {

// GhostBoundaryModifier flags, same bit layout as the reference (ghosts_comm_scheme.h:46-81); mirrored in xnb_common.cuh
enum : uint32_t { H_SHIFT_X = 1u << 0, H_SIDE_X = 1u << 2, H_SHIFT_Y = 1u << 3, H_SIDE_Y = 1u << 5, H_SHIFT_Z = 1u << 6, H_SIDE_Z = 1u << 8 };

Block simple_block_rcb(Block b, size_t n_parts, size_t part)
{
  while (n_parts > 1)
  {
    const size_t pivot = n_parts / 2;
    const bool side = part >= pivot;
    const int64_t d[3] = {b.e[0] - b.s[0], b.e[1] - b.s[1], b.e[2] - b.s[2]};
    int ax = 2;
    if (d[0] >= d[1] && d[0] >= d[2]) ax = 0; else if (d[1] >= d[0] && d[1] >= d[2]) ax = 1;
    if (side) b.s[ax] = b.s[ax] + d[ax] / 2; else b.e[ax] = b.s[ax] + d[ax] / 2;
    if (side) { part -= pivot; n_parts -= pivot; } else n_parts = pivot;
  }
  return b;
}

// ---- cost-weighted recursive bisection (load_balance_rcb.cpp, non-Zoltan path) ---------------------------------------------
namespace {
struct Split1D { size_t position = 0; double worst_balance = 0.0; long surf = 0; int axis = 0; bool valid = false; };

// :510-545  best cut of a 1-D cost profile between `left` and `right` ranks: minimise max(cost_left/left, cost_right/right)
Split1D best_split(const std::vector<double>& v, size_t left, size_t right)
{
  Split1D r;
  const size_t n = v.size();
  double sum_right = 0., sum_left = 0.;
  for (double x : v) sum_right += x;
  if (n == 0) { r.position = 0; r.worst_balance = sum_right; return r; }
  if (right == 0) { r.position = n - 1; r.worst_balance = sum_right / (double)n; return r; }
  if (left == 0) { r.position = 0; r.worst_balance = sum_right / (double)n; return r; }
  double best = sum_right / (double)right; size_t best_p = 0;
  for (size_t p = 1; p < n; p++)
  {
    sum_left += v[p - 1]; sum_right -= v[p - 1];
    const double wb = std::max(sum_left / (double)left, sum_right / (double)right);
    if (wb < best) { best = wb; best_p = p; }
  }
  r.position = best_p; r.worst_balance = best; r.surf = -1; r.axis = -1;
  return r;
}
} // namespace

Block load_balance_rcb(const int64_t ddims[3], const double* costs, size_t n_parts, size_t part)
{
  Block b{{0, 0, 0}, {ddims[0], ddims[1], ddims[2]}};
  size_t group = n_parts, rank_in_group = part;
  auto empty = [](const Block& q) { return q.e[0] <= q.s[0] || q.e[1] <= q.s[1] || q.e[2] <= q.s[2]; };
  while (group > 1 && !empty(b))                                       // :285
  {
    const int64_t d[3] = {b.e[0] - b.s[0], b.e[1] - b.s[1], b.e[2] - b.s[2]};
    std::vector<double> prof[3] = {std::vector<double>((size_t)d[0], 0.), std::vector<double>((size_t)d[1], 0.), std::vector<double>((size_t)d[2], 0.)};
    for (int64_t k = 0; k < d[2]; k++) for (int64_t j = 0; j < d[1]; j++) for (int64_t i = 0; i < d[0]; i++)      // :293-306 cost profiles of the block
    {
      const double c = costs[((b.s[2] + k) * ddims[1] + (b.s[1] + j)) * ddims[0] + (b.s[0] + i)];
      prof[0][(size_t)i] += c; prof[1][(size_t)j] += c; prof[2][(size_t)k] += c;
    }
    const size_t left = group / 2, right = group - left;               // :313-326
    const bool side = rank_in_group >= left;
    Split1D sp[3];
    for (int a = 0; a < 3; a++)                                          // :332-349
    {
      sp[a] = best_split(prof[a], left, right);
      sp[a].surf = (long)(d[(a + 1) % 3] * d[(a + 2) % 3]);
      sp[a].valid = d[a] >= 2 && sp[a].position > 0 && (int64_t)sp[a].position < d[a];
      sp[a].axis = a;
    }
    // :351-362  the better balanced cut wins; within 5 % of each other the smaller cut surface wins (same comparator, same std::sort)
    auto better = [](const Split1D& x, const Split1D& y) -> bool
    {
      if (!x.valid) return false;
      if (!y.valid) return true;
      const double mx = std::max(x.worst_balance, y.worst_balance);
      if (mx == 0.0) return true;
      const double mn = std::min(x.worst_balance, y.worst_balance);
      if (mn / mx > 0.95) return x.surf < y.surf;
      return x.worst_balance < y.worst_balance;
    };
    std::sort(sp, sp + 3, better);
    Block lb = b, rb = b;
    if (sp[0].valid) { lb.e[sp[0].axis] = b.s[sp[0].axis] + (int64_t)sp[0].position; rb.s[sp[0].axis] = lb.e[sp[0].axis]; }      // :368-388
    else
    {
      // :389-412  no cut balances anything (e.g. zero costs): halve the longest axis (ties: i, then j)
      int a = 2;
      if (d[0] >= d[1] && d[0] >= d[2]) a = 0; else if (d[1] >= d[0] && d[1] >= d[2]) a = 1;
      lb.e[a] = b.s[a] + d[a] / 2; rb.s[a] = lb.e[a];
    }
    b = side ? rb : lb;                                                  // :428-440
    if (side) { rank_in_group -= left; group = right; } else group = left;
  }
  return b;
}

void enumerate_sends(const std::vector<Block>& blocks, const int64_t ddims[3], const int periodic[3], int from, int to, int gl,
                     std::vector<HostItem>& out)
{
  const Block& bf = blocks[(size_t)from];
  const Block& bt = blocks[(size_t)to];
  int64_t gs[3], ge[3], tdims[3], fdims[3], foff[3];
  for (int d = 0; d < 3; d++) { gs[d] = bt.s[d] - gl; ge[d] = bt.e[d] + gl; tdims[d] = ge[d] - gs[d]; fdims[d] = bf.e[d] - bf.s[d] + 2 * gl; foff[d] = bf.s[d] - gl; }
  const int lo[3] = {periodic[0] ? -1 : 0, periodic[1] ? -1 : 0, periodic[2] ? -1 : 0};
  const int hi[3] = {periodic[0] ? 1 : 0, periodic[1] ? 1 : 0, periodic[2] ? 1 : 0};
  for (int sk = lo[2]; sk <= hi[2]; sk++) for (int sj = lo[1]; sj <= hi[1]; sj++) for (int si = lo[0]; si <= hi[0]; si++)
  {
    if (si == 0 && sj == 0 && sk == 0 && from == to) continue;
    const int sh[3] = {si, sj, sk};
    uint32_t flags = 0;
    if (si == -1) flags |= H_SHIFT_X; if (si == 1) flags |= H_SHIFT_X | H_SIDE_X;
    if (sj == -1) flags |= H_SHIFT_Y; if (sj == 1) flags |= H_SHIFT_Y | H_SIDE_Y;
    if (sk == -1) flags |= H_SHIFT_Z; if (sk == 1) flags |= H_SHIFT_Z | H_SIDE_Z;
    int64_t a[3], b[3];
    bool empty = false;
    for (int d = 0; d < 3; d++)
    {
      const int64_t s = (int64_t)sh[d] * ddims[d];
      a[d] = std::max(bf.s[d], gs[d] - s); b[d] = std::min(bf.e[d], ge[d] - s);
      if (a[d] >= b[d]) empty = true;
    }
    if (empty) continue;
    for (int64_t k = a[2]; k < b[2]; k++) for (int64_t j = a[1]; j < b[1]; j++) for (int64_t i = a[0]; i < b[0]; i++)
    {
      const int64_t dl[3] = {i, j, k};
      int64_t t[3];
      bool inside_inner = true;
      for (int d = 0; d < 3; d++) { t[d] = dl[d] + (int64_t)sh[d] * ddims[d]; if (t[d] < bt.s[d] || t[d] >= bt.e[d]) inside_inner = false; }
      if (inside_inner) continue;   // can only happen for degenerate tiny domains
      HostItem it;
      it.src_cell = (uint32_t)(((k - foff[2]) * fdims[1] + (j - foff[1])) * fdims[0] + (i - foff[0]));
      it.dst_cell = (uint32_t)(((t[2] - gs[2]) * tdims[1] + (t[1] - gs[1])) * tdims[0] + (t[0] - gs[0]));
      it.flags = flags;
      out.push_back(it);
    }
  }
}

} // namespace xnb

extern "C" int xnb_host_simple_cost_model(int64_t n_cells, const uint32_t* cell_count, double cell_size, const double coefs[4], double* cell_costs)
{
  // simple_cost_model.h:67-146: p = N / cell volume ; cost = coefs[0] p^3 + coefs[1] p^2 + coefs[2] p + coefs[3]   (defaults {0, 0, 1, 0})
  if (n_cells < 0 || (n_cells > 0 && (!cell_count || !cell_costs)) || !coefs || !(cell_size > 0.0)) return XNB_ERR_INVALID;
  const double d3 = coefs[0], d2 = coefs[1], d1 = coefs[2], cc = coefs[3];
  const double cell_volume = cell_size * cell_size * cell_size;
  for (int64_t c = 0; c < n_cells; c++)
  {
    const double pvol = (double)cell_count[c] / cell_volume;
    cell_costs[c] = pvol * d1 + pvol * pvol * d2 + pvol * pvol * pvol * d3 + cc;
  }
  return XNB_OK;
}

extern "C" int xnb_host_load_balance_rcb(const int64_t grid_dims[3], const double* cell_costs, int nranks, int rank, int64_t start[3], int64_t end[3],
                                         double* block_cost)
{
  if (!grid_dims || !cell_costs || !start || !end || nranks < 1 || rank < 0 || rank >= nranks || grid_dims[0] < 1 || grid_dims[1] < 1 || grid_dims[2] < 1) return XNB_ERR_INVALID;
  // costs must be finite and non-negative (a NaN would poison every comparison of the bisection)
  const int64_t nd = grid_dims[0] * grid_dims[1] * grid_dims[2];
  for (int64_t q = 0; q < nd; q++) if (!(cell_costs[q] >= 0.0) || !std::isfinite(cell_costs[q])) return XNB_ERR_INVALID;
  const xnb::Block b = xnb::load_balance_rcb(grid_dims, cell_costs, (size_t)nranks, (size_t)rank);
  // "Assigned grid block is empty" is a fatal error of the reference (load_balance_rcb.cpp:410-411,457): more ranks than cells along the cuts
  for (int d = 0; d < 3; d++) if (b.e[d] <= b.s[d]) return XNB_ERR_INVALID;
  double cost = 0.0;
  for (int64_t k = b.s[2]; k < b.e[2]; k++) for (int64_t j = b.s[1]; j < b.e[1]; j++) for (int64_t i = b.s[0]; i < b.e[0]; i++) cost += cell_costs[(k * grid_dims[1] + j) * grid_dims[0] + i];
  for (int d = 0; d < 3; d++) { start[d] = b.s[d]; end[d] = b.e[d]; }
  if (block_cost) *block_cost = cost;        // :443-452 feeds lb_inbalance = (max - avg) / avg over the ranks
  return XNB_OK;
}

extern "C" int xnb_host_rcb_block(const int64_t grid_dims[3], int nranks, int rank, int64_t start[3], int64_t end[3])
{
  if (!grid_dims || !start || !end || nranks < 1 || rank < 0 || rank >= nranks) return XNB_ERR_INVALID;
  const xnb::Block whole{{0, 0, 0}, {grid_dims[0], grid_dims[1], grid_dims[2]}};
  const xnb::Block b = xnb::simple_block_rcb(whole, (size_t)nranks, (size_t)rank);
  for (int d = 0; d < 3; d++) { start[d] = b.s[d]; end[d] = b.e[d]; }
  return XNB_OK;
}

extern "C" int64_t xnb_host_ghost_items(const int64_t grid_dims[3], const int32_t periodic[3], int ghost_layers, int nranks, int from, int to,
                                        int64_t capacity, uint32_t* src_cell, uint32_t* dst_cell, uint32_t* flags)
{
  if (!grid_dims || !periodic || nranks < 1 || from < 0 || from >= nranks || to < 0 || to >= nranks || ghost_layers < 0) return -1;
  std::vector<xnb::Block> blocks((size_t)nranks);
  const xnb::Block whole{{0, 0, 0}, {grid_dims[0], grid_dims[1], grid_dims[2]}};
  for (int r = 0; r < nranks; r++) blocks[(size_t)r] = xnb::simple_block_rcb(whole, (size_t)nranks, (size_t)r);
  const int per[3] = {periodic[0] ? 1 : 0, periodic[1] ? 1 : 0, periodic[2] ? 1 : 0};
  std::vector<xnb::HostItem> items;
  xnb::enumerate_sends(blocks, grid_dims, per, from, to, ghost_layers, items);
  if ((int64_t)items.size() <= capacity)
    for (size_t q = 0; q < items.size(); q++)
    {
      if (src_cell) src_cell[q] = items[q].src_cell;
      if (dst_cell) dst_cell[q] = items[q].dst_cell;
      if (flags) flags[q] = items[q].flags;
    }
  return (int64_t)items.size();
}

// ---------------------------------------------------------------------------------------------------------------------
// op `amr_grid_pairs`: max_distance_sub_cell_pairs (src/amr/lib/amr_grid_algorithm.cpp:102-218) -> AmrSubCellPairCache
// (amr/include/exanb/amr/amr_grid_algorithm.h:439-453).  For every resolution pair (res_a <= res_b <= max_res, in the order
// res_b outer / res_a inner = unique_pair_id) and every neighbour cell offset (k, j, i in [0, layers], layers = ceil(max_dist / cell_size)):
// the list of (sub-cell a, sub-cell b) pairs, each coded (k << 10) | (j << 5) | i, whose boxes are at most max_dist apart.
// list_offsets: n_lists + 1 entries (u16 words); pairs: a, b interleaved.  Returns the total number of u16 words (pairs may be NULL
// to size the buffer), or -1 on bad arguments.  n_lists = max_res (max_res + 1) / 2 * (layers + 1)^3.
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int64_t xnb_host_amr_sub_cell_pairs(int max_res, double cell_size, double max_dist, uint64_t* list_offsets, uint16_t* pairs)
{
  if (max_res < 1 || max_res >= 32 || !(cell_size > 0.0) || !(max_dist >= 0.0)) return -1;
  const double max_dist2 = max_dist * max_dist;
  const int layers = (int)std::ceil(max_dist / cell_size);
  int64_t total = 0; size_t list = 0;
  auto min_d2_1d = [](double alo, double ahi, double blo, double bhi) { const double d = std::max(0.0, std::max(blo - ahi, alo - bhi)); return d * d; };
  for (int res_b = 1; res_b <= max_res; res_b++) for (int res_a = 1; res_a <= res_b; res_a++)
  {
    const double sa = cell_size / res_a, sb = cell_size / res_b;
    for (int ck = 0; ck <= layers; ck++) for (int cj = 0; cj <= layers; cj++) for (int ci = 0; ci <= layers; ci++)
    {
      if (list_offsets) list_offsets[list] = (uint64_t)total;
      list++;
      for (int ka = 0; ka < res_a; ka++) for (int ja = 0; ja < res_a; ja++) for (int ia = 0; ia < res_a; ia++)
        for (int kb = 0; kb < res_b; kb++) for (int jb = 0; jb < res_b; jb++) for (int ib = 0; ib < res_b; ib++)
        {
          // min_distance2_between of the two boxes (core/geometry.h:125-156): per axis gap, squared and summed x, y, z
          const double dx = min_d2_1d(ia * sa, (ia + 1) * sa, ci * cell_size + ib * sb, ci * cell_size + (ib + 1) * sb);
          const double dy = min_d2_1d(ja * sa, (ja + 1) * sa, cj * cell_size + jb * sb, cj * cell_size + (jb + 1) * sb);
          const double dz = min_d2_1d(ka * sa, (ka + 1) * sa, ck * cell_size + kb * sb, ck * cell_size + (kb + 1) * sb);
          if (dx + dy + dz <= max_dist2)
          {
            if (pairs) { pairs[total] = (uint16_t)((ka << 10) | (ja << 5) | ia); pairs[total + 1] = (uint16_t)((kb << 10) | (jb << 5) | ib); }
            total += 2;
          }
        }
    }
  }
  if (list_offsets) list_offsets[list] = (uint64_t)total;
  return total;
}
