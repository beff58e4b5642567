// xnb_host_decomp.hpp -- host-side spatial decomposition logic shared by the CUDA orchestration (xnb_hotpath.cu) and the
// host-only C entry points (xnb_host_decomp.cpp) that the CPU tests exercise with two gloo ranks.
#pragma once
#include <cstdint>
#include <vector>

namespace xnb {

struct Block { int64_t s[3], e[3]; };                         // GridBlock: [start, end) in domain cells
struct HostItem { uint32_t src_cell, dst_cell, flags; };      // one ghost send item (sender cell, receiver ghost cell, boundary flags)

// reference src/core/lib/simple_block_rcb.cpp:27-59 : recursive bisection, longest axis first (ties: i, then j)
Block simple_block_rcb(Block b, size_t n_parts, size_t part);

// reference src/mpi/load_balance_rcb.cpp:228-452,510-545 (the path without Zoltan): cost-weighted recursive bisection of the domain
// cell grid.  `costs` = the all-reduced cost of every domain cell, index (k*dj + j)*di + i.  Returns the block of `part`.
Block load_balance_rcb(const int64_t ddims[3], const double* costs, size_t n_parts, size_t part);

// ghost send items of rank `from` towards rank `to` in the reference's order
// (update_ghosts_comm_scheme.cpp:168-196 shift loops k,j,i ; :429-443 cell loop k,j,i and ghost-shell membership)
void enumerate_sends(const std::vector<Block>& blocks, const int64_t ddims[3], const int periodic[3], int from, int to, int gl,
                     std::vector<HostItem>& out);

} // namespace xnb
