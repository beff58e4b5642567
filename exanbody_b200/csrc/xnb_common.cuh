// xnb_common.cuh -- shared declarations of the B200-native exaNBody hot path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace xnb {

// Local grid + domain description handed to kernels by value.
// Mirrors the slice of Grid / Domain the hot path reads (reference: src/core/include/exanb/core/grid.h:83-134,
// src/core/include/exanb/core/domain.h:36-129).
struct GridP
{
  double org[3];      // Grid::origin == Domain::bounds().bmin
  double dmin[3];     // Domain bounds
  double dmax[3];
  double cs;          // cell size
  int dims[3];        // local grid dims, ghost layers included
  int off[3];         // Grid::offset (domain location of local cell 0,0,0) = block start - gl
  int ddims[3];       // domain grid dims
  int bstart[3];      // inner block [bstart,bend) in domain cells
  int bend[3];
  int periodic[3];
  int gl;             // ghost layers
  int n_cells;
};

// flat SoA particle arrays (device pointers)
struct ParticlesP
{
  double *rx, *ry, *rz, *vx, *vy, *vz, *fx, *fy, *fz;
  unsigned long long* id;
  uint8_t* type;
};

// GhostBoundaryModifier flags, same bit layout as the reference (src/mpi/include/exanb/mpi/ghosts_comm_scheme.h:46-81)
enum : uint32_t
{
  GB_SHIFT_X = 1u << 0, GB_MIRROR_X = 1u << 1, GB_SIDE_X = 1u << 2,
  GB_SHIFT_Y = 1u << 3, GB_MIRROR_Y = 1u << 4, GB_SIDE_Y = 1u << 5,
  GB_SHIFT_Z = 1u << 6, GB_MIRROR_Z = 1u << 7, GB_SIDE_Z = 1u << 8
};

// error bits written by kernels into the ctx error word
enum : uint32_t
{
  DERR_LOST_PARTICLE = 1u << 0,   // left a non periodic domain
  DERR_CELL_OVERFLOW = 1u << 1,   // > 65535 particles in one cell (u16 stream index)
  DERR_GROUP_OVERFLOW = 1u << 2,  // u16 counter overflow in a stream
  DERR_FAR_MIGRATION = 1u << 3,   // particle jumped to a rank that is not a ghost partner
  DERR_SORT_CAPACITY = 1u << 4,   // cell too large for the in-cell sort
  DERR_ID_RANGE = 1u << 5,        // particle id >= 2^52
  DERR_TILE_CAPACITY = 1u << 6,   // a tile did not fit its shared-memory staging capacity (host falls back to the untiled kernels)
  DERR_PEER_TIMEOUT = 1u << 7     // a partner's halo slab / displacement count never arrived in my mailbox (xnb_peer_halo.cuh)
};

#define XNB_DEVINL __device__ __forceinline__

// squared norm exactly as the oracle defines it: ((x*x) + (y*y)) + (z*z), every operation rounded, no FMA contraction.
XNB_DEVINL double norm2_exact(double x, double y, double z)
{
  return __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
}

// (k,j,i) lexicographic index, reference src/core/include/exanb/core/grid_algorithm.h:84-92
XNB_DEVINL int ijk_to_index(const int* d, int i, int j, int k) { return (k * d[1] + j) * d[0] + i; }

// reference src/core/include/exanb/core/backup_r.h:31-34
XNB_DEVINL double restore_u32_double(uint32_t x, double o, double r)
{
  return __dadd_rn(o, __ddiv_rn(__dmul_rn((double)x, r), 4294967296.0));
}

// reference src/mpi/include/exanb/mpi/ghosts_comm_scheme.h:76-81 (periodic shift only; mirrors are out of scope)
XNB_DEVINL double coord_shift(double x, double rmin, double rmax, uint32_t f3)
{
  if (f3 & 1u) return __dadd_rn(x, __dmul_rn((f3 & 4u) ? 1.0 : -1.0, __dadd_rn(rmax, -rmin)));
  return x;
}

} // namespace xnb
