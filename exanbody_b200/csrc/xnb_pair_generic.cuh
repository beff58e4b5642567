// xnb_pair_generic.cuh -- the general pair sweep: ANY functor of the reference's concept, both call forms, per-neighbour fields.
//
// reference: compute/include/exanb/compute/compute_cell_particle_pairs_impl_default.h:87-239 (the default sweep: stream walk, d2 test,
//            buffer-less call :199-204, buffer call :213-222), compute_pair_buffer.h:39-68,150-243 (ComputePairBuffer2 and its append
//            function), compute_pair_traits.h:24-74 (which form a functor offers), core/grid.h:70-71,678 (cells[c][field][p]),
//            contribs/pi/gravitational_force.cu:48-132,144-150,210-217 (the second functor: needs field::type of the NEIGHBOUR,
//            SimpleNbhComputeBuffer<FieldSet<type>>).
//
// k_lj_sweep_cl is specialised for functors that read nothing but dr and d2 (Lennard-Jones).  This kernel is the other end of the
// trade: one thread per particle walks that particle's list in the reference-format stream, positions and fields come from the
// flat particle arrays through a CellsView (the device counterpart of the reference's `cells` pointer), and the functor is called
// exactly as the reference calls it:
//   * buffer-less form (ComputePairTraits<F>::BufferLessCompatible):  func(dr, d2, central fields..., cells, cell_b, p_b, weight)
//   * buffer form      (ComputePairTraits<F>::ComputeBufferCompatible): neighbours inside the cut are appended to a ComputePairBuffer2
//     (dr, d2, (cell_b, p_b), neighbour fields) and the functor is called once per particle: func(n, buf, central fields..., cells)
// Central fields here: type, fx, fy, fz (CentralParticleFieldSet of gravitational_force.cu:172); weight is
// ComputePairNullWeightIterator's 1.0; xform = identity; ComputePairOptionalLocks<false> (a thread writes its own particle only).
#pragma once
#include "xnb_pair_functor.cuh"

namespace xnb {

// cells[cell][field][p] of the reference (core/grid.h:70-71): per-cell views of the flat SoA arrays
struct CellsView
{
  const uint32_t* cell_start; const uint32_t* cell_count;
  const double *rx, *ry, *rz, *vx, *vy, *vz;
  const unsigned long long* id; const uint8_t* type;
  XNB_DEVINL uint32_t flat(size_t cell, size_t p) const { return cell_start[cell] + (uint32_t)p; }
  XNB_DEVINL int type_of(size_t cell, size_t p) const { return (int)type[flat(cell, p)]; }
  XNB_DEVINL unsigned long long id_of(size_t cell, size_t p) const { return id[flat(cell, p)]; }
};

// ComputePairBuffer2 (compute_pair_buffer.h:150-243) with UseNeighbors = true, NbhFieldSet = FieldSet<type>, no user weights.
// One per thread, in local memory (the reference keeps it on the CPU stack, or in shared memory for its block-cooperative variant).
template <int MAXN>
struct ComputePairBuffer2
{
  static constexpr int MaxNeighbors = MAXN;
  double drx[MAXN], dry[MAXN], drz[MAXN], d2[MAXN];
  uint32_t nbh_cell[MAXN]; uint16_t nbh_part[MAXN];       // ComputePairBuffer2Nbh<true>
  uint8_t nbh_type[MAXN];                                 // ComputePairBuffer2NbhFields<FieldSet<type>>: nbh_pt[i][field::type]
  unsigned long long cell; uint32_t part; int32_t count; uint32_t ta, tb;
  struct Weights { XNB_DEVINL double get(int) const { return 1.0; } XNB_DEVINL double operator[](int) const { return 1.0; } } nbh_data;
  // DefaultComputePairBufferAppendFunc (compute_pair_buffer.h:39-68)
  XNB_DEVINL void process_neighbor(double dx, double dy, double dz, double dd, const CellsView& cells, size_t cell_b, size_t p_b)
  {
    const int w = count++;
    d2[w] = dd; drx[w] = dx; dry[w] = dy; drz[w] = dz;
    nbh_cell[w] = (uint32_t)cell_b; nbh_part[w] = (uint16_t)p_b; nbh_type[w] = (uint8_t)cells.type_of(cell_b, p_b);
  }
};
constexpr int XNB_MAX_PARTICLE_NEIGHBORS = 512;

// ------------------------------------------------------------------------------------------------------------------
// GravitationalForceFunctor (contribs/pi/gravitational_force.cu:48-132): e = -G ma mb / r, de = G ma mb / r^2,
// f += de * w / r * dr; masses through the particle TYPE of the central particle and of the neighbour.
// Both call forms; the buffer form is the one the reference keeps under `#if 0` (:87-132) with mass_b restored.
// ------------------------------------------------------------------------------------------------------------------
struct GravitationalForceFunctor
{
  double G; const double* type_mass;
  XNB_DEVINL void compute_energy(double mass_a, double mass_b, double r, double& e, double& de) const
  {
    const double inv_r = 1.0 / r;
    e = -G * mass_a * mass_b * inv_r;
    de = G * mass_a * mass_b * inv_r * inv_r;
  }
  // buffer-less (:64-85)
  XNB_DEVINL void operator()(double3 dr, double d2, int type_a, double& fx, double& fy, double& fz, const CellsView& cells, size_t neighbor_cell,
                             size_t neighbor_particle, double interaction_weight) const
  {
    const double mass_a = type_mass[type_a];
    const int type_b = cells.type_of(neighbor_cell, neighbor_particle);
    const double mass_b = type_mass[type_b];
    const double r = sqrt(d2);
    double pair_e = 0.0, pair_de = 0.0;
    compute_energy(mass_a, mass_b, r, pair_e, pair_de);
    pair_de *= interaction_weight / r;
    fx += pair_de * dr.x; fy += pair_de * dr.y; fz += pair_de * dr.z;
  }
  // buffer form (:96-131)
  template <class BufT>
  XNB_DEVINL void operator()(int n, const BufT& buffer, int type_a, double& fx, double& fy, double& fz, const CellsView&) const
  {
    double _fx = 0., _fy = 0., _fz = 0.;
    const double mass_a = type_mass[type_a];
    for (int i = 0; i < n; i++)
    {
      const double r = sqrt(buffer.d2[i]);
      const double mass_b = type_mass[buffer.nbh_type[i]];
      double pair_e = 0.0, pair_de = 0.0;
      compute_energy(mass_a, mass_b, r, pair_e, pair_de);
      pair_de *= buffer.nbh_data.get(i) / r;
      _fx += pair_de * buffer.drx[i]; _fy += pair_de * buffer.dry[i]; _fz += pair_de * buffer.drz[i];
    }
    fx += _fx; fy += _fy; fz += _fz;
  }
};
template <> struct ComputePairTraits<GravitationalForceFunctor>
{
  static constexpr bool BufferLessCompatible = true;
  static constexpr bool ComputeBufferCompatible = true;
  static constexpr bool CudaCompatible = true;
  static constexpr bool RequiresNbhOptionalData = false;
  static constexpr bool HasParticleContext = false;
  static constexpr bool Batch4 = false;
};

// ------------------------------------------------------------------------------------------------------------------
// k_pair_sweep_generic<F, BUFFER>: compute_cell_particle_pairs over the inner particles, full (non-symmetric) lists.
// err |= DERR_GROUP_OVERFLOW when a particle has more neighbours inside the cut than the buffer holds (the reference aborts in
// debug builds and overflows silently otherwise: compute_pair_buffer.h:199-208).
// ------------------------------------------------------------------------------------------------------------------
template <class F, bool BUFFER>
__global__ void __launch_bounds__(128)
k_pair_sweep_generic(GridP g, int n_inner, F func, double rcut2, CellsView cells, double* __restrict__ fx, double* __restrict__ fy, double* __restrict__ fz,
                     const uint32_t* __restrict__ atom_cell, const uint16_t* const* __restrict__ cell_stream, uint32_t* __restrict__ err)
{
  static_assert(ComputePairTraits<F>::CudaCompatible, "functor must be device callable");
  static_assert(BUFFER ? ComputePairTraits<F>::ComputeBufferCompatible : ComputePairTraits<F>::BufferLessCompatible, "functor does not offer this call form");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_inner) return;
  const uint32_t ca = atom_cell[i];
  const uint32_t na = cells.cell_count[ca], pa = (uint32_t)i - cells.cell_start[ca];
  const uint16_t* cs = cell_stream[ca];
  if (!cs) return;
  const uint32_t off0 = reinterpret_cast<const uint32_t*>(cs)[pa];
  const uint16_t* lst = cs + 2u * (na + 1u) + off0;          // first word behind the group counter (offsets are biased by the number of tables = 1)
  uint32_t ngrp = (uint32_t)lst[-1];
  const double xa = cells.rx[i], ya = cells.ry[i], za = cells.rz[i];
  const int type_a = (int)cells.type[i];
  double ax = fx[i], ay = fy[i], az = fz[i];                 // the functor adds to the central particle's force fields
  const int dxy = g.dims[0] * g.dims[1];
  ComputePairBuffer2<BUFFER ? XNB_MAX_PARTICLE_NEIGHBORS : 1> tab;
  if (BUFFER) { tab.cell = ca; tab.part = pa; tab.count = 0; tab.ta = (uint32_t)type_a; tab.tb = 0; }
  bool overflow = false;
  for (; ngrp > 0u; ngrp--)
  {
    const uint32_t code = *lst++; uint32_t n = *lst++;
    // chunk_neighbors.h:137-162: 5-bit fields of the code minus 16 = cell of b relative to the cell of a
    const int cb = (int)ca + ((int)(code >> 10) - 16) * dxy + ((int)((code >> 5) & 31u) - 16) * g.dims[0] + ((int)(code & 31u) - 16);
    const uint32_t sb = cells.cell_start[cb];
    for (; n > 0u; n--)
    {
      const uint32_t pb = *lst++;
      const uint32_t j = sb + pb;
      const double dx = __dadd_rn(cells.rx[j], -xa), dy = __dadd_rn(cells.ry[j], -ya), dz = __dadd_rn(cells.rz[j], -za);
      const double d2 = norm2_exact(dx, dy, dz);
      if (d2 > 0.0 && d2 <= rcut2)
      {
        if (BUFFER)
        {
          if (tab.count >= tab.MaxNeighbors) { overflow = true; continue; }
          tab.process_neighbor(dx, dy, dz, d2, cells, (size_t)cb, (size_t)pb);
        }
        else func(make_double3(dx, dy, dz), d2, type_a, ax, ay, az, cells, (size_t)cb, (size_t)pb, 1.0);
      }
    }
  }
  if (BUFFER && tab.count > 0) func((int)tab.count, tab, type_a, ax, ay, az, cells);
  if (overflow) atomicOr(err, DERR_GROUP_OVERFLOW);
  fx[i] = ax; fy[i] = ay; fz[i] = az;
}

// ------------------------------------------------------------------------------------------------------------------
// Particle-context call form (compute_pair_traits.h: HasParticleContextStart / HasParticleContext / HasParticleContextStop;
// impl_default.h:152,199-204,224): the sweep owns a per-particle context (ComputeContextNoBuffer<ExtStorage>), calls
// func(ctx, cells, cell_a, p_a, Start{}) before the particle's first neighbour, func(ctx, dr, d2, cells, cell_b, p_b, weight) per pair
// and func(ctx, cells, cell_a, p_a, Stop{}) after the last.
// ------------------------------------------------------------------------------------------------------------------
struct ComputePairParticleContextStart {};
struct ComputePairParticleContextStop {};
template <class Ext> struct ComputeContextNoBuffer { Ext ext; };        // compute_pair_buffer.h:143-148

// AverageNeighborsFunctor (src/compute/average_neighbors.cu:38-100): avg_field[a] = sum_b w(d) nbh_field[b] / sum_b w(d) over the
// neighbours within rcut, w = a0 + a1 d + a2 d^2 + a3 d^3; a third kind of functor: per-neighbour SCALAR FIELD + particle context
struct AverageNeighborsExtStorage
{
  double m_sum, m_weight_sum;
  XNB_DEVINL void reset() { m_sum = 0.0; m_weight_sum = 0.0; }
  XNB_DEVINL double avg() const { return (m_weight_sum > 0.0) ? (m_sum / m_weight_sum) : (m_sum / 1.0); }
};
struct AverageNeighborsFunctor
{
  double m_rcut_sq, a0, a1, a2, a3;
  double* m_avg_field;                 // flat array of the averaged field (cells[c][avg_field][p] = m_avg_field[cell_start[c] + p])
  const double* m_nbh_field_f64;       // the neighbours' field: one of the f64 particle arrays ...
  const unsigned long long* m_nbh_field_u64; const uint8_t* m_nbh_field_u8;      // ... or id / type (anything convertible to double, :143)
  typedef ComputeContextNoBuffer<AverageNeighborsExtStorage> Context;
  XNB_DEVINL double nbh_value(uint32_t j) const { return m_nbh_field_f64 ? m_nbh_field_f64[j] : m_nbh_field_u64 ? (double)m_nbh_field_u64[j] : (double)m_nbh_field_u8[j]; }
  XNB_DEVINL void operator()(Context& ctx, const CellsView&, size_t, size_t, ComputePairParticleContextStart) const { ctx.ext.reset(); }
  XNB_DEVINL void operator()(Context& ctx, const CellsView& cells, size_t cell_a, size_t p_a, ComputePairParticleContextStop) const { m_avg_field[cells.flat(cell_a, p_a)] = ctx.ext.avg(); }
  XNB_DEVINL void operator()(Context& ctx, double3, double d2, const CellsView& cells, size_t cell_b, size_t p_b, double) const
  {
    if (d2 <= m_rcut_sq)
    {
      double w = a0 + a2 * d2;
      if (a1 != 0.0 || a3 != 0.0) { const double d = sqrt(d2); w += a1 * d + a3 * d2 * d; }
      ctx.ext.m_sum += w * nbh_value(cells.flat(cell_b, p_b));
      ctx.ext.m_weight_sum += w;
    }
  }
};

// k_pair_sweep_context<F>: compute_cell_particle_pairs for functors with a particle context (no central fields, no buffer)
template <class F>
__global__ void __launch_bounds__(128)
k_pair_sweep_context(GridP g, int n_inner, F func, double rcut2, CellsView cells, const uint32_t* __restrict__ atom_cell,
                     const uint16_t* const* __restrict__ cell_stream)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_inner) return;
  const uint32_t ca = atom_cell[i];
  const uint32_t na = cells.cell_count[ca], pa = (uint32_t)i - cells.cell_start[ca];
  typename F::Context ctx;
  func(ctx, cells, (size_t)ca, (size_t)pa, ComputePairParticleContextStart{});
  const uint16_t* cs = cell_stream[ca];
  if (cs)
  {
    const uint32_t off0 = reinterpret_cast<const uint32_t*>(cs)[pa];
    const uint16_t* lst = cs + 2u * (na + 1u) + off0;
    uint32_t ngrp = (uint32_t)lst[-1];
    const double xa = cells.rx[i], ya = cells.ry[i], za = cells.rz[i];
    const int dxy = g.dims[0] * g.dims[1];
    for (; ngrp > 0u; ngrp--)
    {
      const uint32_t code = *lst++; uint32_t n = *lst++;
      const int cb = (int)ca + ((int)(code >> 10) - 16) * dxy + ((int)((code >> 5) & 31u) - 16) * g.dims[0] + ((int)(code & 31u) - 16);
      const uint32_t sb = cells.cell_start[cb];
      for (; n > 0u; n--)
      {
        const uint32_t pb = *lst++;
        const uint32_t j = sb + pb;
        const double dx = __dadd_rn(cells.rx[j], -xa), dy = __dadd_rn(cells.ry[j], -ya), dz = __dadd_rn(cells.rz[j], -za);
        const double d2 = norm2_exact(dx, dy, dz);
        if (d2 > 0.0 && d2 <= rcut2) func(ctx, make_double3(dx, dy, dz), d2, cells, (size_t)cb, (size_t)pb, 1.0);
      }
    }
  }
  func(ctx, cells, (size_t)ca, (size_t)pa, ComputePairParticleContextStop{});
}

} // namespace xnb
