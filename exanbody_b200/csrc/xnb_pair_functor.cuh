// xnb_pair_functor.cuh -- the pair-functor concept of the reference, restated for the sm_100a pair sweeps.
//
// reference: compute/include/exanb/compute/compute_pair_traits.h:24-74 (ComputePairTraits / compute_pair_traits::*_v),
//            compute_cell_particle_pairs_impl_default.h:183-204 (the buffer-less call: func(dr, d2, fields..., cells, cell_b, p_b, weight)),
//            compute_pair_optional_args.h:37-209 (locks, weights, xform), compute_pair_buffer.h:150-243 (ComputePairBuffer2),
//            contribs/md/lennard_jones/lennard_jones.cu:40-56,59-126,133-141 (parameters, functor, traits of the LJ operator).
//
// The sweeps (k_lj_sweep_cl, k_lj_sweep) are templates over a functor type F.  What they require of F:
//
//   double rcut2() const                                      accept a listed candidate iff 0 < d2 <= rcut2()  (impl_default.h:186)
//   void operator()(double3 dr, double d2, double& fx, double& fy, double& fz, PairNbh b, double weight) const
//                                                             the BUFFER-LESS call form: adds the force of b on a to (fx,fy,fz), dr = r_b - r_a (:183)
//   double pair_energy(double d2) const                       energy of the pair, only for the EV instantiations (oracle-defined observables)
//
// and what ComputePairTraits<F> tells them (same names and meaning as the reference's traits; those the sweeps do not
// consult are fixed by static_asserts in the kernels):
//
//   BufferLessCompatible    must be true: the sweeps keep dr and d2 in registers and never materialise a ComputePairBuffer2
//                           (the reference's 256-neighbour scratch, 8 KB per thread: compute_pair_buffer.h:199-208 -- it is what
//                           overflows unchecked at rc = 5 sigma; not having it is a documented divergence)
//   ComputeBufferCompatible ignored (the buffer form `(n, buf, fields..., cells)` is never called)
//   CudaCompatible          must be true
//   RequiresNbhOptionalData must be false: weights come from ComputePairNullWeightIterator (compute_pair_optional_args.h), i.e. 1.0
//   HasParticleContext      must be false (no per-particle start/stop hooks on this path)
//   Batch4                  OURS: F also offers pairs4<EV>(dx, dy, dz, d2, ok, acc), four candidates evaluated phase by phase with
//                           independent FP64 chains; the sweeps call it instead of operator() when present.  Same pair set, same
//                           order of accumulation.
//
// Optional arguments of compute_cell_particle_pairs that the LJ operator passes (lennard_jones.cu:196-201) and how they appear here:
// ComputePairOptionalLocks<false> (zero-size fakes: a thread only ever writes its own particle -- the full, non-symmetric list
// makes the sweep atomics-free), ComputePairNullWeightIterator (weight = 1.0), LinearXForm{domain.xform()} (identity only:
// the deformable box is out of scope), no cell / particle filter.
#pragma once
#include "xnb_common.cuh"

namespace xnb {

// neighbour b as the sweep knows it (the reference passes cells, neighbor_cell, neighbor_particle): its index in the block's
// staged position arrays, or in the flat particle arrays on the global-memory path
struct PairNbh { uint32_t index; };

// per-particle accumulators of a sweep thread: force, and (EV) energy + virial partial sums
struct PairAcc { double ax, ay, az, e, wxx, wyy, wzz, wxy, wxz, wyz; };
typedef PairAcc LJAcc;

template <class F> struct ComputePairTraits
{
  static constexpr bool BufferLessCompatible = true;
  static constexpr bool ComputeBufferCompatible = false;
  static constexpr bool CudaCompatible = true;
  static constexpr bool RequiresNbhOptionalData = false;
  static constexpr bool HasParticleContext = false;
  static constexpr bool Batch4 = false;
};

// Lennard-Jones parameters as the kernels use them (LennardJonesParms {epsilon, sigma} + rcut, lennard_jones.cu:40-44,180)
struct LJP { double eps24; double sig2; double rcut2; double eps4; double neg_eps48; double eps; double sig; };

// 1/x for normal x > 0: hardware seed (MUFU.RCP64H, ~2^-20) + one cubically convergent step: y (1 + e + e^2), e = 1 - x y
XNB_DEVINL double fast_rcp(double x)
{
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  const double t = fma(e, e, e);
  return fma(y, t, y);                      // relative error e^3 ~ 2^-60, + rounding
}

// scheduling fence: everything that produces v0..v3 is placed before this point and everything that consumes them after
// it, so the four candidates advance phase by phase (independent FP64 chains in flight) instead of one after the other
XNB_DEVINL void fence4(double& v0, double& v1, double& v2, double& v3) { asm volatile("" : "+d"(v0), "+d"(v1), "+d"(v2), "+d"(v3)); }

// ------------------------------------------------------------------------------------------------------------------
// LennardJonesForceFunctor: the functor of lennard_jones.cu:59-126 restated with one reciprocal instead of sqrt + two
// divisions:  de/r = -24 eps (2 s12 - s6) / d2 = (24 eps - 48 eps s6) (s6 / d2),  s6 = (sigma^2/d2)^3  (<= 3e-13 relative on
// forces against the literal form below).  This is the instantiation the C-ABI runs by default.
// ------------------------------------------------------------------------------------------------------------------
struct LennardJonesForceFunctor
{
  LJP p;
  XNB_DEVINL double rcut2() const { return p.rcut2; }
  XNB_DEVINL double pair_energy(double d2) const { const double s2 = p.sig2 * fast_rcp(d2), s6 = s2 * s2 * s2; return p.eps4 * (s6 * s6 - s6); }
  XNB_DEVINL void operator()(double3 dr, double d2, double& fx, double& fy, double& fz, PairNbh, double weight) const
  {
    const double inv = fast_rcp(d2);
    const double s2 = p.sig2 * inv;
    const double s6 = s2 * s2 * s2;
    const double de = fma(p.neg_eps48, s6, p.eps24) * (s6 * inv) * weight;
    fx = fma(de, dr.x, fx); fy = fma(de, dr.y, fy); fz = fma(de, dr.z, fz);
  }
  // one pair with the observables (global-memory fallback path of k_lj_sweep)
  template <bool EV>
  XNB_DEVINL void pair1(double dx, double dy, double dz, double d2, PairAcc& a) const
  {
    const double inv = fast_rcp(d2);
    const double s2 = p.sig2 * inv;
    const double s6 = s2 * s2 * s2;
    const double de = fma(p.neg_eps48, s6, p.eps24) * (s6 * inv);
    a.ax = fma(de, dx, a.ax); a.ay = fma(de, dy, a.ay); a.az = fma(de, dz, a.az);
    if (EV)
    {
      a.e += 0.5 * p.eps4 * (s6 * s6 - s6);
      const double px = de * dx, py = de * dy, pz = de * dz;
      a.wxx -= 0.5 * dx * px; a.wyy -= 0.5 * dy * py; a.wzz -= 0.5 * dz * pz;
      a.wxy -= 0.5 * dx * py; a.wxz -= 0.5 * dx * pz; a.wyz -= 0.5 * dy * pz;
    }
  }
  // four candidates at once.  ok[u] = false (not a candidate, or outside the cut) contributes exactly zero; its d2 may be
  // anything (the coefficient is replaced, not multiplied).
  template <bool EV>
  XNB_DEVINL void pairs4(const double (&dx)[4], const double (&dy)[4], const double (&dz)[4], double (&d2)[4], const bool (&ok)[4], PairAcc& a) const
  {
    double inv[4], s6[4], de[4];
    fence4(d2[0], d2[1], d2[2], d2[3]);
#pragma unroll
    for (int u = 0; u < 4; u++) inv[u] = fast_rcp(d2[u]);
    fence4(inv[0], inv[1], inv[2], inv[3]);
#pragma unroll
    for (int u = 0; u < 4; u++) { const double s2 = p.sig2 * inv[u]; s6[u] = s2 * s2 * s2; }
    fence4(s6[0], s6[1], s6[2], s6[3]);
#pragma unroll
    for (int u = 0; u < 4; u++) { de[u] = fma(p.neg_eps48, s6[u], p.eps24) * (s6[u] * inv[u]); if (!ok[u]) de[u] = 0.0; }
    fence4(de[0], de[1], de[2], de[3]);
#pragma unroll
    for (int u = 0; u < 4; u++) { a.ax = fma(de[u], dx[u], a.ax); a.ay = fma(de[u], dy[u], a.ay); a.az = fma(de[u], dz[u], a.az); }
    if (EV)
    {
#pragma unroll
      for (int u = 0; u < 4; u++)
      {
        if (ok[u]) a.e += 0.5 * p.eps4 * (s6[u] * s6[u] - s6[u]);
        const double px = de[u] * dx[u], py = de[u] * dy[u], pz = de[u] * dz[u];
        a.wxx -= 0.5 * dx[u] * px; a.wyy -= 0.5 * dy[u] * py; a.wzz -= 0.5 * dz[u] * pz;
        a.wxy -= 0.5 * dx[u] * py; a.wxz -= 0.5 * dx[u] * pz; a.wyz -= 0.5 * dy[u] * pz;
      }
    }
  }
};
template <> struct ComputePairTraits<LennardJonesForceFunctor>
{
  static constexpr bool BufferLessCompatible = true;
  static constexpr bool ComputeBufferCompatible = false;
  static constexpr bool CudaCompatible = true;
  static constexpr bool RequiresNbhOptionalData = false;
  static constexpr bool HasParticleContext = false;
  static constexpr bool Batch4 = true;
};

// ------------------------------------------------------------------------------------------------------------------
// LennardJonesForceFunctorRef: the literal form of the reference -- lj_compute_energy (lennard_jones.cu:46-56) and the
// buffer-less operator() (:106-124): r = sqrt(d2); ir = 1/r; s = sigma ir; s6 = (s^2)^3; s12 = s6^2; e = 4 eps (s12 - s6);
// de = -24 eps (2 s12 - s6) ir; de *= w / r; f += de dr.  It has no Batch4 hook: the sweeps drive it through the generic
// buffer-less call, one candidate at a time -- the path any other functor of the concept would take
// (xnb_set_pair_functor(ctx, XNB_FUNCTOR_LJ_REFERENCE_FORM)).
// ------------------------------------------------------------------------------------------------------------------
struct LennardJonesForceFunctorRef
{
  LJP p;
  XNB_DEVINL double rcut2() const { return p.rcut2; }
  XNB_DEVINL void lj_compute_energy(double r, double& e, double& de) const
  {
    const double inv_r = 1.0 / r;
    const double ratio = p.sig * inv_r;
    const double ratio2 = ratio * ratio;
    const double ratio6 = ratio2 * ratio2 * ratio2;
    const double ratio12 = ratio6 * ratio6;
    e = 4. * p.eps * (ratio12 - ratio6);
    de = (-24. * p.eps * (2. * ratio12 - ratio6)) * inv_r;
  }
  XNB_DEVINL double pair_energy(double d2) const { double e, de; lj_compute_energy(sqrt(d2), e, de); return e; }
  XNB_DEVINL void operator()(double3 dr, double d2, double& fx, double& fy, double& fz, PairNbh, double weight) const
  {
    const double r = sqrt(d2);
    double pair_e = 0.0, pair_de = 0.0;
    lj_compute_energy(r, pair_e, pair_de);
    pair_de *= weight / r;
    fx += pair_de * dr.x;
    fy += pair_de * dr.y;
    fz += pair_de * dr.z;
  }
};

// ------------------------------------------------------------------------------------------------------------------
// how a sweep applies F to the candidates it has gathered
// ------------------------------------------------------------------------------------------------------------------
// one accepted pair (d2 already tested against the cut)
template <bool EV, class F>
XNB_DEVINL void pair_apply1(const F& f, double dx, double dy, double dz, double d2, uint32_t nbh, PairAcc& a)
{
  static_assert(ComputePairTraits<F>::BufferLessCompatible && ComputePairTraits<F>::CudaCompatible, "the sweeps call the buffer-less device form");
  static_assert(!ComputePairTraits<F>::RequiresNbhOptionalData && !ComputePairTraits<F>::HasParticleContext, "no per-neighbour data / particle context on this path");
  if constexpr (ComputePairTraits<F>::Batch4) f.template pair1<EV>(dx, dy, dz, d2, a);
  else if (!EV) f(make_double3(dx, dy, dz), d2, a.ax, a.ay, a.az, PairNbh{nbh}, 1.0);
  else
  {
    double tx = 0., ty = 0., tz = 0.;
    f(make_double3(dx, dy, dz), d2, tx, ty, tz, PairNbh{nbh}, 1.0);
    a.ax += tx; a.ay += ty; a.az += tz;
    a.e += 0.5 * f.pair_energy(d2);
    a.wxx -= 0.5 * dx * tx; a.wyy -= 0.5 * dy * ty; a.wzz -= 0.5 * dz * tz;
    a.wxy -= 0.5 * dx * ty; a.wxz -= 0.5 * dx * tz; a.wyz -= 0.5 * dy * tz;
  }
}

// four gathered candidates, ok[u] = listed and inside the cut
template <bool EV, class F>
XNB_DEVINL void pair_apply4(const F& f, const double (&dx)[4], const double (&dy)[4], const double (&dz)[4], double (&d2)[4], const bool (&ok)[4],
                            const uint32_t (&nbh)[4], PairAcc& a)
{
  if constexpr (ComputePairTraits<F>::Batch4) f.template pairs4<EV>(dx, dy, dz, d2, ok, a);
  else
  {
#pragma unroll
    for (int u = 0; u < 4; u++) if (ok[u]) pair_apply1<EV>(f, dx[u], dy[u], dz[u], d2[u], nbh[u], a);
  }
}

} // namespace xnb
