// xnb_host_inputs.cpp -- host-side input operators of the LJ decks: `lattice` (FCC) and `gaussian_noise_r` / velocities.
// They only create the initial particle set handed to xnb_set_particles; nothing here is on the timed path.
//   reference: src/grid_cell_particles/include/exanb/grid_cell_particles/generate_particle_lattice.h:247-388 (lattice walk,
//              inclusion tests, deterministic ids), lattice_generator.h:162-170 (FCC basis),
//              src/compute/include/exanb/compute/gaussian_noise.h:61-80,150-161 (per-domain-cell reseeded std::mt19937_64)
#include "../../include/xnb_hotpath.h"
#include <cmath>
#include <cstdint>
#include <random>
#include <vector>

namespace {
struct Atom { double r[3]; double v[3]; };
}

extern "C" int64_t xnb_host_lattice_fcc(const xnb_lattice_cfg* cfg, int64_t capacity, double* rx, double* ry, double* rz,
                                        double* vx, double* vy, double* vz, uint64_t* id, uint8_t* type)
{
  static const double basis[4][3] = {{0., 0., 0.}, {0., .5, .5}, {.5, 0., .5}, {.5, .5, 0.}};
  const double a = cfg->lattice_a, cs = cfg->cell_size;
  const int64_t* gd = cfg->grid_dims;
  const int64_t n_cells = gd[0] * gd[1] * gd[2];
  std::vector<std::vector<Atom>> cells((size_t)n_cells);
  // C5 clusters: spheres from std::mt19937(12345), explicit arithmetic
  struct Sph { double c[3], r, d[3]; };
  std::vector<Sph> sph;
  {
    std::mt19937 g(12345);
    auto u01 = [&]() { return (double)g() / 4294967296.0; };
    for (int s = 0; s < cfg->n_spheres; s++)
    {
      Sph q;
      for (int d = 0; d < 3; d++) q.c[d] = cfg->bounds_min[d] + u01() * (cfg->bounds_max[d] - cfg->bounds_min[d]);
      q.r = cfg->sphere_rmin + u01() * (cfg->sphere_rmax - cfg->sphere_rmin);
      for (int d = 0; d < 3; d++) q.d[d] = (g() & 1u) ? cfg->drift_speed : -cfg->drift_speed;
      sph.push_back(q);
    }
  }
  int64_t lo[3], hi[3];
  for (int d = 0; d < 3; d++) { lo[d] = (int64_t)std::floor(cfg->bounds_min[d] / a) - 1; hi[d] = (int64_t)std::ceil(cfg->bounds_max[d] / a) + 1; }
  for (int64_t k = lo[2]; k <= hi[2]; k++) for (int64_t j = lo[1]; j <= hi[1]; j++) for (int64_t i = lo[0]; i <= hi[0]; i++)
    for (int l = 0; l < 4; l++)
    {
      const double p[3] = {((double)i + basis[l][0]) * a, ((double)j + basis[l][1]) * a, ((double)k + basis[l][2]) * a};
      int64_t loc[3]; bool ok = true;
      for (int d = 0; d < 3; d++)
      {
        loc[d] = (int64_t)std::floor((p[d] - cfg->bounds_min[d]) / cs);
        const double gmax = cfg->bounds_min[d] + (double)gd[d] * cs;
        if (loc[d] < 0 || loc[d] >= gd[d] || p[d] < cfg->bounds_min[d] || p[d] > cfg->bounds_max[d] || p[d] > gmax) ok = false;
      }
      if (!ok) continue;
      Atom at{{p[0], p[1], p[2]}, {0., 0., 0.}};
      if (!sph.empty())
      {
        bool keep = false;
        for (const Sph& q : sph)
        {
          const double dx = p[0] - q.c[0], dy = p[1] - q.c[1], dz = p[2] - q.c[2];
          if (dx * dx + dy * dy + dz * dz <= q.r * q.r) { keep = true; at.v[0] = q.d[0]; at.v[1] = q.d[1]; at.v[2] = q.d[2]; break; }
        }
        if (!keep) continue;
      }
      cells[(size_t)((loc[2] * gd[1] + loc[1]) * gd[0] + loc[0])].push_back(at);
    }
  // noise: one engine per domain cell, seed = cell_index*1023 (+1 for velocities), draws in in-cell order x,y,z
  for (int64_t c = 0; c < n_cells; c++)
  {
    if (cfg->noise_sigma > 0.)
    {
      std::mt19937_64 re; re.seed((uint64_t)c * 1023u);
      std::normal_distribution<double> gs(0.0, cfg->noise_sigma);
      for (Atom& at : cells[(size_t)c]) { at.r[0] += gs(re); at.r[1] += gs(re); at.r[2] += gs(re); }
    }
    if (cfg->vel_sigma > 0.)
    {
      std::mt19937_64 re; re.seed((uint64_t)c * 1023u + 1u);
      std::normal_distribution<double> gs(0.0, cfg->vel_sigma);
      for (Atom& at : cells[(size_t)c]) { at.v[0] += gs(re); at.v[1] += gs(re); at.v[2] += gs(re); }
    }
  }
  int64_t n = 0;
  for (const auto& c : cells) n += (int64_t)c.size();
  if (n > capacity) return -n;
  if (cfg->vel_sigma > 0. && n > 0)
  {
    double s[3] = {0, 0, 0};
    for (const auto& c : cells) for (const Atom& at : c) { s[0] += at.v[0]; s[1] += at.v[1]; s[2] += at.v[2]; }
    for (int d = 0; d < 3; d++) s[d] /= (double)n;
    for (auto& c : cells) for (Atom& at : c) { at.v[0] -= s[0]; at.v[1] -= s[1]; at.v[2] -= s[2]; }
  }
  int64_t o = 0;
  for (const auto& c : cells) for (const Atom& at : c)
  {
    rx[o] = at.r[0]; ry[o] = at.r[1]; rz[o] = at.r[2]; vx[o] = at.v[0]; vy[o] = at.v[1]; vz[o] = at.v[2];
    if (id) id[o] = (uint64_t)o + 1u;      // ids start right after the greatest existing id (0 for an empty grid) :139-158
    if (type) type[o] = 0;
    o++;
  }
  return n;
}
