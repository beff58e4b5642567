/*
 * xnb_oracle.cpp -- CPU ORACLE (test infrastructure, see xnb_oracle.h) for the exaNBody LJ hot path.
 *
 * Own-written restatement of the reference algorithms; nothing here is copied from the reference.
 * Each function cites the reference file:line (relative to the reference tree) whose behaviour it restates.
 * Compile with -ffp-contract=off: distance tests are defined as plain left-to-right x*x + y*y + z*z
 * (the reference's norm2 lives in onika, which is not in the tree: "parity unpinned" at the bit level).
 */
#include "xnb_oracle.h"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <random>
#include <string>
#include <vector>
#include <omp.h>

namespace {

using i64 = int64_t;
struct V3 { double x, y, z; };
struct IJK { i64 i, j, k; };
static inline IJK operator+(IJK a, IJK b) { return {a.i + b.i, a.j + b.j, a.k + b.k}; }
static inline IJK operator-(IJK a, IJK b) { return {a.i - b.i, a.j - b.j, a.k - b.k}; }
static inline IJK operator+(IJK a, i64 s) { return {a.i + s, a.j + s, a.k + s}; }
static inline bool operator!=(IJK a, IJK b) { return a.i != b.i || a.j != b.j || a.k != b.k; }

struct AABB { V3 lo, hi; };

/* core/grid_algorithm.h:84-92 : (k,j,i) lexicographic cell index */
static inline i64 ijk_to_index(IJK d, IJK p) { return (p.k * d.j + p.j) * d.i + p.i; }
/* core/grid_algorithm.h:104-112 */
static inline IJK index_to_ijk(IJK d, i64 idx) { i64 i = idx % d.i; idx /= d.i; i64 j = idx % d.j; idx /= d.j; return {i, j, idx}; }
/* core/grid_algorithm.h:96-101 */
static inline bool grid_contains(IJK d, IJK p) { return p.i >= 0 && p.i < d.i && p.j >= 0 && p.j < d.j && p.k >= 0 && p.k < d.k; }
/* core/grid_algorithm.h:142-150 : distance to nearest border in [margin, margin+thickness) */
static inline bool inside_grid_shell(IJK d, i64 margin, i64 thickness, IJK c)
{
  i64 di = std::min(c.i, d.i - 1 - c.i), dj = std::min(c.j, d.j - 1 - c.j), dk = std::min(c.k, d.k - 1 - c.k);
  i64 b = std::min(std::min(di, dj), dk);
  return b >= margin && b < margin + thickness;
}
/* core/domain.h:143-148 */
static inline i64 pmod(i64 i, i64 n) { return ((i % n) + n) % n; }

/* core/geometry.h:140-156 : squared distance from a point to a box */
static inline double min_dist2_point_box(V3 p, const AABB& b)
{
  double dx = 0., dy = 0., dz = 0.;
  if (p.x < b.lo.x) dx = b.lo.x - p.x; else if (p.x > b.hi.x) dx = p.x - b.hi.x;
  if (p.y < b.lo.y) dy = b.lo.y - p.y; else if (p.y > b.hi.y) dy = p.y - b.hi.y;
  if (p.z < b.lo.z) dz = b.lo.z - p.z; else if (p.z > b.hi.z) dz = p.z - b.hi.z;
  return dx * dx + dy * dy + dz * dz;
}
/* core/geometry.h:107-131 : 1-D gap between two ranges, then box/box squared min distance */
static inline double range_min_dist(double amin, double amax, double bmin, double bmax)
{
  if (amax < bmin) return bmin - amax;
  if (bmax < amin) return amin - bmax;
  return 0.0;
}
static inline double min_dist2_box_box(const AABB& a, const AABB& b)
{
  double x = range_min_dist(a.lo.x, a.hi.x, b.lo.x, b.hi.x);
  double y = range_min_dist(a.lo.y, a.hi.y, b.lo.y, b.hi.y);
  double z = range_min_dist(a.lo.z, a.hi.z, b.lo.z, b.hi.z);
  return x * x + y * y + z * z;
}
static inline bool in_range_incl(double lo, double hi, double x) { return x >= lo && x <= hi; }        /* geometry.h:184-188 */
static inline bool is_inside_incl(const AABB& b, V3 p)                                                 /* geometry.h:198-203 */
{ return in_range_incl(b.lo.x, b.hi.x, p.x) && in_range_incl(b.lo.y, b.hi.y, p.y) && in_range_incl(b.lo.z, b.hi.z, p.z); }

/* per-cell SoA block: the default field set rx..rz, vx..vz, fx..fz(=ax..az), id, type (core/grid_fields.h:83-99,114) */
struct Cell
{
  std::vector<double> rx, ry, rz, vx, vy, vz, fx, fy, fz;
  std::vector<uint64_t> id;
  std::vector<uint8_t> type;
  size_t size() const { return rx.size(); }
  void clear() { rx.clear(); ry.clear(); rz.clear(); vx.clear(); vy.clear(); vz.clear(); fx.clear(); fy.clear(); fz.clear(); id.clear(); type.clear(); }
  void resize(size_t n) { rx.resize(n); ry.resize(n); rz.resize(n); vx.resize(n); vy.resize(n); vz.resize(n); fx.resize(n); fy.resize(n); fz.resize(n); id.resize(n); type.resize(n); }
};
struct Tuple { double rx, ry, rz, vx, vy, vz, fx, fy, fz; uint64_t id; uint8_t type; };
static inline Tuple get_tuple(const Cell& c, size_t p) { return {c.rx[p], c.ry[p], c.rz[p], c.vx[p], c.vy[p], c.vz[p], c.fx[p], c.fy[p], c.fz[p], c.id[p], c.type[p]}; }
static inline void set_tuple(Cell& c, size_t p, const Tuple& t)
{ c.rx[p] = t.rx; c.ry[p] = t.ry; c.rz[p] = t.rz; c.vx[p] = t.vx; c.vy[p] = t.vy; c.vz[p] = t.vz; c.fx[p] = t.fx; c.fy[p] = t.fy; c.fz[p] = t.fz; c.id[p] = t.id; c.type[p] = t.type; }
static inline void push_tuple(Cell& c, const Tuple& t)
{ c.rx.push_back(t.rx); c.ry.push_back(t.ry); c.rz.push_back(t.rz); c.vx.push_back(t.vx); c.vy.push_back(t.vy); c.vz.push_back(t.vz); c.fx.push_back(t.fx); c.fy.push_back(t.fy); c.fz.push_back(t.fz); c.id.push_back(t.id); c.type.push_back(t.type); }

/* core/grid.h:57-688 (the slice the hot path touches) */
struct Grid
{
  V3 origin{0, 0, 0};
  IJK offset{0, 0, 0};
  IJK dims{0, 0, 0};
  double cell_size = 0.;
  double max_nbh_dist = 0.;
  std::vector<Cell> cells;
  i64 ghost_layers() const { return (i64)std::ceil(max_nbh_dist / cell_size); }                       /* grid.h:110 */
  i64 n_cells() const { return dims.i * dims.j * dims.k; }
  V3 cell_position(IJK loc) const                                                                      /* grid.h:113-116 */
  { return {origin.x + (double)(offset.i + loc.i) * cell_size, origin.y + (double)(offset.j + loc.j) * cell_size, origin.z + (double)(offset.k + loc.k) * cell_size}; }
  AABB cell_bounds(IJK loc) const { return {cell_position(loc), cell_position(loc + 1)}; }             /* grid.h:119-122 */
  double eps_cell_size2() const { double e = (1.0 / (double)(1ull << 48)) * cell_size; return e * e; } /* grid.h:60,89-90 */
  bool is_ghost_cell(IJK loc) const { return inside_grid_shell(dims, 0, ghost_layers(), loc); }        /* grid.h:202-206 */
};

struct SubCellPairs { std::vector<uint16_t> ab; };
/* amr/amr_grid_algorithm.h:439-453 */
struct PairCache
{
  double cell_size = 0., max_dist = 0.; size_t max_res = 0; std::vector<SubCellPairs> pairs;
  size_t layers() const { return (size_t)std::ceil(max_dist / cell_size); }
  size_t n_nbh_cells() const { size_t x = layers() + 1; return x * x * x; }
};

struct GhostSend { i64 cell_i; i64 partner_cell_i; uint32_t flags; std::vector<uint32_t> particle_i; };

enum : uint32_t { SHIFT_X = 1u << 0, SIDE_X = 1u << 2, SHIFT_Y = 1u << 3, SIDE_Y = 1u << 5, SHIFT_Z = 1u << 6, SIDE_Z = 1u << 8 };

static thread_local std::string g_err;

} // namespace

struct xo_sim
{
  xo_config cfg;
  /* Domain */
  V3 dmin, dmax; IJK ddims; double cell_size; bool periodic[3];
  Grid grid;
  /* nbh_dist.cpp:47-71 */
  double nbh_dist = 0., max_displ = 0., ghost_dist = 0.;
  /* AmrGrid */
  std::vector<size_t> sub_grid_start; std::vector<uint32_t> sub_grid_cells;
  PairCache pair_cache;
  /* PositionBackupData */
  std::vector<std::vector<uint32_t>> backup;
  /* GhostCommunicationScheme, self partner only */
  std::vector<GhostSend> sends;
  /* GridChunkNeighbors */
  std::vector<std::vector<uint16_t>> streams;
  unsigned max_neighbors = 0;
  i64 rebuilds = 0;
  /* ChunkNeighborsConfig::half_symmetric / skip_ghosts (chunk_neighbors_config.h:35-36) */
  bool nbh_half_symmetric = false, nbh_skip_ghosts = false;
};

namespace {

/* core/domain.h:160-191 : wrap r into the periodic domain, return the domain-grid location of its cell */
static IJK domain_periodic_location(const xo_sim& s, V3& r)
{
  V3 rel{r.x - s.dmin.x, r.y - s.dmin.y, r.z - s.dmin.z};
  IJK loc{(i64)std::floor(rel.x / s.cell_size), (i64)std::floor(rel.y / s.cell_size), (i64)std::floor(rel.z / s.cell_size)};
  if ((loc.i < 0 || loc.i >= s.ddims.i) && s.periodic[0]) { i64 o = loc.i; loc.i = pmod(loc.i, s.ddims.i); r.x += (double)(loc.i - o) * s.cell_size; }
  if ((loc.j < 0 || loc.j >= s.ddims.j) && s.periodic[1]) { i64 o = loc.j; loc.j = pmod(loc.j, s.ddims.j); r.y += (double)(loc.j - o) * s.cell_size; }
  if ((loc.k < 0 || loc.k >= s.ddims.k) && s.periodic[2]) { i64 o = loc.k; loc.k = pmod(loc.k, s.ddims.k); r.z += (double)(loc.k - o) * s.cell_size; }
  return loc;
}

/* ------------------------------------------------------------------------------------------------
 * input_data: lattice (generate_particle_lattice.h:247-388, FCC basis lattice_generator.h:162-170),
 * gaussian_noise_r (gaussian_noise.h:61-80,150-161).  Grid has NO ghost layers at this point
 * (init_rcb_grid.cpp:65-77 leaves max_neighbor_distance = 0).
 * ---------------------------------------------------------------------------------------------- */
static void init_rcb_grid(xo_sim& s)
{
  s.grid = Grid{};
  s.grid.offset = {0, 0, 0};               /* single rank: simple_block_rcb(block,1,0) = whole domain */
  s.grid.origin = s.dmin;
  s.grid.cell_size = s.cell_size;
  s.grid.dims = s.ddims;
  s.grid.max_nbh_dist = 0.;
  s.grid.cells.assign((size_t)s.grid.n_cells(), Cell{});
}

struct Sphere { V3 c; double r; V3 drift; };
static std::vector<Sphere> make_spheres(const xo_sim& s)
{
  /* C5 input: spheres from std::mt19937(12345); explicit arithmetic so any host restates it bit-exactly */
  std::vector<Sphere> sp;
  std::mt19937 g(12345);
  auto u01 = [&]() { return (double)g() / 4294967296.0; };
  for (int i = 0; i < s.cfg.n_spheres; i++)
  {
    Sphere q;
    q.c.x = s.dmin.x + u01() * (s.dmax.x - s.dmin.x);
    q.c.y = s.dmin.y + u01() * (s.dmax.y - s.dmin.y);
    q.c.z = s.dmin.z + u01() * (s.dmax.z - s.dmin.z);
    q.r = s.cfg.sphere_rmin + u01() * (s.cfg.sphere_rmax - s.cfg.sphere_rmin);
    q.drift.x = (g() & 1u) ? s.cfg.drift_speed : -s.cfg.drift_speed;
    q.drift.y = (g() & 1u) ? s.cfg.drift_speed : -s.cfg.drift_speed;
    q.drift.z = (g() & 1u) ? s.cfg.drift_speed : -s.cfg.drift_speed;
    sp.push_back(q);
  }
  return sp;
}

static void lattice_fcc(xo_sim& s)
{
  static const double basis[4][3] = {{0., 0., 0.}, {0., .5, .5}, {.5, 0., .5}, {.5, .5, 0.}};   /* lattice_generator.h:162-170 */
  Grid& g = s.grid;
  const double a = s.cfg.lattice_a;
  const AABB dom{s.dmin, s.dmax};
  const AABB gb{g.cell_position({0, 0, 0}), g.cell_position(g.dims)};
  /* generate_particle_lattice.h:247-262 : lattice index box padded by one */
  IJK lo{(i64)std::floor(s.dmin.x / a), (i64)std::floor(s.dmin.y / a), (i64)std::floor(s.dmin.z / a)};
  IJK hi{(i64)std::ceil(s.dmax.x / a), (i64)std::ceil(s.dmax.y / a), (i64)std::ceil(s.dmax.z / a)};
  const std::vector<Sphere> spheres = make_spheres(s);
  const uint64_t no_id = std::numeric_limits<uint64_t>::max();
  for (i64 k = lo.k - 1; k <= hi.k + 1; k++)
    for (i64 j = lo.j - 1; j <= hi.j + 1; j++)
      for (i64 i = lo.i - 1; i <= hi.i + 1; i++)
        for (int l = 0; l < 4; l++)
        {
          /* :289-297 : position, locate_cell (floor), inclusive inside tests */
          V3 p{((double)i + basis[l][0]) * a, ((double)j + basis[l][1]) * a, ((double)k + basis[l][2]) * a};
          IJK loc{(i64)std::floor((p.x - g.origin.x) / g.cell_size) - g.offset.i,
                  (i64)std::floor((p.y - g.origin.y) / g.cell_size) - g.offset.j,
                  (i64)std::floor((p.z - g.origin.z) / g.cell_size) - g.offset.k};
          if (!(grid_contains(g.dims, loc) && is_inside_incl(dom, p) && is_inside_incl(gb, p))) continue;
          V3 drift{0, 0, 0};
          if (!spheres.empty())
          {
            bool keep = false;
            for (const Sphere& q : spheres)
            {
              double dx = p.x - q.c.x, dy = p.y - q.c.y, dz = p.z - q.c.z;
              if (dx * dx + dy * dy + dz * dz <= q.r * q.r) { keep = true; drift = q.drift; break; }
            }
            if (!keep) continue;
          }
          Cell& c = g.cells[(size_t)ijk_to_index(g.dims, loc)];
          push_tuple(c, Tuple{p.x, p.y, p.z, drift.x, drift.y, drift.z, 0., 0., 0., no_id, 0});
        }
  /* :139-158 next_id = (max existing id, 0 for an empty grid) + 1  => ids start at 1 ;
     :328-388 deterministic ids: exclusive scan of per-domain-cell counts, then in-cell order */
  uint64_t next = 1;
  for (Cell& c : g.cells) for (size_t p = 0; p < c.size(); p++) c.id[p] = next++;
}

/* gaussian_noise.h:61-80 (apply), :150-161 (per-cell reseed with domain_cell_index*1023 [+ seed_shift, ours]) */
static void gaussian_noise(xo_sim& s, double sigma, bool velocity, uint64_t seed_shift)
{
  Grid& g = s.grid;
  const i64 gl = g.ghost_layers();
  for (i64 k = gl; k < g.dims.k - gl; k++) for (i64 j = gl; j < g.dims.j - gl; j++) for (i64 i = gl; i < g.dims.i - gl; i++)
  {
    IJK loc{i, j, k};
    Cell& c = g.cells[(size_t)ijk_to_index(g.dims, loc)];
    const i64 dom_idx = ijk_to_index(s.ddims, loc + g.offset);
    std::mt19937_64 re;
    re.seed((uint64_t)dom_idx * 1023u + seed_shift);
    std::normal_distribution<double> gauss(0.0, sigma);
    for (size_t p = 0; p < c.size(); p++)
    {
      if (!velocity) { c.rx[p] += gauss(re); c.ry[p] += gauss(re); c.rz[p] += gauss(re); }
      else           { c.vx[p] += gauss(re); c.vy[p] += gauss(re); c.vz[p] += gauss(re); }
    }
  }
}

/* synthetic configs only: remove net momentum (equal masses) */
static void zero_momentum(xo_sim& s)
{
  double sx = 0, sy = 0, sz = 0; i64 n = 0;
  for (const Cell& c : s.grid.cells) for (size_t p = 0; p < c.size(); p++) { sx += c.vx[p]; sy += c.vy[p]; sz += c.vz[p]; n++; }
  if (n == 0) return;
  sx /= (double)n; sy /= (double)n; sz /= (double)n;
  for (Cell& c : s.grid.cells) for (size_t p = 0; p < c.size(); p++) { c.vx[p] -= sx; c.vy[p] -= sy; c.vz[p] -= sz; }
}

/* ------------------------------------------------------------------------------------------------
 * move_particles  (grid_cell_particles/.../move_particles_across_cells.h:104-229)
 * ---------------------------------------------------------------------------------------------- */
static int move_particles(xo_sim& s)
{
  Grid& g = s.grid;
  const i64 n_cells = g.n_cells();
  const i64 gl = g.ghost_layers();
  const IJK dng{g.dims.i - 2 * gl, g.dims.j - 2 * gl, g.dims.k - 2 * gl};
  const double eps2 = g.eps_cell_size2();
  struct Mover { i64 dst; Tuple t; };
  std::vector<std::vector<Mover>> movers((size_t)n_cells);       /* per SOURCE cell, in p order */
  std::vector<std::vector<int32_t>> removed((size_t)n_cells);
  /* :104-156 : detect movers.  The reference pushes into destination buffers inside an omp critical, so its
     multi-thread in-cell arrival order is unspecified; we collect per source cell and merge in (k,j,i,p) order,
     which IS the reference's single-thread order. */
# pragma omp parallel for collapse(3) schedule(dynamic)
  for (i64 k = 0; k < dng.k; k++) for (i64 j = 0; j < dng.j; j++) for (i64 i = 0; i < dng.i; i++)
  {
    IJK src{i + gl, j + gl, k + gl};
    const i64 ci = ijk_to_index(g.dims, src);
    const Cell& c = g.cells[(size_t)ci];
    for (size_t p = 0; p < c.size(); p++)
    {
      V3 r{c.rx[p], c.ry[p], c.rz[p]};
      const V3 ro = r;
      IJK dst = domain_periodic_location(s, r) - g.offset;
      if (src != dst || min_dist2_point_box(ro, g.cell_bounds(dst)) >= eps2)
      {
        i64 cj = -1;   /* -1 : outside the grid or ghost => otb_particles (:124-138) */
        if (grid_contains(g.dims, dst) && !inside_grid_shell(g.dims, 0, gl, dst)) cj = ijk_to_index(g.dims, dst);
        Tuple t = get_tuple(c, p); t.rx = r.x; t.ry = r.y; t.rz = r.z;
        movers[(size_t)ci].push_back({cj, t});
        removed[(size_t)ci].push_back((int32_t)p);
      }
    }
  }
  std::vector<std::vector<Tuple>> inbox((size_t)n_cells);
  i64 otb = 0;
  for (i64 ci = 0; ci < n_cells; ci++) for (const Mover& m : movers[(size_t)ci]) { if (m.dst >= 0) inbox[(size_t)m.dst].push_back(m.t); else otb++; }
  /* :170-216 : compact sources back-to-front (holes take the current last element), then append arrivals */
# pragma omp parallel for schedule(dynamic)
  for (i64 ci = 0; ci < n_cells; ci++)
  {
    Cell& c = g.cells[(size_t)ci];
    int32_t n = (int32_t)c.size();
    const int32_t n_out = (int32_t)removed[(size_t)ci].size();
    if (n_out > 0)
    {
      std::vector<int32_t> pack((size_t)n);
      for (int32_t j = 0; j < n; j++) pack[(size_t)j] = j;
      for (int32_t j = 0; j < n_out; j++) pack[(size_t)removed[(size_t)ci][(size_t)j]] = -1;
      for (int32_t j = n - 1; j >= 0; j--)
        if (pack[(size_t)j] == -1) { if (j < n - 1) pack[(size_t)j] = pack[(size_t)(n - 1)]; --n; }
      for (int32_t j = 0; j < n; j++) if (j != pack[(size_t)j]) set_tuple(c, (size_t)j, get_tuple(c, (size_t)pack[(size_t)j]));
      c.resize((size_t)n);
    }
    for (const Tuple& t : inbox[(size_t)ci]) push_tuple(c, t);
  }
  if (otb > 0) { g_err = "move_particles: " + std::to_string(otb) + " particles left the single-rank grid (non-periodic exit)"; return 1; }
  return 0;
}

/* "stable" variant used when cfg.serial_order == 0: same cell membership, arrivals merged by (src cell, p) after stayers in
   their previous relative order.  (Membership is what the reference pins; order inside a cell is not, see above.) */

/* ------------------------------------------------------------------------------------------------
 * migrate_cell_particles, single rank (mpi/migrate_cell_particles.cpp:101-143): keep inner cells, set the
 * ghost shell thickness from ghost_dist, leave ghost cells EMPTY.
 * ---------------------------------------------------------------------------------------------- */
static void migrate_single_rank(xo_sim& s)
{
  Grid& g = s.grid;
  const i64 old_gl = g.ghost_layers();
  Grid ng;
  ng.origin = g.origin; ng.cell_size = g.cell_size; ng.max_nbh_dist = s.ghost_dist;
  const i64 gl = ng.ghost_layers();
  ng.offset = {-gl, -gl, -gl};
  ng.dims = {s.ddims.i + 2 * gl, s.ddims.j + 2 * gl, s.ddims.k + 2 * gl};
  ng.cells.assign((size_t)ng.n_cells(), Cell{});
  for (i64 k = 0; k < s.ddims.k; k++) for (i64 j = 0; j < s.ddims.j; j++) for (i64 i = 0; i < s.ddims.i; i++)
  {
    Cell& src = g.cells[(size_t)ijk_to_index(g.dims, IJK{i + old_gl, j + old_gl, k + old_gl})];
    ng.cells[(size_t)ijk_to_index(ng.dims, IJK{i + gl, j + gl, k + gl})] = std::move(src);
  }
  s.grid = std::move(ng);
}

/* ------------------------------------------------------------------------------------------------
 * rebuild_amr  (amr/rebuild_amr.cpp:42-61 ; amr_grid_algorithm.h:66-78, 378-434, 112-374)
 * ---------------------------------------------------------------------------------------------- */
static size_t sub_grid_size(size_t n, double avg_density)
{
  if (n == 0) return 0;
  double side = std::cbrt((double)n / avg_density);
  if (side < 2.0) return 1;
  return std::min((size_t)std::floor(side), (size_t)16);
}
static inline i64 icbrt(i64 n) { i64 r = (i64)std::floor(std::cbrt((double)n) + 0.5); return r; }
static inline i64 clampi(i64 v, i64 lo, i64 hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void rebuild_amr(xo_sim& s)
{
  Grid& g = s.grid;
  const i64 n_cells = g.n_cells();
  s.sub_grid_start.assign((size_t)n_cells + 1, 0);
  for (i64 c = 0; c < n_cells; c++)
  {
    i64 side = (i64)sub_grid_size(g.cells[(size_t)c].size(), s.cfg.sub_grid_density);
    s.sub_grid_start[(size_t)c + 1] = (size_t)std::max<i64>(side * side * side - 1, 0);
  }
  std::partial_sum(s.sub_grid_start.begin(), s.sub_grid_start.end(), s.sub_grid_start.begin());
  s.sub_grid_cells.assign(s.sub_grid_start[(size_t)n_cells], 0);
  /* project_particles_in_sub_grids: stable counting sort by sub-cell index (k*s+j)*s+i of trunc(pcoord*s) clamped */
# pragma omp parallel for schedule(dynamic)
  for (i64 ci = 0; ci < n_cells; ci++)
  {
    Cell& c = g.cells[(size_t)ci];
    const i64 n = (i64)c.size();
    const i64 sgstart = (i64)s.sub_grid_start[(size_t)ci];
    const i64 sgsize = (i64)s.sub_grid_start[(size_t)ci + 1] - sgstart;
    if (sgsize <= 0) continue;
    const i64 side = icbrt(sgsize + 1);
    const V3 low = g.cell_position(index_to_ijk(g.dims, ci));
    std::vector<i64> sgidx((size_t)n);
    std::vector<uint32_t> cnt((size_t)sgsize + 1, 0);
    for (i64 p = 0; p < n; p++)
    {
      /* grid.h:184-190 particle_pcoord */
      double px = (c.rx[(size_t)p] - low.x) / g.cell_size, py = (c.ry[(size_t)p] - low.y) / g.cell_size, pz = (c.rz[(size_t)p] - low.z) / g.cell_size;
      i64 si = clampi((i64)(px * (double)side), 0, side - 1);
      i64 sj = clampi((i64)(py * (double)side), 0, side - 1);
      i64 sk = clampi((i64)(pz * (double)side), 0, side - 1);
      sgidx[(size_t)p] = (sk * side + sj) * side + si;
      cnt[(size_t)sgidx[(size_t)p]]++;
    }
    std::vector<uint32_t> start((size_t)sgsize + 2, 0);
    for (i64 q = 0; q <= sgsize; q++) start[(size_t)q + 1] = start[(size_t)q] + cnt[(size_t)q];
    /* the reference realises this permutation in place with cycle swaps (:256-281); the result is the stable sort */
    Cell out; out.resize((size_t)n);
    std::vector<uint32_t> cur(start.begin(), start.end() - 1);
    for (i64 p = 0; p < n; p++) set_tuple(out, cur[(size_t)sgidx[(size_t)p]]++, get_tuple(c, (size_t)p));
    c = std::move(out);
    /* sub_grid_cells[sgstart+q] = number of particles in sub-cells 0..q (cumulative END offsets), q < side^3-1 (:283-292) */
    for (i64 q = 0; q < sgsize; q++) s.sub_grid_cells[(size_t)(sgstart + q)] = start[(size_t)q + 1];
  }
}

/* ------------------------------------------------------------------------------------------------
 * backup_r  (core/backup_r.h:31-51 ; io/backup_r.cpp:56-78)
 * ---------------------------------------------------------------------------------------------- */
static inline double restore_u32_double(uint32_t x, double o, double r) { return o + ((double)x * r) / (double)(1ull << 32); }
static inline uint32_t encode_double_u32(double x, double o, double r)
{
  double xo = x - o;
  i64 q = (i64)((xo * (double)(1ull << 32)) / r);
  uint32_t a = (uint32_t)clampi(q, 0, (i64)std::numeric_limits<uint32_t>::max());
  uint32_t b = a - 1, c = a + 1;
  double ea = std::fabs(restore_u32_double(a, o, r) - x), eb = std::fabs(restore_u32_double(b, o, r) - x), ec = std::fabs(restore_u32_double(c, o, r) - x);
  if (eb < ea) return b;
  if (ec < ea) return c;
  return a;
}
static void backup_r(xo_sim& s)
{
  Grid& g = s.grid;
  const i64 gl = g.ghost_layers();
  s.backup.assign((size_t)g.n_cells(), {});
# pragma omp parallel for collapse(3) schedule(dynamic)
  for (i64 k = gl; k < g.dims.k - gl; k++) for (i64 j = gl; j < g.dims.j - gl; j++) for (i64 i = gl; i < g.dims.i - gl; i++)
  {
    IJK loc{i, j, k};
    const i64 ci = ijk_to_index(g.dims, loc);
    const Cell& c = g.cells[(size_t)ci];
    const V3 o = g.cell_position(loc);
    std::vector<uint32_t>& rb = s.backup[(size_t)ci];
    rb.resize(c.size() * 3);
    for (size_t p = 0; p < c.size(); p++)
    {
      rb[p * 3 + 0] = encode_double_u32(c.rx[p], o.x, g.cell_size);
      rb[p * 3 + 1] = encode_double_u32(c.ry[p], o.y, g.cell_size);
      rb[p * 3 + 2] = encode_double_u32(c.rz[p], o.z, g.cell_size);
    }
  }
}

/* particle_displ_over.cu:47-65,172-176 : count inner atoms with |r - restore(backup)|^2 >= threshold^2 */
static i64 displ_over(const xo_sim& s)
{
  const Grid& g = s.grid;
  const i64 gl = g.ghost_layers();
  const double thr2 = s.max_displ * s.max_displ;
  i64 count = 0;
# pragma omp parallel for collapse(3) schedule(dynamic) reduction(+ : count)
  for (i64 k = gl; k < g.dims.k - gl; k++) for (i64 j = gl; j < g.dims.j - gl; j++) for (i64 i = gl; i < g.dims.i - gl; i++)
  {
    IJK loc{i, j, k};
    const i64 ci = ijk_to_index(g.dims, loc);
    const Cell& c = g.cells[(size_t)ci];
    const V3 o = g.cell_position(loc);
    const std::vector<uint32_t>& rb = s.backup[(size_t)ci];
    for (size_t p = 0; p < c.size(); p++)
    {
      double dx = c.rx[p] - restore_u32_double(rb[p * 3 + 0], o.x, g.cell_size);
      double dy = c.ry[p] - restore_u32_double(rb[p * 3 + 1], o.y, g.cell_size);
      double dz = c.rz[p] - restore_u32_double(rb[p * 3 + 2], o.z, g.cell_size);
      if (dx * dx + dy * dy + dz * dz >= thr2) ++count;
    }
  }
  return count;
}

/* ------------------------------------------------------------------------------------------------
 * ghosts, self partner (periodic images)  mpi/update_ghosts_comm_scheme.cpp:130-145,168-196,403-481 ;
 * coordinate modifier ghosts_comm_scheme.h:61-81 ; creation/unpack update_ghost_functors.h:339-349,369-452
 * ---------------------------------------------------------------------------------------------- */
static inline double coord_shift(double x, double rmin, double rmax, uint32_t flags3)
{
  if (flags3 & 1u) return x + ((flags3 & 4u) ? 1.0 : -1.0) * (rmax - rmin);
  return x;
}
static inline V3 apply_r_modifier(const xo_sim& s, V3 r, uint32_t flags)
{ return {coord_shift(r.x, s.dmin.x, s.dmax.x, flags >> 0), coord_shift(r.y, s.dmin.y, s.dmax.y, flags >> 3), coord_shift(r.z, s.dmin.z, s.dmax.z, flags >> 6)}; }

static void ghost_comm_scheme(xo_sim& s)
{
  Grid& g = s.grid;
  const i64 gl = g.ghost_layers();
  s.sends.clear();
  /* partner block = my own inner block (single rank) */
  const IJK pstart{0, 0, 0}, pend = s.ddims;
  const IJK gstart{pstart.i - gl, pstart.j - gl, pstart.k - gl}, gend{pend.i + gl, pend.j + gl, pend.k + gl};
  const IJK gdims{gend.i - gstart.i, gend.j - gstart.j, gend.k - gstart.k};
  const AABB inner{{g.origin.x + (double)pstart.i * g.cell_size, g.origin.y + (double)pstart.j * g.cell_size, g.origin.z + (double)pstart.k * g.cell_size},
                   {g.origin.x + (double)pend.i * g.cell_size, g.origin.y + (double)pend.j * g.cell_size, g.origin.z + (double)pend.k * g.cell_size}};
  const double e = g.max_nbh_dist;
  const AABB outer{{inner.lo.x - e, inner.lo.y - e, inner.lo.z - e}, {inner.hi.x + e, inner.hi.y + e, inner.hi.z + e}};
  const int i0 = s.periodic[0] ? -1 : 0, i1 = s.periodic[0] ? 1 : 0;
  const int j0 = s.periodic[1] ? -1 : 0, j1 = s.periodic[1] ? 1 : 0;
  const int k0 = s.periodic[2] ? -1 : 0, k1 = s.periodic[2] ? 1 : 0;
  for (int sk = k0; sk <= k1; sk++) for (int sj = j0; sj <= j1; sj++) for (int si = i0; si <= i1; si++)
  {
    if (si == 0 && sj == 0 && sk == 0) continue;   /* :176 self partner only with a non-null shift */
    uint32_t flags = 0;
    if (si == -1) flags |= SHIFT_X; if (si == 1) flags |= SHIFT_X | SIDE_X;
    if (sj == -1) flags |= SHIFT_Y; if (sj == 1) flags |= SHIFT_Y | SIDE_Y;
    if (sk == -1) flags |= SHIFT_Z; if (sk == 1) flags |= SHIFT_Z | SIDE_Z;
    for (i64 k = 0; k < g.dims.k - 2 * gl; k++) for (i64 j = 0; j < g.dims.j - 2 * gl; j++) for (i64 i = 0; i < g.dims.i - 2 * gl; i++)
    {
      const IJK my{i + gl, j + gl, k + gl};
      const IJK dloc = my + g.offset;
      const IJK sh{dloc.i + si * s.ddims.i, dloc.j + sj * s.ddims.j, dloc.k + sk * s.ddims.k};
      if (!(sh.i >= gstart.i && sh.i < gend.i && sh.j >= gstart.j && sh.j < gend.j && sh.k >= gstart.k && sh.k < gend.k)) continue;
      GhostSend snd;
      snd.cell_i = ijk_to_index(g.dims, my);
      snd.partner_cell_i = ijk_to_index(gdims, sh - gstart);
      snd.flags = flags;
      const Cell& c = g.cells[(size_t)snd.cell_i];
      for (size_t p = 0; p < c.size(); p++)
      {
        V3 gr = apply_r_modifier(s, V3{c.rx[p], c.ry[p], c.rz[p]}, flags);
        if (is_inside_incl(outer, gr)) snd.particle_i.push_back((uint32_t)p);     /* :462-470 */
      }
      if (!snd.particle_i.empty()) s.sends.push_back(std::move(snd));            /* empty cells dropped :214-236 */
    }
  }
}

/* grid_update_ghosts.h:63-202 with self loop-back (:176-187).  create=true : ghost_update_all (all fields, cells resized),
   create=false : ghost_update_r (positions only, update_ghosts.cu:46,58) */
static void ghost_update(xo_sim& s, bool create)
{
  Grid& g = s.grid;
# pragma omp parallel for schedule(dynamic)
  for (size_t q = 0; q < s.sends.size(); q++)
  {
    const GhostSend& snd = s.sends[q];
    const Cell& src = g.cells[(size_t)snd.cell_i];
    Cell& dst = g.cells[(size_t)snd.partner_cell_i];
    const size_t n = snd.particle_i.size();
    if (create) { dst.clear(); dst.resize(n); }
    for (size_t t = 0; t < n; t++)
    {
      const size_t p = snd.particle_i[t];
      V3 r = apply_r_modifier(s, V3{src.rx[p], src.ry[p], src.rz[p]}, snd.flags);
      dst.rx[t] = r.x; dst.ry[t] = r.y; dst.rz[t] = r.z;
      if (create)
      {
        dst.vx[t] = src.vx[p]; dst.vy[t] = src.vy[p]; dst.vz[t] = src.vz[p];
        dst.fx[t] = src.fx[p]; dst.fy[t] = src.fy[p]; dst.fz[t] = src.fz[p];
        dst.id[t] = src.id[p]; dst.type[t] = src.type[p];
      }
    }
  }
}

static void clear_ghost_cells(xo_sim& s)
{
  Grid& g = s.grid;
  for (i64 ci = 0; ci < g.n_cells(); ci++) if (g.is_ghost_cell(index_to_ijk(g.dims, ci))) g.cells[(size_t)ci].clear();
}

/* ------------------------------------------------------------------------------------------------
 * amr_grid_pairs  (amr/lib/amr_grid_algorithm.cpp:102-218)
 * ---------------------------------------------------------------------------------------------- */
static inline unsigned unique_pair_id(unsigned a, unsigned b) { if (a > b) std::swap(a, b); return (b * (b + 1)) / 2 + a; }   /* core/particle_type_pair.h:39-54 */

static void amr_grid_pairs(xo_sim& s)
{
  const i64 n_cells = s.grid.n_cells();
  int max_res = 0;
  for (i64 c = 0; c < n_cells; c++) max_res = std::max(max_res, (int)icbrt((i64)(s.sub_grid_start[(size_t)c + 1] - s.sub_grid_start[(size_t)c]) + 1));
  PairCache& pc = s.pair_cache;
  if ((size_t)max_res <= pc.max_res && pc.cell_size == s.cell_size && pc.max_dist == s.nbh_dist) return;   /* :129-133 */
  pc.max_res = (size_t)max_res; pc.cell_size = s.cell_size; pc.max_dist = s.nbh_dist; pc.pairs.clear();
  const double cs = s.cell_size, md2 = s.nbh_dist * s.nbh_dist;
  const int L = (int)std::ceil(s.nbh_dist / cs);
  for (int rb = 1; rb <= max_res; rb++) for (int ra = 1; ra <= rb; ra++)
  {
    const double sa = cs / ra, sb = cs / rb;
    for (int ck = 0; ck <= L; ck++) for (int cj = 0; cj <= L; cj++) for (int ci = 0; ci <= L; ci++)
    {
      pc.pairs.emplace_back();
      std::vector<uint16_t>& out = pc.pairs.back().ab;
      for (int ka = 0; ka < ra; ka++) for (int ja = 0; ja < ra; ja++) for (int ia = 0; ia < ra; ia++)
        for (int kb = 0; kb < rb; kb++) for (int jb = 0; jb < rb; jb++) for (int ib = 0; ib < rb; ib++)
        {
          AABB A{{ia * sa, ja * sa, ka * sa}, {(ia + 1) * sa, (ja + 1) * sa, (ka + 1) * sa}};
          AABB B{{ci * cs + ib * sb, cj * cs + jb * sb, ck * cs + kb * sb}, {ci * cs + (ib + 1) * sb, cj * cs + (jb + 1) * sb, ck * cs + (kb + 1) * sb}};
          if (min_dist2_box_box(A, B) <= md2) { out.push_back((uint16_t)((ka << 10) | (ja << 5) | ia)); out.push_back((uint16_t)((kb << 10) | (jb << 5) | ib)); }
        }
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * chunk_neighbors  (particle_neighbors/.../chunk_neighbors_execute.h:100-413 ; codec chunk_neighbors.h:137-162 ;
 * filter neighbor_filter_func.h:36-52).  Config = update-particles.msp:13-21 : build_particle_offset, chunk_size 1.
 * ---------------------------------------------------------------------------------------------- */
static inline uint16_t encode_cell_index(IJK rel) { return (uint16_t)((((rel.k + 16) << 5) + (rel.j + 16)) << 5) + (uint16_t)(rel.i + 16); }

static void chunk_neighbors(xo_sim& s)
{
  Grid& g = s.grid;
  const IJK dims = g.dims;
  const i64 n_cells = g.n_cells();
  const double max_dist2 = s.nbh_dist * s.nbh_dist;
  const i64 gap = (i64)s.pair_cache.layers();
  const i64 side = gap + 1;
  const size_t n_nbh_cell = s.pair_cache.n_nbh_cells();
  s.streams.assign((size_t)n_cells, {});
  unsigned gmax = 0;
# pragma omp parallel
  {
    std::vector<std::vector<std::pair<uint16_t, uint16_t>>> pn;
    std::vector<uint16_t> cc;
    unsigned tmax = 0;
#   pragma omp for collapse(3) schedule(dynamic)
    for (i64 ak = 0; ak < dims.k; ak++) for (i64 aj = 0; aj < dims.j; aj++) for (i64 ai = 0; ai < dims.i; ai++)
    {
      const IJK la{ai, aj, ak};
      const i64 cell_a = ijk_to_index(dims, la);
      const Cell& A = g.cells[(size_t)cell_a];
      const size_t na = A.size();
      pn.resize(na);
      for (size_t p = 0; p < na; p++) pn[p].clear();
      const i64 sgstart_a = (i64)s.sub_grid_start[(size_t)cell_a];
      const i64 sgsize_a = (i64)s.sub_grid_start[(size_t)cell_a + 1] - sgstart_a;
      const i64 side_a = icbrt(sgsize_a + 1);
      /* :134-143 neighbour-cell box clamped to the local grid, no wrap */
      for (i64 bk = std::max<i64>(ak - gap, 0); bk <= std::min<i64>(ak + gap, dims.k - 1); bk++)
      for (i64 bj = std::max<i64>(aj - gap, 0); bj <= std::min<i64>(aj + gap, dims.j - 1); bj++)
      for (i64 bi = std::max<i64>(ai - gap, 0); bi <= std::min<i64>(ai + gap, dims.i - 1); bi++)
      {
        const IJK lb{bi, bj, bk};
        const i64 cell_b = ijk_to_index(dims, lb);
        const bool ghost_b = g.is_ghost_cell(lb);
        const Cell& B = g.cells[(size_t)cell_b];
        const size_t nb = B.size();
        const i64 sgstart_b = (i64)s.sub_grid_start[(size_t)cell_b];
        const i64 sgsize_b = (i64)s.sub_grid_start[(size_t)cell_b + 1] - sgstart_b;
        const i64 side_b = icbrt(sgsize_b + 1);
        const IJK rloc = lb - la;
        const uint16_t enc = encode_cell_index(rloc);
        /* :160-182 : locate the cached sub-cell pair list (resolution-ordered, axis-reflected) */
        const size_t res_pair = unique_pair_id((unsigned)(side_a - 1), (unsigned)(side_b - 1));
        IJK rl = rloc;
        const bool rev_ab = side_a > side_b;
        if (rev_ab) rl = {-rloc.i, -rloc.j, -rloc.k};
        const bool rev_i = rl.i < 0, rev_j = rl.j < 0, rev_k = rl.k < 0;
        if (rev_i) rl.i = -rl.i; if (rev_j) rl.j = -rl.j; if (rev_k) rl.k = -rl.k;
        const size_t block = (size_t)(rl.k * side * side + rl.j * side + rl.i);
        const std::vector<uint16_t>& pab = s.pair_cache.pairs[res_pair * n_nbh_cell + block].ab;
        const size_t n_pairs = pab.size() / 2;
        const unsigned IA = rev_ab ? 1u : 0u, IB = rev_ab ? 0u : 1u;
        for (size_t sp = 0; sp < n_pairs; sp++)
        {
          uint16_t ax = pab[sp * 2 + IA], bx = pab[sp * 2 + IB];
          i64 sai = ax & 31, saj = (ax >> 5) & 31, sak = ax >> 10;
          i64 sbi = bx & 31, sbj = (bx >> 5) & 31, sbk = bx >> 10;
          if (rev_i) { sai = side_a - 1 - sai; sbi = side_b - 1 - sbi; }
          if (rev_j) { saj = side_a - 1 - saj; sbj = side_b - 1 - sbj; }
          if (rev_k) { sak = side_a - 1 - sak; sbk = side_b - 1 - sbk; }
          const i64 sga = (sak * side_a + saj) * side_a + sai;
          const i64 sgb = (sbk * side_b + sbj) * side_b + sbi;
          unsigned pa0 = 0, pa1 = (unsigned)na, pb0 = 0, pb1 = (unsigned)nb;
          if (sga > 0) pa0 = s.sub_grid_cells[(size_t)(sgstart_a + sga - 1)];
          if (sga < sgsize_a) pa1 = s.sub_grid_cells[(size_t)(sgstart_a + sga)];
          if (sgb > 0) pb0 = s.sub_grid_cells[(size_t)(sgstart_b + sgb - 1)];
          if (sgb < sgsize_b) pb1 = s.sub_grid_cells[(size_t)(sgstart_b + sgb)];
          if (pb1 <= pb0) continue;
          for (unsigned pa = pa0; pa < pa1; pa++)
            for (unsigned pb = pb0; pb < pb1; pb++)
            {
              /* :225-227 : dr = r_a - r_b ; identity LinearXForm is exact ; d2 = x*x + y*y + z*z ; filter d2 > 0 && d2 <= max_dist2 */
              const double dx = A.rx[pa] - B.rx[pb], dy = A.ry[pa] - B.ry[pb], dz = A.rz[pa] - B.rz[pb];
              const double d2 = dx * dx + dy * dy + dz * dz;
              if ((cell_a != cell_b || pa != pb) && d2 > 0.0 && d2 <= max_dist2)
              {
                /* NeighborFilterHalfSymGhost (neighbor_filter_func.h:36-52) */
                if (s.nbh_half_symmetric && (cell_a < cell_b || (cell_a == cell_b && pa < pb))) continue;
                if (s.nbh_skip_ghosts && ghost_b) continue;
                pn[pa].push_back({enc, (uint16_t)pb});
              }
            }
        }
      }
      /* :265-398 : encode  [ (N+1) x u32 offsets ]  { n_groups { enc n p_b... } } */
      cc.clear();
      const size_t off_sz = (na + 1) * 2;
      if (na > 0) cc.assign(off_sz, 0);
      for (size_t pa = 0; pa < na; pa++)
      {
        uint32_t off = (uint32_t)(cc.size() - off_sz + 1);
        cc[pa * 2] = (uint16_t)off; cc[pa * 2 + 1] = (uint16_t)(off >> 16);
        auto& v = pn[pa];
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        tmax = std::max(tmax, (unsigned)v.size());
        const size_t cnt_idx = cc.size();
        cc.push_back(0);
        size_t chunk_idx = 0; uint16_t last = 0;
        for (const auto& nb : v)
        {
          if (nb.first != last) { ++cc[cnt_idx]; last = nb.first; cc.push_back(last); chunk_idx = cc.size(); cc.push_back(0); }
          ++cc[chunk_idx];
          cc.push_back(nb.second);
        }
      }
      if (na > 0) { uint32_t off = (uint32_t)(cc.size() - off_sz + 1); cc[na * 2] = (uint16_t)off; cc[na * 2 + 1] = (uint16_t)(off >> 16); }
      s.streams[(size_t)cell_a] = cc;
    }
#   pragma omp critical
    gmax = std::max(gmax, tmax);
  }
  s.max_neighbors = gmax;
}

/* chunk_neighbors.h:164-186 */
struct StreamInfo { const uint16_t* stream; const uint32_t* offset; int shift; };
static inline StreamInfo stream_info(const std::vector<uint16_t>& st, size_t n)
{
  if (st.empty() || n == 0) return {nullptr, nullptr, 0};
  if (st[0] <= 2 && st[1] == 0)
  {
    const int t = st[0];
    const size_t tsz = (t >= 1) ? ((n * (size_t)t + 1) * 2) : 0;
    return {st.data() + tsz, (t >= 1) ? reinterpret_cast<const uint32_t*>(st.data()) : nullptr, -t};
  }
  return {st.data(), nullptr, 0};
}

/* walk the stream of (cell_a,p_a); F(cell_b, p_b).  Decoding per compute_cell_particle_pairs_impl_default.h:141-179 */
template <class F>
static inline void for_each_listed(const xo_sim& s, i64 cell_a, size_t pa, const StreamInfo& si, F&& f)
{
  const IJK dims = s.grid.dims;
  const uint16_t* st = si.stream + si.offset[pa] + si.shift;
  int groups = *st++;
  for (; groups > 0; --groups)
  {
    const uint16_t enc = *st++;
    int n = *st++;
    const int ri = (int)(enc & 31) - 16, rj = (int)((enc >> 5) & 31) - 16, rk = (int)((enc >> 10) & 31) - 16;
    const i64 cell_b = cell_a + ((i64)rk * dims.j + rj) * dims.i + ri;
    for (; n > 0; --n) f(cell_b, (size_t)*st++);
  }
}

/* ------------------------------------------------------------------------------------------------
 * compute_all_forces_energy of the LJ deck (input_lj_Ni.msp:88-92):
 *   zero_particle_force{ghost:true} (zero_particle_force.cu:15-46)
 *   lennard_jones_force : compute_cell_particle_pairs, CPU = compute-buffer mode (compute_cell_particle_pairs.h:60-65,
 *     impl_default.h:148-239, buffer compute_pair_buffer.h:39-68), functor lennard_jones.cu:46-56,68-102
 *   update_force_from_ghost : value-wise no-op for a full (non symmetric) list
 *   divide_force_by_type_scalar: mass (vec3_typescalar_op.cu:118-122, math_functors.h:136-141)
 * ---------------------------------------------------------------------------------------------- */
static int compute_force(xo_sim& s, double* epot_out, double* vir_out)
{
  Grid& g = s.grid;
  const i64 gl = g.ghost_layers();
  const double rcut2 = s.cfg.rcut * s.cfg.rcut;
  const double eps = s.cfg.epsilon, sig = s.cfg.sigma, mass = s.cfg.mass;
  const int maxn = s.cfg.max_neighbors;
  for (Cell& c : g.cells) { std::fill(c.fx.begin(), c.fx.end(), 0.); std::fill(c.fy.begin(), c.fy.end(), 0.); std::fill(c.fz.begin(), c.fz.end(), 0.); }
  int overflow = 0;
  const bool want_ev = epot_out != nullptr;
  std::vector<double> cell_e, cell_w;
  if (want_ev) { cell_e.assign((size_t)g.n_cells(), 0.); cell_w.assign((size_t)g.n_cells() * 6, 0.); }
# pragma omp parallel
  {
    std::vector<double> bx((size_t)maxn), by((size_t)maxn), bz((size_t)maxn), bd2((size_t)maxn);
#   pragma omp for collapse(3) schedule(dynamic)
    for (i64 k = gl; k < g.dims.k - gl; k++) for (i64 j = gl; j < g.dims.j - gl; j++) for (i64 i = gl; i < g.dims.i - gl; i++)
    {
      const i64 cell_a = ijk_to_index(g.dims, IJK{i, j, k});
      Cell& A = g.cells[(size_t)cell_a];
      const size_t na = A.size();
      const StreamInfo si = stream_info(s.streams[(size_t)cell_a], na);
      double ce = 0., cw[6] = {0, 0, 0, 0, 0, 0};
      for (size_t pa = 0; pa < na; pa++)
      {
        int cnt = 0;
        const double xa = A.rx[pa], ya = A.ry[pa], za = A.rz[pa];
        for_each_listed(s, cell_a, pa, si, [&](i64 cell_b, size_t pb) {
          const Cell& B = g.cells[(size_t)cell_b];
          const double dx = B.rx[pb] - xa, dy = B.ry[pb] - ya, dz = B.rz[pb] - za;   /* impl_default.h:183 dr = r_b - r_a */
          const double d2 = dx * dx + dy * dy + dz * dz;
          if (d2 > 0.0 && d2 <= rcut2)
          {
            if (cnt >= maxn) { overflow = 1; return; }
            bx[(size_t)cnt] = dx; by[(size_t)cnt] = dy; bz[(size_t)cnt] = dz; bd2[(size_t)cnt] = d2; ++cnt;
          }
        });
        double tfx = 0., tfy = 0., tfz = 0.;
        for (int q = 0; q < cnt; q++)
        {
          const double r = std::sqrt(bd2[(size_t)q]);
          const double inv_r = 1.0 / r;
          const double ratio = sig * inv_r;
          const double ratio2 = ratio * ratio;
          const double ratio6 = ratio2 * ratio2 * ratio2;
          const double ratio12 = ratio6 * ratio6;
          const double e = 4. * eps * (ratio12 - ratio6);
          double de = (-24. * eps * (2. * ratio12 - ratio6)) * inv_r;
          de *= 1.0 / r;
          tfx += de * bx[(size_t)q]; tfy += de * by[(size_t)q]; tfz += de * bz[(size_t)q];
          if (want_ev)
          {
            ce += 0.5 * e;
            const double fx = de * bx[(size_t)q], fy = de * by[(size_t)q], fz = de * bz[(size_t)q];
            cw[0] -= 0.5 * bx[(size_t)q] * fx; cw[1] -= 0.5 * by[(size_t)q] * fy; cw[2] -= 0.5 * bz[(size_t)q] * fz;
            cw[3] -= 0.5 * bx[(size_t)q] * fy; cw[4] -= 0.5 * bx[(size_t)q] * fz; cw[5] -= 0.5 * by[(size_t)q] * fz;
          }
        }
        if (cnt > 0) { A.fx[pa] += tfx; A.fy[pa] += tfy; A.fz[pa] += tfz; }
        A.fx[pa] /= mass; A.fy[pa] /= mass; A.fz[pa] /= mass;
      }
      if (want_ev) { cell_e[(size_t)cell_a] = ce; for (int q = 0; q < 6; q++) cell_w[(size_t)cell_a * 6 + q] = cw[q]; }
    }
  }
  if (overflow) { g_err = "compute buffer overflow: raise max_neighbors (reference default 256, unchecked in release: compute_pair_buffer.h:199-208)"; return 1; }
  if (want_ev)
  {
    /* fixed cell order + Neumaier compensation: thread-count independent */
    auto ksum = [](const double* v, size_t n, size_t stride) { double sum = 0., c = 0.; for (size_t q = 0; q < n; q++) { double x = v[q * stride]; double t = sum + x; c += (std::fabs(sum) >= std::fabs(x)) ? (sum - t) + x : (x - t) + sum; sum = t; } return sum + c; };
    *epot_out = ksum(cell_e.data(), cell_e.size(), 1);
    for (int q = 0; q < 6; q++) vir_out[q] = ksum(cell_w.data() + q, cell_e.size(), 6);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * gravitational_force (contribs/pi/gravitational_force.cu:161-217): compute_cell_particle_pairs over the inner cells with the
 * functor of :48-85 (buffer-less form; CentralParticleFieldSet = type, fx, fy, fz; the neighbour's type is read through
 * cells[cell_b][type][p_b] :76, masses through the type property 'mass' :75,77):
 *   r = sqrt(d2); inv_r = 1/r; de = G ma mb inv_r inv_r; de *= w / r; f += de * dr          (:50-55, :78-84)
 * ADDS to the force fields (no zeroing, no division by mass: neither is part of the operator).  Not pinned by any reference test.
 * ---------------------------------------------------------------------------------------------- */
static int gravitational_force(xo_sim& s, double G, double rcut, const double* type_mass, int n_types)
{
  Grid& g = s.grid;
  const i64 gl = g.ghost_layers();
  const double rcut2 = rcut * rcut;
  int bad_type = 0;
# pragma omp parallel for collapse(3) schedule(dynamic)
  for (i64 k = gl; k < g.dims.k - gl; k++) for (i64 j = gl; j < g.dims.j - gl; j++) for (i64 i = gl; i < g.dims.i - gl; i++)
  {
    const i64 cell_a = ijk_to_index(g.dims, IJK{i, j, k});
    Cell& A = g.cells[(size_t)cell_a];
    const size_t na = A.size();
    const StreamInfo si = stream_info(s.streams[(size_t)cell_a], na);
    for (size_t pa = 0; pa < na; pa++)
    {
      const double xa = A.rx[pa], ya = A.ry[pa], za = A.rz[pa];
      const int type_a = (int)A.type[pa];
      if (type_a >= n_types) { bad_type = 1; continue; }
      const double mass_a = type_mass[type_a];
      double fx = A.fx[pa], fy = A.fy[pa], fz = A.fz[pa];
      for_each_listed(s, cell_a, pa, si, [&](i64 cell_b, size_t pb) {
        const Cell& B = g.cells[(size_t)cell_b];
        const double dx = B.rx[pb] - xa, dy = B.ry[pb] - ya, dz = B.rz[pb] - za;
        const double d2 = dx * dx + dy * dy + dz * dz;
        if (d2 > 0.0 && d2 <= rcut2)
        {
          const int type_b = (int)B.type[pb];
          if (type_b >= n_types) { bad_type = 1; return; }
          const double mass_b = type_mass[type_b];
          const double r = std::sqrt(d2);
          const double inv_r = 1.0 / r;
          double de = G * mass_a * mass_b * inv_r * inv_r;
          de *= 1.0 / r;
          fx += de * dx; fy += de * dy; fz += de * dz;
        }
      });
      A.fx[pa] = fx; A.fy[pa] = fy; A.fz[pa] = fz;
    }
  }
  if (bad_type) { g_err = "gravitational_force: particle type without a mass"; return 1; }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * average_neighbors_scalar (src/compute/average_neighbors.cu:38-100 functor, :123-175 operator): compute_cell_particle_pairs over
 * the inner cells with a functor that has a PARTICLE CONTEXT: Start resets (m_sum, m_weight_sum) (:41-45), every neighbour with
 * d2 <= rcut^2 adds w * nbh_field[b] and w, w = a0 + a1 d + a2 d^2 + a3 d^3 (:85-99), Stop writes sum / weight_sum (:47-51,73-77).
 * nbh_field: 0..8 = rx ry rz vx vy vz fx fy fz, 9 = id, 10 = type.  out = all particles in cell order, ghost cells untouched (0).
 * Parity unpinned by the reference's tests (it ships none for this operator); pinned here by a brute-force all-pairs evaluation
 * in tests/test_gpu_parity.py.
 * ------------------------------------------------------------------------------------------------ */
static int average_neighbors(xo_sim& s, double rcut, const double wf[4], int nbh_field, double* out)
{
  Grid& g = s.grid;
  if (nbh_field < 0 || nbh_field > 10) { g_err = "average_neighbors: unknown field"; return 1; }
  const i64 gl = g.ghost_layers();
  const double rcut2 = rcut * rcut;
  std::vector<size_t> first(g.cells.size() + 1, 0);
  for (size_t c = 0; c < g.cells.size(); c++) first[c + 1] = first[c] + g.cells[c].size();
  std::fill(out, out + first.back(), 0.0);
  auto value = [&](const Cell& B, size_t p) -> double {
    switch (nbh_field) {
      case 0: return B.rx[p]; case 1: return B.ry[p]; case 2: return B.rz[p];
      case 3: return B.vx[p]; case 4: return B.vy[p]; case 5: return B.vz[p];
      case 6: return B.fx[p]; case 7: return B.fy[p]; case 8: return B.fz[p];
      case 9: return (double)B.id[p]; default: return (double)B.type[p];
    }
  };
# pragma omp parallel for collapse(3) schedule(dynamic)
  for (i64 k = gl; k < g.dims.k - gl; k++) for (i64 j = gl; j < g.dims.j - gl; j++) for (i64 i = gl; i < g.dims.i - gl; i++)
  {
    const i64 cell_a = ijk_to_index(g.dims, IJK{i, j, k});
    const Cell& A = g.cells[(size_t)cell_a];
    const size_t na = A.size();
    const StreamInfo si = stream_info(s.streams[(size_t)cell_a], na);
    for (size_t pa = 0; pa < na; pa++)
    {
      const double xa = A.rx[pa], ya = A.ry[pa], za = A.rz[pa];
      double sum = 0.0, wsum = 0.0;
      for_each_listed(s, cell_a, pa, si, [&](i64 cell_b, size_t pb) {
        const Cell& B = g.cells[(size_t)cell_b];
        const double dx = B.rx[pb] - xa, dy = B.ry[pb] - ya, dz = B.rz[pb] - za;
        const double d2 = dx * dx + dy * dy + dz * dz;
        if (d2 > 0.0 && d2 <= rcut2)
        {
          double w = wf[0] + wf[2] * d2;
          if (wf[1] != 0.0 || wf[3] != 0.0) { const double d = std::sqrt(d2); w += wf[1] * d + wf[3] * d2 * d; }
          sum += w * value(B, pb);
          wsum += w;
        }
      });
      out[first[(size_t)cell_a] + pa] = (wsum > 0.0) ? sum / wsum : sum;
    }
  }
  return 0;
}

/* zero_particle_force{ghost:true} alone (zero_particle_force.cu:15-46) */
static void zero_force(xo_sim& s)
{
  for (Cell& c : s.grid.cells) { std::fill(c.fx.begin(), c.fx.end(), 0.); std::fill(c.fy.begin(), c.fy.end(), 0.); std::fill(c.fz.begin(), c.fz.end(), 0.); }
}

/* ------------------------------------------------------------------------------------------------
 * Symmetric (Newton-3) pair sweep over half_symmetric lists -- SURVEY 8(f) rank 2.  The reference has no symmetric LJ
 * operator; this restates what its machinery does for a symmetric pair functor:
 *   zero_particle_force{ghost:true}
 *   compute_cell_particle_pairs<Symmetric=true> over the inner cells (impl_default.h:143-239; :181 stops a particle's walk
 *     at the first entry that lies "after" it -- with half_symmetric lists there is none), functor of lennard_jones.cu:46-56
 *     applied once per pair: f_a += de*dr, f_b -= de*dr (b may be a ghost; ComputePairOptionalLocks<true> serialises the
 *     concurrent updates of a cell, compute_pair_optional_args.h:152-161 -- here the sweep is sequential instead)
 *   update_force_from_ghost (UpdateFromGhosts<fx,fy,fz, UpdateValueAdd>, mpi/update_force_from_ghost.cu:44,
 *     update_from_ghost_functors.h:36-120): every ghost's force is ADDED to the particle it is an image of
 *   divide_force_by_type_scalar: mass
 * Parity unpinned by the reference's tests (like energy / virial); pinned here by: forces equal those of the full-list
 * sweep to rounding (tests/test_oracle_properties.py).
 * ---------------------------------------------------------------------------------------------- */
static void update_force_from_ghost(xo_sim& s)
{
  Grid& g = s.grid;
  for (size_t q = 0; q < s.sends.size(); q++)
  {
    const GhostSend& snd = s.sends[q];
    Cell& src = g.cells[(size_t)snd.cell_i];
    const Cell& gh = g.cells[(size_t)snd.partner_cell_i];
    for (size_t t = 0; t < snd.particle_i.size(); t++)
    {
      const size_t p = snd.particle_i[t];
      src.fx[p] += gh.fx[t]; src.fy[p] += gh.fy[t]; src.fz[p] += gh.fz[t];
    }
  }
}

static int compute_force_symmetric(xo_sim& s)
{
  Grid& g = s.grid;
  const i64 gl = g.ghost_layers();
  const double rcut2 = s.cfg.rcut * s.cfg.rcut;
  const double eps = s.cfg.epsilon, sig = s.cfg.sigma, mass = s.cfg.mass;
  if (!s.nbh_half_symmetric) { g_err = "compute_force_symmetric needs half_symmetric lists"; return 1; }
  for (Cell& c : g.cells) { std::fill(c.fx.begin(), c.fx.end(), 0.); std::fill(c.fy.begin(), c.fy.end(), 0.); std::fill(c.fz.begin(), c.fz.end(), 0.); }
  for (i64 k = gl; k < g.dims.k - gl; k++) for (i64 j = gl; j < g.dims.j - gl; j++) for (i64 i = gl; i < g.dims.i - gl; i++)
  {
    const i64 cell_a = ijk_to_index(g.dims, IJK{i, j, k});
    Cell& A = g.cells[(size_t)cell_a];
    const size_t na = A.size();
    const StreamInfo si = stream_info(s.streams[(size_t)cell_a], na);
    for (size_t pa = 0; pa < na; pa++)
    {
      const double xa = A.rx[pa], ya = A.ry[pa], za = A.rz[pa];
      bool stop = false;
      for_each_listed(s, cell_a, pa, si, [&](i64 cell_b, size_t pb) {
        if (stop) return;
        if (cell_b > cell_a || (cell_b == cell_a && pb > pa)) { stop = true; return; }      /* impl_default.h:181 */
        Cell& B = g.cells[(size_t)cell_b];
        const double dx = B.rx[pb] - xa, dy = B.ry[pb] - ya, dz = B.rz[pb] - za;
        const double d2 = dx * dx + dy * dy + dz * dz;
        if (d2 > 0.0 && d2 <= rcut2)
        {
          const double r = std::sqrt(d2);
          const double inv_r = 1.0 / r;
          const double ratio = sig * inv_r;
          const double ratio2 = ratio * ratio;
          const double ratio6 = ratio2 * ratio2 * ratio2;
          const double ratio12 = ratio6 * ratio6;
          double de = (-24. * eps * (2. * ratio12 - ratio6)) * inv_r;
          de *= 1.0 / r;
          const double fx = de * dx, fy = de * dy, fz = de * dz;
          A.fx[pa] += fx; A.fy[pa] += fy; A.fz[pa] += fz;
          B.fx[pb] -= fx; B.fy[pb] -= fy; B.fz[pb] -= fz;
        }
      });
    }
  }
  update_force_from_ghost(s);
  for (i64 k = gl; k < g.dims.k - gl; k++) for (i64 j = gl; j < g.dims.j - gl; j++) for (i64 i = gl; i < g.dims.i - gl; i++)
  {
    Cell& A = g.cells[(size_t)ijk_to_index(g.dims, IJK{i, j, k})];
    for (size_t pa = 0; pa < A.size(); pa++) { A.fx[pa] /= mass; A.fy[pa] /= mass; A.fz[pa] /= mass; }
  }
  return 0;
}

/* defbox/push_vec3_2nd_order.h:29-39,85-88 via compute_cell_particles (inner cells) */
static void push_f_v_r(xo_sim& s)
{
  Grid& g = s.grid; const i64 gl = g.ghost_layers();
  const double dt = s.cfg.dt * 1.0, dt2 = dt * dt * 0.5;
# pragma omp parallel for collapse(3) schedule(dynamic)
  for (i64 k = gl; k < g.dims.k - gl; k++) for (i64 j = gl; j < g.dims.j - gl; j++) for (i64 i = gl; i < g.dims.i - gl; i++)
  {
    Cell& c = g.cells[(size_t)ijk_to_index(g.dims, IJK{i, j, k})];
    for (size_t p = 0; p < c.size(); p++)
    { c.rx[p] += c.vx[p] * dt + c.fx[p] * dt2; c.ry[p] += c.vy[p] * dt + c.fy[p] * dt2; c.rz[p] += c.vz[p] * dt + c.fz[p] * dt2; }
  }
}
/* defbox/push_vec3_1st_order.h:29-38,79-81 */
static void push_f_v(xo_sim& s, double dt_scale)
{
  Grid& g = s.grid; const i64 gl = g.ghost_layers();
  const double dt = s.cfg.dt * dt_scale;
# pragma omp parallel for collapse(3) schedule(dynamic)
  for (i64 k = gl; k < g.dims.k - gl; k++) for (i64 j = gl; j < g.dims.j - gl; j++) for (i64 i = gl; i < g.dims.i - gl; i++)
  {
    Cell& c = g.cells[(size_t)ijk_to_index(g.dims, IJK{i, j, k})];
    for (size_t p = 0; p < c.size(); p++) { c.vx[p] += c.fx[p] * dt; c.vy[p] += c.fy[p] * dt; c.vz[p] += c.fz[p] * dt; }
  }
}

/* update-particles.msp:47-53 parallel_update_particles (+ :42-45 update_particle_neighbors) */
static int update_particles_full(xo_sim& s)
{
  clear_ghost_cells(s);            /* migrate_cell_particles leaves ghost cells empty */
  if (s.grid.max_nbh_dist != s.ghost_dist) migrate_single_rank(s);
  rebuild_amr(s);
  backup_r(s);
  ghost_comm_scheme(s);
  ghost_update(s, true);
  amr_grid_pairs(s);
  chunk_neighbors(s);
  s.rebuilds++;
  return 0;
}

} // namespace

/* =================================================================================================
 * C interface
 * ================================================================================================= */
extern "C" {

const char* xo_last_error(void) { return g_err.c_str(); }
int xo_num_threads(void) { return omp_get_max_threads(); }
void xo_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

xo_sim* xo_create(const xo_config* cfg)
{
  xo_sim* s = new xo_sim{};
  s->cfg = *cfg;
  s->dmin = {cfg->bounds_min[0], cfg->bounds_min[1], cfg->bounds_min[2]};
  s->dmax = {cfg->bounds_max[0], cfg->bounds_max[1], cfg->bounds_max[2]};
  s->ddims = {cfg->grid_dims[0], cfg->grid_dims[1], cfg->grid_dims[2]};
  s->cell_size = cfg->cell_size;
  for (int d = 0; d < 3; d++) s->periodic[d] = cfg->periodic[d] != 0;
  if (s->cfg.sub_grid_density <= 0.) s->cfg.sub_grid_density = 6.5;
  if (s->cfg.max_neighbors <= 0) s->cfg.max_neighbors = 256;
  if (s->cfg.mass <= 0.) s->cfg.mass = 1.0;
  /* nbh_dist.cpp:47-71 (identity xform: scale 1) ; rcut_max = max(rcut, rcut_max) lennard_jones.cu:193 */
  s->nbh_dist = cfg->rcut + cfg->rcut_inc;
  s->max_displ = cfg->rcut_inc / 2.0;
  s->ghost_dist = cfg->rcut + cfg->rcut_inc;
  init_rcb_grid(*s);
  return s;
}
void xo_destroy(xo_sim* s) { delete s; }

/* input_data of the deck: lattice + gaussian_noise_r (+ synthetic velocities) */
int xo_generate(xo_sim* s)
{
  lattice_fcc(*s);
  if (s->cfg.noise_sigma > 0.) gaussian_noise(*s, s->cfg.noise_sigma, false, 0);
  if (s->cfg.vel_sigma > 0.) { gaussian_noise(*s, s->cfg.vel_sigma, true, 1); zero_momentum(*s); }
  return 0;
}

/* init_particles (update-particles.msp:55-60) then first force (compute-loop.msp:1-7) */
int xo_first_iteration(xo_sim* s)
{
  if (move_particles(*s)) return 1;
  if (update_particles_full(*s)) return 1;
  return compute_force(*s, nullptr, nullptr);
}

int xo_init(xo_sim* s)
{
  if (xo_generate(s)) return 1;
  return xo_first_iteration(s);
}

int xo_run(xo_sim* s, int nsteps)
{
  int rebuilds = 0;
  for (int it = 0; it < nsteps; it++)
  {
    /* numerical-scheme.msp:13-25 */
    push_f_v_r(*s);
    push_f_v(*s, 0.5);
    /* check_and_update_particles (update-particles.msp:75-78) */
    if (displ_over(*s) > 0) { if (move_particles(*s)) return -1; if (update_particles_full(*s)) return -1; rebuilds++; }
    else ghost_update(*s, false);
    if (compute_force(*s, nullptr, nullptr)) return -1;
    push_f_v(*s, 0.5);
  }
  return rebuilds;
}

int xo_move_particles(xo_sim* s) { return move_particles(*s); }
int xo_update_particles_full(xo_sim* s) { return update_particles_full(*s); }
int xo_ghost_update_r(xo_sim* s) { ghost_update(*s, false); return 0; }
int xo_build_neighbors(xo_sim* s) { amr_grid_pairs(*s); chunk_neighbors(*s); return 0; }
int xo_compute_force(xo_sim* s) { return compute_force(*s, nullptr, nullptr); }
int xo_compute_force_symmetric(xo_sim* s) { return compute_force_symmetric(*s); }
int xo_zero_force(xo_sim* s) { zero_force(*s); return 0; }
int64_t xo_amr_pair_cache(const xo_sim* s, int64_t* max_res, uint64_t* list_offsets, uint16_t* pairs)
{
  const PairCache& pc = s->pair_cache;
  if (max_res) *max_res = (int64_t)pc.max_res;
  int64_t total = 0; size_t q = 0;
  for (const SubCellPairs& l : pc.pairs)
  {
    if (list_offsets) list_offsets[q] = (uint64_t)total;
    if (pairs) for (size_t t = 0; t < l.ab.size(); t++) pairs[(size_t)total + t] = l.ab[t];
    total += (int64_t)l.ab.size(); q++;
  }
  if (list_offsets) list_offsets[q] = (uint64_t)total;
  return total;
}
int xo_average_neighbors(xo_sim* s, double rcut, const double wf[4], int nbh_field, double* out) { return average_neighbors(*s, rcut, wf, nbh_field, out); }
int xo_gravitational_force(xo_sim* s, double G, double rcut, const double* type_mass, int n_types) { return gravitational_force(*s, G, rcut, type_mass, n_types); }
void xo_set_nbh_config(xo_sim* s, int half_symmetric, int skip_ghosts) { s->nbh_half_symmetric = half_symmetric != 0; s->nbh_skip_ghosts = skip_ghosts != 0; }
int xo_push_f_v_r(xo_sim* s) { push_f_v_r(*s); return 0; }
int xo_push_f_v(xo_sim* s, double sc) { push_f_v(*s, sc); return 0; }
int64_t xo_displ_over(xo_sim* s) { return displ_over(*s); }
int64_t xo_rebuild_count(const xo_sim* s) { return s->rebuilds; }

void xo_grid_info(const xo_sim* s, int64_t dims[3], int64_t offset[3], int64_t* gl, int64_t* n_cells)
{
  dims[0] = s->grid.dims.i; dims[1] = s->grid.dims.j; dims[2] = s->grid.dims.k;
  offset[0] = s->grid.offset.i; offset[1] = s->grid.offset.j; offset[2] = s->grid.offset.k;
  *gl = s->grid.ghost_layers(); *n_cells = s->grid.n_cells();
}
int64_t xo_total_particles(const xo_sim* s) { int64_t n = 0; for (const Cell& c : s->grid.cells) n += (int64_t)c.size(); return n; }
int64_t xo_inner_particles(const xo_sim* s)
{
  int64_t n = 0; const Grid& g = s->grid;
  for (i64 c = 0; c < g.n_cells(); c++) if (!g.is_ghost_cell(index_to_ijk(g.dims, c))) n += (int64_t)g.cells[(size_t)c].size();
  return n;
}
void xo_cell_counts(const xo_sim* s, int32_t* counts) { for (size_t c = 0; c < s->grid.cells.size(); c++) counts[c] = (int32_t)s->grid.cells[c].size(); }

void xo_get_particles(const xo_sim* s, double* rx, double* ry, double* rz, double* vx, double* vy, double* vz,
                      double* fx, double* fy, double* fz, uint64_t* id, uint8_t* type)
{
  size_t o = 0;
  for (const Cell& c : s->grid.cells)
  {
    const size_t n = c.size();
    if (n == 0) continue;
    if (rx) memcpy(rx + o, c.rx.data(), n * 8); if (ry) memcpy(ry + o, c.ry.data(), n * 8); if (rz) memcpy(rz + o, c.rz.data(), n * 8);
    if (vx) memcpy(vx + o, c.vx.data(), n * 8); if (vy) memcpy(vy + o, c.vy.data(), n * 8); if (vz) memcpy(vz + o, c.vz.data(), n * 8);
    if (fx) memcpy(fx + o, c.fx.data(), n * 8); if (fy) memcpy(fy + o, c.fy.data(), n * 8); if (fz) memcpy(fz + o, c.fz.data(), n * 8);
    if (id) memcpy(id + o, c.id.data(), n * 8); if (type) memcpy(type + o, c.type.data(), n);
    o += n;
  }
}

int xo_set_particles(xo_sim* s, const int32_t* counts, const double* rx, const double* ry, const double* rz,
                     const double* vx, const double* vy, const double* vz, const double* fx, const double* fy, const double* fz,
                     const uint64_t* id, const uint8_t* type)
{
  size_t o = 0;
  for (size_t ci = 0; ci < s->grid.cells.size(); ci++)
  {
    Cell& c = s->grid.cells[ci];
    const size_t n = (size_t)counts[ci];
    c.clear(); c.resize(n);
    for (size_t p = 0; p < n; p++)
    {
      c.rx[p] = rx[o + p]; c.ry[p] = ry[o + p]; c.rz[p] = rz[o + p];
      c.vx[p] = vx ? vx[o + p] : 0.; c.vy[p] = vy ? vy[o + p] : 0.; c.vz[p] = vz ? vz[o + p] : 0.;
      c.fx[p] = fx ? fx[o + p] : 0.; c.fy[p] = fy ? fy[o + p] : 0.; c.fz[p] = fz ? fz[o + p] : 0.;
      c.id[p] = id ? id[o + p] : 0; c.type[p] = type ? type[o + p] : 0;
    }
    o += n;
  }
  return 0;
}

/* recompute AMR offset tables for the CURRENT in-cell order without moving particles (used after xo_set_particles with
   a grid that is already sub-cell sorted): validates the order and returns 1 if some cell is not sorted by sub-cell */
int64_t xo_amr_tables(const xo_sim* s, int64_t* sgs, uint32_t* sgc)
{
  if (sgs) for (size_t q = 0; q < s->sub_grid_start.size(); q++) sgs[q] = (int64_t)s->sub_grid_start[q];
  if (sgc) for (size_t q = 0; q < s->sub_grid_cells.size(); q++) sgc[q] = s->sub_grid_cells[q];
  return (int64_t)s->sub_grid_cells.size();
}

int64_t xo_get_backup(const xo_sim* s, uint32_t* out)
{
  int64_t n = 0;
  for (const auto& v : s->backup) { if (out && !v.empty()) memcpy(out + n, v.data(), v.size() * 4); n += (int64_t)v.size(); }
  return n;
}

int64_t xo_stream_total_u16(const xo_sim* s) { int64_t n = 0; for (const auto& v : s->streams) n += (int64_t)v.size(); return n; }
void xo_stream_sizes(const xo_sim* s, uint32_t* sz) { for (size_t c = 0; c < s->streams.size(); c++) sz[c] = (uint32_t)s->streams[c].size(); }
void xo_stream_data(const xo_sim* s, uint16_t* out) { size_t o = 0; for (const auto& v : s->streams) { if (!v.empty()) memcpy(out + o, v.data(), v.size() * 2); o += v.size(); } }
int64_t xo_max_neighbors(const xo_sim* s) { return s->max_neighbors; }

int64_t xo_pairs(const xo_sim* s, uint64_t* out)
{
  const Grid& g = s->grid;
  const i64 gl = g.ghost_layers();
  std::vector<std::pair<uint64_t, uint64_t>> pairs;
  for (i64 k = gl; k < g.dims.k - gl; k++) for (i64 j = gl; j < g.dims.j - gl; j++) for (i64 i = gl; i < g.dims.i - gl; i++)
  {
    const i64 cell_a = ijk_to_index(g.dims, IJK{i, j, k});
    const Cell& A = g.cells[(size_t)cell_a];
    const StreamInfo si = stream_info(s->streams[(size_t)cell_a], A.size());
    for (size_t pa = 0; pa < A.size(); pa++)
      for_each_listed(*s, cell_a, pa, si, [&](i64 cell_b, size_t pb) { pairs.push_back({A.id[pa], g.cells[(size_t)cell_b].id[pb]}); });
  }
  if (out)
  {
    std::sort(pairs.begin(), pairs.end());
    for (size_t q = 0; q < pairs.size(); q++) { out[q * 2] = pairs[q].first; out[q * 2 + 1] = pairs[q].second; }
  }
  return (int64_t)pairs.size();
}

/* restated invariants: verify_chunk_neighbors.cpp:89-130 (monotone (cell_b,p_b), no self, in bounds),
   chunk_neighbors_stream_check.h:52-99 (offsets consistent with a sequential walk), chunk_neighbors.h:134-135,175-185 */
int xo_check_streams(const xo_sim* s)
{
  const Grid& g = s->grid;
  for (i64 ca = 0; ca < g.n_cells(); ca++)
  {
    const size_t na = g.cells[(size_t)ca].size();
    const auto& st = s->streams[(size_t)ca];
    if (na == 0) { if (!st.empty()) { g_err = "non-empty stream for empty cell"; return 1; } continue; }
    if (!(st[0] == 1 && st[1] == 0)) { g_err = "missing offset-table signature (1,0)"; return 2; }
    const StreamInfo si = stream_info(st, na);
    size_t cur = 0;
    for (size_t pa = 0; pa < na; pa++)
    {
      if ((size_t)(si.offset[pa] + si.shift) != cur) { g_err = "offset table disagrees with sequential walk"; return 3; }
      const uint16_t* p = si.stream + cur;
      int groups = *p++;
      uint16_t last_enc = 0;
      for (; groups > 0; --groups)
      {
        const uint16_t enc = *p++;
        if (enc < 1057) { g_err = "encoded cell below 1057"; return 4; }
        if (enc <= last_enc) { g_err = "cell groups not strictly ascending"; return 5; }
        last_enc = enc;
        int n = *p++;
        if (n <= 0) { g_err = "empty cell group"; return 6; }
        const int ri = (int)(enc & 31) - 16, rj = (int)((enc >> 5) & 31) - 16, rk = (int)((enc >> 10) & 31) - 16;
        const IJK la = index_to_ijk(g.dims, ca);
        const IJK lb{la.i + ri, la.j + rj, la.k + rk};
        if (!grid_contains(g.dims, lb)) { g_err = "neighbour cell out of grid"; return 7; }
        const i64 cb = ijk_to_index(g.dims, lb);
        int last_p = -1;
        for (; n > 0; --n)
        {
          const int pb = *p++;
          if (pb <= last_p) { g_err = "p_b not strictly ascending"; return 8; }
          last_p = pb;
          if ((size_t)pb >= g.cells[(size_t)cb].size()) { g_err = "p_b out of cell"; return 9; }
          if (cb == ca && (size_t)pb == pa) { g_err = "self neighbour"; return 10; }
        }
      }
      cur = (size_t)(p - si.stream);
    }
    if ((size_t)(si.offset[na] + si.shift) != cur) { g_err = "closing offset wrong"; return 11; }
    if ((size_t)((si.stream - st.data()) + cur) != st.size()) { g_err = "stream size mismatch"; return 12; }
  }
  return 0;
}

void xo_energy_virial(const xo_sim* cs, double* epot, double vir[6], double* ekin)
{
  xo_sim* s = const_cast<xo_sim*>(cs);
  /* forces are recomputed identically; energy/virial are by-products */
  compute_force(*s, epot, vir);
  const Grid& g = s->grid; const i64 gl = g.ghost_layers();
  double ke = 0., c = 0.;
  for (i64 k = gl; k < g.dims.k - gl; k++) for (i64 j = gl; j < g.dims.j - gl; j++) for (i64 i = gl; i < g.dims.i - gl; i++)
  {
    const Cell& C = g.cells[(size_t)ijk_to_index(g.dims, IJK{i, j, k})];
    for (size_t p = 0; p < C.size(); p++)
    {
      double x = 0.5 * s->cfg.mass * (C.vx[p] * C.vx[p] + C.vy[p] * C.vy[p] + C.vz[p] * C.vz[p]);
      double t = ke + x; c += (std::fabs(ke) >= std::fabs(x)) ? (ke - t) + x : (x - t) + ke; ke = t;
    }
  }
  if (ekin) *ekin = ke + c;
}

} // extern "C"
