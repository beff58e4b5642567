// xnb_operators.cpp -- the operators of the LJ hot path behind the reference's operator interface (see xnb_operators.hpp).
// Each operator cites the reference operator whose name, slots and YAML keys it mirrors; execute() forwards to the C-ABI.
#include "xnb_operators.hpp"
#include <regex>
#include <cstring>
#include <cstdio>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>
#include <set>
#include <random>
#include <ctime>
#include <cstdio>

namespace xnb { namespace host {

void fatal_error(const std::string& msg) { std::fprintf(stderr, "*** fatal error: %s\n", msg.c_str()); std::fflush(stderr); std::abort(); }

static void ck(xnb_ctx* c, int rc, const char* what) { if (rc != XNB_OK) fatal_error(std::string(what) + ": " + xnb_last_error(c)); }

Grid::~Grid() { if (ctx) xnb_destroy(ctx); }
int64_t Grid::number_of_particles() const { return ctx ? xnb_num_total(ctx) : 0; }
int64_t Grid::number_of_cells() const { if (!ctx) return 0; xnb_grid_info gi; return xnb_get_grid_info(ctx, &gi) == XNB_OK ? gi.n_cells : 0; }
size_t GridChunkNeighbors::number_of_cells() const { if (!ctx) return 0; xnb_grid_info gi; return xnb_get_grid_info(ctx, &gi) == XNB_OK ? (size_t)gi.n_cells : 0; }

// ---- parameters -----------------------------------------------------------------------------------------------------
static std::string trim(const std::string& s) { size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n"); return a == std::string::npos ? "" : s.substr(a, b - a + 1); }

// internal units of the LJ decks (input_lj_Ni.msp:15-24): angstrom, picosecond, Dalton.  Constants as recovered from the
// reference's golden file (tests/golden/README.md): e = 1.6021892e-19 C, u = 1.66053904e-27 kg.
double convert_quantity(const std::string& text)
{
  std::istringstream is(text);
  double v = 0; std::string unit;
  if (!(is >> v)) fatal_error("not a quantity: '" + text + "'");
  is >> unit;
  static const double J_INTERNAL = 1.0 / (1.66053904e-27 * 1e-20 / 1e-24);
  if (unit.empty()) return v;
  if (unit == "ang" || unit == "angstrom") return v;
  if (unit == "nm") return v * 10.0;
  if (unit == "um") return v * 1e4;
  if (unit == "m") return v * 1e10;
  if (unit == "ps") return v;
  if (unit == "fs") return v * 1e-3;
  if (unit == "ns") return v * 1e3;
  if (unit == "s") return v * 1e12;
  if (unit == "Da" || unit == "Dalton") return v;
  if (unit == "eV") return v * 1.6021892e-19 * J_INTERNAL;
  if (unit == "J") return v * J_INTERNAL;
  fatal_error("unknown unit '" + unit + "' in '" + text + "'");
}

Params Params::parse(const std::string& text)
{
  Params p;
  std::string s = trim(text);
  if (s.empty()) return p;
  if (s.front() == '{' && s.back() == '}') s = s.substr(1, s.size() - 2);
  // split on top level commas (brackets and braces nest)
  std::vector<std::string> items; int depth = 0; std::string cur;
  for (char ch : s)
  {
    if (ch == '[' || ch == '{') depth++;
    if (ch == ']' || ch == '}') depth--;
    if (ch == ',' && depth == 0) { items.push_back(cur); cur.clear(); } else cur += ch;
  }
  if (!trim(cur).empty()) items.push_back(cur);
  for (const std::string& it : items)
  {
    const size_t c = it.find(':');
    if (c == std::string::npos) fatal_error("expected 'key: value' in '" + it + "'");
    const std::string k = trim(it.substr(0, c)), v = trim(it.substr(c + 1));
    if (!v.empty() && v.front() == '{') { Params sub = parse(v); for (auto& kv : sub.kv) p.kv[k + "." + kv.first] = kv.second; }
    else p.kv[k] = v;
  }
  return p;
}
double Params::quantity(const std::string& k) const { auto it = kv.find(k); if (it == kv.end()) fatal_error("missing parameter '" + k + "'"); return convert_quantity(it->second); }
bool Params::boolean(const std::string& k) const { const std::string v = kv.at(k); return v == "true" || v == "True" || v == "yes" || v == "1"; }
std::vector<double> Params::quantities(const std::string& k) const
{
  std::string v = kv.at(k); std::vector<double> out; std::string cur;
  for (char ch : v) { if (ch == '[' || ch == ']') continue; if (ch == ',') { if (!trim(cur).empty()) out.push_back(convert_quantity(trim(cur))); cur.clear(); } else cur += ch; }
  if (!trim(cur).empty()) out.push_back(convert_quantity(trim(cur)));
  return out;
}

// ---- slots / factory / batch ---------------------------------------------------------------------------------------
SlotBase::SlotBase(OperatorNode* op, const char* n, SlotDirection d, bool req, const char* doc_, std::type_index t)
  : name(n), dir(d), required(req), doc(doc_), type(t) { op->slots.push_back(this); }

OperatorNodeFactory* OperatorNodeFactory::instance() { static OperatorNodeFactory f; return &f; }
std::unique_ptr<OperatorNode> OperatorNodeFactory::make_operator(const std::string& name) const
{
  auto it = makers_.find(name);
  if (it == makers_.end()) fatal_error("no operator factory registered for '" + name + "'");
  auto op = it->second(); op->name = name; return op;
}
std::vector<std::string> OperatorNodeFactory::available_operators() const { std::vector<std::string> v; for (auto& kv : makers_) v.push_back(kv.first); return v; }

OperatorNode* Batch::add(const std::string& op_name, const std::string& params, const std::map<std::string, std::string>& rebind)
{
  register_hot_path_operators();
  std::unique_ptr<OperatorNode> op = OperatorNodeFactory::instance()->make_operator(op_name);
  op->stream = stream_;
  const Params p = Params::parse(params);
  for (SlotBase* s : op->slots)
  {
    if (s->dir == PRIVATE) { s->bind(s->make_default()); continue; }
    auto rb = rebind.find(s->name);
    const std::string gname = rb == rebind.end() ? s->name : rb->second;
    Entry& e = table()[gname];
    if (!e.p)
    {
      // nobody produced it yet: outputs and defaulted inputs create the graph value, a REQUIRED input without producer
      // is resolved at execute() time (a later yaml_initialize may still provide it)
      std::shared_ptr<void> d = s->make_default();
      if (d) { e.p = d; e.t = s->type; }
    }
    if (e.p)
    {
      if (e.t != s->type) fatal_error("slot '" + s->name + "' of operator '" + op_name + "' is connected to '" + gname + "' of another type");
      s->bind(e.p);
    }
  }
  op->yaml_initialize(p);
  ops_.push_back(std::move(op));
  return ops_.back().get();
}

void Batch::execute()
{
  for (auto& op : ops_)
  {
    for (SlotBase* s : op->slots)
      if (s->required && !s->has_value() && s->dir != OUTPUT) fatal_error("operator '" + op->name + "': required slot '" + s->name + "' is not connected");
    op->execute();
  }
}

// =====================================================================================================================
// operators
// =====================================================================================================================
namespace {

struct MpiRank { int rank = 0, nranks = 1; };

// scalar parameter of an operator: YAML value if given, else the connected slot
// (a YAML value is private to the operator: it must not write through to the graph value other operators are bound to,
// e.g. push_f_v_r{dt_scale: 1.0} followed by push_f_v{dt_scale: 0.5})
#define XNB_PARAM_QUANTITY(p, slotname) do { if (p.has(#slotname)) { slotname.value = std::make_shared<double>(p.quantity(#slotname)); } } while (0)
#define XNB_PARAM_BOOL(p, slotname) do { if (p.has(#slotname)) { slotname.value = std::make_shared<bool>(p.boolean(#slotname)); } } while (0)

// op `domain` (core/lib/domain.cpp:359-..., YAML keys of input_lj_Ni.msp:63-69)
struct DomainOp : OperatorNode
{
  ADD_SLOT(Domain, domain, INPUT_OUTPUT);
  void yaml_initialize(const Params& p) override
  {
    Domain& d = *domain;
    if (p.has("cell_size")) d.cell_size = p.quantity("cell_size");
    if (p.has("grid_dims")) { auto v = p.quantities("grid_dims"); d.grid_dims = IJK{(int64_t)v.at(0), (int64_t)v.at(1), (int64_t)v.at(2)}; }
    if (p.has("bounds"))
    {
      auto v = p.quantities("bounds");           // [ [ xmin, ymin, zmin ] , [ xmax, ymax, zmax ] ]
      d.bounds.bmin = Vec3d{v.at(0), v.at(1), v.at(2)}; d.bounds.bmax = Vec3d{v.at(3), v.at(4), v.at(5)};
    }
    if (p.has("periodic"))
    {
      std::string s = p.str("periodic"); int q = 0; std::string cur;
      for (char ch : s + ",") { if (ch == '[' || ch == ']' || ch == ' ') continue; if (ch == ',') { if (!cur.empty() && q < 3) d.periodic[q++] = (cur == "true"); cur.clear(); } else cur += ch; }
    }
    if (p.has("expandable")) d.expandable = p.boolean("expandable");
  }
  void execute() override {}
};

// op `init_rcb_grid` (grid_cell_particles/init_rcb_grid.cpp:37-84): creates the local grid = this rank's RCB block
struct InitRcbGrid : OperatorNode
{
  ADD_SLOT(MpiRank, mpi, INPUT, MpiRank{});
  ADD_SLOT(Domain, domain, INPUT, REQUIRED);
  ADD_SLOT(Grid, grid, INPUT_OUTPUT);
  void execute() override
  {
    Grid& g = *grid;
    if (!g.ctx) { int rc = xnb_create(&g.ctx, g.device); if (rc != XNB_OK) fatal_error(std::string("init_rcb_grid: ") + xnb_last_error(nullptr)); }
    const Domain& d = *domain;
    const double bmin[3] = {d.bounds.bmin.x, d.bounds.bmin.y, d.bounds.bmin.z}, bmax[3] = {d.bounds.bmax.x, d.bounds.bmax.y, d.bounds.bmax.z};
    const int64_t gd[3] = {d.grid_dims.i, d.grid_dims.j, d.grid_dims.k};
    const int32_t per[3] = {d.periodic[0], d.periodic[1], d.periodic[2]};
    ck(g.ctx, xnb_set_domain(g.ctx, bmin, bmax, d.cell_size, gd, per), "domain");
    ck(g.ctx, xnb_init_rcb_grid(g.ctx, mpi->rank, mpi->nranks), "init_rcb_grid");
  }
};

// ops `particle_types` / `particle_type_add_properties` reduced to what the path reads: the per-type mass
struct ParticleTypeAddProperties : OperatorNode
{
  ADD_SLOT(ParticleTypeProperties, particle_type_properties, INPUT_OUTPUT);
  void yaml_initialize(const Params& p) override
  {
    for (auto& kv : p.kv) { const size_t dot = kv.first.find(".mass"); if (dot != std::string::npos) { particle_type_properties->mass.push_back(convert_quantity(kv.second)); particle_type_properties->names.push_back(kv.first.substr(0, dot)); } }
  }
  void execute() override {}
};

// ops `lattice` (structure FCC) and `gaussian_noise_r` with deterministic_noise
// (generate_particle_lattice.h:247-388, gaussian_noise.h:61-80,150-161).  The particles stay on the host until
// move_particles bins them (the reference fills cells directly).
struct LatticeRecipe { bool have = false; double a = 0, noise_sigma = 0, vel_sigma = 0; };
static void generate_particles(const Domain& d, const LatticeRecipe& r, ParticleSet& ps)
{
  xnb_lattice_cfg cfg{};
  cfg.bounds_min[0] = d.bounds.bmin.x; cfg.bounds_min[1] = d.bounds.bmin.y; cfg.bounds_min[2] = d.bounds.bmin.z;
  cfg.bounds_max[0] = d.bounds.bmax.x; cfg.bounds_max[1] = d.bounds.bmax.y; cfg.bounds_max[2] = d.bounds.bmax.z;
  cfg.cell_size = d.cell_size; cfg.grid_dims[0] = d.grid_dims.i; cfg.grid_dims[1] = d.grid_dims.j; cfg.grid_dims[2] = d.grid_dims.k;
  cfg.lattice_a = r.a; cfg.noise_sigma = r.noise_sigma; cfg.vel_sigma = r.vel_sigma;
  const double vol = (cfg.bounds_max[0] - cfg.bounds_min[0]) * (cfg.bounds_max[1] - cfg.bounds_min[1]) * (cfg.bounds_max[2] - cfg.bounds_min[2]);
  const int64_t cap = (int64_t)(4.0 * vol / (r.a * r.a * r.a) * 1.1) + 1024;
  for (auto* v : {&ps.rx, &ps.ry, &ps.rz, &ps.vx, &ps.vy, &ps.vz}) v->assign((size_t)cap, 0.0);
  ps.id.assign((size_t)cap, 0); ps.type.assign((size_t)cap, 0);
  const int64_t n = xnb_host_lattice_fcc(&cfg, cap, ps.rx.data(), ps.ry.data(), ps.rz.data(), ps.vx.data(), ps.vy.data(), ps.vz.data(), ps.id.data(), ps.type.data());
  if (n < 0) fatal_error("lattice: capacity");
  for (auto* v : {&ps.rx, &ps.ry, &ps.rz, &ps.vx, &ps.vy, &ps.vz}) v->resize((size_t)n);
  ps.id.resize((size_t)n); ps.type.resize((size_t)n);
}
struct Lattice : OperatorNode
{
  ADD_SLOT(Domain, domain, INPUT, REQUIRED);
  ADD_SLOT(LatticeRecipe, lattice_recipe, INPUT_OUTPUT);
  ADD_SLOT(ParticleSet, pending_particles, INPUT_OUTPUT);
  void yaml_initialize(const Params& p) override
  {
    if (p.has("structure") && p.str("structure") != "FCC") fatal_error("lattice: only structure FCC is on the LJ hot path");
    if (p.has("size")) lattice_recipe->a = p.quantities("size").at(0);
  }
  void execute() override { lattice_recipe->have = true; generate_particles(*domain, *lattice_recipe, *pending_particles); }
};
struct GaussianNoiseR : OperatorNode
{
  ADD_SLOT(Domain, domain, INPUT, REQUIRED);
  ADD_SLOT(double, sigma, INPUT, 1.0);
  ADD_SLOT(bool, deterministic_noise, INPUT, false);
  ADD_SLOT(LatticeRecipe, lattice_recipe, INPUT_OUTPUT);
  ADD_SLOT(ParticleSet, pending_particles, INPUT_OUTPUT);
  void yaml_initialize(const Params& p) override { XNB_PARAM_QUANTITY(p, sigma); }
  void execute() override
  {
    if (!lattice_recipe->have) fatal_error("gaussian_noise_r: no lattice to perturb");
    if (!*deterministic_noise) fatal_error("gaussian_noise_r: only deterministic_noise: true is reproducible (gaussian_noise.h:150-161)");
    lattice_recipe->noise_sigma = *sigma; generate_particles(*domain, *lattice_recipe, *pending_particles);
  }
};

// op `nbh_dist` (particle_neighbors/nbh_dist.cpp:30-88), identity xform: grid space == lab space
struct NeighborDistance : OperatorNode
{
  ADD_SLOT(double, rcut_max, INPUT, 0.0, DocString{"maximum search distance for the neighborhood, in physical space"});
  ADD_SLOT(double, rcut_inc, INPUT, 0.0, DocString{"value added to the search distance to update neighbor list less frequently"});
  ADD_SLOT(double, ghost_dist_max, INPUT_OUTPUT, 0.0, DocString{"maximum distance needed for ghost particles out of sub domain"});
  ADD_SLOT(Domain, domain, INPUT, REQUIRED);
  ADD_SLOT(double, nbh_dist_lab, INPUT_OUTPUT, DocString{"neighborhood distance, in lab space"});
  ADD_SLOT(double, nbh_dist, INPUT_OUTPUT, DocString{"neighborhood distance, in grid space"});
  ADD_SLOT(double, ghost_dist, INPUT_OUTPUT, DocString{"thickness of ghost particle layer, in grid space"});
  ADD_SLOT(double, max_displ, INPUT_OUTPUT, DocString{"move threshold, in grid space, that must trigger a neighbor list update"});
  ADD_SLOT(double, max_displ_lab, INPUT_OUTPUT, DocString{"move threshold, in lab space"});
  ADD_SLOT(Grid, grid, INPUT_OUTPUT, DocString{"(mirror only) the ctx that receives the distances"});
  void execute() override
  {
    *nbh_dist_lab = *rcut_max + *rcut_inc; *nbh_dist = *nbh_dist_lab;                      // :47
    *max_displ_lab = *rcut_inc / 2.0; *max_displ = *max_displ_lab;                        // :48
    *ghost_dist_max = std::max(*ghost_dist_max, *rcut_max); *ghost_dist = *ghost_dist_max + *rcut_inc;   // :60-62
    ck(grid->ctx, xnb_set_nbh_dist(grid->ctx, *rcut_max, *rcut_inc), "nbh_dist");
  }
};

// op `move_particles` (grid_cell_particles/move_particles.cpp:42-48 -> move_particles_across_cells.h:78-235);
// with several ranks it also performs the migrate_cell_particles hand-off
struct MoveParticles : OperatorNode
{
  ADD_SLOT(Domain, domain, INPUT, REQUIRED);
  ADD_SLOT(Grid, grid, INPUT_OUTPUT);
  ADD_SLOT(ParticleSet, pending_particles, INPUT_OUTPUT);
  ADD_SLOT(ParticleTypeProperties, particle_type_properties, INPUT, ParticleTypeProperties{});
  void execute() override
  {
    ParticleSet& ps = *pending_particles;
    if (!ps.rx.empty())
    {
      if (!particle_type_properties->mass.empty()) ck(grid->ctx, xnb_set_type_mass(grid->ctx, particle_type_properties->mass.data(), (int)particle_type_properties->mass.size()), "particle masses");
      ck(grid->ctx, xnb_set_particles(grid->ctx, (int64_t)ps.rx.size(), ps.rx.data(), ps.ry.data(), ps.rz.data(), ps.vx.data(), ps.vy.data(), ps.vz.data(), ps.id.data(), ps.type.data()), "lattice upload");
      ps = ParticleSet{};
    }
    ck(grid->ctx, xnb_move_particles(grid->ctx, stream), "move_particles");
  }
};
struct Nop : OperatorNode { void execute() override {} };

// op `amr_grid_pairs` (amr/amr_grid_pairs.cpp -> max_distance_sub_cell_pairs, amr/lib/amr_grid_algorithm.cpp:102-218): the
// AmrSubCellPairCache, rebuilt only when the maximum sub-grid resolution, the cell size or the distance changed (:129-133)
struct AmrGridPairs : OperatorNode
{
  ADD_SLOT(Grid, grid, INPUT);
  ADD_SLOT(AmrGrid, amr, INPUT, AmrGrid{});
  ADD_SLOT(Domain, domain, INPUT, REQUIRED);
  ADD_SLOT(double, nbh_dist, INPUT, REQUIRED);
  ADD_SLOT(AmrSubCellPairCache, amr_grid_pairs, INPUT_OUTPUT);
  void execute() override
  {
    if (!grid->ctx) return;
    xnb_grid_info gi; ck(grid->ctx, xnb_get_grid_info(grid->ctx, &gi), "amr_grid_pairs");
    std::vector<int64_t> sgs((size_t)gi.n_cells + 1, 0);
    if (xnb_get_amr(grid->ctx, sgs.data(), nullptr) < 0) fatal_error(std::string("amr_grid_pairs: ") + xnb_last_error(grid->ctx));
    size_t max_res = 1;
    for (int64_t c = 0; c < gi.n_cells; c++) { const size_t side = (size_t)std::llround(std::cbrt((double)(sgs[(size_t)c + 1] - sgs[(size_t)c] + 1))); max_res = std::max(max_res, side); }
    AmrSubCellPairCache& pc = *amr_grid_pairs;
    pc.ctx = grid->ctx;
    if (max_res <= pc.m_max_res && domain->cell_size == pc.m_cell_size && *nbh_dist == pc.m_max_dist) return;       // cache up to date
    pc.m_max_res = max_res; pc.m_cell_size = domain->cell_size; pc.m_max_dist = *nbh_dist;
    const int64_t n = xnb_host_amr_sub_cell_pairs((int)max_res, pc.m_cell_size, pc.m_max_dist, nullptr, nullptr);
    if (n < 0) fatal_error("amr_grid_pairs: bad resolution / distance");
    const size_t layers = (size_t)std::ceil(pc.m_max_dist / pc.m_cell_size);
    pc.m_list_offsets.assign(max_res * (max_res + 1) / 2 * (layers + 1) * (layers + 1) * (layers + 1) + 1, 0);
    pc.m_pair_ab.assign((size_t)n, 0);
    xnb_host_amr_sub_cell_pairs((int)max_res, pc.m_cell_size, pc.m_max_dist, pc.m_list_offsets.data(), pc.m_pair_ab.data());
  }
};     // operators whose work is folded into a neighbour (see registration)

struct RebuildAmr : OperatorNode        // amr/rebuild_amr.cpp:35-68 (slot default 5.0; the default decks set 6.5, update-particles.msp:9)
{
  ADD_SLOT(double, sub_grid_density, INPUT, 5.0);
  ADD_SLOT(long, enforced_ordering, INPUT, 1L);
  ADD_SLOT(Grid, grid, INPUT_OUTPUT);
  ADD_SLOT(AmrGrid, amr, INPUT_OUTPUT);
  void yaml_initialize(const Params& p) override { XNB_PARAM_QUANTITY(p, sub_grid_density); }
  void execute() override { ck(grid->ctx, xnb_set_sub_grid_density(grid->ctx, *sub_grid_density), "rebuild_amr"); ck(grid->ctx, xnb_rebuild_amr(grid->ctx, stream), "rebuild_amr"); amr->ctx = grid->ctx; }
};
struct BackupR : OperatorNode           // io/backup_r.cpp:36-78
{
  ADD_SLOT(Grid, grid, INPUT);
  ADD_SLOT(Domain, domain, INPUT);
  ADD_SLOT(PositionBackupData, backup_r, INPUT_OUTPUT);
  void execute() override { ck(grid->ctx, xnb_backup_r(grid->ctx, stream), "backup_r"); backup_r->ctx = grid->ctx; }
};
struct GhostCommSchemeOp : OperatorNode // mpi/update_ghosts_comm_scheme.cpp:58-492
{
  ADD_SLOT(Grid, grid, INPUT_OUTPUT);
  ADD_SLOT(Domain, domain, INPUT, REQUIRED);
  ADD_SLOT(GhostCommunicationScheme, ghost_comm_scheme, INPUT_OUTPUT);
  void execute() override { ck(grid->ctx, xnb_ghost_comm_scheme(grid->ctx, stream), "ghost_comm_scheme"); ghost_comm_scheme->ctx = grid->ctx; }
};
template <bool ALL>
struct GhostUpdate : OperatorNode       // mpi/update_ghosts.cu:45-64, update_ghosts.h:53-190
{
  ADD_SLOT(Grid, grid, INPUT_OUTPUT);
  ADD_SLOT(GhostCommunicationScheme, ghost_comm_scheme, INPUT_OUTPUT);
  ADD_SLOT(UpdateGhostConfig, update_ghost_config, INPUT, UpdateGhostConfig{});
  void execute() override { ck(grid->ctx, ALL ? xnb_ghost_update_all(grid->ctx, stream) : xnb_ghost_update_r(grid->ctx, stream), ALL ? "ghost_update_all" : "ghost_update_r"); }
};
struct BuildChunkNeighbors : OperatorNode   // particle_neighbors/chunk_neighbors.cpp:48-74
{
  ADD_SLOT(Grid, grid, INPUT);
  ADD_SLOT(AmrGrid, amr, INPUT, AmrGrid{});
  ADD_SLOT(AmrSubCellPairCache, amr_grid_pairs, INPUT, AmrSubCellPairCache{});
  ADD_SLOT(Domain, domain, INPUT, REQUIRED);
  ADD_SLOT(double, nbh_dist_lab, INPUT, REQUIRED);
  ADD_SLOT(GridChunkNeighbors, chunk_neighbors, INPUT_OUTPUT);
  ADD_SLOT(ChunkNeighborsConfig, config, INPUT_OUTPUT, ChunkNeighborsConfig{});
  void yaml_initialize(const Params& p) override
  {
    if (p.has("config.chunk_size") && (unsigned)p.quantity("config.chunk_size") != 1u) fatal_error("chunk_neighbors: chunk_size is frozen at 1 (chunk_neighbors_config.h:49,74)");
    if (p.has("config.half_symmetric")) config->half_symmetric = p.boolean("config.half_symmetric");      // chunk_neighbors_config.h:72
    if (p.has("config.skip_ghosts")) config->skip_ghosts = p.boolean("config.skip_ghosts");
    if (p.has("config.build_particle_offset") && !p.boolean("config.build_particle_offset")) fatal_error("chunk_neighbors: build_particle_offset: false is not supported");
  }
  void execute() override
  {
    config->chunk_size = 1;
    ck(grid->ctx, xnb_set_chunk_neighbors_config(grid->ctx, config->half_symmetric ? 1 : 0, config->skip_ghosts ? 1 : 0), "chunk_neighbors config");
    ck(grid->ctx, xnb_chunk_neighbors(grid->ctx, stream), "chunk_neighbors");
    GridChunkNeighbors& n = *chunk_neighbors; n.ctx = grid->ctx;
    ck(grid->ctx, xnb_view_chunk_neighbors(grid->ctx, &n.cell_stream, &n.cell_stream_size, &n.max_neighbors), "chunk_neighbors view");
  }
};
struct ZeroParticleForce : OperatorNode     // compute/zero_particle_force.cu:32-53
{
  ADD_SLOT(Grid, grid, INPUT_OUTPUT);
  ADD_SLOT(bool, ghost, INPUT, false);
  void yaml_initialize(const Params& p) override { XNB_PARAM_BOOL(p, ghost); }
  void execute() override { ck(grid->ctx, xnb_zero_particle_force(grid->ctx, *ghost ? 1 : 0, stream), "zero_particle_force"); }
};
// SYMMETRIC = true is an extension registered as `lennard_jones_force_symmetric` (same slots): the Newton-3 sweep over
// half_symmetric lists, to be followed by update_force_from_ghost
template <bool SYMMETRIC>
struct LennardJonesForce : OperatorNode     // contribs/md/lennard_jones/lennard_jones.cu:171-215
{
  ADD_SLOT(LennardJonesParms, config, INPUT, REQUIRED, DocString{"Lennard-Jones potential parameters"});
  ADD_SLOT(double, rcut, INPUT, 0.0, DocString{"Cutoff distance"});
  ADD_SLOT(GridChunkNeighbors, chunk_neighbors, INPUT, GridChunkNeighbors{}, DocString{"neighbor list"});
  ADD_SLOT(bool, ghost, INPUT, false, DocString{"Enables computation in ghost cells"});
  ADD_SLOT(bool, experimental_ccb, INPUT, false, DocString{"accepted and ignored: the sweep kernel is always block cooperative"});
  ADD_SLOT(Domain, domain, INPUT, REQUIRED, DocString{"Simulation domain"});
  ADD_SLOT(double, rcut_max, INPUT_OUTPUT, 0.0, DocString{"Updated max rcut"});
  ADD_SLOT(Grid, grid, INPUT_OUTPUT, DocString{"Local sub-domain particles grid"});
  void yaml_initialize(const Params& p) override
  {
    if (p.has("config.epsilon")) { if (!config.value) config.value = std::make_shared<LennardJonesParms>(); config->epsilon = p.quantity("config.epsilon"); config->sigma = p.quantity("config.sigma"); }
    if (p.has("rcut")) { rcut.value = std::make_shared<double>(p.quantity("rcut")); }
    XNB_PARAM_BOOL(p, ghost);
    *rcut_max = std::max(*rcut, *rcut_max);        // the reference raises rcut_max when the graph is initialised so that nbh_dist sees it (:193)
  }
  void execute() override
  {
    *rcut_max = std::max(*rcut, *rcut_max);
    if (grid->number_of_cells() == 0) return;
    if (SYMMETRIC) ck(grid->ctx, xnb_lennard_jones_force_symmetric(grid->ctx, config->epsilon, config->sigma, *rcut, stream), "lennard_jones_force (symmetric)");
    else ck(grid->ctx, xnb_lennard_jones_force(grid->ctx, config->epsilon, config->sigma, *rcut, *ghost ? 1 : 0, stream), "lennard_jones_force");
  }
};
// a second functor of the pair-functor concept, with a per-neighbour field (the neighbour's type): runs through the general pair sweep
struct GravitationalForce : OperatorNode    // contribs/pi/gravitational_force.cu:161-217
{
  ADD_SLOT(GravitationalParms, config, INPUT, REQUIRED, DocString{"gravitational constant G"});
  ADD_SLOT(ParticleTypeProperties, particle_type_properties, INPUT, ParticleTypeProperties{});
  ADD_SLOT(double, rcut, INPUT, 0.0, DocString{"Cutoff distance"});
  ADD_SLOT(GridChunkNeighbors, chunk_neighbors, INPUT, GridChunkNeighbors{}, DocString{"neighbor list"});
  ADD_SLOT(bool, ghost, INPUT, false, DocString{"Enables computation in ghost cells"});
  ADD_SLOT(bool, compute_buffer, INPUT, false, DocString{"OURS: call the functor's ComputePairBuffer2 form instead of the buffer-less one"});
  ADD_SLOT(Domain, domain, INPUT, REQUIRED, DocString{"Simulation domain"});
  ADD_SLOT(double, rcut_max, INPUT_OUTPUT, 0.0, DocString{"Updated max rcut"});
  ADD_SLOT(Grid, grid, INPUT_OUTPUT, DocString{"Local sub-domain particles grid"});
  void yaml_initialize(const Params& p) override
  {
    if (p.has("config.G")) { if (!config.value) config.value = std::make_shared<GravitationalParms>(); config->G = p.quantity("config.G"); }
    if (p.has("rcut")) { rcut.value = std::make_shared<double>(p.quantity("rcut")); }
    XNB_PARAM_BOOL(p, ghost); XNB_PARAM_BOOL(p, compute_buffer);
    *rcut_max = std::max(*rcut, *rcut_max);
  }
  void execute() override
  {
    *rcut_max = std::max(*rcut, *rcut_max);            // :184
    if (grid->number_of_cells() == 0) return;
    if (particle_type_properties->mass.empty()) fatal_error("particle type property 'mass' is missing");      // :187-190
    ck(grid->ctx, xnb_set_type_mass(grid->ctx, particle_type_properties->mass.data(), (int)particle_type_properties->mass.size()), "particle masses");
    ck(grid->ctx, xnb_gravitational_force(grid->ctx, config->G, *rcut, *ghost ? 1 : 0, *compute_buffer ? 1 : 0, stream), "gravitational_force");
  }
};
// a third functor: per-neighbour scalar field + particle context (start / pair / stop), through the general pair sweep
struct AverageNeighborsScalar : OperatorNode    // src/compute/average_neighbors.cu:114-215
{
  ADD_SLOT(double, rcut, INPUT, 0.0, DocString{"Cutoff distance for average operation."});
  ADD_SLOT(std::vector<double>, weight_function, INPUT, std::vector<double>{1.0}, DocString{"[a0,...,an] coefficients of the polynomial distance weighting function"});
  ADD_SLOT(GridChunkNeighbors, chunk_neighbors, INPUT, GridChunkNeighbors{}, DocString{"neighbor list"});
  ADD_SLOT(Domain, domain, INPUT, REQUIRED, DocString{"Simulation domain"});
  ADD_SLOT(ParticleTypeProperties, particle_type_properties, INPUT, ParticleTypeProperties{});
  ADD_SLOT(std::string, avg_field, INPUT, REQUIRED, DocString{"Name of the resulting averaged field."});
  ADD_SLOT(std::string, nbh_field, INPUT, REQUIRED, DocString{"Name of the neighbors field to be averaged."});
  ADD_SLOT(double, rcut_max, INPUT_OUTPUT, 0.0, DocString{"Updated max rcut"});
  ADD_SLOT(Grid, grid, INPUT_OUTPUT, DocString{"Local sub-domain particles grid"});
  void yaml_initialize(const Params& p) override
  {
    if (p.has("rcut")) rcut.value = std::make_shared<double>(p.quantity("rcut"));
    if (p.has("weight_function")) weight_function.value = std::make_shared<std::vector<double>>(p.quantities("weight_function"));
    if (p.has("avg_field")) avg_field.value = std::make_shared<std::string>(p.str("avg_field"));
    if (p.has("nbh_field")) nbh_field.value = std::make_shared<std::string>(p.str("nbh_field"));
  }
  void execute() override
  {
    *rcut_max = std::max(*rcut, *rcut_max);            // :197
    if (grid->number_of_cells() == 0) return;
    if (weight_function->size() > 4) fatal_error("weighting function polynomial has a maximum degree of 3 (maximum 4 coefficients)");   // :199-202
    static const char* names[] = {"rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz", "id", "type"};
    int field = -1;
    for (int i = 0; i <= XNB_FIELD_TYPE; i++) if (*nbh_field == names[i]) field = i;
    if (field < 0) return;                             // like the reference (:133): no grid field of that name, nothing is computed
    double coefs[4] = {1.0, 0.0, 0.0, 0.0};
    for (size_t i = 0; i < weight_function->size(); i++) coefs[i] = (*weight_function)[i];
    ck(grid->ctx, xnb_average_neighbors(grid->ctx, *rcut, coefs, field, stream), "average_neighbors_scalar");
    grid->generic_real_field = *avg_field;             // field::mk_generic_real(avg_field): the ctx holds one generic real field
  }
};
// dynamic re-partition: cost model + cost-weighted RCB + migration in one collective call
struct LoadBalanceRCB : OperatorNode        // mpi/load_balance_rcb.cpp:51-601 (+ simple_cost_model.h, migrate_cell_particles.cpp:101-143)
{
  ADD_SLOT(Grid, grid, INPUT_OUTPUT);
  ADD_SLOT(Domain, domain, INPUT, REQUIRED);
  ADD_SLOT(double, lb_inbalance, INPUT_OUTPUT, 0.0, DocString{"(max - avg) / avg of the block costs after the re-partition"});
  void execute() override
  {
    double before = 0.0, after = 0.0;
    ck(grid->ctx, xnb_load_balance_rcb(grid->ctx, nullptr, &before, &after, stream), "load_balance_rcb");
    *lb_inbalance = after;
  }
};
struct UpdateForceFromGhost : OperatorNode  // mpi/update_force_from_ghost.cu:44 (UpdateFromGhosts<fx,fy,fz, UpdateValueAdd>)
{
  ADD_SLOT(Grid, grid, INPUT_OUTPUT);
  void execute() override { ck(grid->ctx, xnb_update_force_from_ghost(grid->ctx, stream), "update_force_from_ghost"); }
};
struct DivideForceByTypeScalar : OperatorNode   // compute/vec3_typescalar_op.cu:71-122 (`divide_force_by_type_scalar: mass`)
{
  ADD_SLOT(Grid, grid, INPUT_OUTPUT);
  void execute() override { ck(grid->ctx, xnb_divide_force_by_mass(grid->ctx, stream), "divide_force_by_type_scalar"); }
};
struct PushFVR : OperatorNode                   // defbox/push_vec3_2nd_order.h:77-117
{
  ADD_SLOT(Grid, grid, INPUT_OUTPUT);
  ADD_SLOT(double, dt, INPUT);
  ADD_SLOT(double, dt_scale, INPUT, 1.0);
  ADD_SLOT(Domain, domain, INPUT, REQUIRED);
  void yaml_initialize(const Params& p) override { XNB_PARAM_QUANTITY(p, dt_scale); }
  void execute() override { ck(grid->ctx, xnb_push_f_v_r(grid->ctx, *dt, *dt_scale, stream), "push_f_v_r"); }
};
struct PushFV : OperatorNode                    // defbox/push_vec3_1st_order.h:71-106
{
  ADD_SLOT(Grid, grid, INPUT_OUTPUT);
  ADD_SLOT(double, dt, INPUT);
  ADD_SLOT(double, dt_scale, INPUT, 1.0);
  ADD_SLOT(Domain, domain, INPUT, REQUIRED);
  void yaml_initialize(const Params& p) override { XNB_PARAM_QUANTITY(p, dt_scale); }
  void execute() override { ck(grid->ctx, xnb_push_f_v(grid->ctx, *dt, *dt_scale, stream), "push_f_v"); }
};
struct ParticleDisplOver : OperatorNode         // mpi/particle_displ_over.cu:98-178
{
  ADD_SLOT(Grid, grid, INPUT);
  ADD_SLOT(Domain, domain, INPUT);
  ADD_SLOT(PositionBackupData, backup_r, INPUT);
  ADD_SLOT(double, threshold, INPUT, 0.0);
  ADD_SLOT(double, threshold_lab, INPUT, 0.0);
  ADD_SLOT(bool, async, INPUT, false);
  ADD_SLOT(bool, result, OUTPUT);
  void execute() override { uint64_t n = 0; ck(grid->ctx, xnb_particle_displ_over(grid->ctx, &n, stream), "particle_displ_over"); *result = n > 0; }
};

// op `check_values` (debug/check_values.cpp:52-405), reader side: [id, r(3), a(3), v(3)] hex-float rows
// ---------------------------------------------------------------------------------------------------------------------
// on-disk formats either side of the loop.  Single sub-domain per file (the reference writes one file from all
// ranks through MPI-IO: io/include/exanb/io/mpi_file_io.h).
// ---------------------------------------------------------------------------------------------------------------------
// fetch this rank's inner particles (cell order)
struct HostParticles { std::vector<double> f[9]; std::vector<uint64_t> id; std::vector<uint8_t> type; int64_t n = 0; };
static void fetch_particles(xnb_ctx* ctx, HostParticles& hp, const char* who)
{
  hp.n = xnb_num_inner(ctx);
  for (auto& v : hp.f) v.resize((size_t)hp.n);
  hp.id.resize((size_t)hp.n); hp.type.resize((size_t)hp.n);
  ck(ctx, xnb_get_particles(ctx, 0, hp.n, hp.f[0].data(), hp.f[1].data(), hp.f[2].data(), hp.f[3].data(), hp.f[4].data(), hp.f[5].data(), hp.f[6].data(), hp.f[7].data(), hp.f[8].data(),
                            hp.id.data(), hp.type.data(), nullptr), who);
}

// op `write_xyz` (io/write_xyz.cpp:28-116, io/include/exanb/io/write_xyz.h:200-415): extended-XYZ text.
//   line 1: particle count; line 2: Lattice="<9 x %10.12e>" Properties=species:S:1:pos:R:3[:vel:R:3][:force:R:3][:processor_id:I:1][:id:I:1][:type:I:1] Time=<t>
//   then one FIXED-WIDTH line per particle: type name "%-8s", ' ', position "% .10e % .10e % .10e", then every selected field preceded by
//   ' ' (reals "% .10e", integers "% 10d").  Fields are selected by the regular expressions of `fields` (default ".*"), position always
//   first; `field_alias` defaults position -> pos, velocity -> vel.  (The reference's `units` map is parsed but its formatter is called
//   without a field id, write_xyz.h:316,368, so no conversion is applied; none is applied here either.)
struct WriteXYZ : OperatorNode
{
  ADD_SLOT(Grid, grid, INPUT);
  ADD_SLOT(Domain, domain, INPUT, REQUIRED);
  ADD_SLOT(bool, ghost, INPUT, false);
  ADD_SLOT(std::string, filename, INPUT, std::string("output"));
  ADD_SLOT(std::vector<std::string>, fields, INPUT, std::vector<std::string>({".*"}), DocString{"List of regular expressions to select fields to write"});
  ADD_SLOT(ParticleTypeProperties, particle_type_properties, INPUT, ParticleTypeProperties{});
  ADD_SLOT(double, physical_time, INPUT, 0.0);
  void yaml_initialize(const Params& p) override
  {
    if (p.has("filename")) filename.value = std::make_shared<std::string>(p.str("filename"));
    if (p.has("fields"))
    {
      std::vector<std::string> l; std::string cur;
      for (char ch : p.str("fields")) { if (ch == '[' || ch == ']' || ch == ' ' || ch == '"') continue; if (ch == ',') { if (!cur.empty()) l.push_back(cur); cur.clear(); } else cur += ch; }
      if (!cur.empty()) l.push_back(cur);
      fields.value = std::make_shared<std::vector<std::string>>(l);
    }
    XNB_PARAM_BOOL(p, ghost);
  }
  void execute() override
  {
    if (*ghost) fatal_error("write_xyz: ghost: true is not offered (ghosts are copies of inner particles)");
    HostParticles hp; fetch_particles(grid->ctx, hp, "write_xyz");
    auto selected = [&](const std::string& name) { for (const auto& f : *fields) if (std::regex_match(name, std::regex(f))) return true; return false; };
    const bool w_vel = selected("velocity"), w_force = selected("force"), w_rank = selected("processor_id"), w_id = selected("id"), w_type = selected("type");
    const Domain& d = *domain;
    std::FILE* out = std::fopen(filename->c_str(), "w");
    if (!out) fatal_error("write_xyz: cannot write '" + *filename + "'");
    const double L[3] = {d.bounds.bmax.x - d.bounds.bmin.x, d.bounds.bmax.y - d.bounds.bmin.y, d.bounds.bmax.z - d.bounds.bmin.z};
    std::fprintf(out, "%ld\nLattice=\"%10.12e %10.12e %10.12e %10.12e %10.12e %10.12e %10.12e %10.12e %10.12e\"", (long)hp.n, L[0], 0., 0., 0., L[1], 0., 0., 0., L[2]);
    std::ostringstream prop; prop << " Properties=species:S:1:pos:R:3";
    if (w_vel) prop << ":vel:R:3";
    if (w_force) prop << ":force:R:3";
    if (w_rank) prop << ":processor_id:I:1";
    if (w_id) prop << ":id:I:1";
    if (w_type) prop << ":type:I:1";
    prop << " Time=" << *physical_time << "\n";
    std::fputs(prop.str().c_str(), out);
    const auto& names = particle_type_properties->names;
    for (int64_t q = 0; q < hp.n; q++)
    {
      const size_t i = (size_t)q;
      const char* tn = hp.type[i] < names.size() ? names[hp.type[i]].c_str() : "XX";
      std::fprintf(out, "%-8s % .10e % .10e % .10e", tn, hp.f[0][i], hp.f[1][i], hp.f[2][i]);
      if (w_vel) std::fprintf(out, " % .10e % .10e % .10e", hp.f[3][i], hp.f[4][i], hp.f[5][i]);
      if (w_force) std::fprintf(out, " % .10e % .10e % .10e", hp.f[6][i], hp.f[7][i], hp.f[8][i]);
      if (w_rank) std::fprintf(out, " % 10d", 0);
      if (w_id) std::fprintf(out, " % 10ld", (long)hp.id[i]);
      if (w_type) std::fprintf(out, " % 10d", (int)hp.type[i]);
      std::fputc('\n', out);
    }
    std::fclose(out);
  }
};

// ops `write_dump` / `read_dump` (io/write_dump.cpp:36-46, read_dump.cpp, io/include/exanb/io/sim_dump_io.h:105-190,
// sim_dump_writer.h:100-490, sim_dump_reader.h): checkpoint of positions, velocities, id, type (SimDumpWriteAllButForce).
// File = SimDumpHeader fields in the reference's order (version 1.3 = 1003, nb_fields, data_flags, tuple_size, field_size[128],
// fields[128][32], nb_particles, time_step, time, domain, optional/table/data offsets, optional_header_size, chunk_count), a table of
// DataChunkItem {u64 global_offset, u32 data_size, i32 n_particles}, then the chunks: arrays of tuples {rx ry rz vx vy vz (f64), id (u64),
// type (u8) + 7 pad bytes}, uncompressed (negative n_particles = "size is the raw size", sim_dump_writer.h:87-92).  The byte layouts
// of the domain record and of the tuple are THIS MIRROR'S (onika's soatl::FieldTuple and the reference Domain class are not
// reproducible byte for byte without onika): files round-trip through these two operators, not through the reference.
struct DumpDomainRecord { double bmin[3], bmax[3], cell_size; int64_t grid_dims[3]; double xform[9]; uint32_t flags; uint32_t pad; };
struct DumpTuple { double r[3], v[3]; uint64_t id; uint8_t type; uint8_t pad[7]; };
struct DumpChunkItem { uint64_t global_offset; uint32_t data_size; int32_t n_particles; };
struct DumpHeader
{
  uint64_t version = 1003; uint32_t nb_fields = 8; uint8_t data_flags[4] = {0, 0, 0, 0}; uint64_t tuple_size = sizeof(DumpTuple);
  uint64_t field_size[128]; char fields[128][32];
  uint64_t nb_particles = 0, time_step = 0; double time = 0.0;
  DumpDomainRecord domain;
  uint64_t optional_offset = 0, table_offset = 0, data_offset = 0, optional_header_size = 0, chunk_count = 0;
};
static const size_t DUMP_CHUNK_PARTICLES = 1048576;      // WRITE_BUFFER_SIZE of sim_dump_writer.h
struct WriteDump : OperatorNode
{
  ADD_SLOT(Grid, grid, INPUT);
  ADD_SLOT(Domain, domain, INPUT, REQUIRED);
  ADD_SLOT(std::string, filename, INPUT, std::string("output.dump"));
  ADD_SLOT(long, timestep, INPUT, 0L);
  ADD_SLOT(double, physical_time, INPUT, 0.0);
  ADD_SLOT(long, compression_level, INPUT, 0L, DocString{"accepted; chunks are stored raw"});
  void yaml_initialize(const Params& p) override { if (p.has("filename")) filename.value = std::make_shared<std::string>(p.str("filename")); }
  void execute() override
  {
    HostParticles hp; fetch_particles(grid->ctx, hp, "write_dump");
    DumpHeader h;
    std::memset(h.field_size, 0, sizeof h.field_size); std::memset(h.fields, 0, sizeof h.fields);
    const char* fn[8] = {"rx", "ry", "rz", "vx", "vy", "vz", "id", "type"}; const uint64_t fs[8] = {8, 8, 8, 8, 8, 8, 8, 1};
    for (int q = 0; q < 8; q++) { std::strncpy(h.fields[q], fn[q], 31); h.field_size[q] = fs[q]; }
    h.nb_particles = (uint64_t)hp.n; h.time_step = (uint64_t)*timestep; h.time = *physical_time;
    const Domain& d = *domain;
    std::memset(&h.domain, 0, sizeof h.domain);
    h.domain.bmin[0] = d.bounds.bmin.x; h.domain.bmin[1] = d.bounds.bmin.y; h.domain.bmin[2] = d.bounds.bmin.z;
    h.domain.bmax[0] = d.bounds.bmax.x; h.domain.bmax[1] = d.bounds.bmax.y; h.domain.bmax[2] = d.bounds.bmax.z;
    h.domain.cell_size = d.cell_size; h.domain.grid_dims[0] = d.grid_dims.i; h.domain.grid_dims[1] = d.grid_dims.j; h.domain.grid_dims[2] = d.grid_dims.k;
    h.domain.xform[0] = h.domain.xform[4] = h.domain.xform[8] = 1.0;
    h.domain.flags = (d.periodic[0] ? 1u : 0u) | (d.periodic[1] ? 2u : 0u) | (d.periodic[2] ? 4u : 0u) | (d.expandable ? 8u : 0u);
    const size_t nchunks = ((size_t)hp.n + DUMP_CHUNK_PARTICLES - 1) / DUMP_CHUNK_PARTICLES;
    h.chunk_count = nchunks + 1;                                       // a free slot at the end stores the file size (sim_dump_writer.h:173)
    h.optional_offset = sizeof(DumpHeader); h.table_offset = h.optional_offset; h.data_offset = h.table_offset + h.chunk_count * sizeof(DumpChunkItem);
    std::vector<DumpChunkItem> table(h.chunk_count);
    uint64_t off = h.data_offset;
    for (size_t c = 0; c < nchunks; c++)
    {
      const size_t n = std::min(DUMP_CHUNK_PARTICLES, (size_t)hp.n - c * DUMP_CHUNK_PARTICLES);
      table[c] = DumpChunkItem{off, (uint32_t)(n * sizeof(DumpTuple)), -(int32_t)n};
      off += n * sizeof(DumpTuple);
    }
    table[nchunks] = DumpChunkItem{off, 0, 0};
    std::FILE* out = std::fopen(filename->c_str(), "wb");
    if (!out) fatal_error("write_dump: cannot write '" + *filename + "'");
    std::fwrite(&h, sizeof h, 1, out); std::fwrite(table.data(), sizeof(DumpChunkItem), table.size(), out);
    std::vector<DumpTuple> buf;
    for (size_t c = 0; c < nchunks; c++)
    {
      const size_t q0 = c * DUMP_CHUNK_PARTICLES, n = (size_t)(-table[c].n_particles);
      buf.assign(n, DumpTuple{});
      for (size_t q = 0; q < n; q++)
      {
        DumpTuple& t = buf[q]; const size_t i = q0 + q;
        for (int a = 0; a < 3; a++) { t.r[a] = hp.f[a][i]; t.v[a] = hp.f[3 + a][i]; }
        t.id = hp.id[i]; t.type = hp.type[i];
      }
      std::fwrite(buf.data(), sizeof(DumpTuple), n, out);
    }
    std::fclose(out);
    std::printf("write_dump: %lld particles, %zu chunks -> %s\n", (long long)hp.n, nchunks, filename->c_str());
  }
};
struct ReadDump : OperatorNode
{
  ADD_SLOT(std::string, filename, INPUT, std::string("output.dump"));
  ADD_SLOT(Domain, domain, INPUT_OUTPUT);
  ADD_SLOT(ParticleSet, pending_particles, INPUT_OUTPUT);
  ADD_SLOT(long, timestep, INPUT_OUTPUT, 0L);
  ADD_SLOT(double, physical_time, INPUT_OUTPUT, 0.0);
  void yaml_initialize(const Params& p) override { if (p.has("filename")) filename.value = std::make_shared<std::string>(p.str("filename")); }
  void execute() override
  {
    std::FILE* in = std::fopen(filename->c_str(), "rb");
    if (!in) fatal_error("read_dump: cannot read '" + *filename + "'");
    DumpHeader h;
    if (std::fread(&h, sizeof h, 1, in) != 1) fatal_error("read_dump: short header");
    if (h.version > 1003) fatal_error("SimDumpHeader::check : bad version number");                       // sim_dump_io.h:181-185
    if (h.tuple_size != sizeof(DumpTuple) || h.nb_fields != 8) fatal_error("SimDumpHeader::check : bad tuple size");      // :187-191
    std::vector<DumpChunkItem> table(h.chunk_count);
    std::fseek(in, (long)h.table_offset, SEEK_SET);
    if (std::fread(table.data(), sizeof(DumpChunkItem), table.size(), in) != table.size()) fatal_error("read_dump: short chunk table");
    Domain& d = *domain;
    d.bounds.bmin = {h.domain.bmin[0], h.domain.bmin[1], h.domain.bmin[2]}; d.bounds.bmax = {h.domain.bmax[0], h.domain.bmax[1], h.domain.bmax[2]};
    d.cell_size = h.domain.cell_size; d.grid_dims = {h.domain.grid_dims[0], h.domain.grid_dims[1], h.domain.grid_dims[2]};
    for (int a = 0; a < 3; a++) d.periodic[a] = (h.domain.flags >> a) & 1u;
    d.expandable = (h.domain.flags >> 3) & 1u;
    ParticleSet& ps = *pending_particles;
    ps = ParticleSet{};
    std::vector<DumpTuple> buf;
    for (const DumpChunkItem& c : table)
    {
      const size_t n = (size_t)std::abs(c.n_particles);
      if (n == 0) continue;
      buf.resize(n);
      std::fseek(in, (long)c.global_offset, SEEK_SET);
      if (std::fread(buf.data(), sizeof(DumpTuple), n, in) != n) fatal_error("read_dump: short chunk");
      for (const DumpTuple& t : buf)
      {
        ps.rx.push_back(t.r[0]); ps.ry.push_back(t.r[1]); ps.rz.push_back(t.r[2]); ps.vx.push_back(t.v[0]); ps.vy.push_back(t.v[1]); ps.vz.push_back(t.v[2]);
        ps.id.push_back(t.id); ps.type.push_back(t.type);
      }
    }
    std::fclose(in);
    if (ps.rx.size() != h.nb_particles) fatal_error("read_dump: particle count does not match the header");
    *timestep = (long)h.time_step; *physical_time = h.time;
    std::printf("read_dump: %zu particles, time step %ld <- %s\n", ps.rx.size(), *timestep, filename->c_str());
  }
};

struct CheckValues : OperatorNode
{
  ADD_SLOT(Grid, grid, INPUT);
  ADD_SLOT(Domain, domain, INPUT, REQUIRED);
  ADD_SLOT(std::string, file, INPUT, std::string("check_values.dat"));
  ADD_SLOT(long, samples, INPUT, 128L);
  ADD_SLOT(double, pos_threshold, INPUT, 1e-5);
  ADD_SLOT(double, acc_threshold, INPUT, 1e-5);
  ADD_SLOT(double, vel_threshold, INPUT, 1e-5);
  ADD_SLOT(double, max_error, OUTPUT);
  void yaml_initialize(const Params& p) override
  {
    if (p.has("file")) { file.value = std::make_shared<std::string>(p.str("file")); }
    if (p.has("samples")) { samples.value = std::make_shared<long>((long)p.quantity("samples")); }
    XNB_PARAM_QUANTITY(p, pos_threshold); XNB_PARAM_QUANTITY(p, acc_threshold); XNB_PARAM_QUANTITY(p, vel_threshold);
  }
  void execute() override
  {
    std::ifstream in(*file);
    const int64_t n = xnb_num_inner(grid->ctx);
    std::vector<double> f[9]; for (auto& v : f) v.resize((size_t)n);
    std::vector<uint64_t> ids((size_t)n);
    ck(grid->ctx, xnb_get_particles(grid->ctx, 0, n, f[0].data(), f[1].data(), f[2].data(), f[3].data(), f[4].data(), f[5].data(), f[6].data(), f[7].data(), f[8].data(), ids.data(), nullptr, nullptr), "check_values");
    if (!in)
    {
      // no reference file yet: sample `samples` distinct inner particles and write "[id, r, a, v]" rows of hex floats sorted by
      // id (check_values.cpp:146-212,296-303,363-386; row format core/yaml_check_particles.h:68-87).  The reference draws the
      // samples from its run-time random engine; a fixed seed makes the file reproducible here.
      const size_t want = (size_t)std::min<int64_t>(std::max<long>(*samples, 1L), n);
      std::mt19937_64 re(20240613u);
      std::set<uint64_t> chosen;
      std::vector<size_t> rows_q;
      while (rows_q.size() < want) { const size_t q = (size_t)(re() % (uint64_t)n); if (chosen.insert(ids[q]).second) rows_q.push_back(q); }
      std::sort(rows_q.begin(), rows_q.end(), [&](size_t a, size_t b) { return ids[a] < ids[b]; });
      std::ofstream fout(*file);
      if (!fout) fatal_error("check_values: cannot write '" + *file + "'");
      char date[64]; { time_t now; time(&now); strftime(date, sizeof date, "%d-%m-%Y %H:%M:%S", localtime(&now)); }
      fout << "date: '" << date << "'\nlength_unit: 1.0 ang\nvalues:\n";
      for (size_t q : rows_q)
      {
        fout << "- [" << ids[q];
        const int order[9] = {0, 1, 2, 6, 7, 8, 3, 4, 5};      // r, a, v
        for (int c : order) { char buf[64]; std::snprintf(buf, sizeof buf, ", %a", f[c][q]); fout << buf; }
        fout << "]\n";
      }
      *max_error = 0.0;
      std::printf("check_values: wrote %zu reference particles to %s\n", rows_q.size(), file->c_str());
      return;
    }
    std::map<uint64_t, size_t> where; for (size_t q = 0; q < (size_t)n; q++) where[ids[q]] = q;
    const Domain& d = *domain;
    const double L[3] = {d.bounds.bmax.x - d.bounds.bmin.x, d.bounds.bmax.y - d.bounds.bmin.y, d.bounds.bmax.z - d.bounds.bmin.z};
    // the file is a YAML list of rows "[ id , x , y , z , ax , ay , az , vx , vy , vz ]" of hex floats: take every number
    std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    std::vector<std::string> rows; { std::string cur; int depth = 0; for (char ch : text) { if (ch == '[') { depth++; if (depth == 1) cur.clear(); else cur += ' '; } else if (ch == ']') { depth--; if (depth == 0) rows.push_back(cur); } else if (depth >= 1) cur += (ch == ',' ? ' ' : ch); } }
    double worst = 0; size_t checked = 0;
    for (const std::string& row : rows)
    {
      std::istringstream is(row); std::vector<double> v; std::string tok;
      while (is >> tok) { char* end = nullptr; const double x = std::strtod(tok.c_str(), &end); if (end != tok.c_str()) v.push_back(x); }
      if (v.size() != 10) continue;
      auto it = where.find((uint64_t)v[0]);
      if (it == where.end()) fatal_error("check_values: particle id missing");
      const size_t q = it->second;
      for (int c = 0; c < 3; c++)
      {
        double dr = f[c][q] - v[1 + c]; dr -= L[c] * std::round(dr / L[c]);          // periodic un-wrap (check_values.cpp:244,255-257)
        const double da = f[6 + c][q] - v[4 + c], dv = f[3 + c][q] - v[7 + c];
        if (std::fabs(dr) > *pos_threshold || std::fabs(da) > *acc_threshold || std::fabs(dv) > *vel_threshold)
        {
          char msg[256];
          std::snprintf(msg, sizeof msg, "check_values: particle %llu differs from the reference values beyond the thresholds (axis %d: dr %.3e da %.3e dv %.3e)",
                        (unsigned long long)v[0], c, dr, da, dv);
          fatal_error(msg);
        }
        worst = std::max(worst, std::max(std::fabs(dr), std::max(std::fabs(da), std::fabs(dv))));
      }
      checked++;
    }
    if (checked == 0) fatal_error("check_values: no sample rows in '" + *file + "'");
    *max_error = worst;
    std::printf("check_values: %zu particles within thresholds, max abs error %.3e\n", checked, worst);
  }
};

} // namespace

void register_hot_path_operators()
{
  static bool done = false;
  if (done) return;
  done = true;
  OperatorNodeFactory* f = OperatorNodeFactory::instance();
  f->register_factory("domain", make_simple_operator<DomainOp>());
  f->register_factory("init_rcb_grid", make_simple_operator<InitRcbGrid>());
  f->register_factory("particle_type_add_properties", make_simple_operator<ParticleTypeAddProperties>());
  f->register_factory("lattice", make_simple_operator<Lattice>());
  f->register_factory("gaussian_noise_r", make_simple_operator<GaussianNoiseR>());
  f->register_factory("nbh_dist", make_simple_operator<NeighborDistance>());
  f->register_factory("move_particles", make_simple_operator<MoveParticles>());
  f->register_factory("migrate_cell_particles", make_simple_operator<Nop>());    // hand-off happens inside move_particles (static blocks)
  f->register_factory("rebuild_amr", make_simple_operator<RebuildAmr>());
  f->register_factory("backup_r", make_simple_operator<BackupR>());
  f->register_factory("ghost_comm_scheme", make_simple_operator<GhostCommSchemeOp>());
  f->register_factory("ghost_update_all", make_simple_operator<GhostUpdate<true>>());
  f->register_factory("ghost_update_r", make_simple_operator<GhostUpdate<false>>());
  f->register_factory("amr_grid_pairs", make_simple_operator<AmrGridPairs>());   // the cache object for consumers; the builds themselves prune with measured boxes (xnb_nbh_big.cuh)
  f->register_factory("chunk_neighbors", make_simple_operator<BuildChunkNeighbors>());
  f->register_factory("resize_particle_locks", make_simple_operator<Nop>());     // ComputePairOptionalLocks<false>: LJ takes no locks
  f->register_factory("zero_particle_force", make_simple_operator<ZeroParticleForce>());
  f->register_factory("lennard_jones_force", make_simple_operator<LennardJonesForce<false>>());
  f->register_factory("load_balance_rcb", make_simple_operator<LoadBalanceRCB>());
  f->register_factory("gravitational_force", make_simple_operator<GravitationalForce>());
  f->register_factory("average_neighbors_scalar", make_simple_operator<AverageNeighborsScalar>());
  f->register_factory("lennard_jones_force_symmetric", make_simple_operator<LennardJonesForce<true>>());
  f->register_factory("update_force_from_ghost", make_simple_operator<UpdateForceFromGhost>());   // adds zeros after a full-list sweep (ghost forces are 0 then)
  f->register_factory("divide_force_by_type_scalar", make_simple_operator<DivideForceByTypeScalar>());
  f->register_factory("push_f_v_r", make_simple_operator<PushFVR>());
  f->register_factory("push_f_v", make_simple_operator<PushFV>());
  f->register_factory("particle_displ_over", make_simple_operator<ParticleDisplOver>());
  f->register_factory("check_values", make_simple_operator<CheckValues>());
  f->register_factory("write_xyz", make_simple_operator<WriteXYZ>());
  f->register_factory("write_dump", make_simple_operator<WriteDump>());
  f->register_factory("read_dump", make_simple_operator<ReadDump>());
}

}} // namespace xnb::host
