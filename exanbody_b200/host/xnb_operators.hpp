// xnb_operators.hpp -- C++17 host-side mirror of the reference's operator interface for the LJ hot path.
//
// The reference composes a simulation from OperatorNode subclasses declared with ADD_SLOT(Type, name, DIRECTION, default,
// DocString) and registered by name in OperatorNodeFactory (onika::scg; e.g. contribs/md/lennard_jones/lennard_jones.cu:171-223,
// src/particle_neighbors/chunk_neighbors.cpp:48-90).  This header keeps that surface for the operators of the hot path:
// same operator names, same slot names / directions / defaults / YAML keys, slots dereferenced with *slot and slot->,
// connections made BY NAME in the enclosing graph (as the YAML batches do), fatal_error()-style abort on failure.
// Every execute() is a thin call into the C-ABI of include/xnb_hotpath.h: no arithmetic of the path lives here.
// It is NOT a re-implementation of onika (no YAML graph loader, no plugin loader, no task scheduler): `Batch` runs a list
// of operators in order, which is all the default LJ decks need (data/config/*.msp).
#pragma once
#include "../../include/xnb_hotpath.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <typeindex>
#include <utility>
#include <vector>

namespace xnb { namespace host {

enum SlotDirection { INPUT, OUTPUT, INPUT_OUTPUT, PRIVATE };
struct Required {};
constexpr Required REQUIRED{};
struct Optional {};
constexpr Optional OPTIONAL{};
struct DocString { const char* text; };

[[noreturn]] void fatal_error(const std::string& msg);     // reference: fatal_error() << msg << std::endl (aborts)

// ---- values that travel through slots ------------------------------------------------------------------------------
struct IJK { int64_t i = 0, j = 0, k = 0; };
struct Vec3d { double x = 0, y = 0, z = 0; };
struct AABB { Vec3d bmin, bmax; };

// core/domain.h:36-129 (identity xform only)
struct Domain
{
  AABB bounds; double cell_size = 0; IJK grid_dims; bool periodic[3] = {true, true, true}; bool expandable = false;
};
struct LennardJonesParms { double epsilon = 0, sigma = 0; };                 // lennard_jones.cu:40-44
struct GravitationalParms { double G = 0.0; };                              // contribs/pi/gravitational_force.cu:42-45
struct ChunkNeighborsConfig                                                  // chunk_neighbors_config.h:27-39
{
  bool free_scratch_memory = false, build_particle_offset = true, subcell_compaction = true, half_symmetric = false, skip_ghosts = false;
  unsigned chunk_size = 1; double stream_prealloc_factor = 1.05;
};
struct UpdateGhostConfig { bool gpu_buffer_pack = true, staging_buffer = false; };   // update_ghost_config.h:28-37
// The grid and everything derived from it live in one xnb_ctx (one sub-domain on one GPU); the slot types below are
// views on it, so that operators keep the reference's slot signature.
struct Grid { xnb_ctx* ctx = nullptr; int device = 0; ~Grid(); Grid() = default; Grid(const Grid&) = delete; Grid& operator=(const Grid&) = delete;
              int64_t number_of_particles() const; int64_t number_of_cells() const;
              std::string generic_real_field; /* name of the generic real field the ctx holds (field::mk_generic_real) */ };
struct GridChunkNeighbors { xnb_ctx* ctx = nullptr; const uint16_t* const* cell_stream = nullptr; const uint32_t* cell_stream_size = nullptr;
                            uint32_t max_neighbors = 0; size_t number_of_cells() const; };     // chunk_neighbors.h:42-120 (device views)
struct AmrGrid { xnb_ctx* ctx = nullptr; };
// amr_grid_algorithm.h:439-453: per (resolution pair, neighbour cell offset) the sub-cell pairs closer than max_dist, (a, b) codes interleaved
struct AmrSubCellPairCache { xnb_ctx* ctx = nullptr; size_t m_max_res = 0; double m_cell_size = 0.0, m_max_dist = 0.0;
                             std::vector<uint64_t> m_list_offsets; std::vector<uint16_t> m_pair_ab; size_t number_of_lists() const { return m_list_offsets.empty() ? 0 : m_list_offsets.size() - 1; } };
struct PositionBackupData { xnb_ctx* ctx = nullptr; };
struct GhostCommunicationScheme { xnb_ctx* ctx = nullptr; };
struct ParticleTypeProperties { std::vector<double> mass; std::vector<std::string> names; };                // per-type scalars (vec3_typescalar_op.cu:71-122)
struct ParticleSet { std::vector<double> rx, ry, rz, vx, vy, vz; std::vector<uint64_t> id; std::vector<uint8_t> type; };

// ---- minimal parameter node: the `{ key: value, ... }` flow maps operators take in .msp decks -------------------------
struct Params
{
  std::map<std::string, std::string> kv;
  static Params parse(const std::string& flow_map);            // "{ epsilon: 0.3729 eV , sigma: 2.2808 ang }", nested maps flattened with '.'
  bool has(const std::string& k) const { return kv.count(k) != 0; }
  double quantity(const std::string& k) const;                  // number with optional unit (ang, nm, um, ps, fs, Da, eV, J), internal units ang/ps/Da
  bool boolean(const std::string& k) const;
  std::string str(const std::string& k) const { return kv.at(k); }
  std::vector<double> quantities(const std::string& k) const;   // "[ a , b , c ]"
};
double convert_quantity(const std::string& text);              // onika::physics::Quantity::convert() for the units the LJ decks use

// ---- slots / operators / factory -----------------------------------------------------------------------------------
class OperatorNode;
struct SlotBase
{
  std::string name; SlotDirection dir; bool required; const char* doc; std::type_index type;
  SlotBase(OperatorNode* op, const char* n, SlotDirection d, bool req, const char* doc_, std::type_index t);
  virtual ~SlotBase() = default;
  virtual std::shared_ptr<void> make_default() const = 0;
  virtual void bind(const std::shared_ptr<void>& p) = 0;
  virtual bool has_value() const = 0;
};
template <class T>
struct Slot : SlotBase
{
  std::shared_ptr<T> value; std::function<T*()> dflt;
  Slot(OperatorNode* op, const char* n, SlotDirection d) : SlotBase(op, n, d, false, "", typeid(T)) { dflt = [] { return new T(); }; }
  Slot(OperatorNode* op, const char* n, SlotDirection d, DocString ds) : SlotBase(op, n, d, false, ds.text, typeid(T)) { dflt = [] { return new T(); }; }
  Slot(OperatorNode* op, const char* n, SlotDirection d, Required, DocString ds = {""}) : SlotBase(op, n, d, true, ds.text, typeid(T)) {}
  Slot(OperatorNode* op, const char* n, SlotDirection d, Optional, DocString ds = {""}) : SlotBase(op, n, d, false, ds.text, typeid(T)) {}
  template <class U, class = decltype(T(std::declval<U>()))>
  Slot(OperatorNode* op, const char* n, SlotDirection d, U v, DocString ds = {""}) : SlotBase(op, n, d, false, ds.text, typeid(T)) { dflt = [v] { return new T(v); }; }
  std::shared_ptr<void> make_default() const override { return dflt ? std::shared_ptr<void>(std::shared_ptr<T>(dflt())) : nullptr; }
  void bind(const std::shared_ptr<void>& p) override { value = std::static_pointer_cast<T>(p); }
  bool has_value() const override { return (bool)value; }
  T& operator*() const { if (!value) fatal_error("slot '" + name + "' has no value"); return *value; }
  T* operator->() const { return &**this; }
  T* get_pointer() const { return value.get(); }
};
#define ADD_SLOT(T, name, ...) ::xnb::host::Slot<T> name { this, #name, __VA_ARGS__ }

class OperatorNode
{
public:
  virtual ~OperatorNode() = default;
  virtual void execute() = 0;
  virtual void yaml_initialize(const Params&) {}          // operator specific parameters (`op: { ... }` in a deck)
  std::vector<SlotBase*> slots;                            // filled by ADD_SLOT in declaration order
  std::string name;
  SlotBase* slot(const std::string& n) const { for (auto* s : slots) if (s->name == n) return s; return nullptr; }
  void* stream = nullptr;                                  // cudaStream_t the operator enqueues on (parallel_execution_context())
};

class OperatorNodeFactory
{
public:
  using Maker = std::function<std::unique_ptr<OperatorNode>()>;
  static OperatorNodeFactory* instance();
  void register_factory(const std::string& name, Maker m) { makers_[name] = std::move(m); }
  std::unique_ptr<OperatorNode> make_operator(const std::string& name) const;
  std::vector<std::string> available_operators() const;
private:
  std::map<std::string, Maker> makers_;
};
template <class Op> OperatorNodeFactory::Maker make_simple_operator() { return [] { return std::unique_ptr<OperatorNode>(new Op()); }; }

// ---- a batch = operators run in sequence, slots connected by name (what the YAML batches of data/config/*.msp express) --
class Batch
{
public:
  // add an operator by its registered name; params = the deck's `{ ... }` for it ("" = none); rebind = { slot: graph name, ... }
  OperatorNode* add(const std::string& op_name, const std::string& params = "", const std::map<std::string, std::string>& rebind = {});
  void execute();                                          // run every operator in order
  struct Entry { std::shared_ptr<void> p; std::type_index t = std::type_index(typeid(void)); };
  template <class T> std::shared_ptr<T> value(const std::string& graph_name)   // access (and create) a named value of the graph
  {
    Entry& e = table()[graph_name];
    if (!e.p) { e.p = std::shared_ptr<void>(std::make_shared<T>()); e.t = std::type_index(typeid(T)); }
    if (e.t != std::type_index(typeid(T))) fatal_error("graph value '" + graph_name + "' requested with another type");
    return std::static_pointer_cast<T>(e.p);
  }
  std::shared_ptr<Batch> sub_batch() { auto b = std::make_shared<Batch>(); b->parent_ = this; b->stream_ = stream_; return b; }   // shares the named values
  void set_stream(void* s) { stream_ = s; }
  const std::vector<std::unique_ptr<OperatorNode>>& operators() const { return ops_; }
private:
  std::map<std::string, Entry>& table() { return parent_ ? parent_->table() : values_; }
  std::vector<std::unique_ptr<OperatorNode>> ops_;
  std::map<std::string, Entry> values_;
  Batch* parent_ = nullptr;
  void* stream_ = nullptr;
};

void register_hot_path_operators();       // idempotent; the reference does this with ONIKA_AUTORUN_INIT static initialisers

}} // namespace xnb::host
