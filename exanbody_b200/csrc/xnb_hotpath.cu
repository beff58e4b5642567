// xnb_hotpath.cu -- C-ABI (include/xnb_hotpath.h) and host-side orchestration of the sm_100a kernels.
// One xnb_ctx = one sub-domain on one GPU.  No CPU compute path exists here: without a CUDA device every entry fails.
#include "../../include/xnb_hotpath.h"
#include "xnb_kernels.cuh"
#include "xnb_sweep_cl.cuh"
#include "xnb_nbh_bits.cuh"
#include "xnb_nbh_big.cuh"
#include "xnb_sweep_pl.cuh"
#include "xnb_pair_generic.cuh"
#include "xnb_peer_halo.cuh"
#include "xnb_host_decomp.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>

using namespace xnb;

// ---------------------------------------------------------------------------------------------------------------------
// NCCL through dlsym: the library is resolved from the process (torch loads its bundled libnccl) or dlopen'ed; there is
// no link-time dependency so the C-ABI loads on a box without NCCL/GPU.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
typedef void* ncclComm_t;
typedef int ncclResult_t;
enum { nccl_uint8 = 1, nccl_uint32 = 3, nccl_uint64 = 5, nccl_float64 = 8 };
enum { nccl_sum = 0 };
struct NcclApi
{
  bool loaded = false, ok = false;
  ncclResult_t (*GetUniqueId)(void*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, /* ncclUniqueId by value: 128 bytes */ struct Id128, int) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
struct Id128 { char b[128]; };
NcclApi g_nccl;

bool nccl_load()
{
  if (g_nccl.loaded) return g_nccl.ok;
  g_nccl.loaded = true;
  void* h = RTLD_DEFAULT;
  if (!dlsym(h, "ncclSend"))
  {
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return false;
  }
#define XNB_SYM(field, name) *(void**)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) return false;
  XNB_SYM(GetUniqueId, "ncclGetUniqueId") XNB_SYM(CommInitRank, "ncclCommInitRank") XNB_SYM(GroupStart, "ncclGroupStart")
  XNB_SYM(GroupEnd, "ncclGroupEnd") XNB_SYM(Send, "ncclSend") XNB_SYM(Recv, "ncclRecv") XNB_SYM(AllReduce, "ncclAllReduce")
  XNB_SYM(AllGather, "ncclAllGather") XNB_SYM(CommDestroy, "ncclCommDestroy") XNB_SYM(GetErrorString, "ncclGetErrorString")
#undef XNB_SYM
  g_nccl.ok = true;
  return true;
}

std::string g_create_error;
bool env_flag(const char* name) { const char* v = getenv(name); return v && *v && *v != '0'; }
int env_int(const char* name) { const char* v = getenv(name); return v ? atoi(v) : 0; }

constexpr int XNB_MAX_DEVICES = 64;
long long g_device_allocs = 0;      // cudaMalloc calls (xnb_device_allocations)

template <class T>
struct DBuf
{
  T* p = nullptr; size_t cap = 0;
  ~DBuf() { if (p) cudaFree(p); }
  // grow to hold n elements; keep: preserve the first `keep` elements
  cudaError_t ensure(size_t n, size_t keep = 0, double slack = 1.0)
  {
    if (n <= cap) return cudaSuccess;
    size_t ncap = (size_t)((double)n * slack) + 16;
    T* q = nullptr;
    cudaError_t e = cudaMalloc(&q, ncap * sizeof(T));
    if (e != cudaSuccess) return e;
    g_device_allocs++;
    if (p && keep) { e = cudaMemcpy(q, p, std::min(keep, cap) * sizeof(T), cudaMemcpyDeviceToDevice); if (e != cudaSuccess) { cudaFree(q); return e; } }
    if (p) cudaFree(p);
    p = q; cap = ncap;
    return cudaSuccess;
  }
};

} // namespace

struct xnb_ctx
{
  int device = 0;
  std::string err;
  // ---- Domain / decomposition
  double dmin[3] = {0, 0, 0}, dmax[3] = {0, 0, 0}, cs = 0;
  int64_t ddims[3] = {0, 0, 0};
  int periodic[3] = {1, 1, 1};
  bool have_domain = false, have_block = false, have_dist = false, grid_ready = false;
  int rank = 0, nranks = 1;
  std::vector<Block> blocks;
  double rcut_max = 0, rcut_inc = 0, nbh_dist = 0, max_displ = 0, ghost_dist = 0;
  double sub_grid_density = 6.5;
  GridP g{};
  // ---- particles (double buffered SoA)
  DBuf<double> f64[2][9];
  DBuf<unsigned long long> idb[2];
  DBuf<uint8_t> typeb[2];
  int cur = 0;
  int64_t n_inner = 0, n_total = 0;
  DBuf<uint32_t> atom_cell[2]; int cur_ac = 0;
  DBuf<uint32_t> key, rnk, perm, perm2, leave_list;
  DBuf<unsigned long long> sort_keys; DBuf<uint32_t> sort_srcs;     // in-cell sort scratch for cells of more than CELLSORT_MAX particles
  DBuf<uint32_t> cell_start, cell_count;
  DBuf<uint32_t> backup;
  DBuf<double> mass; int n_types = 0;
  // ---- AMR
  DBuf<uint8_t> side_lut; double side_lut_density = 0;
  DBuf<uint32_t> sg_size; DBuf<unsigned long long> sub_grid_start; DBuf<uint32_t> sub_grid_cells;
  int64_t n_sub_grid_cells = 0; uint32_t max_side = 1; bool amr_current = false;   // amr_current: the tables describe the present in-cell order
  // ---- ghosts
  std::vector<int> send_first, recv_first;            // per partner rank [nranks+1] ranges into the item arrays
  int n_send_items = 0, n_recv_items = 0;
  DBuf<uint32_t> it_src_cell, it_flags, it_partner, it_count, it_offset;
  DBuf<double> it_outer;
  DBuf<uint32_t> rc_dst_cell, rc_count, rc_offset;
  DBuf<uint32_t> send_src; DBuf<uint16_t> send_flags;
  DBuf<double> stage, rstage, lb_costs, generic_field; int64_t generic_field_n = -1;
  DBuf<uint32_t> d_ghost_base;                          // device copy of h_send_base | h_recv_base (nranks + 1 entries each)
  std::vector<uint32_t> h_send_base, h_recv_base, h_ghost_base;      // per partner particle offsets [nranks+1]; both, back to back
  int64_t n_send = 0, n_ghost = 0;
  // ---- neighbours
  DBuf<uint32_t> nb_len, nb_cnt, nb_off, stream_size, stream_size_padded, cell_stream_bytes;
  DBuf<unsigned long long> stream_off;
  DBuf<uint16_t> pool; DBuf<uint16_t*> cell_stream;
  bool nbh_half_symmetric = false, nbh_skip_ghosts = false;       // ChunkNeighborsConfig (xnb_set_chunk_neighbors_config)
  int nbh_cap_l = 0; uint32_t nbh_slot_words = 0; bool nbh_full_cap = false;   // capacities of the tiled build (grow on demand)
  // ---- compiled lists of the pair sweep (xnb_sweep_cl.cuh): derived from the streams after every rebuild
  struct ClCfg { bool valid = false, ghost = false; int planes = 0, cap_pl = 0; ClTileP tp{}; int threads = 0, var = 0; size_t smem = 0; unsigned blocks = 0; uint32_t rows = 0; int64_t candidates = 0;
                 unsigned n_interior = 0, n_boundary = 0; };    // tiles whose halo box holds no ghost cell / the others (cl_tile_list: interior first)
  ClCfg cl;
  bool in_rebuild_chain = false;           // move_and_update_full: the operators' small read-backs are merged (one host wait each for binning, ghosts, lists)
  bool cell_stats_valid = false;           // max_cell_count / n_nonempty_inner describe the current cell counts (ghost cells included)
  NextHalfP next_half{};                   // operands of a MODE 2 sweep (xnb_run_steps sets them right before the launch)
  DBuf<uint2> cl_groups; DBuf<uint16_t> cl_rows; uint32_t cl_cap_rows = 0; DBuf<uint32_t> cl_tile_list;
  // ---- k_nbh_bits (xnb_nbh_bits.cuh): masks parked between its two phases, capacities that worked last time, lazily built ghost-cell lists
  int nb_cap_l = 0, nb_cap_trips = 0; bool ghost_lists = false;       // ghost_lists: the streams of the ghost cells are current
  int nb_cap32 = 0, nb_scratch_rows = 0; DBuf<uint32_t> nb_scratch;
  int64_t nbh_builds = 0, steps_since_nbh = 0, last_nbh_interval = -1;   // first-half kicks since the last neighbour build / between the last two builds (-1: unknown)          // k_nbh_big: accept masks parked between its count and fill passes
  struct NbGhostCfg { ClTileP tp{}; int cap_l = 0, cap32 = 0; bool have = false; } nb_ghost;
  cudaStream_t st_comm = nullptr; cudaEvent_t ev_pos = nullptr, ev_ghost = nullptr;      // halo exchange overlapped with the interior tiles
  int64_t n_nonempty_inner = 0;
  int64_t pool_used = 0; uint32_t max_neighbors = 0, max_cell_count = 0, max_stream = 0; double avg_stream = 0; bool have_nbh = false;
  // ---- misc device scalars
  DBuf<unsigned long long> scan_tmp64; DBuf<uint32_t> scan_tmp32;
  DBuf<unsigned long long> d_scalars64;   // [0] displacement counter, [1] scan total
  DBuf<uint32_t> d_scalars32;             // [0] error word, [1] leave_count, [2] max_side, [3] max_nbh, [4] scan total, [8..8+64) migrate counts
  DBuf<double> ev_partials, ev_scratch, ev_ekin;
  DBuf<int> d_blocks;
  DBuf<uint32_t> mig_rank, mig_pos, mig_base; std::vector<uint32_t> h_mig_base;     // mig_base: per destination send offsets | per source receive offsets (nranks + 1 each)
  // ---- xnb_step_host: positions (and ids) go back to the host on their own stream as soon as they are final
  struct HostOut { bool active = false, issued = false, id_always = false, id_copied = false, overflow = false; double* r[3] = {nullptr, nullptr, nullptr}; uint64_t* id = nullptr; size_t capacity = 0; };
  HostOut hout; cudaStream_t st_d2h = nullptr; cudaEvent_t ev_d2h_go = nullptr, ev_d2h_done = nullptr;
  void* h_pinned = nullptr;               // 4 KB pinned scratch for small read-backs ([1024..1032): displacement count of xnb_run_steps)
  cudaEvent_t ev_flag = nullptr;
  // ---- NCCL
  ncclComm_t comm = nullptr; bool own_comm = false;
  // ---- halo over NVLink peer memory (xnb_peer_halo.cuh): my mailbox, the partners' mailboxes mapped here, exchange counters
  struct PeerHalo
  {
    bool tried = false, enabled = false;
    void* box = nullptr; size_t cap_words = 0; unsigned long long gen = 0;   // my mailbox: header + 2 halves of cap_words words
    std::vector<void*> retired;                                             // outgrown mailboxes (partners may still have them mapped)
    std::vector<void*> peer_ptr; std::vector<unsigned long long> peer_gen;   // partner mailboxes as mapped in this process
    DBuf<PeerSlot> d_slots; DBuf<unsigned char> d_rec;                      // device table for the kernels; all-gather scratch
    std::vector<PeerSlot> h_slots;                                          // source of the asynchronous upload of d_slots
    unsigned long long epoch = 0, epoch_over = 0;                           // exchanges done (halo / displacement sum): same on every rank
    unsigned long long timeout_ns = PEER_TIMEOUT_NS;
  };
  PeerHalo peer;
  // ---- counters
  int64_t launches = 0, rebuilds = 0;
  int pair_functor = XNB_FUNCTOR_LJ;   // xnb_set_pair_functor
  // device-side timing without synchronisation: per category a pool of event pairs recorded on the launching stream and
  // summed at xnb_timing_read (categories XNB_T_* of the header)
  bool timing = false;
  struct TPool { std::vector<cudaEvent_t> ev; size_t used = 0; double ms = 0; int64_t n = 0; int64_t dropped = 0; };
  TPool tpool[XNB_T_COUNT];

  int fail(int code, const std::string& m) { err = m; return code; }
  ParticlesP P(int which)
  {
    ParticlesP p;
    p.rx = f64[which][0].p; p.ry = f64[which][1].p; p.rz = f64[which][2].p;
    p.vx = f64[which][3].p; p.vy = f64[which][4].p; p.vz = f64[which][5].p;
    p.fx = f64[which][6].p; p.fy = f64[which][7].p; p.fz = f64[which][8].p;
    p.id = idb[which].p; p.type = typeb[which].p;
    return p;
  }
};

#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return c->fail(XNB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } while (0)
#define NK(call) do { ncclResult_t r__ = (call); if (r__ != 0) return c->fail(XNB_ERR_NCCL, std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error")); } while (0)
// timing scope helpers: record an event pair around a group of launches (no host synchronisation)
static int verlet_first_half(xnb_ctx* c, double dt, cudaStream_t st, unsigned long long* counter);
static int t_begin(xnb_ctx* c, int cat, cudaStream_t st);
static int t_end(xnb_ctx* c, int cat, cudaStream_t st);
#define LAUNCH(kernel, grid, block, stream, ...) do { kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__); c->launches++; CK(cudaGetLastError()); } while (0)
static inline unsigned nblk(int64_t n, int b) { return (unsigned)std::max<int64_t>((n + b - 1) / b, 1); }

static const size_t T_POOL_MAX = 8192;   // event pairs per category between two xnb_timing_read calls
static int t_begin(xnb_ctx* c, int cat, cudaStream_t st)
{
  if (!c->timing) return 0;
  xnb_ctx::TPool& t = c->tpool[cat];
  if (t.used >= T_POOL_MAX) { t.dropped++; return 0; }
  if (t.ev.size() < 2 * (t.used + 1)) { cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); t.ev.push_back(a); t.ev.push_back(b); }
  CK(cudaEventRecord(t.ev[2 * t.used], st));
  return 0;
}
static int t_end(xnb_ctx* c, int cat, cudaStream_t st)
{
  if (!c->timing) return 0;
  xnb_ctx::TPool& t = c->tpool[cat];
  if (t.used >= T_POOL_MAX || t.ev.size() < 2 * (t.used + 1)) return 0;
  CK(cudaEventRecord(t.ev[2 * t.used + 1], st));
  t.used++;
  return 0;
}
static int t_collect(xnb_ctx* c)
{
  for (int cat = 0; cat < XNB_T_COUNT; cat++)
  {
    xnb_ctx::TPool& t = c->tpool[cat];
    for (size_t q = 0; q < t.used; q++)
    {
      CK(cudaEventSynchronize(t.ev[2 * q + 1]));
      float ms = 0; CK(cudaEventElapsedTime(&ms, t.ev[2 * q], t.ev[2 * q + 1]));
      t.ms += ms; t.n++;
    }
    t.used = 0;
  }
  return 0;
}

namespace {

// exclusive scan of n elements; `total` (device pointer, may be null) receives the grand total
template <class TIn, class T>
int scan_exclusive(xnb_ctx* c, const TIn* in, T* out, size_t n, T* d_total, DBuf<T>& tmp, cudaStream_t st)
{
  const size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  CK(tmp.ensure(std::max<size_t>(tiles, 1)));
  if (n == 0) { if (d_total) CK(cudaMemsetAsync(d_total, 0, sizeof(T), st)); return 0; }
  LAUNCH((k_scan_tiles<TIn, T>), (unsigned)tiles, SCAN_BLOCK, st, in, out, tmp.p, n);
  LAUNCH((k_scan_sums<T>), 1, SCAN_BLOCK, st, tmp.p, tiles, d_total);
  LAUNCH((k_scan_add<T>), (unsigned)tiles, SCAN_BLOCK, st, out, tmp.p, n);
  return 0;
}

int ensure_particle_capacity(xnb_ctx* c, size_t n, size_t keep)
{
  for (int w = 0; w < 2; w++)
  {
    const size_t k = (w == c->cur) ? keep : 0;
    for (int f = 0; f < 9; f++) CK(c->f64[w][f].ensure(n, k, 1.1));
    CK(c->idb[w].ensure(n, k, 1.1));
    CK(c->typeb[w].ensure(n, k, 1.1));
    CK(c->atom_cell[w].ensure(n, (w == c->cur_ac) ? keep : 0, 1.1));
  }
  return 0;
}

// builds GridP, the AMR side table and the static ghost item lists once domain, block and distances are known
// sub_grid_size(n, density) table (amr_grid_algorithm.h:66-78), computed on the host with std::cbrt like the reference;
// re-uploaded when rebuild_amr's sub_grid_density changes (it must not invalidate the binned grid)
int ensure_side_lut(xnb_ctx* c)
{
  if (c->side_lut.p && c->side_lut_density == c->sub_grid_density) return 0;
  std::vector<uint8_t> lut(65536);
  for (size_t n = 0; n < 65536; n++)
  {
    size_t side = 0;
    if (n > 0) { const double s = std::cbrt((double)n / c->sub_grid_density); side = (s < 2.0) ? 1 : std::min((size_t)std::floor(s), (size_t)16); }
    lut[n] = (uint8_t)side;
  }
  CK(c->side_lut.ensure(65536));
  CK(cudaMemcpy(c->side_lut.p, lut.data(), 65536, cudaMemcpyHostToDevice));
  c->side_lut_density = c->sub_grid_density;
  return 0;
}

int ensure_grid(xnb_ctx* c)
{
  if (c->grid_ready) return ensure_side_lut(c);
  if (!c->have_domain) return c->fail(XNB_ERR_INVALID, "xnb_set_domain has not been called");
  if (!c->have_dist) return c->fail(XNB_ERR_INVALID, "xnb_set_nbh_dist has not been called");
  if (!c->have_block)
  {
    c->rank = 0; c->nranks = 1; c->blocks.assign(1, Block{{0, 0, 0}, {c->ddims[0], c->ddims[1], c->ddims[2]}}); c->have_block = true;
  }
  GridP& g = c->g;
  const int gl = (int)std::ceil(c->ghost_dist / c->cs);          // grid.h:110
  const int gap = (int)std::ceil(c->nbh_dist / c->cs);           // amr_grid_algorithm.h:451
  if (gap > 15) return c->fail(XNB_ERR_CAPACITY, "neighbour cell offset beyond +-15 cells (chunk_neighbors.h:140-142)");
  const Block& b = c->blocks[(size_t)c->rank];
  for (int d = 0; d < 3; d++) if (b.e[d] <= b.s[d]) return c->fail(XNB_ERR_INVALID, "Assigned grid block is empty (more ranks than cells along a cut)");
  int64_t ncell = 1;
  for (int d = 0; d < 3; d++)
  {
    g.org[d] = c->dmin[d]; g.dmin[d] = c->dmin[d]; g.dmax[d] = c->dmax[d];
    g.dims[d] = (int)(b.e[d] - b.s[d] + 2 * gl); g.off[d] = (int)(b.s[d] - gl);
    g.ddims[d] = (int)c->ddims[d]; g.bstart[d] = (int)b.s[d]; g.bend[d] = (int)b.e[d]; g.periodic[d] = c->periodic[d];
    ncell *= g.dims[d];
    if (c->periodic[d] && c->ddims[d] < 2 * gl) return c->fail(XNB_ERR_INVALID, "periodic domain thinner than two ghost layers is not supported");
  }
  if (ncell > 0x7fffffff) return c->fail(XNB_ERR_CAPACITY, "too many cells");
  g.cs = c->cs; g.gl = gl; g.n_cells = (int)ncell;
  CK(c->cell_start.ensure((size_t)ncell)); CK(c->cell_count.ensure((size_t)ncell));
  CK(c->d_scalars64.ensure(16)); CK(c->d_scalars32.ensure(128));
  CK(cudaMemset(c->d_scalars64.p, 0, 8 * 8)); CK(cudaMemset(c->d_scalars32.p, 0, 128 * 4));
  CK(cudaMemset(c->cell_count.p, 0, (size_t)ncell * 4)); CK(cudaMemset(c->cell_start.p, 0, (size_t)ncell * 4));
  if (!c->h_pinned) CK(cudaMallocHost(&c->h_pinned, 4096));
  { int rc = ensure_side_lut(c); if (rc) return rc; }
  // decomposition table for migration
  {
    std::vector<int> hb((size_t)c->nranks * 6);
    for (int r = 0; r < c->nranks; r++) for (int d = 0; d < 3; d++) { hb[(size_t)r * 6 + d] = (int)c->blocks[(size_t)r].s[d]; hb[(size_t)r * 6 + 3 + d] = (int)c->blocks[(size_t)r].e[d]; }
    CK(c->d_blocks.ensure(hb.size()));
    CK(cudaMemcpy(c->d_blocks.p, hb.data(), hb.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  // static ghost items
  {
    std::vector<uint32_t> s_src, s_flags, s_partner, r_dst;
    std::vector<double> outer((size_t)c->nranks * 6);
    c->send_first.assign((size_t)c->nranks + 1, 0); c->recv_first.assign((size_t)c->nranks + 1, 0);
    for (int p = 0; p < c->nranks; p++)
    {
      const Block& bp = c->blocks[(size_t)p];
      for (int d = 0; d < 3; d++)
      {
        // block_to_bounds then enlarge by Grid::max_neighbor_distance (update_ghosts_comm_scheme.cpp:405-406)
        outer[(size_t)p * 6 + d] = (c->dmin[d] + (double)bp.s[d] * c->cs) - c->ghost_dist;
        outer[(size_t)p * 6 + 3 + d] = (c->dmin[d] + (double)bp.e[d] * c->cs) + c->ghost_dist;
      }
      std::vector<HostItem> snd, rcv;
      enumerate_sends(c->blocks, c->ddims, c->periodic, c->rank, p, gl, snd);
      enumerate_sends(c->blocks, c->ddims, c->periodic, p, c->rank, gl, rcv);
      c->send_first[(size_t)p] = (int)s_src.size(); c->recv_first[(size_t)p] = (int)r_dst.size();
      for (const HostItem& it : snd) { s_src.push_back(it.src_cell); s_flags.push_back(it.flags); s_partner.push_back((uint32_t)p); }
      for (const HostItem& it : rcv) r_dst.push_back(it.dst_cell);
    }
    c->send_first[(size_t)c->nranks] = (int)s_src.size(); c->recv_first[(size_t)c->nranks] = (int)r_dst.size();
    c->n_send_items = (int)s_src.size(); c->n_recv_items = (int)r_dst.size();
    const size_t ns = std::max<size_t>(s_src.size(), 1), nr = std::max<size_t>(r_dst.size(), 1);
    CK(c->it_src_cell.ensure(ns)); CK(c->it_flags.ensure(ns)); CK(c->it_partner.ensure(ns)); CK(c->it_count.ensure(ns + 1)); CK(c->it_offset.ensure(ns + 1));
    CK(c->rc_dst_cell.ensure(nr)); CK(c->rc_count.ensure(nr + 1)); CK(c->rc_offset.ensure(nr + 1));
    CK(c->it_outer.ensure(outer.size()));
    if (!s_src.empty())
    {
      CK(cudaMemcpy(c->it_src_cell.p, s_src.data(), s_src.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(c->it_flags.p, s_flags.data(), s_flags.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(c->it_partner.p, s_partner.data(), s_partner.size() * 4, cudaMemcpyHostToDevice));
    }
    if (!r_dst.empty()) CK(cudaMemcpy(c->rc_dst_cell.p, r_dst.data(), r_dst.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->it_outer.p, outer.data(), outer.size() * 8, cudaMemcpyHostToDevice));
    c->h_send_base.assign((size_t)c->nranks + 1, 0); c->h_recv_base.assign((size_t)c->nranks + 1, 0);
  }
  c->grid_ready = true;
  return 0;
}

// what the device error word says (and clear it)
int decode_device_errors(xnb_ctx* c, uint32_t e, cudaStream_t st)
{
  if (!e) return 0;
  CK(cudaMemsetAsync(c->d_scalars32.p, 0, 4, st));
  if (e & DERR_LOST_PARTICLE) return c->fail(XNB_ERR_LOST_PARTICLE, "a particle left a non periodic domain (reference: stays in otb_particles)");
  if (e & DERR_CELL_OVERFLOW) return c->fail(XNB_ERR_CAPACITY, "more than 65535 particles in a cell (u16 stream index, chunk_neighbors_execute.h:229)");
  if (e & DERR_GROUP_OVERFLOW) return c->fail(XNB_ERR_CAPACITY, "u16 counter overflow in a neighbour stream (chunk_neighbors_execute.h:362,369)");
  if (e & DERR_SORT_CAPACITY) return c->fail(XNB_ERR_CAPACITY, "in-cell sort: no scratch for a cell of more than 2048 particles");
  if (e & DERR_ID_RANGE) return c->fail(XNB_ERR_CAPACITY, "particle id >= 2^52");
  if (e & DERR_TILE_CAPACITY) return c->fail(XNB_ERR_CAPACITY, "a tile exceeded its shared-memory staging capacity");
  if (e & DERR_PEER_TIMEOUT) return c->fail(XNB_ERR_NCCL, "peer-memory halo: a partner's data did not arrive in time (XNB_PEER_TIMEOUT_MS, default 20 s)");
  return c->fail(XNB_ERR_INVALID, "device error word " + std::to_string(e));
}

int check_device_errors(xnb_ctx* c, cudaStream_t st)
{
  uint32_t* h = (uint32_t*)c->h_pinned;
  CK(cudaMemcpyAsync(h, c->d_scalars32.p, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return decode_device_errors(c, h[0], st);
}

// small synchronous read-back through the pinned scratch page
template <class T>
int read_back(xnb_ctx* c, const T* d, size_t n, T* out, cudaStream_t st)
{
  CK(cudaMemcpyAsync(c->h_pinned, d, n * sizeof(T), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  memcpy(out, c->h_pinned, n * sizeof(T));
  return 0;
}


} // namespace

// =====================================================================================================================
extern "C" {

const char* xnb_version(void) { return "exanbody_b200 hot path 0.1 (sm_100a)"; }

const char* xnb_last_error(const xnb_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int xnb_create(xnb_ctx** out, int device)
{
  if (!out) return XNB_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) { g_create_error = "no CUDA device: the exaNBody B200 hot path has no CPU fallback"; return XNB_ERR_NO_DEVICE; }
  if (device < 0 || device >= ndev) { g_create_error = "invalid CUDA device index"; return XNB_ERR_INVALID; }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return XNB_ERR_CUDA; }
  xnb_ctx* c = new xnb_ctx();
  c->device = device;
  double one = 1.0;
  if (c->mass.ensure(256) != cudaSuccess) { g_create_error = "cudaMalloc failed"; delete c; return XNB_ERR_CUDA; }
  std::vector<double> m(256, one);
  cudaMemcpy(c->mass.p, m.data(), 256 * 8, cudaMemcpyHostToDevice);
  c->n_types = 256;
  *out = c;
  return XNB_OK;
}

void xnb_destroy(xnb_ctx* c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (int cat = 0; cat < XNB_T_COUNT; cat++) for (cudaEvent_t e : c->tpool[cat].ev) cudaEventDestroy(e);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  if (c->ev_flag) cudaEventDestroy(c->ev_flag);
  if (c->ev_pos) cudaEventDestroy(c->ev_pos);
  if (c->ev_ghost) cudaEventDestroy(c->ev_ghost);
  if (c->st_comm) cudaStreamDestroy(c->st_comm);
  if (c->ev_d2h_go) cudaEventDestroy(c->ev_d2h_go);
  if (c->ev_d2h_done) cudaEventDestroy(c->ev_d2h_done);
  if (c->st_d2h) cudaStreamDestroy(c->st_d2h);
  for (size_t p = 0; p < c->peer.peer_ptr.size(); p++) if (c->peer.peer_ptr[p] && (int)p != c->rank) cudaIpcCloseMemHandle(c->peer.peer_ptr[p]);
  if (c->peer.box) cudaFree(c->peer.box);
  for (void* q : c->peer.retired) cudaFree(q);
  if (c->comm && c->own_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  delete c;
}

int xnb_set_domain(xnb_ctx* c, const double bmin[3], const double bmax[3], double cell_size, const int64_t grid_dims[3], const int32_t periodic[3])
{
  if (!c) return XNB_ERR_INVALID;
  if (!(cell_size > 0)) return c->fail(XNB_ERR_INVALID, "cell_size must be > 0");
  for (int d = 0; d < 3; d++)
  {
    c->dmin[d] = bmin[d]; c->dmax[d] = bmax[d]; c->ddims[d] = grid_dims[d]; c->periodic[d] = periodic[d] ? 1 : 0;
    if (grid_dims[d] <= 0) return c->fail(XNB_ERR_INVALID, "grid_dims must be > 0");
    // check_domain (src/core/lib/domain.cpp:82-84): bounds must match grid_dims*cell_size
    if (std::fabs(1.0 - ((double)grid_dims[d] * cell_size) / (bmax[d] - bmin[d])) > 1e-12) return c->fail(XNB_ERR_INVALID, "domain bounds do not match grid_dims*cell_size");
  }
  c->cs = cell_size; c->have_domain = true; c->grid_ready = false;
  return XNB_OK;
}

int xnb_init_rcb_grid(xnb_ctx* c, int rank, int nranks)
{
  if (!c) return XNB_ERR_INVALID;
  if (!c->have_domain) return c->fail(XNB_ERR_INVALID, "xnb_set_domain first");
  if (nranks < 1 || rank < 0 || rank >= nranks) return c->fail(XNB_ERR_INVALID, "bad rank/nranks");
  c->rank = rank; c->nranks = nranks;
  c->blocks.resize((size_t)nranks);
  const Block whole{{0, 0, 0}, {c->ddims[0], c->ddims[1], c->ddims[2]}};
  for (int r = 0; r < nranks; r++) c->blocks[(size_t)r] = simple_block_rcb(whole, (size_t)nranks, (size_t)r);
  c->have_block = true; c->grid_ready = false;
  return XNB_OK;
}

int xnb_set_nbh_dist(xnb_ctx* c, double rcut_max, double rcut_inc)
{
  if (!c) return XNB_ERR_INVALID;
  c->rcut_max = rcut_max; c->rcut_inc = rcut_inc;
  c->nbh_dist = rcut_max + rcut_inc;          // nbh_dist.cpp:47
  c->max_displ = rcut_inc / 2.0;              // :48
  c->ghost_dist = rcut_max + rcut_inc;        // :60-62 with ghost_dist_max = rcut_max, no bonds
  c->have_dist = true; c->grid_ready = false;
  return XNB_OK;
}

int xnb_set_type_mass(xnb_ctx* c, const double* m, int n)
{
  if (!c || !m || n < 1 || n > 256) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpy(c->mass.p, m, (size_t)n * 8, cudaMemcpyHostToDevice));
  c->n_types = n;
  return XNB_OK;
}

int xnb_set_sub_grid_density(xnb_ctx* c, double d) { if (!c || !(d > 0)) return XNB_ERR_INVALID; c->sub_grid_density = d; return XNB_OK; }

int xnb_set_nccl_comm(xnb_ctx* c, void* comm)
{
  if (!c) return XNB_ERR_INVALID;
  if (comm && !nccl_load()) return c->fail(XNB_ERR_NCCL, "NCCL library not found");
  c->comm = comm; c->own_comm = false;
  return XNB_OK;
}

int xnb_nccl_unique_id(uint8_t id[128])
{
  if (!nccl_load()) return XNB_ERR_NCCL;
  return g_nccl.GetUniqueId(id) == 0 ? XNB_OK : XNB_ERR_NCCL;
}

int xnb_nccl_init_rank(xnb_ctx* c, const uint8_t id[128], int rank, int nranks)
{
  if (!c) return XNB_ERR_INVALID;
  if (!nccl_load()) return c->fail(XNB_ERR_NCCL, "NCCL library not found");
  CK(cudaSetDevice(c->device));
  Id128 u; memcpy(u.b, id, 128);
  NK(g_nccl.CommInitRank(&c->comm, nranks, u, rank));
  c->own_comm = true;
  return XNB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
int xnb_set_particles(xnb_ctx* c, int64_t n, const double* rx, const double* ry, const double* rz,
                      const double* vx, const double* vy, const double* vz, const uint64_t* id, const uint8_t* type)
{
  if (!c || n < 0 || (n > 0 && (!rx || !ry || !rz))) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  int rc = ensure_grid(c); if (rc) return rc;
  const Block& b = c->blocks[(size_t)c->rank];
  // keep the particles whose (periodically wrapped) cell lies in my block, as each rank's `lattice` does
  std::vector<int64_t> keep; keep.reserve((size_t)n);
  for (int64_t i = 0; i < n; i++)
  {
    const double r[3] = {rx[i], ry[i], rz[i]};
    bool mine = true;
    for (int d = 0; d < 3 && mine; d++)
    {
      int64_t loc = (int64_t)std::floor((r[d] - c->dmin[d]) / c->cs);
      if (c->periodic[d]) loc = ((loc % c->ddims[d]) + c->ddims[d]) % c->ddims[d];
      if (loc < b.s[d] || loc >= b.e[d]) mine = false;
    }
    if (mine || c->nranks == 1) keep.push_back(i);
  }
  const size_t m = keep.size();
  rc = ensure_particle_capacity(c, std::max<size_t>((size_t)((double)m * 1.6), 1024), 0); if (rc) return rc;
  std::vector<double> tmp(m);
  const double* src[9] = {rx, ry, rz, vx, vy, vz, nullptr, nullptr, nullptr};
  for (int f = 0; f < 9; f++)
  {
    for (size_t k = 0; k < m; k++) tmp[k] = src[f] ? src[f][keep[k]] : 0.0;
    if (m) CK(cudaMemcpy(c->f64[c->cur][f].p, tmp.data(), m * 8, cudaMemcpyHostToDevice));
  }
  std::vector<unsigned long long> ids(m); std::vector<uint8_t> ty(m);
  for (size_t k = 0; k < m; k++) { ids[k] = id ? id[keep[k]] : (unsigned long long)keep[k]; ty[k] = type ? type[keep[k]] : 0; }
  if (m) { CK(cudaMemcpy(c->idb[c->cur].p, ids.data(), m * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(c->typeb[c->cur].p, ty.data(), m, cudaMemcpyHostToDevice)); }
  c->n_inner = (int64_t)m; c->n_total = (int64_t)m; c->have_nbh = false; c->n_ghost = 0; c->amr_current = false;
  return XNB_OK;
}

int64_t xnb_num_inner(const xnb_ctx* c) { return c ? c->n_inner : 0; }
int64_t xnb_num_total(const xnb_ctx* c) { return c ? c->n_total : 0; }

int xnb_get_particles(xnb_ctx* c, int64_t first, int64_t n, double* rx, double* ry, double* rz, double* vx, double* vy, double* vz,
                      double* fx, double* fy, double* fz, uint64_t* id, uint8_t* type, uint32_t* cell)
{
  if (!c || first < 0 || n < 0 || first + n > c->n_total) return c ? c->fail(XNB_ERR_INVALID, "range out of bounds") : XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  double* dst[9] = {rx, ry, rz, vx, vy, vz, fx, fy, fz};
  for (int f = 0; f < 9; f++) if (dst[f] && n) CK(cudaMemcpy(dst[f], c->f64[c->cur][f].p + first, (size_t)n * 8, cudaMemcpyDeviceToHost));
  if (id && n) CK(cudaMemcpy(id, c->idb[c->cur].p + first, (size_t)n * 8, cudaMemcpyDeviceToHost));
  if (type && n) CK(cudaMemcpy(type, c->typeb[c->cur].p + first, (size_t)n, cudaMemcpyDeviceToHost));
  if (cell && n) CK(cudaMemcpy(cell, c->atom_cell[c->cur_ac].p + first, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return XNB_OK;
}

int xnb_upload_rv(xnb_ctx* c, const double* rx, const double* ry, const double* rz, const double* vx, const double* vy, const double* vz, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  c->amr_current = false;      // positions change: the sub-cell tables no longer bound them
  cudaStream_t st = (cudaStream_t)stream;
  const double* src[6] = {rx, ry, rz, vx, vy, vz};
  for (int f = 0; f < 6; f++) if (src[f] && c->n_inner) CK(cudaMemcpyAsync(c->f64[c->cur][f].p, src[f], (size_t)c->n_inner * 8, cudaMemcpyHostToDevice, st));
  return XNB_OK;
}

int xnb_download_rvf(xnb_ctx* c, double* rx, double* ry, double* rz, double* vx, double* vy, double* vz, double* fx, double* fy, double* fz, uint64_t* id, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  double* dst[9] = {rx, ry, rz, vx, vy, vz, fx, fy, fz};
  for (int f = 0; f < 9; f++) if (dst[f] && c->n_inner) CK(cudaMemcpyAsync(dst[f], c->f64[c->cur][f].p, (size_t)c->n_inner * 8, cudaMemcpyDeviceToHost, st));
  if (id && c->n_inner) CK(cudaMemcpyAsync(id, c->idb[c->cur].p, (size_t)c->n_inner * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return XNB_OK;
}

int xnb_get_grid_info(const xnb_ctx* cc, xnb_grid_info* out)
{
  xnb_ctx* c = const_cast<xnb_ctx*>(cc);
  if (!c || !out) return XNB_ERR_INVALID;
  int rc = ensure_grid(c); if (rc) return rc;
  for (int d = 0; d < 3; d++) { out->dims[d] = c->g.dims[d]; out->offset[d] = c->g.off[d]; out->block_start[d] = c->g.bstart[d]; out->block_end[d] = c->g.bend[d]; }
  out->ghost_layers = c->g.gl; out->n_cells = c->g.n_cells;
  return XNB_OK;
}

int xnb_get_sweep_info(const xnb_ctx* c, xnb_sweep_info* out)
{
  if (!c || !out) return XNB_ERR_INVALID;
  memset(out, 0, sizeof *out);
  const bool compiled = c->cl.valid && !env_flag("XNB_SWEEP_STREAMS");
  out->compiled = compiled ? 1 : 0;
  if (compiled)
  {
    out->tile[0] = c->cl.tp.ti; out->tile[1] = c->cl.tp.tj; out->tile[2] = c->cl.tp.tk;
    out->threads = c->cl.threads; out->blocks = c->cl.blocks; out->smem_bytes = (int64_t)c->cl.smem;
    out->rows = c->cl.rows; out->candidates = c->cl.candidates; out->ghost = c->cl.ghost ? 1 : 0;
    out->interior_tiles = c->cl.n_interior; out->boundary_tiles = c->cl.n_boundary;
  }
  return XNB_OK;
}

int xnb_get_cells(xnb_ctx* c, uint32_t* cell_start, uint32_t* cell_count)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  int rc = ensure_grid(c); if (rc) return rc;
  CK(cudaDeviceSynchronize());
  if (cell_start) CK(cudaMemcpy(cell_start, c->cell_start.p, (size_t)c->g.n_cells * 4, cudaMemcpyDeviceToHost));
  if (cell_count) CK(cudaMemcpy(cell_count, c->cell_count.p, (size_t)c->g.n_cells * 4, cudaMemcpyDeviceToHost));
  return XNB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// move_particles (+ migrate for nranks > 1)
// ---------------------------------------------------------------------------------------------------------------------
int xnb_move_particles(xnb_ctx* c, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  int rc = ensure_grid(c); if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const GridP& g = c->g;
  int64_t n = c->n_inner;
  uint32_t* s32 = c->d_scalars32.p;
  CK(c->key.ensure((size_t)n + 16, 0, 1.2)); CK(c->rnk.ensure((size_t)n + 16, 0, 1.2)); CK(c->leave_list.ensure((size_t)n + 16, 0, 1.2));
  CK(cudaMemsetAsync(c->cell_count.p, 0, (size_t)g.n_cells * 4, st));
  CK(cudaMemsetAsync(s32 + 1, 0, 4, st));
  ParticlesP A = c->P(c->cur);
  if (n) LAUNCH(k_bin_locate, nblk(n, 256), 256, st, g, (int)n, A.rx, A.ry, A.rz, c->key.p, c->rnk.p, c->cell_count.p, c->leave_list.p, s32 + 1, s32);
  int64_t n_src = n;        // entries of the source arrays (stayers + holes + arrivals)
  int64_t n_leave = 0, n_arrive = 0;
  if (c->nranks > 1)
  {
    if (!c->comm) return c->fail(XNB_ERR_NCCL, "nranks > 1 needs an NCCL communicator (xnb_nccl_init_rank)");
    uint32_t hl = 0;
    rc = read_back(c, s32 + 1, 1, &hl, st); if (rc) return rc;
    n_leave = hl;
    // destination ranks and per-destination counts
    uint32_t* d_cnt = s32 + 8;
    CK(cudaMemsetAsync(d_cnt, 0, 64 * 4, st));
    CK(c->mig_rank.ensure((size_t)n_leave + 16)); CK(c->mig_pos.ensure((size_t)n_leave + 16));
    if (c->nranks > 64) return c->fail(XNB_ERR_INVALID, "more than 64 ranks not supported");
    if (n_leave) LAUNCH(k_migrate_dest, nblk(n_leave, 128), 128, st, g, (int)n_leave, c->leave_list.p, A.rx, A.ry, A.rz, c->d_blocks.p, c->nranks,
                        c->mig_rank.p, c->mig_pos.p, d_cnt, s32);
    // all ranks learn the full count matrix
    DBuf<uint32_t>& mat = c->scan_tmp32;
    CK(mat.ensure((size_t)c->nranks * 64 + 64));
    // the diagonal of the matrix is free (nobody migrates to itself): it carries every rank's device error word, so that an error one
    // rank ran into (a particle left a non periodic domain, jumped further than a neighbour block ...) ends the call on ALL ranks here,
    // before the exchange below would wait for the rank that gave up (the reference's fatal_error aborts every rank)
    CK(cudaMemcpyAsync(d_cnt + c->rank, s32, 4, cudaMemcpyDeviceToDevice, st));
    NK(g_nccl.AllGather(d_cnt, mat.p, 64, nccl_uint32, c->comm, st));
    std::vector<uint32_t> hmat((size_t)c->nranks * 64);
    CK(cudaMemcpyAsync(hmat.data(), mat.p, hmat.size() * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    {
      int bad_rank = -1;
      for (int p = 0; p < c->nranks; p++) { if (hmat[(size_t)p * 64 + p] && bad_rank < 0) bad_rank = p; }
      if (bad_rank >= 0)
      {
        const uint32_t mine = hmat[(size_t)c->rank * 64 + c->rank];
        if (mine) return decode_device_errors(c, mine, st);
        CK(cudaMemsetAsync(s32, 0, 4, st));
        return c->fail(XNB_ERR_INVALID, "rank " + std::to_string(bad_rank) + " reported a device error (word " + std::to_string(hmat[(size_t)bad_rank * 64 + bad_rank]) + ") while binning: the step is abandoned on every rank");
      }
    }
    std::vector<uint32_t> sbase((size_t)c->nranks + 1, 0), rbase((size_t)c->nranks + 1, 0);
    for (int p = 0; p < c->nranks; p++) { sbase[(size_t)p + 1] = sbase[(size_t)p] + hmat[(size_t)c->rank * 64 + p]; rbase[(size_t)p + 1] = rbase[(size_t)p] + hmat[(size_t)p * 64 + c->rank]; }
    n_arrive = rbase[(size_t)c->nranks];
    if ((int64_t)sbase[(size_t)c->nranks] != n_leave) return c->fail(XNB_ERR_LOST_PARTICLE, "migration: particles without owner");
    // one message per partner and direction (like the halo): slabs of GHOST_WORDS_ALL words per particle, scattered by k_ghost_unpack
    c->h_mig_base.assign(sbase.begin(), sbase.end()); c->h_mig_base.insert(c->h_mig_base.end(), rbase.begin(), rbase.end());
    CK(c->mig_base.ensure(2 * ((size_t)c->nranks + 1) + 4));
    CK(cudaMemcpyAsync(c->mig_base.p, c->h_mig_base.data(), c->h_mig_base.size() * 4, cudaMemcpyHostToDevice, st));
    CK(c->stage.ensure((size_t)n_leave * GHOST_WORDS_ALL + 16, 0, 1.5)); CK(c->rstage.ensure((size_t)n_arrive * GHOST_WORDS_ALL + 16, 0, 1.5));
    if (n_leave) LAUNCH(k_migrate_pack, nblk(n_leave, 128), 128, st, (int)n_leave, c->leave_list.p, c->mig_rank.p, c->mig_pos.p, c->mig_base.p, A, c->stage.p);
    rc = ensure_particle_capacity(c, (size_t)(n + n_arrive), (size_t)n); if (rc) return rc;
    A = c->P(c->cur);
    CK(c->key.ensure((size_t)(n + n_arrive) + 16, (size_t)n, 1.2)); CK(c->rnk.ensure((size_t)(n + n_arrive) + 16, (size_t)n, 1.2));
    NK(g_nccl.GroupStart());
    for (int p = 0; p < c->nranks; p++)
    {
      const size_t ns = sbase[(size_t)p + 1] - sbase[(size_t)p], nr = rbase[(size_t)p + 1] - rbase[(size_t)p];
      if (p == c->rank) continue;
      if (ns) NK(g_nccl.Send(c->stage.p + (size_t)GHOST_WORDS_ALL * sbase[(size_t)p], (size_t)GHOST_WORDS_ALL * ns, nccl_float64, p, c->comm, st));
      if (nr) NK(g_nccl.Recv(c->rstage.p + (size_t)GHOST_WORDS_ALL * rbase[(size_t)p], (size_t)GHOST_WORDS_ALL * nr, nccl_float64, p, c->comm, st));
    }
    NK(g_nccl.GroupEnd());
    if (n_arrive) LAUNCH((k_ghost_unpack<true>), nblk(n_arrive, 256), 256, st, (int)n_arrive, (uint32_t)n, A, c->mig_base.p + (size_t)c->nranks + 1, c->nranks, -1, c->rstage.p);
    // locate the arrivals (they are inside my block by construction)
    if (n_arrive) LAUNCH(k_bin_locate, nblk(n_arrive, 256), 256, st, g, (int)n_arrive, A.rx + n, A.ry + n, A.rz + n, c->key.p + n, c->rnk.p + n, c->cell_count.p,
                         (uint32_t*)nullptr, s32 + 1, s32);
    n_src = n + n_arrive;
  }
  const int64_t n_new = n - n_leave + n_arrive;
  CK(c->perm.ensure((size_t)n_src + 16, 0, 1.2)); CK(c->perm2.ensure((size_t)n_src + 16, 0, 1.2));
  rc = scan_exclusive<uint32_t, uint32_t>(c, c->cell_count.p, c->cell_start.p, (size_t)g.n_cells, (uint32_t*)nullptr, c->scan_tmp32, st); if (rc) return rc;
  CK(cudaMemsetAsync(c->perm2.p, 0xFF, ((size_t)n_src + 16) * 4, st));      // sentinel: slots no kernel fills (error paths) are skipped by k_gather
  if (n_src) LAUNCH(k_bin_scatter, nblk(n_src, 256), 256, st, (int)n_src, c->key.p, c->rnk.p, c->cell_start.p, c->perm.p);
  // cells beyond the shared-memory sort capacity rank through global scratch (12 bytes per particle)
  CK(c->sort_keys.ensure((size_t)n_src + 16, 0, 1.2)); CK(c->sort_srcs.ensure((size_t)n_src + 16, 0, 1.2));
  // the fullest cell at the last neighbour build tells which kernel suits (either one is correct for any occupancy)
  if ((c->max_cell_count > 0 && c->max_cell_count <= 56 && !env_flag("XNB_CELLSORT_BLOCK")) || env_flag("XNB_CELLSORT_WARP"))
    LAUNCH(k_cell_sort_warp, std::min(nblk(g.n_cells, 8), 148u * 16u), 256, st, g.n_cells, c->cell_start.p, c->cell_count.p, c->perm.p, c->perm2.p, A.id, c->sort_keys.p, c->sort_srcs.p, s32);
  else
  LAUNCH((k_cell_sort<false>), (unsigned)g.n_cells, CELLSORT_THREADS, st, g, c->cell_start.p, c->cell_count.p, c->perm.p, c->perm2.p,
         A.rx, A.ry, A.rz, A.id, c->side_lut.p, (const unsigned long long*)nullptr, (uint32_t*)nullptr, c->sort_keys.p, c->sort_srcs.p, s32);
  rc = ensure_particle_capacity(c, (size_t)std::max<int64_t>(n_new, 1), (size_t)n_src); if (rc) return rc;
  A = c->P(c->cur);
  ParticlesP B = c->P(1 - c->cur);
  if (n_new) LAUNCH(k_gather, nblk(n_new, 256), 256, st, (int)n_new, (uint32_t)n_src, c->perm2.p, A, B, c->key.p, c->atom_cell[1 - c->cur_ac].p);
  c->cur = 1 - c->cur; c->cur_ac = 1 - c->cur_ac;
  c->n_inner = n_new; c->n_total = n_new; c->n_ghost = 0; c->have_nbh = false; c->amr_current = false;
  LAUNCH(k_ghost_cells_clear, nblk(g.n_cells, 256), 256, st, g, (uint32_t)n_new, c->cell_start.p, c->cell_count.p);
  c->cell_stats_valid = false;
  if (c->in_rebuild_chain) return XNB_OK;          // rebuild_amr, next in the chain, reads the error word with its own results
  return check_device_errors(c, st);
}

// ---------------------------------------------------------------------------------------------------------------------
// op `load_balance_rcb` on a live context (src/mpi/load_balance_rcb.cpp:51-601 without Zoltan, cost model simple_cost_model.h:67-146,
// followed by migrate_cell_particles.cpp:101-143): per-cell costs on the device -> sum over the ranks -> cost-weighted recursive
// bisection (every rank computes the same table) -> the new block replaces the old one and xnb_move_particles' migration hand-off
// carries every particle to its new owner.  Collective.  The caller continues with the rebuild chain (rebuild_amr ... chunk_neighbors).
// ---------------------------------------------------------------------------------------------------------------------
int xnb_load_balance_rcb(xnb_ctx* c, const double coefs[4], double* inbalance_before, double* inbalance_after, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  int rc = ensure_grid(c); if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  static const double default_coefs[4] = {0.0, 0.0, 1.0, 0.0};
  const double* k = coefs ? coefs : default_coefs;
  const size_t nd = (size_t)c->ddims[0] * (size_t)c->ddims[1] * (size_t)c->ddims[2];
  CK(c->lb_costs.ensure(nd + 8));
  CK(cudaMemsetAsync(c->lb_costs.p, 0, nd * 8, st));
  LAUNCH(k_cell_costs, nblk(c->g.n_cells, 256), 256, st, c->g, c->cell_count.p, k[0], k[1], k[2], k[3], c->lb_costs.p);
  if (c->nranks > 1)
  {
    if (!c->comm) return c->fail(XNB_ERR_NCCL, "nranks > 1 needs an NCCL communicator");
    NK(g_nccl.AllReduce(c->lb_costs.p, c->lb_costs.p, nd, nccl_float64, nccl_sum, c->comm, st));       // MPI_Allreduce(SUM), load_balance_rcb.cpp:270
  }
  std::vector<double> costs(nd);
  CK(cudaMemcpyAsync(costs.data(), c->lb_costs.p, nd * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  auto block_cost = [&](const Block& b) {
    double s = 0.0;
    for (int64_t kk = b.s[2]; kk < b.e[2]; kk++) for (int64_t jj = b.s[1]; jj < b.e[1]; jj++) for (int64_t ii = b.s[0]; ii < b.e[0]; ii++)
      s += costs[(size_t)((kk * c->ddims[1] + jj) * c->ddims[0] + ii)];
    return s;
  };
  auto inbalance = [&](const std::vector<Block>& bl) {        // lb_inbalance = (max - avg) / avg (load_balance_rcb.cpp:443-452)
    double mx = 0.0, sum = 0.0;
    for (const Block& b : bl) { const double v = block_cost(b); mx = std::max(mx, v); sum += v; }
    const double avg = sum / (double)bl.size();
    return avg > 0.0 ? (mx - avg) / avg : 0.0;
  };
  std::vector<Block> nb((size_t)c->nranks);
  for (int r = 0; r < c->nranks; r++)
  {
    nb[(size_t)r] = load_balance_rcb(c->ddims, costs.data(), (size_t)c->nranks, (size_t)r);
    for (int d = 0; d < 3; d++) if (nb[(size_t)r].e[d] <= nb[(size_t)r].s[d]) return c->fail(XNB_ERR_INVALID, "load_balance_rcb: Assigned grid block is empty");      // :457
  }
  if (inbalance_before) *inbalance_before = inbalance(c->blocks);
  if (inbalance_after) *inbalance_after = inbalance(nb);
  bool same = true;
  for (int r = 0; r < c->nranks; r++) for (int d = 0; d < 3; d++) same = same && nb[(size_t)r].s[d] == c->blocks[(size_t)r].s[d] && nb[(size_t)r].e[d] == c->blocks[(size_t)r].e[d];
  if (same) return XNB_OK;
  // the new decomposition: grid geometry, ghost items and the migration table are rebuilt; lists and tile shapes of the old grid are void
  c->blocks = nb; c->grid_ready = false;
  c->n_total = c->n_inner; c->n_ghost = 0; c->n_send = 0;
  c->have_nbh = false; c->cl.valid = false; c->nb_ghost.have = false; c->ghost_lists = false; c->amr_current = false; c->nbh_slot_words = 0;
  return xnb_move_particles(c, stream);
}

int xnb_get_block(const xnb_ctx* c, int rank, int64_t start[3], int64_t end[3])
{
  if (!c || !c->have_block || rank < 0 || rank >= c->nranks) return XNB_ERR_INVALID;
  for (int d = 0; d < 3; d++) { start[d] = c->blocks[(size_t)rank].s[d]; end[d] = c->blocks[(size_t)rank].e[d]; }
  return XNB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
int xnb_rebuild_amr(xnb_ctx* c, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  int rc = ensure_grid(c); if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const GridP& g = c->g;
  uint32_t* s32 = c->d_scalars32.p;
  CK(c->sg_size.ensure((size_t)g.n_cells + 1)); CK(c->sub_grid_start.ensure((size_t)g.n_cells + 1));
  CK(cudaMemsetAsync(c->sg_size.p + g.n_cells, 0, 4, st));
  CK(cudaMemsetAsync(s32 + 2, 0, 4, st));
  LAUNCH(k_amr_sizes, nblk(g.n_cells, 256), 256, st, g, c->cell_count.p, c->side_lut.p, c->sg_size.p, s32 + 2);
  rc = scan_exclusive<uint32_t, unsigned long long>(c, c->sg_size.p, c->sub_grid_start.p, (size_t)g.n_cells + 1, c->d_scalars64.p + 1, c->scan_tmp64, st); if (rc) return rc;
  uint32_t ms = 0; unsigned long long tot = 0;
  {
    // one host wait: error word (binning before me may have deferred its check), largest sub-grid side, sub-grid cell total
    char* hp = static_cast<char*>(c->h_pinned);
    CK(cudaMemcpyAsync(hp + 256, s32, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hp + 288, c->d_scalars64.p + 1, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    uint32_t w[4]; memcpy(w, hp + 256, 16); memcpy(&tot, hp + 288, 8);
    if ((rc = decode_device_errors(c, w[0], st))) return rc;
    ms = w[2];
  }
  c->max_side = std::max(ms, 1u); c->n_sub_grid_cells = (int64_t)tot; c->amr_current = true;
  if (ms <= 1) return XNB_OK;          // every cell has a 1x1x1 sub grid: nothing to reorder (C2/C3)
  const int64_t n = c->n_inner;
  CK(c->sub_grid_cells.ensure((size_t)tot + 16));
  CK(c->perm.ensure((size_t)n + 16)); CK(c->perm2.ensure((size_t)n + 16));
  ParticlesP A = c->P(c->cur), B = c->P(1 - c->cur);
  LAUNCH(k_iota, nblk(n, 256), 256, st, (int)n, c->perm.p);
  CK(c->sort_keys.ensure((size_t)n + 16, 0, 1.2)); CK(c->sort_srcs.ensure((size_t)n + 16, 0, 1.2));
  LAUNCH((k_cell_sort<true>), (unsigned)g.n_cells, CELLSORT_THREADS, st, g, c->cell_start.p, c->cell_count.p, c->perm.p, c->perm2.p,
         A.rx, A.ry, A.rz, A.id, c->side_lut.p, c->sub_grid_start.p, c->sub_grid_cells.p, c->sort_keys.p, c->sort_srcs.p, s32);
  LAUNCH(k_gather, nblk(n, 256), 256, st, (int)n, (uint32_t)n, c->perm2.p, A, B, c->atom_cell[c->cur_ac].p, c->atom_cell[1 - c->cur_ac].p);
  c->cur = 1 - c->cur; c->cur_ac = 1 - c->cur_ac;
  return XNB_OK;
}

int xnb_backup_r(xnb_ctx* c, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  int rc = ensure_grid(c); if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  CK(c->backup.ensure((size_t)c->n_inner * 3 + 16, 0, 1.2));
  ParticlesP A = c->P(c->cur);
  if (c->n_inner) LAUNCH(k_backup_r, nblk(c->n_inner, 256), 256, st, c->g, (int)c->n_inner, A.rx, A.ry, A.rz, c->atom_cell[c->cur_ac].p, c->backup.p);
  return XNB_OK;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// halo over peer memory: mailbox management (xnb_peer_halo.cuh).  Called from xnb_ghost_comm_scheme, i.e. collectively and once per
// rebuild: the per-step path (ghost_update_r, the displacement sum) then runs without any host-side communication call.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
struct PeerRecord                      // what every rank tells every other rank at a rebuild (all-gathered, 512 bytes)
{
  cudaIpcMemHandle_t handle; unsigned long long gen, cap_words; uint32_t ok, pad; uint32_t recv_base[PEER_MAX_RANKS + 1];
  unsigned char fill[512 - sizeof(cudaIpcMemHandle_t) - 16 - 8 - 4 * (PEER_MAX_RANKS + 1)];
};
static_assert(sizeof(PeerRecord) == 512, "PeerRecord");

// my mailbox holds two halves of at least need_words words
int peer_ensure_box(xnb_ctx* c, size_t need_words, cudaStream_t st)
{
  xnb_ctx::PeerHalo& P = c->peer;
  if (P.box && need_words <= P.cap_words) return 0;
  const size_t cap = std::max<size_t>((size_t)((double)need_words * 1.5), (size_t)1 << 16);
  void* q = nullptr;
  CK(cudaMalloc(&q, PEER_HDR_BYTES + 2 * cap * 8));
  g_device_allocs++;
  CK(cudaMemsetAsync(q, 0, PEER_HDR_BYTES, st));
  if (P.box) P.retired.push_back(P.box);
  P.box = q; P.cap_words = cap; P.gen++;
  return 0;
}

// all-gather the records, map what changed, refresh the device table.  *all_ok = every rank could do it.
int peer_exchange(xnb_ctx* c, bool my_ok, bool* all_ok, cudaStream_t st)
{
  xnb_ctx::PeerHalo& P = c->peer;
  const size_t n = (size_t)c->nranks;
  PeerRecord mine; memset(&mine, 0, sizeof(mine));
  mine.ok = my_ok ? 1u : 0u;
  if (my_ok)
  {
    if (cudaIpcGetMemHandle(&mine.handle, P.box) != cudaSuccess) { cudaGetLastError(); mine.ok = 0u; }
    mine.gen = P.gen; mine.cap_words = P.cap_words;
    for (size_t p = 0; p <= n; p++) mine.recv_base[p] = c->h_recv_base.size() > p ? c->h_recv_base[p] : 0u;
  }
  CK(P.d_rec.ensure((n + 1) * sizeof(PeerRecord)));
  CK(cudaMemcpyAsync(P.d_rec.p + n * sizeof(PeerRecord), &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
  NK(g_nccl.AllGather(P.d_rec.p + n * sizeof(PeerRecord), P.d_rec.p, sizeof(PeerRecord), nccl_uint8, c->comm, st));
  std::vector<PeerRecord> all(n);
  CK(cudaMemcpyAsync(all.data(), P.d_rec.p, n * sizeof(PeerRecord), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  bool ok = true;
  for (size_t p = 0; p < n; p++) ok = ok && all[p].ok != 0u;
  *all_ok = ok;
  if (!ok) return 0;
  P.peer_ptr.resize(n, nullptr); P.peer_gen.resize(n, 0);
  std::vector<PeerSlot>& slots = P.h_slots; slots.resize(n);
  for (size_t p = 0; p < n; p++)
  {
    if ((int)p == c->rank) P.peer_ptr[p] = P.box;
    else if (P.peer_gen[p] != all[p].gen || !P.peer_ptr[p])
    {
      if (P.peer_ptr[p]) cudaIpcCloseMemHandle(P.peer_ptr[p]);
      P.peer_ptr[p] = nullptr;
      if (cudaIpcOpenMemHandle(&P.peer_ptr[p], all[p].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); P.peer_ptr[p] = nullptr; *all_ok = false; }
    }
    P.peer_gen[p] = all[p].gen;
    slots[p] = PeerSlot{(unsigned long long)(uintptr_t)P.peer_ptr[p], all[p].cap_words, all[p].recv_base[(size_t)c->rank], 0u};
  }
  CK(P.d_slots.ensure(n + 1));
  CK(cudaMemcpyAsync(P.d_slots.p, slots.data(), n * sizeof(PeerSlot), cudaMemcpyHostToDevice, st));
  return 0;
}

// every rebuild (from xnb_ghost_comm_scheme, after the receive layout is known).  The first call decides, for all ranks alike, whether
// the peer path is usable (same node, IPC permitted, XNB_GHOST_NCCL unset); afterwards a failure is an error.
int peer_refresh(xnb_ctx* c, cudaStream_t st)
{
  xnb_ctx::PeerHalo& P = c->peer;
  if (c->nranks < 2 || (P.tried && !P.enabled)) return 0;
  const bool first = !P.tried;
  P.tried = true;
  if (first && env_int("XNB_PEER_TIMEOUT_MS") > 0) P.timeout_ns = (unsigned long long)env_int("XNB_PEER_TIMEOUT_MS") * 1000000ull;
  bool my_ok = c->nranks <= PEER_MAX_RANKS && !(first && env_flag("XNB_GHOST_NCCL"));
  if (my_ok)
  {
    const int rc = peer_ensure_box(c, (size_t)c->n_ghost * GHOST_WORDS_ALL + 16, st);
    if (rc) { if (!first) return rc; my_ok = false; cudaGetLastError(); }
  }
  bool all_ok = false;
  int rc = peer_exchange(c, my_ok, &all_ok, st); if (rc) return rc;
  if (first)
  {
    // opening a handle may have failed on some rank only: agree once more
    unsigned long long* flag = c->d_scalars64.p + 8;
    const unsigned long long mine = all_ok ? 0ull : 1ull;
    CK(cudaMemcpyAsync(flag, &mine, 8, cudaMemcpyHostToDevice, st));
    NK(g_nccl.AllReduce(flag, flag, 1, nccl_uint64, nccl_sum, c->comm, st));
    unsigned long long bad = 0;
    rc = read_back(c, flag, 1, &bad, st); if (rc) return rc;
    P.enabled = bad == 0;
    return 0;
  }
  if (!all_ok) return c->fail(XNB_ERR_NCCL, "peer-memory halo: a partner's mailbox could not be mapped");
  return 0;
}
} // namespace

extern "C" {
// ---------------------------------------------------------------------------------------------------------------------
// ghosts
// ---------------------------------------------------------------------------------------------------------------------
int xnb_ghost_comm_scheme(xnb_ctx* c, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  int rc = ensure_grid(c); if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const GridP& g = c->g;
  ParticlesP A = c->P(c->cur);
  GhostItemsP it{c->it_src_cell.p, c->it_flags.p, c->it_partner.p, c->it_outer.p, c->n_send_items};
  const int ns = c->n_send_items, nr = c->n_recv_items;
  // the previous ghosts are discarded (migrate_cell_particles leaves ghost cells empty)
  LAUNCH(k_ghost_cells_clear, nblk(g.n_cells, 256), 256, st, g, (uint32_t)c->n_inner, c->cell_start.p, c->cell_count.p);
  c->n_total = c->n_inner; c->n_ghost = 0; c->n_send = 0;
  if (ns) LAUNCH(k_ghost_count, nblk((int64_t)ns * 32, 128), 128, st, g, it, c->cell_start.p, c->cell_count.p, A.rx, A.ry, A.rz, c->it_count.p);
  CK(cudaMemsetAsync(c->it_count.p + ns, 0, 4, st));
  rc = scan_exclusive<uint32_t, uint32_t>(c, c->it_count.p, c->it_offset.p, (size_t)ns + 1, (uint32_t*)nullptr, c->scan_tmp32, st); if (rc) return rc;
  // exchange per-item counts with the partners (update_ghosts_comm_scheme.cpp:257-303)
  if (c->nranks > 1)
  {
    if (!c->comm) return c->fail(XNB_ERR_NCCL, "nranks > 1 needs an NCCL communicator");
    NK(g_nccl.GroupStart());
    for (int p = 0; p < c->nranks; p++)
    {
      if (p == c->rank) continue;
      const int s0 = c->send_first[(size_t)p], s1 = c->send_first[(size_t)p + 1], r0 = c->recv_first[(size_t)p], r1 = c->recv_first[(size_t)p + 1];
      if (s1 > s0) NK(g_nccl.Send(c->it_count.p + s0, (size_t)(s1 - s0), nccl_uint32, p, c->comm, st));
      if (r1 > r0) NK(g_nccl.Recv(c->rc_count.p + r0, (size_t)(r1 - r0), nccl_uint32, p, c->comm, st));
    }
    NK(g_nccl.GroupEnd());
  }
  {
    const int s0 = c->send_first[(size_t)c->rank], s1 = c->send_first[(size_t)c->rank + 1], r0 = c->recv_first[(size_t)c->rank];
    if (s1 > s0) CK(cudaMemcpyAsync(c->rc_count.p + r0, c->it_count.p + s0, (size_t)(s1 - s0) * 4, cudaMemcpyDeviceToDevice, st));
  }
  CK(cudaMemsetAsync(c->rc_count.p + nr, 0, 4, st));
  rc = scan_exclusive<uint32_t, uint32_t>(c, c->rc_count.p, c->rc_offset.p, (size_t)nr + 1, (uint32_t*)nullptr, c->scan_tmp32, st); if (rc) return rc;
  // per partner particle bases (host needs them for the NCCL calls and the totals for allocation)
  std::vector<uint32_t> hso((size_t)ns + 1), hro((size_t)nr + 1);
  CK(cudaMemcpyAsync(hso.data(), c->it_offset.p, ((size_t)ns + 1) * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(hro.data(), c->rc_offset.p, ((size_t)nr + 1) * 4, cudaMemcpyDeviceToHost, st));
  // occupancy of the grid for the neighbour build that follows (its tile shape and capacities), read in the same host wait: the inner
  // cells as binned, the ghost cells from the receive counts (one receive item per ghost cell)
  uint32_t cstats[2] = {0, 0};
  CK(cudaMemsetAsync(c->d_scalars32.p + 5, 0, 8, st));
  LAUNCH(k_cell_stats, nblk(g.n_cells, 256), 256, st, g, c->cell_count.p, c->d_scalars32.p + 5);
  if (nr) LAUNCH(k_max_u32, nblk(nr, 256), 256, st, nr, c->rc_count.p, c->d_scalars32.p + 5);
  CK(cudaMemcpyAsync(cstats, c->d_scalars32.p + 5, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  c->max_cell_count = std::max<uint32_t>(cstats[0], 1); c->n_nonempty_inner = cstats[1]; c->cell_stats_valid = true;
  for (int p = 0; p <= c->nranks; p++) { c->h_send_base[(size_t)p] = hso[(size_t)c->send_first[(size_t)p]]; c->h_recv_base[(size_t)p] = hro[(size_t)c->recv_first[(size_t)p]]; }
  c->n_send = hso[(size_t)ns]; c->n_ghost = hro[(size_t)nr];
  {
    // the pack / unpack kernels find an entry's partner in these (one slab per partner on the wire)
    CK(c->d_ghost_base.ensure(2 * ((size_t)c->nranks + 1) + 4));
    c->h_ghost_base.resize(2 * ((size_t)c->nranks + 1));
    for (int p = 0; p <= c->nranks; p++) { c->h_ghost_base[(size_t)p] = c->h_send_base[(size_t)p]; c->h_ghost_base[(size_t)c->nranks + 1 + p] = c->h_recv_base[(size_t)p]; }
    CK(cudaMemcpyAsync(c->d_ghost_base.p, c->h_ghost_base.data(), c->h_ghost_base.size() * 4, cudaMemcpyHostToDevice, st));
  }
  if (c->nranks > 1) { rc = peer_refresh(c, st); if (rc) return rc; }
  CK(c->send_src.ensure((size_t)c->n_send + 16, 0, 1.2)); CK(c->send_flags.ensure((size_t)c->n_send + 16, 0, 1.2));
  if (ns) LAUNCH(k_ghost_fill, nblk((int64_t)ns * 32, 128), 128, st, g, it, c->cell_start.p, c->cell_count.p, A.rx, A.ry, A.rz, c->it_offset.p, c->send_src.p, c->send_flags.p);
  rc = ensure_particle_capacity(c, (size_t)(c->n_inner + c->n_ghost), (size_t)c->n_inner); if (rc) return rc;
  if (nr) LAUNCH(k_ghost_cells, nblk((int64_t)nr * 32, 128), 128, st, nr, c->rc_dst_cell.p, c->rc_count.p, c->rc_offset.p, (uint32_t)c->n_inner,
                 c->cell_start.p, c->cell_count.p, c->atom_cell[c->cur_ac].p);
  c->n_total = c->n_inner + c->n_ghost;
  c->have_nbh = false;
  return XNB_OK;
}

} // extern "C"
static int ghost_update(xnb_ctx* c, bool all, cudaStream_t st)
{
  const GridP& g = c->g;
  // (the exchange counter of the peer transport advances on every rank alike, also on one that has nothing to send or receive)
  if (c->n_send == 0 && c->n_ghost == 0 && !(c->nranks > 1 && c->peer.enabled)) return XNB_OK;
  ParticlesP A = c->P(c->cur);
  const int self_first = (int)c->h_send_base[(size_t)c->rank], self_end = (int)c->h_send_base[(size_t)c->rank + 1];
  const uint32_t self_dst = (uint32_t)(c->n_inner + c->h_recv_base[(size_t)c->rank]);
  const size_t ns = (size_t)c->n_send, ng = (size_t)c->n_ghost;
  const size_t nw = all ? GHOST_WORDS_ALL : GHOST_WORDS_R;          // 8-byte words per ghost on the wire
  if (c->nranks > 1 && c->peer.enabled)
  {
    // pack + NVLink stores + flag in one kernel, unpack waits on the flags: no host-side communication call on this path
    xnb_ctx::PeerHalo& P = c->peer;
    const unsigned long long epoch = ++P.epoch;
    PeerHdr* hdr = static_cast<PeerHdr*>(P.box);
    const unsigned grid_cap = 2u * 148u;
    if (ns)
    {
      const unsigned nb = std::min(nblk((int64_t)ns, 256), grid_cap);
      if (all) LAUNCH((k_ghost_push<true>), nb, 256, st, g, (int)ns, c->send_src.p, c->send_flags.p, A, self_first, self_end, self_dst, c->d_ghost_base.p, c->nranks, c->rank, P.d_slots.p, epoch, hdr);
      else     LAUNCH((k_ghost_push<false>), nb, 256, st, g, (int)ns, c->send_src.p, c->send_flags.p, A, self_first, self_end, self_dst, c->d_ghost_base.p, c->nranks, c->rank, P.d_slots.p, epoch, hdr);
    }
    if (ng)
    {
      const unsigned nb = std::min(nblk((int64_t)ng, 256), grid_cap);
      const uint32_t* rb = c->d_ghost_base.p + (size_t)c->nranks + 1;
      const double* half = reinterpret_cast<const double*>(static_cast<const char*>(P.box) + PEER_HDR_BYTES) + (epoch & 1ull) * P.cap_words;
      if (all) LAUNCH((k_ghost_pull<true>), nb, 256, st, (int)ng, (uint32_t)c->n_inner, A, rb, c->nranks, c->rank, hdr, half, epoch, c->d_scalars32.p, P.timeout_ns);
      else     LAUNCH((k_ghost_pull<false>), nb, 256, st, (int)ng, (uint32_t)c->n_inner, A, rb, c->nranks, c->rank, hdr, half, epoch, c->d_scalars32.p, P.timeout_ns);
    }
    return XNB_OK;
  }
  if (c->nranks > 1) { CK(c->stage.ensure(ns * nw + 16, 0, 1.2)); CK(c->rstage.ensure(ng * nw + 16, 0, 1.2)); }
  if (ns)
  {
    if (all) LAUNCH((k_ghost_pack<true>), nblk((int64_t)ns, 256), 256, st, g, (int)ns, c->send_src.p, c->send_flags.p, A, self_first, self_end, self_dst, c->d_ghost_base.p, c->nranks, c->stage.p);
    else     LAUNCH((k_ghost_pack<false>), nblk((int64_t)ns, 256), 256, st, g, (int)ns, c->send_src.p, c->send_flags.p, A, self_first, self_end, self_dst, c->d_ghost_base.p, c->nranks, c->stage.p);
  }
  if (c->nranks > 1)
  {
    // one message per partner and direction (update_ghosts_comm_manager.h:267-281,390,436)
    NK(g_nccl.GroupStart());
    for (int p = 0; p < c->nranks; p++)
    {
      if (p == c->rank) continue;
      const size_t s0 = c->h_send_base[(size_t)p], sn = c->h_send_base[(size_t)p + 1] - s0;
      const size_t r0 = c->h_recv_base[(size_t)p], rn = c->h_recv_base[(size_t)p + 1] - r0;
      if (sn) NK(g_nccl.Send(c->stage.p + nw * s0, nw * sn, nccl_float64, p, c->comm, st));
      if (rn) NK(g_nccl.Recv(c->rstage.p + nw * r0, nw * rn, nccl_float64, p, c->comm, st));
    }
    NK(g_nccl.GroupEnd());
    if (ng)
    {
      const uint32_t* rb = c->d_ghost_base.p + (size_t)c->nranks + 1;
      if (all) LAUNCH((k_ghost_unpack<true>), nblk((int64_t)ng, 256), 256, st, (int)ng, (uint32_t)c->n_inner, A, rb, c->nranks, c->rank, c->rstage.p);
      else     LAUNCH((k_ghost_unpack<false>), nblk((int64_t)ng, 256), 256, st, (int)ng, (uint32_t)c->n_inner, A, rb, c->nranks, c->rank, c->rstage.p);
    }
  }
  return XNB_OK;
}

extern "C" {
// 0: NCCL send / recv per partner; 1: NVLink peer-memory mailboxes (xnb_peer_halo.cuh); decided at the first xnb_ghost_comm_scheme
int xnb_ghost_transport(const xnb_ctx* c) { return c && c->peer.enabled ? 1 : 0; }
int xnb_ghost_update_all(xnb_ctx* c, void* stream) { if (!c) return XNB_ERR_INVALID; CK(cudaSetDevice(c->device)); return ghost_update(c, true, (cudaStream_t)stream); }
int xnb_ghost_update_r(xnb_ctx* c, void* stream) { if (!c) return XNB_ERR_INVALID; CK(cudaSetDevice(c->device)); return ghost_update(c, false, (cudaStream_t)stream); }

} // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// tiles of the pair sweep (and of the neighbour build, which works on the same tiles): candidate shapes ranked by the
// occupancy they give the sweep kernel
// ---------------------------------------------------------------------------------------------------------------------
struct ClCand { int t[3]; int threads, var; size_t smem; int cap, nh, tc; double score; };
static const size_t XNB_SM_BYTES = 227 * 1024;

static std::vector<ClCand> cl_tile_candidates(const xnb_ctx* c, const int lo[3], const int hi[3], int gap)
{
  const GridP& g = c->g;
  const int nc[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
  std::vector<ClCand> cands;
  if (nc[0] <= 0 || nc[1] <= 0 || nc[2] <= 0) return cands;
  const double ne = (double)std::max<int64_t>(c->n_nonempty_inner, 1);
  const double avg = std::max((double)c->n_inner / ne, 1.0);
  const double mx = (double)std::max<uint32_t>(c->max_cell_count, 1);
  static const int shapes[][3] = {{4, 2, 2}, {4, 4, 1}, {2, 2, 2}, {4, 2, 1}, {3, 3, 2}, {4, 3, 1}, {2, 2, 1}, {2, 1, 1}, {1, 1, 1}, {4, 4, 2}, {3, 2, 2}, {8, 2, 1}, {8, 2, 2}};
  int et[3] = {0, 0, 0};
  if (const char* e = getenv("XNB_CL_TILE")) sscanf(e, "%d,%d,%d", &et[0], &et[1], &et[2]);
  for (const auto& sh : shapes)
  {
    ClCand k{};
    for (int d = 0; d < 3; d++) k.t[d] = std::min(sh[d], nc[d]);
    if (et[0] > 0 && (k.t[0] != std::min(et[0], nc[0]) || k.t[1] != std::min(et[1], nc[1]) || k.t[2] != std::min(et[2], nc[2]))) continue;
    k.tc = k.t[0] * k.t[1] * k.t[2];
    if (k.tc > 32) continue;
    bool dup = false; for (const ClCand& o : cands) dup |= (o.t[0] == k.t[0] && o.t[1] == k.t[1] && o.t[2] == k.t[2]);
    if (dup) continue;
    k.nh = std::min(k.t[0] + 2 * gap, g.dims[0]) * std::min(k.t[1] + 2 * gap, g.dims[1]) * std::min(k.t[2] + 2 * gap, g.dims[2]);
    const double tile_avg = avg * k.tc;
    k.threads = 32 * (int)std::ceil(std::min(mx * k.tc, tile_avg * 1.04 + 8.0) / 32.0);
    if (k.threads > 1024) continue;
    k.threads = std::max(k.threads, 64);
    k.var = k.threads <= 576 ? 0 : 1;
    const double capd = std::min(mx * k.nh, avg * k.nh * 1.08 + 64.0);
    if (capd > 8191.0) continue;
    k.cap = ((int)capd + 1) & ~1;
    k.smem = cl_sweep_smem_bytes(k.nh, k.tc, k.cap, k.threads / 32);
    if (k.smem + 1024 + 2048 > XNB_SM_BYTES) continue;
    const int by_smem = (int)(XNB_SM_BYTES / (k.smem + 1024 + 256));
    const int by_regs = 65536 / (k.threads * (k.var == 0 ? 56 : 64));
    const int by_warps = 64 / (k.threads / 32);
    const int resident = std::min(std::min(by_smem, by_regs), std::min(by_warps, 32));
    if (resident < 1) continue;
    const double useful = resident * (tile_avg / 32.0);               // useful resident warps per SM
    k.score = std::min(useful, 36.0) - 0.25 * (double)k.nh / (double)k.tc;
    cands.push_back(k);
  }
  std::stable_sort(cands.begin(), cands.end(), [](const ClCand& a, const ClCand& b) { return a.score > b.score; });
  return cands;
}

static void cl_tile_grid(ClTileP& tp, const ClCand& k, const int lo[3], const int hi[3], int gap)
{
  tp.gap = gap;
  for (int d = 0; d < 3; d++) { tp.lo[d] = lo[d]; tp.hi[d] = hi[d]; }
  tp.ti = k.t[0]; tp.tj = k.t[1]; tp.tk = k.t[2];
  tp.tiles_i = (hi[0] - lo[0] + tp.ti - 1) / tp.ti; tp.tiles_j = (hi[1] - lo[1] + tp.tj - 1) / tp.tj; tp.tiles_k = (hi[2] - lo[2] + tp.tk - 1) / tp.tk;
  tp.nh_max = (tp.ti + 2 * gap) * (tp.tj + 2 * gap) * (tp.tk + 2 * gap); tp.tc_max = k.tc;
  tp.cap = k.cap; tp.gmax = k.threads / 32;
}

// the sweep configuration once the compiled rows of tile grid `tp` exist (built by k_nbh_bits or by k_cl_compile)
static int cl_finish(xnb_ctx* c, const ClTileP& tp, bool ghost, unsigned blocks, uint32_t rows, int64_t candidates, uint32_t max_groups, cudaStream_t st, int planes = 0, int cap_pl = 0)
{
  const GridP& g = c->g;
  const bool same_grid = c->cl.valid && c->cl.ghost == ghost && c->cl.blocks == blocks && c->cl.tp.ti == tp.ti && c->cl.tp.tj == tp.tj && c->cl.tp.tk == tp.tk &&
                         c->cl.tp.tiles_i == tp.tiles_i && c->cl.tp.tiles_j == tp.tiles_j && c->cl.tp.tiles_k == tp.tiles_k;
  c->cl.tp = tp; c->cl.ghost = ghost; c->cl.blocks = blocks; c->cl.rows = rows; c->cl.candidates = candidates;
  // every tile is swept in one pass: one warp per group of the fullest tile
  c->cl.threads = std::max(32 * (int)std::max<uint32_t>(max_groups, 1u), 64);
  if (env_int("XNB_CL_THREADS") > 0) c->cl.threads = std::min(env_int("XNB_CL_THREADS") & ~31, 1024);
  c->cl.var = c->cl.threads <= 576 ? 0 : 1;
  if (getenv("XNB_CL_VAR")) { const int v = env_int("XNB_CL_VAR"); if ((v == 2 && c->cl.threads <= 288) || (v == 3 && c->cl.threads <= 576)) c->cl.var = v; }   // experiments: 2 = three blocks per SM (<= 72 registers), 3 = one block per SM
  c->cl.smem = cl_sweep_smem_bytes(tp.nh_max, tp.tc_max, tp.cap, std::max(tp.gmax, c->cl.threads / 32));
  c->cl.planes = planes; c->cl.cap_pl = cap_pl;
  if (planes) { c->cl.threads = std::max(32 * (int)std::max<uint32_t>(max_groups, 1u), 64); c->cl.smem = pl_sweep_smem_bytes(tp.nh_max, tp.tc_max, cap_pl); }      // k_lj_sweep_pl: one warp per group, one plane of the halo staged at a time
  if (!same_grid)
  {
    // interior tiles: the halo box touches no ghost cell, so they can be swept while the halo exchange is in flight
    std::vector<uint32_t> inner, outer;
    for (unsigned b = 0; b < blocks; b++)
    {
      const int t_i = (int)(b % tp.tiles_i), t_j = (int)((b / tp.tiles_i) % tp.tiles_j), t_k = (int)(b / ((unsigned)tp.tiles_i * tp.tiles_j));
      const int c0[3] = {tp.lo[0] + t_i * tp.ti, tp.lo[1] + t_j * tp.tj, tp.lo[2] + t_k * tp.tk};
      const int tc[3] = {std::min(tp.ti, tp.hi[0] - c0[0]), std::min(tp.tj, tp.hi[1] - c0[1]), std::min(tp.tk, tp.hi[2] - c0[2])};
      bool in = !ghost;
      for (int d = 0; d < 3; d++) in = in && c0[d] - tp.gap >= g.gl && c0[d] + tc[d] - 1 + tp.gap < g.dims[d] - g.gl;
      (in ? inner : outer).push_back(b);
    }
    c->cl.n_interior = (unsigned)inner.size(); c->cl.n_boundary = (unsigned)outer.size();
    inner.insert(inner.end(), outer.begin(), outer.end());
    CK(c->cl_tile_list.ensure(inner.size() + 16));
    // (pageable source: the call returns once the data sits in the driver's staging buffer, `inner` may go out of scope)
    CK(cudaMemcpyAsync(c->cl_tile_list.p, inner.data(), inner.size() * 4, cudaMemcpyHostToDevice, st));
  }
  c->cl.valid = true;
  if (getenv("XNB_TILE_DEBUG"))
    fprintf(stderr, "[xnb] compiled lists: tiles %dx%dx%d threads %d var %d smem %zu cap %d gmax %d blocks %u rows %u\n", tp.ti, tp.tj, tp.tk,
            c->cl.threads, c->cl.var, c->cl.smem, tp.cap, tp.gmax, blocks, rows);
  return XNB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// neighbour build on the tiles of the sweep: k_nbh_bits.  mode 0: lists of the inner cells + compiled rows of the sweep (what
// the step loop needs); mode 1: lists of the ghost cells only, streams only -- built lazily, when somebody asks for them
// (xnb_view_chunk_neighbors / xnb_get_streams / a sweep with ghost = true).  *done = false: no tile shape fits shared memory
// (large cells) and the caller falls back to the per-particle two-pass kernels.
// ---------------------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------------------
// the same for LARGE cells (more than 32 * NBH_CELL_BLOCKS particles): k_nbh_big, one block per cell, sub-cell pruning through the
// bounding boxes of 32-particle blocks, two passes (count / fill) instead of lists in shared memory
// ---------------------------------------------------------------------------------------------------------------------
static int nbh_big_run(xnb_ctx* c, int mode, cudaStream_t st, bool* done)
{
  *done = false;
  const GridP& g = c->g;
  const int gap = (int)std::ceil(c->nbh_dist / c->cs);
  const int n1 = 2 * gap + 1;
  if (n1 * n1 * n1 > 125 || env_flag("XNB_NBH_NO_BIG")) return XNB_OK;
  int lo[3], hi[3];
  for (int d = 0; d < 3; d++) { lo[d] = mode == 0 ? g.gl : 0; hi[d] = mode == 0 ? g.dims[d] - g.gl : g.dims[d]; }
  if (mode == 1 && g.gl == 0) { *done = true; return XNB_OK; }
  ParticlesP A = c->P(c->cur);
  const uint32_t mcc = std::max<uint32_t>(c->max_cell_count, 1);
  if (c->nbh_slot_words == 0)
  {
    const double frac = std::min(1.0, 4.19 * c->nbh_dist * c->nbh_dist * c->nbh_dist / std::pow((2.0 * gap + 1.0) * c->cs, 3.0));
    const double per = 1.0 + 2.0 * std::pow(2.0 * gap + 1.0, 3.0) * 0.6 + frac * std::pow(2.0 * gap + 1.0, 3.0) * (double)mcc * 1.3;
    c->nbh_slot_words = (uint32_t)((size_t)(2 * (mcc + 1) + (double)mcc * per + 64 + 7) & ~(size_t)7);
  }
  // the sweep's tile for such cells: one cell
  ClCand k{};
  bool have = false;
  for (const ClCand& q : cl_tile_candidates(c, lo, hi, gap)) if (q.t[0] == 1 && q.t[1] == 1 && q.t[2] == 1) { k = q; have = true; break; }
  if (!have) return XNB_OK;
  static bool attr_done_dev[XNB_MAX_DEVICES] = {};
  if (!attr_done_dev[c->device % XNB_MAX_DEVICES])
  {
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k_nbh_big));
    CK(cudaFuncSetAttribute(k_nbh_big, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024 - (int)fa.sharedSizeBytes));
    attr_done_dev[c->device % XNB_MAX_DEVICES] = true;
  }
  uint32_t* counters = c->d_scalars32.p + 64;
  unsigned long long* totals = c->d_scalars64.p + 4;
  ClTileP tp{};
  cl_tile_grid(tp, k, lo, hi, gap);
  const ClTileP& prev = mode == 0 ? c->cl.tp : c->nb_ghost.tp;
  const bool have_prev = mode == 0 ? (c->cl.valid && !c->cl.ghost) : c->nb_ghost.have;
  if (have_prev && prev.ti == 1 && prev.tj == 1 && prev.tk == 1) { tp.gmax = std::max(tp.gmax, prev.gmax); tp.cap = std::max(tp.cap, prev.cap); }
  tp.gmax = std::max(tp.gmax, (int)((mcc + 31u) / 32u));
  int& cap32 = mode == 0 ? c->nb_cap32 : c->nb_ghost.cap32;
  // staged candidates of ONE plane of halo cells, every cell padded to a multiple of 32
  cap32 = std::max(cap32, (int)std::min<double>((double)(n1 * n1) * (double)((mcc + 31u) & ~31u), (double)tp.cap / n1 * 1.10 + 32.0 * n1 * n1 + 31.0) & ~31);
  const unsigned blocks = (unsigned)((int64_t)tp.tiles_i * tp.tiles_j * tp.tiles_k);
  for (int attempt = 0; attempt < 6; attempt++)
  {
    if (tp.cap > 8191 || tp.gmax > NBH_BIG_MAX_THREADS / 32) return XNB_OK;          // one warp per group of the cell
    const int nwarp = std::max(tp.gmax, 2);
    const size_t smem = nbig_smem_bytes(tp.nh_max, tp.tc_max, tp.gmax, cap32, n1 * n1 * n1, nwarp);
    if (smem + 2048 > XNB_SM_BYTES) return XNB_OK;
    if ((size_t)g.n_cells * c->nbh_slot_words > ((size_t)24 << 30)) return XNB_OK;
    CK(c->pool.ensure((size_t)g.n_cells * c->nbh_slot_words + 64));
    if (mode == 0)
    {
      if (c->nb_cap_trips == 0) c->nb_cap_trips = c->max_neighbors ? (int)(c->max_neighbors / 4 + 8) : (int)(c->nbh_slot_words / mcc / 4 + 8);
      CK(c->cl_rows.ensure((size_t)blocks * tp.gmax * c->nb_cap_trips * 128 + 64, 0, 1.3));
      CK(c->cl_groups.ensure((size_t)blocks * tp.gmax + 16, 0, 1.3));
    }
    CK(cudaMemsetAsync(counters, 0, NB_U32_COUNT * 4, st)); CK(cudaMemsetAsync(totals, 0, 3 * 8, st));
    NbhBitsP bp{};
    bp.emit_rows = mode == 0 ? 1 : 0; bp.sel_mode = mode; bp.slot_words = (int)c->nbh_slot_words; bp.cap_trips = c->nb_cap_trips; bp.max_dist2 = c->nbh_dist * c->nbh_dist;
    // compiled rows in one segment per z-plane of halo cells: the sweep then stages a third of the halo at a time (k_lj_sweep_pl)
    // (plane segments are padded per plane: about a third more rows.  Measured at C4, which rebuilds on every step: build 2.97 -> 3.03 ms,
    // sweep 0.816 -> 0.644 ms, step 4.19 -> 4.08 ms -- it pays even when a list is swept once.)
    const bool planes = mode == 0 && gap == 1 && tp.gmax <= SWEEP_PL_MAX_THREADS / 32 && cap32 <= 8190 && !env_flag("XNB_CL_NO_PLANES");
    bp.planes = planes ? 1 : 0; bp.cap_pl = cap32;
    NbhBitsOut o{c->pool.p, c->cell_stream.p, c->stream_size.p, c->cell_stream_bytes.p, c->stream_off.p, c->cl_groups.p, reinterpret_cast<uint2*>(c->cl_rows.p), counters, totals};
    if (getenv("XNB_TILE_DEBUG")) fprintf(stderr, "[xnb] nbh_big mode %d cap %d cap32 %d gmax %d trips %d slot_words %u smem %zu blocks %u (max cell %u)\n", mode, tp.cap, cap32, tp.gmax, c->nb_cap_trips, c->nbh_slot_words, smem, blocks, mcc);
    // scratch rows per group: every 32-candidate block of the neighbourhood could survive
    // (two blocks of slack per cell: the occupancy of the fullest cell changes from rebuild to rebuild and a reallocation costs milliseconds)
    // first build: every block of the neighbourhood could accept something; afterwards what the last build needed + 30 % (the rows of all
    // resident groups should stay in L2: 128 bytes per row)
    const int scratch_full = n1 * n1 * n1 * (int)std::min<uint32_t>((mcc + 95u) / 32u + 1u, 16u);
    if (c->nb_scratch_rows == 0 || c->nb_scratch_rows > scratch_full) c->nb_scratch_rows = scratch_full;
    const int scratch_rows = c->nb_scratch_rows;
    if ((mcc + 31u) / 32u > 16u) return XNB_OK;      // tag = slot * 16 + block
    CK(c->nb_scratch.ensure((size_t)blocks * tp.gmax * scratch_rows * 33 + 64, 0, 1.3));
    k_nbh_big<<<blocks, 32 * nwarp, smem, st>>>(g, tp, bp, cap32, c->nb_scratch.p, scratch_rows, A.rx, A.ry, A.rz, c->cell_start.p, c->cell_count.p, o, c->d_scalars32.p);
    c->launches++; CK(cudaGetLastError());
    uint32_t h[NB_U32_COUNT]; unsigned long long tot[3];
    {
      char* hp = static_cast<char*>(c->h_pinned);
      CK(cudaMemcpyAsync(hp, counters, NB_U32_COUNT * 4, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(hp + 64, totals, 3 * 8, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(hp + 96, c->d_scalars32.p, 4, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      memcpy(h, hp, NB_U32_COUNT * 4); memcpy(tot, hp + 64, 24);
      uint32_t e; memcpy(&e, hp + 96, 4);
      if (e) { const int rce = decode_device_errors(c, e, st); if (rce) return rce; }
    }
    bool again = false;
    if ((int)h[NB_GMAX] > tp.gmax) { tp.gmax = (int)h[NB_GMAX]; again = true; }
    if ((int)h[NB_CAP] > tp.cap) { tp.cap = ((int)h[NB_CAP] + 1) & ~1; again = true; }
    if ((int)h[NB_SLOTS] > cap32) { cap32 = ((int)h[NB_SLOTS] + 31) & ~31; again = true; }
    if (h[NB_SLOT_WORDS] > c->nbh_slot_words) { c->nbh_slot_words = (uint32_t)(((size_t)(h[NB_SLOT_WORDS] * 1.12) + 64 + 7) & ~(size_t)7); again = true; }
    if (mode == 0 && (int)h[NB_TRIPS] > c->nb_cap_trips) { c->nb_cap_trips = (int)(h[NB_TRIPS] * 1.25) + 8; again = true; }      // lists that grow from rebuild to rebuild (a heating system) must not overflow every time
    if (h[NB_OVERFLOW] & 4u) return XNB_OK;
    if ((int)h[NB_SURV] > scratch_rows) { c->nb_scratch_rows = (int)(h[NB_SURV] * 1.3) + 8; again = true; }
    if (again || h[NB_OVERFLOW]) { if (!again) return c->fail(XNB_ERR_CAPACITY, "chunk_neighbors: large-cell build reported an overflow it cannot size"); continue; }
    c->nb_scratch_rows = std::min(scratch_rows, (int)(h[NB_SURV] * 1.3) + 8);
    if (mode == 0)
    {
      c->pool_used = (int64_t)tot[1]; c->max_neighbors = h[NB_MAX_NBH]; c->n_nonempty_inner = h[NB_NONEMPTY]; c->max_stream = h[NB_MAX_STREAM];
      c->avg_stream = c->n_inner ? (double)tot[2] / (double)c->n_inner : 0.0;
      c->have_nbh = true; c->ghost_lists = g.gl == 0;
      c->nb_cap_trips = std::min(c->nb_cap_trips, (int)(h[NB_TRIPS] * 1.15) + 8);      // room for the next rebuild's longest list: a launch that overflows is run twice
      int rc = cl_finish(c, tp, false, blocks, h[NB_ROWS], (int64_t)tot[0], h[NB_GMAX], st, planes ? 1 : 0, cap32); if (rc) return rc;
    }
    else
    {
      c->pool_used += (int64_t)tot[1]; c->max_neighbors = std::max(c->max_neighbors, h[NB_MAX_NBH]); c->max_stream = std::max(c->max_stream, h[NB_MAX_STREAM]);
      c->nb_ghost.tp = tp; c->nb_ghost.have = true;
      c->ghost_lists = true;
    }
    *done = true;
    return XNB_OK;
  }
  return c->fail(XNB_ERR_CAPACITY, "chunk_neighbors: large-cell build did not converge");
}

static int nbh_bits_run(xnb_ctx* c, int mode, cudaStream_t st, bool* done)
{
  *done = false;
  const GridP& g = c->g;
  const int gap = (int)std::ceil(c->nbh_dist / c->cs);
  int lo[3], hi[3];
  for (int d = 0; d < 3; d++) { lo[d] = mode == 0 ? g.gl : 0; hi[d] = mode == 0 ? g.dims[d] - g.gl : g.dims[d]; }
  if (mode == 1 && g.gl == 0) { *done = true; return XNB_OK; }
  ParticlesP A = c->P(c->cur);
  const uint32_t mcc = std::max<uint32_t>(c->max_cell_count, 1);
  // stream capacity per cell: first guess from the list radius (volume ratio of the sphere to the neighbourhood), grows on demand
  if (c->nbh_slot_words == 0)
  {
    const double frac = std::min(1.0, 4.19 * c->nbh_dist * c->nbh_dist * c->nbh_dist / std::pow((2.0 * gap + 1.0) * c->cs, 3.0));
    const double per = 1.0 + 2.0 * std::pow(2.0 * gap + 1.0, 3.0) * 0.6 + frac * std::pow(2.0 * gap + 1.0, 3.0) * (double)mcc * 1.3;
    c->nbh_slot_words = (uint32_t)((size_t)(2 * (mcc + 1) + (double)mcc * per + 64 + 7) & ~(size_t)7);
  }
  std::vector<ClCand> cands = cl_tile_candidates(c, lo, hi, gap);
  const int n1 = 2 * gap + 1;
  if (n1 * n1 * n1 > 128) return XNB_OK;                    // byte lists: a group header holds 0x80 | neighbour slot -> two-pass build
  const bool u8 = true;
  if (mcc + 3u > 32u * NBH_CELL_BLOCKS && !env_flag("XNB_NBH_FORCE_BITS")) return nbh_big_run(c, mode, st, done);   // a neighbour cell would need more accept masks than k_nbh_bits keeps in registers
  static bool attr_done_dev[XNB_MAX_DEVICES] = {};
  if (!attr_done_dev[c->device % XNB_MAX_DEVICES])
  {
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k_nbh_bits));
    CK(cudaFuncSetAttribute(k_nbh_bits, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024 - (int)fa.sharedSizeBytes));
    attr_done_dev[c->device % XNB_MAX_DEVICES] = true;
  }
  uint32_t* counters = c->d_scalars32.p + 64;                  // NB_U32_COUNT u32
  unsigned long long* totals = c->d_scalars64.p + 4;           // 3 u64
  // list capacity per particle: first guess from the list radius, then what the last build needed
  const double frac = std::min(1.0, 4.19 * c->nbh_dist * c->nbh_dist * c->nbh_dist / std::pow((double)n1 * c->cs, 3.0));
  int& cap_l_state = mode == 0 ? c->nb_cap_l : c->nb_ghost.cap_l;
  if (cap_l_state == 0) cap_l_state = (int)(1.0 + 2.0 * std::min(27.0, (double)n1 * n1 * n1) + frac * (double)(n1 * n1 * n1) * (double)std::max(mcc, 8u) * 1.25 + 16.0);
  for (const ClCand& k : cands)
  {
    ClTileP tp{};
    cl_tile_grid(tp, k, lo, hi, gap);
    const ClTileP& prev = mode == 0 ? c->cl.tp : c->nb_ghost.tp;
    const bool have_prev = mode == 0 ? (c->cl.valid && !c->cl.ghost) : c->nb_ghost.have;
    if (have_prev && prev.ti == tp.ti && prev.tj == tp.tj && prev.tk == tp.tk)
    {
      // same shape as last time: start from the capacities that worked
      tp.gmax = std::max(tp.gmax, prev.gmax); tp.cap = std::max(tp.cap, prev.cap);
    }
    const unsigned blocks = (unsigned)((int64_t)tp.tiles_i * tp.tiles_j * tp.tiles_k);
    bool ok = false, shape_fails = false;
    for (int attempt = 0; attempt < 6; attempt++)
    {
      if (tp.cap > 8191 || tp.gmax > 32) { shape_fails = true; break; }
      const int cap_l = ((cap_l_state + 7) & ~7) + 4;                  // cap_l / 4 odd: the 32 list areas of a warp start in 32 different banks
      // one warp per group of 32 tile particles, the groups of a tile in even rounds over the warps
      const int rounds = (tp.gmax + NBH_BITS_MAX_THREADS / 32 - 1) / (NBH_BITS_MAX_THREADS / 32);
      const int nwarp = std::max(env_int("XNB_NBH_WARPS") > 0 ? std::min(env_int("XNB_NBH_WARPS"), NBH_BITS_MAX_THREADS / 32) : (tp.gmax + rounds - 1) / rounds, 2);
      const size_t smem = nb_smem_bytes(tp.nh_max, tp.tc_max, tp.gmax, tp.cap, cap_l, u8 ? 1 : 2, nwarp);
      if (smem + 2048 > XNB_SM_BYTES) { shape_fails = true; break; }
      if ((size_t)g.n_cells * c->nbh_slot_words > ((size_t)24 << 30)) return XNB_OK;      // streams would not fit: two-pass build with a compact pool
      CK(c->pool.ensure((size_t)g.n_cells * c->nbh_slot_words + 64));
      if (mode == 0)
      {
        // a group owns a fixed block of cap_trips rows (4 list entries per lane and row)
        if (c->nb_cap_trips == 0) c->nb_cap_trips = c->max_neighbors ? (int)(c->max_neighbors / 4 + 4) : (cap_l_state + 3) / 4;
        CK(c->cl_rows.ensure((size_t)blocks * tp.gmax * c->nb_cap_trips * 128 + 64, 0, 1.05));
        CK(c->cl_groups.ensure((size_t)blocks * tp.gmax + 16));
      }
      CK(cudaMemsetAsync(counters, 0, NB_U32_COUNT * 4, st)); CK(cudaMemsetAsync(totals, 0, 3 * 8, st));
      NbhBitsP bp{};
      bp.cap_l = cap_l; bp.emit_rows = mode == 0 ? 1 : 0; bp.sel_mode = mode; bp.slot_words = (int)c->nbh_slot_words; bp.cap_trips = c->nb_cap_trips;
      bp.max_dist2 = c->nbh_dist * c->nbh_dist;
      NbhBitsOut o{c->pool.p, c->cell_stream.p, c->stream_size.p, c->cell_stream_bytes.p, c->stream_off.p, c->cl_groups.p, reinterpret_cast<uint2*>(c->cl_rows.p),
                   counters, totals};
      if (getenv("XNB_TILE_DEBUG")) fprintf(stderr, "[xnb] nbh_bits mode %d tiles %dx%dx%d cap %d gmax %d cap_l %d u8 %d warps %d trips %d slot_words %u smem %zu blocks %u (max cell %u)\n", mode, tp.ti, tp.tj, tp.tk, tp.cap, tp.gmax, cap_l, (int)u8, nwarp, c->nb_cap_trips, c->nbh_slot_words, smem, blocks, mcc);
      k_nbh_bits<<<blocks, 32 * nwarp, smem, st>>>(g, tp, bp, A.rx, A.ry, A.rz, c->cell_start.p, c->cell_count.p, o, c->d_scalars32.p);
      c->launches++; CK(cudaGetLastError());
      uint32_t h[NB_U32_COUNT]; unsigned long long tot[3];
      {
        // one host synchronisation for the small results
        char* hp = static_cast<char*>(c->h_pinned);
        CK(cudaMemcpyAsync(hp, counters, NB_U32_COUNT * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(hp + 64, totals, 3 * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(hp + 96, c->d_scalars32.p, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        memcpy(h, hp, NB_U32_COUNT * 4); memcpy(tot, hp + 64, 24);
        uint32_t e; memcpy(&e, hp + 96, 4);
        if (e) { const int rce = decode_device_errors(c, e, st); if (rce) return rce; }
      }
      if (h[NB_OVERFLOW] & 4u) return XNB_OK;                      // a row needs more masks than the kernel holds: two-pass build
      bool again = false;
      if ((int)h[NB_GMAX] > tp.gmax) { tp.gmax = (int)h[NB_GMAX]; again = true; }
      if ((int)h[NB_CAP] > tp.cap) { tp.cap = ((int)h[NB_CAP] + 1) & ~1; again = true; }
      if ((int)h[NB_SLOTS] > cap_l) { cap_l_state = (int)(h[NB_SLOTS] * 1.1) + 8; again = true; }
      if (h[NB_SLOT_WORDS] > c->nbh_slot_words) { c->nbh_slot_words = (uint32_t)(((size_t)(h[NB_SLOT_WORDS] * 1.12) + 64 + 7) & ~(size_t)7); again = true; }
      if (mode == 0 && (int)h[NB_TRIPS] > c->nb_cap_trips) { c->nb_cap_trips = (int)h[NB_TRIPS] + 3; again = true; }
      if (again || h[NB_OVERFLOW]) { if (!again) return c->fail(XNB_ERR_CAPACITY, "chunk_neighbors: tiled build reported an overflow it cannot size"); continue; }
      cap_l_state = std::max(cap_l_state, std::min(cap_l, (int)(h[NB_SLOTS] * 1.08) + 8));     // keep some room for the next rebuild, no more
      if ((int)(h[NB_SLOTS] * 1.08) + 8 < cap_l_state) cap_l_state = (int)(h[NB_SLOTS] * 1.08) + 8;
      if (mode == 0)
      {
        c->pool_used = (int64_t)tot[1]; c->max_neighbors = h[NB_MAX_NBH]; c->n_nonempty_inner = h[NB_NONEMPTY]; c->max_stream = h[NB_MAX_STREAM];
        c->avg_stream = c->n_inner ? (double)tot[2] / (double)c->n_inner : 0.0;
        c->have_nbh = true; c->ghost_lists = g.gl == 0;
        c->nb_cap_trips = std::min(c->nb_cap_trips, (int)(h[NB_TRIPS] * 1.1) + 4);       // rows of a group closer together next time, with room for the next rebuild
        int rc = cl_finish(c, tp, false, blocks, h[NB_ROWS], (int64_t)tot[0], h[NB_GMAX], st); if (rc) return rc;
      }
      else
      {
        c->pool_used += (int64_t)tot[1]; c->max_neighbors = std::max(c->max_neighbors, h[NB_MAX_NBH]); c->max_stream = std::max(c->max_stream, h[NB_MAX_STREAM]);
        c->nb_ghost.tp = tp; c->nb_ghost.have = true;
        c->ghost_lists = true;
      }
      ok = true;
      break;
    }
    if (ok) { *done = true; return XNB_OK; }
    if (!shape_fails) return c->fail(XNB_ERR_CAPACITY, "chunk_neighbors: tiled build did not converge");
  }
  return XNB_OK;
}

// lists of the ghost-cell particles, on demand (the step loop with ghost = false never reads them)
static int ensure_ghost_lists(xnb_ctx* c, cudaStream_t st)
{
  if (!c->have_nbh || c->ghost_lists) return XNB_OK;
  bool done = false;
  int rc = nbh_bits_run(c, 1, st, &done); if (rc) return rc;
  if (!done) return c->fail(XNB_ERR_CAPACITY, "chunk_neighbors: the ghost-cell lists could not be built on tiles");
  return check_device_errors(c, st);
}

// ---------------------------------------------------------------------------------------------------------------------
// compiled lists derived from existing streams (k_cl_compile): for lists built by the two-pass kernels and for sweeps over
// the ghost cells too (ghost = true).  Not finding a tile shape that fits is not an error: the sweep then reads the streams
// directly (k_lj_sweep).
// ---------------------------------------------------------------------------------------------------------------------
static int cl_prepare(xnb_ctx* c, bool ghost, cudaStream_t st)
{
  c->cl.valid = false;
  if (env_flag("XNB_SWEEP_STREAMS") || c->n_total == 0 || !c->have_nbh) return XNB_OK;
  if (ghost) { int rc = ensure_ghost_lists(c, st); if (rc) return rc; }
  const GridP& g = c->g;
  const int gap = (int)std::ceil(c->nbh_dist / c->cs);
  int lo[3], hi[3];
  for (int d = 0; d < 3; d++) { lo[d] = ghost ? 0 : g.gl; hi[d] = ghost ? g.dims[d] : g.dims[d] - g.gl; }
  std::vector<ClCand> cands = cl_tile_candidates(c, lo, hi, gap);
  uint32_t* counters = c->d_scalars32.p + 112;
  for (const ClCand& k : cands)
  {
    ClTileP tp{};
    cl_tile_grid(tp, k, lo, hi, gap);
    const unsigned blocks = (unsigned)((int64_t)tp.tiles_i * tp.tiles_j * tp.tiles_k);
    size_t smem = cl_sweep_smem_bytes(tp.nh_max, tp.tc_max, tp.cap, tp.gmax);
    bool ok = false;
    for (int attempt = 0; attempt < 5; attempt++)
    {
      if (c->cl_cap_rows == 0) c->cl_cap_rows = (uint32_t)std::min<double>(4.0e9, (double)c->pool_used / 128.0 * 1.10 + 4096.0);
      CK(c->cl_rows.ensure((size_t)c->cl_cap_rows * 128 + 64));
      CK(c->cl_groups.ensure((size_t)blocks * tp.gmax + 16));
      CK(cudaMemsetAsync(counters, 0, 8 * 4, st));      // [0..2] u32 counters, [4..5] one u64 (8-byte aligned): list entries
      const size_t tbytes = (((size_t)(2 * tp.nh_max + 2 * tp.tc_max + 2) * 4 + 15) & ~(size_t)15);
      k_cl_compile<<<blocks, 32 * std::min(tp.gmax, 32), tbytes, st>>>(g, tp, c->cell_start.p, c->cell_count.p, (const uint16_t* const*)c->cell_stream.p, c->cl_groups.p,
                                                                     reinterpret_cast<uint2*>(c->cl_rows.p), c->cl_cap_rows, counters, reinterpret_cast<unsigned long long*>(counters + 4));
      c->launches++; CK(cudaGetLastError());
      uint32_t h[8]; int rc = read_back(c, counters, 8, h, st); if (rc) return rc;
      bool again = false;
      if ((int)h[1] > tp.gmax) { if (h[1] * 32u > 1024u) break; tp.gmax = (int)h[1]; again = true; }
      if ((int)h[2] > tp.cap) { tp.cap = ((int)h[2] + 1) & ~1; again = true; }
      smem = cl_sweep_smem_bytes(tp.nh_max, tp.tc_max, tp.cap, tp.gmax);
      if (tp.cap > 8191 || smem + 1024 + 2048 > XNB_SM_BYTES) break;
      if (!again && h[0] > c->cl_cap_rows) { c->cl_cap_rows = (uint32_t)((double)h[0] * 1.05) + 1024u; again = true; }
      if (again) continue;
      rc = cl_finish(c, tp, ghost, blocks, h[0], (int64_t)(((unsigned long long)h[5] << 32) | h[4]), h[1], st); if (rc) return rc;
      ok = true;
      break;
    }
    if (ok) break;
  }
  if (getenv("XNB_TILE_DEBUG") && !c->cl.valid) fprintf(stderr, "[xnb] compiled lists: no tile shape fits, sweeping the streams\n");
  return XNB_OK;
}

extern "C" {
// ---------------------------------------------------------------------------------------------------------------------
// chunk_neighbors
// ---------------------------------------------------------------------------------------------------------------------
int xnb_chunk_neighbors(xnb_ctx* c, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  int rc = ensure_grid(c); if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const GridP& g = c->g;
  const int64_t n = c->n_total;
  const int gap = (int)std::ceil(c->nbh_dist / c->cs);
  const double md2 = c->nbh_dist * c->nbh_dist;
  uint32_t* s32 = c->d_scalars32.p;
  CK(c->nb_len.ensure((size_t)n + 16, 0, 1.1)); CK(c->nb_cnt.ensure((size_t)n + 16, 0, 1.1)); CK(c->nb_off.ensure((size_t)n + 16, 0, 1.1));
  CK(c->stream_size.ensure((size_t)g.n_cells + 1)); CK(c->stream_size_padded.ensure((size_t)g.n_cells + 1)); CK(c->stream_off.ensure((size_t)g.n_cells + 1));
  CK(c->cell_stream.ensure((size_t)g.n_cells)); CK(c->cell_stream_bytes.ensure((size_t)g.n_cells));
  ParticlesP A = c->P(c->cur);
  if ((rc = t_begin(c, XNB_T_NBH, st))) return rc;
  NbhOut out{c->nb_len.p, c->nb_cnt.p, c->nb_off.p, c->cell_stream.p};

  if (c->nbh_builds > 0) c->last_nbh_interval = c->steps_since_nbh;      // (have_nbh is already false here: ghost_comm_scheme voided the lists)
  c->steps_since_nbh = 0; c->nbh_builds++;
  // ---- tiled form (k_nbh_bits, one kernel: streams of the inner cells + compiled rows of the sweep) when a tile fits shared
  // memory, else the per-particle two-pass kernels.  Lists filtered by ChunkNeighborsConfig::half_symmetric / skip_ghosts are
  // built by the two-pass kernels.
  c->ghost_lists = false;
  if (n > 0 && !env_flag("XNB_NBH_UNTILED") && !c->nbh_half_symmetric && !c->nbh_skip_ghosts)
  {
    // occupancy of the cells (tile shape and capacities follow from it); ghost_comm_scheme has usually read it already
    if (!c->cell_stats_valid)
    {
      CK(cudaMemsetAsync(s32 + 5, 0, 8, st));
      LAUNCH(k_cell_stats, nblk(g.n_cells, 256), 256, st, g, c->cell_count.p, s32 + 5);
      uint32_t cs[2] = {0, 0};
      rc = read_back(c, s32 + 5, 2, cs, st); if (rc) return rc;
      c->max_cell_count = std::max<uint32_t>(cs[0], 1); c->n_nonempty_inner = cs[1]; c->cell_stats_valid = true;
    }
    bool done = false;
    rc = nbh_bits_run(c, 0, st, &done); if (rc) return rc;
    if (done) return t_end(c, XNB_T_NBH, st);       // (the build's own host wait read the device error word too)
  }
  // ---- per-particle two-pass form (count -> sizes -> scan -> fill)
  // the AMR tables prune whole sub-cells when they describe the current in-cell order (rebuild_amr ran after the last binning)
  const bool prune = c->amr_current && c->max_side > 1 && c->sub_grid_cells.p && !env_flag("XNB_NBH_NO_SUBCELLS");
  const unsigned long long* sgs_prune = prune ? c->sub_grid_start.p : nullptr;
  const uint32_t* sgc_prune = prune ? c->sub_grid_cells.p : nullptr;
  if (getenv("XNB_TILE_DEBUG")) fprintf(stderr, "[xnb] two-pass neighbour build: %lld particles, sub-cell pruning %s (max side %u, tables current %d)\n", (long long)n, prune ? "on" : "off", c->max_side, (int)c->amr_current);
  if (n) LAUNCH((k_nbh_build<false>), nblk(n, 128), 128, st, g, (int)n, gap, md2, (int)c->nbh_half_symmetric, (int)c->nbh_skip_ghosts, sgs_prune, sgc_prune, A.rx, A.ry, A.rz, c->atom_cell[c->cur_ac].p, c->cell_start.p, c->cell_count.p, out, s32);
  CK(cudaMemsetAsync(s32 + 3, 0, 16, st)); CK(cudaMemsetAsync(s32 + 109, 0, 8, st)); CK(cudaMemsetAsync(c->d_scalars64.p + 2, 0, 8, st));
  LAUNCH(k_nbh_cell_sizes, nblk((int64_t)g.n_cells * 32, 128), 128, st, g, g.n_cells, c->cell_start.p, c->cell_count.p, c->nb_len.p, c->nb_cnt.p, c->nb_off.p,
         c->stream_size.p, c->stream_size_padded.p, s32 + 3, s32 + 5, s32 + 6, s32 + 109, c->d_scalars64.p + 2, s32);
  rc = scan_exclusive<uint32_t, unsigned long long>(c, c->stream_size_padded.p, c->stream_off.p, (size_t)g.n_cells, c->d_scalars64.p + 1, c->scan_tmp64, st); if (rc) return rc;
  unsigned long long tot = 0, tot2[2] = {0, 0}; uint32_t mx = 0;
  rc = read_back(c, c->d_scalars64.p + 1, 2, tot2, st); if (rc) return rc;
  tot = tot2[0];
  uint32_t sc[4] = {0, 0, 0, 0}, sc2[2] = {0, 0};
  rc = read_back(c, s32 + 3, 4, sc, st); if (rc) return rc;
  rc = read_back(c, s32 + 109, 2, sc2, st); if (rc) return rc;
  mx = sc[0];
  c->pool_used = (int64_t)tot; c->max_neighbors = mx; c->max_cell_count = sc[2]; c->max_stream = sc[3]; c->n_nonempty_inner = sc2[1];
  c->avg_stream = c->n_inner ? (double)tot2[1] / (double)c->n_inner : 0.0;      // padded u16 words per inner particle
  // realloc_stream_pool (chunk_neighbors.h:70-96): grow with the reference's 5% head-room (update-particles.msp:20)
  CK(c->pool.ensure((size_t)tot + 64, 0, 1.05));
  LAUNCH(k_nbh_pointers, nblk(g.n_cells, 256), 256, st, g.n_cells, c->pool.p, c->stream_off.p, c->stream_size.p, c->cell_stream.p, c->cell_stream_bytes.p);
  if (n) LAUNCH((k_nbh_build<true>), nblk(n, 128), 128, st, g, (int)n, gap, md2, (int)c->nbh_half_symmetric, (int)c->nbh_skip_ghosts, sgs_prune, sgc_prune, A.rx, A.ry, A.rz, c->atom_cell[c->cur_ac].p, c->cell_start.p, c->cell_count.p, out, s32);
  c->have_nbh = true; c->ghost_lists = true;          // the two-pass kernels build the lists of every cell
  if ((rc = cl_prepare(c, false, st))) return rc;
  if ((rc = t_end(c, XNB_T_NBH, st))) return rc;
  return check_device_errors(c, st);
}

// ---------------------------------------------------------------------------------------------------------------------
// force / integrate
// ---------------------------------------------------------------------------------------------------------------------
int xnb_zero_particle_force(xnb_ctx* c, int ghost, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  ParticlesP A = c->P(c->cur);
  const int64_t n = ghost ? c->n_total : c->n_inner;
  if (n) LAUNCH(k_zero_force, nblk(n, 256), 256, st, (int)n, A.fx, A.fy, A.fz);
  return XNB_OK;
}

} // extern "C"
static LJP make_lj(double eps, double sig, double rcut) { LJP p; p.eps24 = 24.0 * eps; p.sig2 = sig * sig; p.rcut2 = rcut * rcut; p.eps4 = 4.0 * eps; p.neg_eps48 = -48.0 * eps; p.eps = eps; p.sig = sig; return p; }

// tile shape of the pair sweep: ti x tj cells per block, thread count and staging capacities from the cell occupancy
struct TileCfg { TileP tp; int threads; size_t smem; unsigned blocks; };
static TileCfg make_tiles(const xnb_ctx* c, bool ghost)
{
  const GridP& g = c->g;
  TileCfg t{};
  TileP& tp = t.tp;
  tp.gap = (int)std::ceil(c->nbh_dist / c->cs);
  for (int d = 0; d < 3; d++) { tp.lo[d] = ghost ? 0 : g.gl; tp.hi[d] = ghost ? g.dims[d] : g.dims[d] - g.gl; }
  const int ni = tp.hi[0] - tp.lo[0], nj = tp.hi[1] - tp.lo[1], nk = tp.hi[2] - tp.lo[2];
  // occupancy of the non-empty inner cells (clusters + voids: the dense tiles set the capacities)
  const double ne = (double)std::max<int64_t>(c->n_nonempty_inner, 1);
  const double avg = std::max((double)c->n_inner / ne, 1.0);
  const double mx = (double)std::max<uint32_t>(c->max_cell_count, 1);
  const double wpp = std::max(c->avg_stream, 8.0);                  // padded stream words per inner particle
  const size_t SM_BYTES = 227 * 1024;
  static const int shapes[][2] = {{4, 2}, {4, 1}, {2, 2}, {2, 1}, {1, 1}, {8, 2}, {4, 4}, {4, 3}, {8, 1}};
  double best_score = -1; int bi = 1, bj = 1, bthreads = 64; size_t bsmem = 0; int bcap = 0, bcaps = 0;
  const int eti = env_int("XNB_TILE_I"), etj = env_int("XNB_TILE_J");
  const double slack = 1.10;
  for (const auto& sh : shapes)
  {
    const int ti = std::min(sh[0], ni), tj = std::min(sh[1], nj);
    if (eti > 0 && ti != std::min(eti, ni)) continue;
    if (etj > 0 && tj != std::min(etj, nj)) continue;
    const int tc = ti * tj;
    if (tc > 32) continue;
    const double tile_avg = avg * tc;
    int threads = 32 * (int)std::ceil(std::min(mx * tc, tile_avg * 1.08 + 8.0) / 32.0);
    if (threads > SWEEP_MAX_THREADS) continue;
    threads = std::max(threads, 64);
    const int hx = ti + 2 * tp.gap, hy = tj + 2 * tp.gap, hz = 2 * tp.gap + 1;
    const int nh = hx * hy * hz;
    const size_t head = (((size_t)(2 * nh + 1 + 4 * tc + 2) * 4 + 15) & ~(size_t)15);
    const size_t cap = (size_t)std::min(mx * nh, avg * nh * slack + 64.0);
    if (cap > 32767) continue;
    size_t cap_s = (size_t)std::min((double)c->max_stream * tc, wpp * tile_avg * slack + 256.0);
    cap_s = (cap_s + 7) & ~(size_t)7;
    const size_t smem = head + (cap_s + 16) * 2 + cap * 24;
    if (smem + 1024 > SM_BYTES) continue;
    const int by_smem = (int)(SM_BYTES / (smem + 1024));
    const int by_regs = 65536 / (threads * 104);
    const int by_warps = 64 / (threads / 32);
    const int resident = std::min(std::min(by_smem, by_regs), std::min(by_warps, 32));
    if (resident < 1) continue;
    const double warps = resident * (tile_avg / 32.0);                // useful resident warps per SM
    const double staged_per_particle = (double)smem / tile_avg;       // tie-break: less staging per particle
    const double score = warps + (resident >= 2 ? 4.0 : 0.0) - 1e-3 * staged_per_particle;
    if (score > best_score) { best_score = score; bi = ti; bj = tj; bthreads = threads; bsmem = smem; bcap = (int)cap; bcaps = (int)cap_s; }
  }
  if (best_score < 0)
  {
    // nothing fits (fat cells): 1x1 tiles without staging, the kernel takes its global-memory path
    bi = bj = 1; bthreads = 32 * (int)std::min((double)(SWEEP_MAX_THREADS / 32), std::max(2.0, std::ceil(avg / 32.0))); bcap = 0; bcaps = 0;
    const int nh = (1 + 2 * tp.gap) * (1 + 2 * tp.gap) * (1 + 2 * tp.gap);
    bsmem = (((size_t)(2 * nh + 1 + 4 + 2) * 4 + 15) & ~(size_t)15) + 16;
  }
  if (env_int("XNB_FORCE_THREADS") > 0) bthreads = std::min(env_int("XNB_FORCE_THREADS"), SWEEP_MAX_THREADS);
  tp.ti = bi; tp.tj = bj;
  tp.tiles_i = (ni + bi - 1) / bi; tp.tiles_j = (nj + bj - 1) / bj;
  tp.hx = bi + 2 * tp.gap; tp.hy = bj + 2 * tp.gap; tp.hz = 2 * tp.gap + 1;
  tp.cap = bcap; tp.cap_s = bcaps;
  t.threads = bthreads; t.smem = bsmem;
  t.blocks = (unsigned)((int64_t)tp.tiles_i * tp.tiles_j * nk);
  if (getenv("XNB_TILE_DEBUG")) fprintf(stderr, "[xnb] force tiles %dx%d threads %d smem %zu cap %d cap_s %d blocks %u (avg %.1f max %.0f wpp %.1f)\n", bi, bj, bthreads, bsmem, bcap, bcaps, t.blocks, avg, mx, wpp);
  return t;
}

// part: 0 = every tile, 1 = interior tiles only, 2 = boundary tiles only (compiled lists; 1 and 2 are timed by the caller)
template <class F, int MODE, bool EV>
static int launch_force_f(xnb_ctx* c, bool ghost, const F& lj, double dth, double* fxo, double* fyo, double* fzo, double** evp_out, unsigned* nblocks_out, cudaStream_t st,
                        const unsigned long long* skip_if_nonzero, int part)
{
  ParticlesP A = c->P(c->cur);
  int rc;
  // a full-list sweep over half_symmetric lists would apply every pair to one of its two particles only (the reference's operator
  // does exactly that, silently); here it is an error: use xnb_lennard_jones_force_symmetric + xnb_update_force_from_ghost
  if (c->nbh_half_symmetric) return c->fail(XNB_ERR_INVALID, "the full-list pair sweep needs full lists: chunk_neighbors was configured half_symmetric");
  if (ghost && (rc = ensure_ghost_lists(c, st))) return rc;
  const bool sweep_streams = env_flag("XNB_SWEEP_STREAMS");         // diagnostic: sweep the reference-format streams directly (k_lj_sweep)
  if (!sweep_streams && (c->cl.ghost != ghost || !c->cl.valid)) { if ((rc = cl_prepare(c, ghost, st))) return rc; c->cl.ghost = ghost; }
  if (c->cl.valid && !sweep_streams)
  {
    const xnb_ctx::ClCfg& k = c->cl;
    if (EV) { CK(c->ev_partials.ensure((size_t)k.blocks * 7 + 16)); if (evp_out) *evp_out = c->ev_partials.p; if (nblocks_out) *nblocks_out = k.blocks; }
    static bool cl_attr_done_dev[XNB_MAX_DEVICES][3][2][4] = {};
    auto& cl_attr_done = cl_attr_done_dev[c->device % XNB_MAX_DEVICES];
#define XNB_CL_LAUNCH(VAR) do { \
      if (!cl_attr_done[MODE][EV ? 1 : 0][VAR]) { \
        cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, (k_lj_sweep_cl<F, MODE, EV, VAR>))); \
        CK(cudaFuncSetAttribute((k_lj_sweep_cl<F, MODE, EV, VAR>), cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024 - (int)fa.sharedSizeBytes)); \
        cl_attr_done[MODE][EV ? 1 : 0][VAR] = true; } \
      if (part == 0 && (rc = t_begin(c, XNB_T_FORCE, st))) return rc; \
      k_lj_sweep_cl<F, MODE, EV, VAR><<<nb, k.threads, k.smem, st>>>(c->g, k.tp, (int)c->n_inner, (int)c->n_total, lj, dth, c->next_half, A.rx, A.ry, A.rz, A.vx, A.vy, A.vz, \
          fxo ? fxo : A.fx, fyo ? fyo : A.fy, fzo ? fzo : A.fz, A.type, c->mass.p, c->cell_start.p, c->cell_count.p, c->cl_groups.p, \
          reinterpret_cast<const uint2*>(c->cl_rows.p), EV ? c->ev_partials.p : nullptr, c->d_scalars32.p, skip_if_nonzero, tl); } while (0)
    const unsigned nb = part == 0 ? k.blocks : part == 1 ? k.n_interior : k.n_boundary;
    const uint32_t* tl = part == 0 ? nullptr : part == 1 ? c->cl_tile_list.p : c->cl_tile_list.p + k.n_interior;
    if (nb == 0) return XNB_OK;
    if (k.planes)
    {
      {
      // large cells: rows in plane segments, one plane of the halo staged at a time (xnb_sweep_pl.cuh)
      static bool pl_attr_done_dev[XNB_MAX_DEVICES][3][2] = {};
      auto& pl_attr_done = pl_attr_done_dev[c->device % XNB_MAX_DEVICES];
      if (!pl_attr_done[MODE][EV ? 1 : 0])
      {
        cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, (k_lj_sweep_pl<F, MODE, EV>)));
        CK(cudaFuncSetAttribute((k_lj_sweep_pl<F, MODE, EV>), cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024 - (int)fa.sharedSizeBytes));
        pl_attr_done[MODE][EV ? 1 : 0] = true;
      }
      if (part == 0 && (rc = t_begin(c, XNB_T_FORCE, st))) return rc;
      k_lj_sweep_pl<F, MODE, EV><<<nb, k.threads, k.smem, st>>>(c->g, k.tp, k.cap_pl, (int)c->n_inner, (int)c->n_total, lj, dth, c->next_half, A.rx, A.ry, A.rz, A.vx, A.vy, A.vz,
          fxo ? fxo : A.fx, fyo ? fyo : A.fy, fzo ? fzo : A.fz, A.type, c->mass.p, c->cell_start.p, c->cell_count.p, c->cl_groups.p,
          reinterpret_cast<const uint2*>(c->cl_rows.p), EV ? c->ev_partials.p : nullptr, c->d_scalars32.p, skip_if_nonzero, tl);
      c->launches++; CK(cudaGetLastError());
      return part == 0 ? t_end(c, XNB_T_FORCE, st) : XNB_OK;
      }
    }
    if (k.var == 0) XNB_CL_LAUNCH(0); else if (k.var == 1) XNB_CL_LAUNCH(1); else if (k.var == 2) XNB_CL_LAUNCH(2); else XNB_CL_LAUNCH(3);
#undef XNB_CL_LAUNCH
    c->launches++; CK(cudaGetLastError());
    return part == 0 ? t_end(c, XNB_T_FORCE, st) : XNB_OK;
  }
  if constexpr (MODE == 2) return c->fail(XNB_ERR_INVALID, "fused next first half: needs the compiled lists");
  else {
  const TileCfg t = make_tiles(c, ghost);
  if (t.blocks == 0) return XNB_OK;
  if (EV) { CK(c->ev_partials.ensure((size_t)t.blocks * 7 + 16)); if (evp_out) *evp_out = c->ev_partials.p; if (nblocks_out) *nblocks_out = t.blocks; }
  static bool attr_done_dev[XNB_MAX_DEVICES][2][2] = {};
  auto& attr_done = attr_done_dev[c->device % XNB_MAX_DEVICES];
  if (!attr_done[MODE][EV ? 1 : 0])
  {
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k_lj_sweep<F, MODE, EV>));
    CK(cudaFuncSetAttribute(k_lj_sweep<F, MODE, EV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024 - (int)fa.sharedSizeBytes));
    attr_done[MODE][EV ? 1 : 0] = true;
  }
  if ((rc = t_begin(c, XNB_T_FORCE, st))) return rc;
  k_lj_sweep<F, MODE, EV><<<t.blocks, t.threads, t.smem, st>>>(c->g, t.tp, (int)c->n_inner, (int)c->n_total, lj, dth, A.rx, A.ry, A.rz, A.vx, A.vy, A.vz,
      fxo ? fxo : A.fx, fyo ? fyo : A.fy, fzo ? fzo : A.fz, A.type, c->mass.p, c->cell_start.p, c->cell_count.p, (const uint16_t* const*)c->cell_stream.p,
      c->stream_size.p, EV ? c->ev_partials.p : nullptr, skip_if_nonzero);
  c->launches++; CK(cudaGetLastError());
  return t_end(c, XNB_T_FORCE, st);
  }
}

// the functor the sweeps are instantiated with: the restated LJ functor (four-candidate hook) or the literal reference form
// through the generic buffer-less call (xnb_set_pair_functor)
template <int MODE, bool EV>
static int launch_force(xnb_ctx* c, bool ghost, const LJP& lj, double dth, double* fxo, double* fyo, double* fzo, double** evp_out, unsigned* nblocks_out, cudaStream_t st,
                        const unsigned long long* skip_if_nonzero = nullptr, int part = 0)
{
  if (c->pair_functor == XNB_FUNCTOR_LJ_REFERENCE_FORM)
    return launch_force_f<LennardJonesForceFunctorRef, MODE, EV>(c, ghost, LennardJonesForceFunctorRef{lj}, dth, fxo, fyo, fzo, evp_out, nblocks_out, st, skip_if_nonzero, part);
  return launch_force_f<LennardJonesForceFunctor, MODE, EV>(c, ghost, LennardJonesForceFunctor{lj}, dth, fxo, fyo, fzo, evp_out, nblocks_out, st, skip_if_nonzero, part);
}

extern "C" {
int xnb_set_pair_functor(xnb_ctx* c, int form)
{
  if (!c) return XNB_ERR_INVALID;
  if (form != XNB_FUNCTOR_LJ && form != XNB_FUNCTOR_LJ_REFERENCE_FORM) return c->fail(XNB_ERR_INVALID, "set_pair_functor: unknown functor");
  c->pair_functor = form;
  return XNB_OK;
}

int xnb_lennard_jones_force(xnb_ctx* c, double eps, double sig, double rcut, int ghost, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  if (!c->have_nbh) return c->fail(XNB_ERR_INVALID, "lennard_jones_force: no neighbour list (run xnb_chunk_neighbors)");
  c->rcut_max = std::max(c->rcut_max, rcut);    // lennard_jones.cu:193
  return launch_force<0, false>(c, ghost != 0, make_lj(eps, sig, rcut), 0.0, nullptr, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

// op `gravitational_force` (contribs/pi/gravitational_force.cu:161-217): the second functor of the concept, with a per-neighbour
// field (the neighbour's type -> mass).  Goes through the general sweep (xnb_pair_generic.cuh); buffer_form selects which of the two
// call forms of the functor the sweep uses (0: buffer-less, 1: ComputePairBuffer2).  Accumulates into fx,fy,fz.
int xnb_gravitational_force(xnb_ctx* c, double G, double rcut, int ghost, int buffer_form, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  if (!c->have_nbh) return c->fail(XNB_ERR_INVALID, "gravitational_force: no neighbour list (run xnb_chunk_neighbors)");
  if (c->nbh_half_symmetric) return c->fail(XNB_ERR_INVALID, "gravitational_force: the full-list pair sweep needs full lists");
  if (ghost) return c->fail(XNB_ERR_INVALID, "gravitational_force: ghost = true is not offered by the general sweep");
  if (!c->mass.p || c->n_types <= 0) return c->fail(XNB_ERR_INVALID, "gravitational_force: particle type property 'mass' is missing (xnb_set_type_mass)");
  c->rcut_max = std::max(c->rcut_max, rcut);    // gravitational_force.cu:184
  cudaStream_t st = (cudaStream_t)stream;
  ParticlesP A = c->P(c->cur);
  int rc;
  if ((rc = t_begin(c, XNB_T_FORCE, st))) return rc;
  if (c->n_inner)
  {
    const CellsView cells{c->cell_start.p, c->cell_count.p, A.rx, A.ry, A.rz, A.vx, A.vy, A.vz, A.id, A.type};
    const GravitationalForceFunctor f{G, c->mass.p};
    if (buffer_form)
      LAUNCH((k_pair_sweep_generic<GravitationalForceFunctor, true>), nblk(c->n_inner, 128), 128, st, c->g, (int)c->n_inner, f, rcut * rcut, cells, A.fx, A.fy, A.fz,
             c->atom_cell[c->cur_ac].p, (const uint16_t* const*)c->cell_stream.p, c->d_scalars32.p);
    else
      LAUNCH((k_pair_sweep_generic<GravitationalForceFunctor, false>), nblk(c->n_inner, 128), 128, st, c->g, (int)c->n_inner, f, rcut * rcut, cells, A.fx, A.fy, A.fz,
             c->atom_cell[c->cur_ac].p, (const uint16_t* const*)c->cell_stream.p, c->d_scalars32.p);
  }
  return t_end(c, XNB_T_FORCE, st);
}

// op `average_neighbors_scalar` (src/compute/average_neighbors.cu:102-215): a functor with a per-neighbour scalar field and a particle
// context (start / pair / stop), through the general sweep.  nbh_field: XNB_FIELD_*.  The result lives in a ctx-owned generic real
// field (one double per particle, current particle order): xnb_get_generic_field / xnb_view_generic_field.
int xnb_average_neighbors(xnb_ctx* c, double rcut, const double weight_function[4], int nbh_field, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  if (!c->have_nbh) return c->fail(XNB_ERR_INVALID, "average_neighbors: no neighbour list (run xnb_chunk_neighbors)");
  if (c->nbh_half_symmetric) return c->fail(XNB_ERR_INVALID, "average_neighbors: needs full lists");
  if (nbh_field < 0 || nbh_field > XNB_FIELD_TYPE) return c->fail(XNB_ERR_INVALID, "average_neighbors: unknown nbh_field");
  c->rcut_max = std::max(c->rcut_max, rcut);
  cudaStream_t st = (cudaStream_t)stream;
  ParticlesP A = c->P(c->cur);
  CK(c->generic_field.ensure((size_t)c->n_total + 16, 0, 1.2));
  CK(cudaMemsetAsync(c->generic_field.p, 0, (size_t)c->n_total * 8, st));
  if (c->n_inner)
  {
    const CellsView cells{c->cell_start.p, c->cell_count.p, A.rx, A.ry, A.rz, A.vx, A.vy, A.vz, A.id, A.type};
    const double* f64[9] = {A.rx, A.ry, A.rz, A.vx, A.vy, A.vz, A.fx, A.fy, A.fz};
    AverageNeighborsFunctor f{rcut * rcut, weight_function ? weight_function[0] : 1.0, weight_function ? weight_function[1] : 0.0, weight_function ? weight_function[2] : 0.0,
                              weight_function ? weight_function[3] : 0.0, c->generic_field.p, nbh_field <= XNB_FIELD_FZ ? f64[nbh_field] : nullptr,
                              nbh_field == XNB_FIELD_ID ? A.id : nullptr, nbh_field == XNB_FIELD_TYPE ? A.type : nullptr};
    LAUNCH((k_pair_sweep_context<AverageNeighborsFunctor>), nblk(c->n_inner, 128), 128, st, c->g, (int)c->n_inner, f, rcut * rcut, cells, c->atom_cell[c->cur_ac].p,
           (const uint16_t* const*)c->cell_stream.p);
  }
  c->generic_field_n = c->n_inner;
  return XNB_OK;
}

int xnb_get_generic_field(xnb_ctx* c, double* out)
{
  if (!c || !out) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  if (!c->generic_field.p || c->generic_field_n != c->n_inner) return c->fail(XNB_ERR_INVALID, "no generic field for the current particle set");
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out, c->generic_field.p, (size_t)c->n_inner * 8, cudaMemcpyDeviceToHost));
  return XNB_OK;
}

int xnb_set_chunk_neighbors_config(xnb_ctx* c, int half_symmetric, int skip_ghosts)
{
  if (!c) return XNB_ERR_INVALID;
  const bool h = half_symmetric != 0, s = skip_ghosts != 0;
  if (h != c->nbh_half_symmetric || s != c->nbh_skip_ghosts) { c->have_nbh = false; c->cl.valid = false; }    // lists of the other kind are void
  c->nbh_half_symmetric = h; c->nbh_skip_ghosts = s;
  return XNB_OK;
}

int xnb_lennard_jones_force_symmetric(xnb_ctx* c, double eps, double sig, double rcut, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  if (!c->have_nbh) return c->fail(XNB_ERR_INVALID, "no neighbour list (run xnb_chunk_neighbors)");
  if (!c->nbh_half_symmetric) return c->fail(XNB_ERR_INVALID, "the symmetric sweep needs half_symmetric lists (xnb_set_chunk_neighbors_config)");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  ParticlesP A = c->P(c->cur);
  int rc;
  if ((rc = t_begin(c, XNB_T_FORCE, st))) return rc;
  const LJP lj = make_lj(eps, sig, rcut);
  if (c->n_inner)
  {
    if (c->pair_functor == XNB_FUNCTOR_LJ_REFERENCE_FORM)
      LAUNCH((k_pair_sweep_sym<LennardJonesForceFunctorRef>), nblk(c->n_inner, 128), 128, st, c->g, (int)c->n_inner, LennardJonesForceFunctorRef{lj}, A.rx, A.ry, A.rz, A.fx, A.fy, A.fz,
             c->atom_cell[c->cur_ac].p, c->cell_start.p, c->cell_count.p, (const uint16_t* const*)c->cell_stream.p);
    else
      LAUNCH((k_pair_sweep_sym<LennardJonesForceFunctor>), nblk(c->n_inner, 128), 128, st, c->g, (int)c->n_inner, LennardJonesForceFunctor{lj}, A.rx, A.ry, A.rz, A.fx, A.fy, A.fz,
             c->atom_cell[c->cur_ac].p, c->cell_start.p, c->cell_count.p, (const uint16_t* const*)c->cell_stream.p);
  }
  return t_end(c, XNB_T_FORCE, st);
}

int xnb_update_force_from_ghost(xnb_ctx* c, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (c->n_send == 0 && c->n_ghost == 0) return XNB_OK;
  const size_t ns = (size_t)c->n_send;
  const int self_first = (int)c->h_send_base[(size_t)c->rank], self_end = (int)c->h_send_base[(size_t)c->rank + 1];
  const uint32_t self_dst = (uint32_t)(c->n_inner + c->h_recv_base[(size_t)c->rank]);
  int rc;
  if ((rc = t_begin(c, XNB_T_GHOST_UPDATE, st))) return rc;
  if (c->nranks > 1)
  {
    // the forward exchange run backwards: every rank returns the forces of the ghosts it holds to the rank that sent them
    if (!c->comm) return c->fail(XNB_ERR_NCCL, "nranks > 1 needs an NCCL communicator");
    CK(c->stage.ensure(ns * 3 + 16, 0, 1.2));
    NK(g_nccl.GroupStart());
    for (int p = 0; p < c->nranks; p++)
    {
      if (p == c->rank) continue;
      const size_t s0 = c->h_send_base[(size_t)p], sn = c->h_send_base[(size_t)p + 1] - s0;
      const size_t r0 = c->h_recv_base[(size_t)p], rn = c->h_recv_base[(size_t)p + 1] - r0;
      if (rn) for (int f = 0; f < 3; f++) NK(g_nccl.Send(c->f64[c->cur][6 + f].p + (size_t)c->n_inner + r0, rn, nccl_float64, p, c->comm, st));
      if (sn) for (int f = 0; f < 3; f++) NK(g_nccl.Recv(c->stage.p + (size_t)f * ns + s0, sn, nccl_float64, p, c->comm, st));
    }
    NK(g_nccl.GroupEnd());
  }
  ParticlesP A = c->P(c->cur);
  if (ns) LAUNCH(k_ghost_unpack_add, nblk((int64_t)ns, 256), 256, st, (int)ns, c->send_src.p, self_first, self_end, self_dst, c->stage.p, A.fx, A.fy, A.fz);
  return t_end(c, XNB_T_GHOST_UPDATE, st);
}

int xnb_divide_force_by_mass(xnb_ctx* c, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  ParticlesP A = c->P(c->cur);
  if (c->n_inner) LAUNCH(k_divide_force_by_mass, nblk(c->n_inner, 256), 256, st, (int)c->n_inner, A.fx, A.fy, A.fz, A.type, c->mass.p);
  return XNB_OK;
}

int xnb_push_f_v_r(xnb_ctx* c, double dt, double dt_scale, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  c->amr_current = false;      // positions change: the sub-cell tables no longer bound them
  cudaStream_t st = (cudaStream_t)stream;
  ParticlesP A = c->P(c->cur);
  const double delta_t = dt * dt_scale, delta_t2 = delta_t * delta_t * 0.5;     // push_vec3_2nd_order.h:87-88
  if (c->n_inner) LAUNCH(k_push_f_v_r, nblk(c->n_inner, 256), 256, st, (int)c->n_inner, delta_t, delta_t2, A.rx, A.ry, A.rz, A.vx, A.vy, A.vz, A.fx, A.fy, A.fz);
  return XNB_OK;
}

int xnb_push_f_v(xnb_ctx* c, double dt, double dt_scale, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  ParticlesP A = c->P(c->cur);
  if (c->n_inner) LAUNCH(k_push_f_v, nblk(c->n_inner, 256), 256, st, (int)c->n_inner, dt * dt_scale, A.vx, A.vy, A.vz, A.fx, A.fy, A.fz);
  return XNB_OK;
}

int xnb_read_displ_over(xnb_ctx* c, uint64_t* count_out, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  // MPI_Allreduce(SUM, 1 x u64) of particle_displ_over.cu:174, into a slot of its own: the local count stays intact, so calling this twice
  // after one xnb_verlet_first_half returns the same number
  const unsigned long long* src = c->d_scalars64.p;
  if (c->nranks > 1) { if (!c->comm) return c->fail(XNB_ERR_NCCL, "no communicator"); NK(g_nccl.AllReduce(c->d_scalars64.p, c->d_scalars64.p + 7, 1, nccl_uint64, nccl_sum, c->comm, st)); src = c->d_scalars64.p + 7; }
  unsigned long long v = 0;
  int rc = read_back(c, src, 1, &v, st); if (rc) return rc;
  if (count_out) *count_out = v;
  return XNB_OK;
}

int xnb_particle_displ_over(xnb_ctx* c, uint64_t* count_out, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  ParticlesP A = c->P(c->cur);
  CK(cudaMemsetAsync(c->d_scalars64.p, 0, 8, st));
  if (c->n_inner) LAUNCH(k_displ_over, nblk(c->n_inner, 256), 256, st, c->g, (int)c->n_inner, A.rx, A.ry, A.rz, c->atom_cell[c->cur_ac].p, c->backup.p,
                         c->max_displ * c->max_displ, c->d_scalars64.p);
  return xnb_read_displ_over(c, count_out, stream);
}

int xnb_verlet_first_half(xnb_ctx* c, double dt, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  return verlet_first_half(c, dt, (cudaStream_t)stream, c->d_scalars64.p);
}

int xnb_force_and_second_half(xnb_ctx* c, double eps, double sig, double rcut, double dth, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  if (!c->have_nbh) return c->fail(XNB_ERR_INVALID, "no neighbour list (run xnb_chunk_neighbors)");
  return launch_force<1, false>(c, false, make_lj(eps, sig, rcut), dth, nullptr, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

} // extern "C"
// xnb_step_host: the inner positions (and the particle order) of this step are final once `after` has happened on its stream;
// their device-to-host copies start then, on a second stream, and overlap whatever the step still has to compute
static int host_out_positions(xnb_ctx* c, cudaEvent_t after, bool rebuilt)
{
  xnb_ctx::HostOut& h = c->hout;
  if (!h.active || h.issued) return XNB_OK;
  h.issued = true;
  CK(cudaStreamWaitEvent(c->st_d2h, after, 0));
  const size_t n = (size_t)c->n_inner;
  if (n > h.capacity) { h.overflow = true; return XNB_OK; }       // (several ranks: migration brought in more particles than the caller's arrays hold)
  for (int f = 0; f < 3; f++) if (h.r[f] && n) CK(cudaMemcpyAsync(h.r[f], c->f64[c->cur][f].p, n * 8, cudaMemcpyDeviceToHost, c->st_d2h));
  h.id_copied = h.id && (rebuilt || h.id_always);
  if (h.id_copied && n) CK(cudaMemcpyAsync(h.id, c->idb[c->cur].p, n * 8, cudaMemcpyDeviceToHost, c->st_d2h));
  return XNB_OK;
}

// move_particles + parallel_update_particles (update-particles.msp:47-68)
static int move_and_update_full(xnb_ctx* c, void* stream)
{
  int rc;
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = t_begin(c, XNB_T_BIN, st))) return rc;
  c->in_rebuild_chain = true;
  rc = xnb_move_particles(c, stream);
  c->in_rebuild_chain = false;
  if (rc) return rc;
  if ((rc = xnb_rebuild_amr(c, stream))) return rc;
  if ((rc = xnb_backup_r(c, stream))) return rc;
  if ((rc = t_end(c, XNB_T_BIN, st))) return rc;
  if (c->hout.active && !c->hout.issued) { CK(cudaEventRecord(c->ev_d2h_go, st)); if ((rc = host_out_positions(c, c->ev_d2h_go, true))) return rc; }
  if ((rc = t_begin(c, XNB_T_GHOST_SCHEME, st))) return rc;
  if ((rc = xnb_ghost_comm_scheme(c, stream))) return rc;
  if ((rc = xnb_ghost_update_all(c, stream))) return rc;
  if ((rc = t_end(c, XNB_T_GHOST_SCHEME, st))) return rc;
  if ((rc = xnb_chunk_neighbors(c, stream))) return rc;
  c->rebuilds++;
  return XNB_OK;
}

extern "C" {
int xnb_first_iteration(xnb_ctx* c, double eps, double sig, double rcut, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  int rc;
  if ((rc = move_and_update_full(c, stream))) return rc;
  return xnb_force_and_second_half(c, eps, sig, rcut, 0.0, stream);
}

} // extern "C"
// the first half kick with its displacement count going to `counter` (xnb_verlet_first_half: slot 0 of d_scalars64)
static int verlet_first_half(xnb_ctx* c, double dt, cudaStream_t st, unsigned long long* counter)
{
  c->amr_current = false;      // positions change: the sub-cell tables no longer bound them
  c->steps_since_nbh++;
  ParticlesP A = c->P(c->cur);
  int rc;
  if ((rc = t_begin(c, XNB_T_FIRST_HALF, st))) return rc;
  CK(cudaMemsetAsync(counter, 0, 8, st));
  if (c->n_inner) LAUNCH(k_verlet_first_half, nblk(c->n_inner, 256), 256, st, c->g, (int)c->n_inner, dt, dt * dt * 0.5, dt * 0.5, A.rx, A.ry, A.rz, A.vx, A.vy, A.vz,
                         A.fx, A.fy, A.fz, c->atom_cell[c->cur_ac].p, c->backup.p, c->max_displ * c->max_displ, counter);
  return t_end(c, XNB_T_FIRST_HALF, st);
}
// sweep + second half kick; fused: + the first half of the NEXT step in the sweep's epilogue (k_lj_sweep_cl MODE 2)
static int sweep_step(xnb_ctx* c, bool fused, const LJP& lj, double dth, cudaStream_t st, const unsigned long long* skip, int part)
{
  return fused ? launch_force<2, false>(c, false, lj, dth, nullptr, nullptr, nullptr, nullptr, nullptr, st, skip, part)
               : launch_force<1, false>(c, false, lj, dth, nullptr, nullptr, nullptr, nullptr, nullptr, st, skip, part);
}
extern "C" {

int xnb_run_steps(xnb_ctx* c, int nsteps, double dt, double eps, double sig, double rcut, void* stream, int* rebuilds_out)
{
  if (!c) return XNB_ERR_INVALID;
  if (!c->have_nbh) return c->fail(XNB_ERR_INVALID, "no neighbour list (run xnb_first_iteration)");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  int rebuilds = 0, rc;
  if (!c->ev_flag) CK(cudaEventCreateWithFlags(&c->ev_flag, cudaEventDisableTiming));
  volatile unsigned long long* h_flag = reinterpret_cast<volatile unsigned long long*>(static_cast<char*>(c->h_pinned) + 1024);
  const bool speculate = !env_flag("XNB_NO_SPECULATION");
  const LJP lj = make_lj(eps, sig, rcut);
  // Between two steps of this loop nothing observes the state, so the sweep of step k also performs the first half of step k + 1 in its
  // epilogue (MODE 2): a and the kicked v are in registers there.  The new positions go to the other position buffer (other blocks still
  // stage the old ones) and the two buffers trade places afterwards; the displacement count of step k + 1 goes to the counter slot step k
  // is not using.  The last step of the call is a plain MODE 1 sweep, so the state a caller sees is the usual one.
  const bool fuse_ok = speculate && !env_flag("XNB_NO_FUSED_FIRST_HALF") && !c->hout.active;
  auto can_fuse = [&](int it) { return fuse_ok && it + 1 < nsteps && c->cl.valid && !c->cl.ghost && !env_flag("XNB_SWEEP_STREAMS") && c->n_inner > 0; };
  auto arm_next_half = [&](unsigned long long* counter) -> int {
    ParticlesP B = c->P(1 - c->cur);
    CK(cudaMemsetAsync(counter, 0, 8, st));
    c->next_half = NextHalfP{dt, dt * dt * 0.5, c->max_displ * c->max_displ, B.rx, B.ry, B.rz, c->atom_cell[c->cur_ac].p, c->backup.p, counter};
    return 0;
  };
  bool pending = false;        // this step's first half was done by the previous sweep
  int par = 0;                 // which counter slot holds this step's displacement count
  for (int it = 0; it < nsteps; it++)
  {
    unsigned long long* const cnt = c->d_scalars64.p + (par ? 9 : 0);
    unsigned long long* const next_cnt = c->d_scalars64.p + (par ? 0 : 9);
    if (!pending) { if ((rc = verlet_first_half(c, dt, st, cnt))) return rc; }
    else { c->amr_current = false; c->steps_since_nbh++; }
    // trigger_move_particles (update-particles.msp:1-6): MPI_Allreduce(SUM, 1 x u64) of particle_displ_over.cu:174, then the
    // host reads the count.  The fast path (ghost_update_r + sweep) is enqueued BEFORE the host waits for that count, so the GPU
    // never idles on the host round trip; the sweep reads the same counter on the device and returns at once if a rebuild is due
    // (ghost_update_r before a rebuild is harmless: the rebuild recreates every ghost).
    if (c->nranks > 1)
    {
      if (!c->comm) return c->fail(XNB_ERR_NCCL, "no communicator");
      if (c->peer.enabled)
        LAUNCH(k_peer_allsum, 1, PEER_MAX_RANKS, st, c->nranks, c->rank, c->peer.d_slots.p, static_cast<const PeerHdr*>(c->peer.box), ++c->peer.epoch_over, cnt, c->d_scalars32.p, c->peer.timeout_ns);
      else NK(g_nccl.AllReduce(cnt, cnt, 1, nccl_uint64, nccl_sum, c->comm, st));
    }
    CK(cudaMemcpyAsync(const_cast<unsigned long long*>(h_flag), cnt, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(c->ev_flag, st));
    size_t force_scopes = c->tpool[XNB_T_FORCE].used;
    bool fused = false;        // the sweep enqueued for this step carries the next first half
    const bool overlap = speculate && c->nranks > 1 && c->cl.valid && !c->cl.ghost && c->cl.n_interior > 0 && c->cl.n_boundary > 0 && !env_flag("XNB_NO_OVERLAP") && !env_flag("XNB_SWEEP_STREAMS");
    if (overlap)
    {
      // halo exchange on its own stream while the interior tiles are swept; the boundary tiles wait for it.  One timing scope spans
      // both sweep launches.
      if (!c->st_comm)
      {
        // highest priority: the exchange kernels must get SMs while the interior sweep fills the GPU
        int prio_lo = 0, prio_hi = 0; CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CK(cudaStreamCreateWithPriority(&c->st_comm, cudaStreamNonBlocking, prio_hi));
        CK(cudaEventCreateWithFlags(&c->ev_pos, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c->ev_ghost, cudaEventDisableTiming));
      }
      fused = can_fuse(it);
      if (fused && (rc = arm_next_half(next_cnt))) return rc;
      CK(cudaEventRecord(c->ev_pos, st));
      CK(cudaStreamWaitEvent(c->st_comm, c->ev_pos, 0));
      if ((rc = t_begin(c, XNB_T_GHOST_UPDATE, c->st_comm))) return rc;
      if ((rc = xnb_ghost_update_r(c, c->st_comm))) return rc;
      if ((rc = t_end(c, XNB_T_GHOST_UPDATE, c->st_comm))) return rc;
      // the boundary tiles follow the halo on ITS stream: they start the moment the ghosts are in, alongside whatever is left of
      // the interior sweep (one tail instead of two), and the main stream joins at the end
      const bool boundary_on_comm = !env_flag("XNB_BOUNDARY_ON_MAIN");
      if (boundary_on_comm && (rc = sweep_step(c, fused, lj, dt * 0.5, c->st_comm, cnt, 2))) return rc;
      CK(cudaEventRecord(c->ev_ghost, c->st_comm));
      if ((rc = t_begin(c, XNB_T_FORCE, st))) return rc;
      if ((rc = sweep_step(c, fused, lj, dt * 0.5, st, cnt, 1))) return rc;
      CK(cudaStreamWaitEvent(st, c->ev_ghost, 0));
      if (!boundary_on_comm && (rc = sweep_step(c, fused, lj, dt * 0.5, st, cnt, 2))) return rc;
      if ((rc = t_end(c, XNB_T_FORCE, st))) return rc;
    }
    else if (speculate)
    {
      if ((rc = t_begin(c, XNB_T_GHOST_UPDATE, st))) return rc;
      if ((rc = xnb_ghost_update_r(c, stream))) return rc;
      if ((rc = t_end(c, XNB_T_GHOST_UPDATE, st))) return rc;
      fused = can_fuse(it);
      if (fused && (rc = arm_next_half(next_cnt))) return rc;
      if ((rc = sweep_step(c, fused, lj, dt * 0.5, st, cnt, 0))) return rc;
    }
    CK(cudaEventSynchronize(c->ev_flag));
    const unsigned long long over = *h_flag;
    if (over > 0)
    {
      if (speculate && c->timing && c->tpool[XNB_T_FORCE].used == force_scopes + 1) c->tpool[XNB_T_FORCE].used = force_scopes;   // the void launch is not a sweep
      // peer transport: a partner whose count never arrived also lands here (k_peer_allsum forces the rebuild path on a timeout) --
      // report it before the rebuild's NCCL exchanges would wait for that partner
      if (c->nranks > 1 && c->peer.enabled && (rc = check_device_errors(c, st))) return rc;
      if ((rc = move_and_update_full(c, stream))) return rc;
      rebuilds++;
      fused = can_fuse(it);                       // the void launch did nothing; the sweep after the rebuild may carry the next first half
      if (fused && (rc = arm_next_half(next_cnt))) return rc;
      if ((rc = sweep_step(c, fused, lj, dt * 0.5, st, nullptr, 0))) return rc;
    }
    else if ((rc = host_out_positions(c, c->ev_flag, false))) return rc;     // no rebuild: positions are final since the first half
    if (over == 0 && !speculate)
    {
      if ((rc = t_begin(c, XNB_T_GHOST_UPDATE, st))) return rc;
      if ((rc = xnb_ghost_update_r(c, stream))) return rc;
      if ((rc = t_end(c, XNB_T_GHOST_UPDATE, st))) return rc;
      if ((rc = xnb_force_and_second_half(c, eps, sig, rcut, dt * 0.5, stream))) return rc;
    }
    pending = fused;
    if (fused)
    {
      // the sweep wrote the next positions into the other buffer: the two trade places (pointers and capacities; the ghost slots of
      // the new current buffer are stale until the next ghost_update_r / rebuild, which always comes first)
      for (int f = 0; f < 3; f++) { std::swap(c->f64[c->cur][f].p, c->f64[1 - c->cur][f].p); std::swap(c->f64[c->cur][f].cap, c->f64[1 - c->cur][f].cap); }
      par ^= 1;
    }
  }
  if (rebuilds_out) *rebuilds_out = rebuilds;
  return XNB_OK;
}

} // extern "C"
static int step_host_impl(xnb_ctx* c, double dt, double eps, double sig, double rcut,
                          const double* const in_r[3], const double* const in_v[3],
                          double* const out_r[3], double* const out_v[3], double* const out_f[3], uint64_t* out_id, int id_always,
                          int64_t capacity, int64_t* n_out, void* stream, int* rebuilt_out);
extern "C" {
int xnb_step_host(xnb_ctx* c, double dt, double eps, double sig, double rcut,
                  const double* const in_r[3], const double* const in_v[3],
                  double* const out_r[3], double* const out_v[3], double* const out_f[3], uint64_t* out_id, int id_always,
                  void* stream, int* rebuilt_out)
{
  if (!c) return XNB_ERR_INVALID;
  if (c->nranks != 1) return c->fail(XNB_ERR_INVALID, "xnb_step_host: single sub-domain only (particles migrate between ranks otherwise: xnb_step_host_n)");
  return step_host_impl(c, dt, eps, sig, rcut, in_r, in_v, out_r, out_v, out_f, out_id, id_always, -1, nullptr, stream, rebuilt_out);
}
int xnb_step_host_n(xnb_ctx* c, double dt, double eps, double sig, double rcut,
                    const double* const in_r[3], const double* const in_v[3],
                    double* const out_r[3], double* const out_v[3], double* const out_f[3], uint64_t* out_id,
                    int64_t capacity, int64_t* n_out, void* stream, int* rebuilt_out)
{
  if (!c || capacity < 0 || !n_out) return XNB_ERR_INVALID;
  return step_host_impl(c, dt, eps, sig, rcut, in_r, in_v, out_r, out_v, out_f, out_id, 0, capacity, n_out, stream, rebuilt_out);
}
} // extern "C"
static int step_host_impl(xnb_ctx* c, double dt, double eps, double sig, double rcut,
                          const double* const in_r[3], const double* const in_v[3],
                          double* const out_r[3], double* const out_v[3], double* const out_f[3], uint64_t* out_id, int id_always,
                          int64_t capacity, int64_t* n_out, void* stream, int* rebuilt_out)
{
  if (!c->have_nbh) return c->fail(XNB_ERR_INVALID, "no neighbour list (run xnb_first_iteration)");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (!c->st_d2h)
  {
    CK(cudaStreamCreateWithFlags(&c->st_d2h, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_d2h_go, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c->ev_d2h_done, cudaEventDisableTiming));
  }
  c->amr_current = false;
  const size_t n = (size_t)c->n_inner;
  for (int f = 0; f < 3; f++)
  {
    if (in_r && in_r[f] && n) CK(cudaMemcpyAsync(c->f64[c->cur][f].p, in_r[f], n * 8, cudaMemcpyHostToDevice, st));
    if (in_v && in_v[f] && n) CK(cudaMemcpyAsync(c->f64[c->cur][3 + f].p, in_v[f], n * 8, cudaMemcpyHostToDevice, st));
  }
  xnb_ctx::HostOut& h = c->hout;
  h = xnb_ctx::HostOut();
  h.active = true; h.id_always = id_always != 0; h.id = out_id;
  h.capacity = capacity < 0 ? n : (size_t)capacity;
  for (int f = 0; f < 3; f++) h.r[f] = out_r ? out_r[f] : nullptr;
  int rebuilds = 0;
  int rc = xnb_run_steps(c, 1, dt, eps, sig, rcut, stream, &rebuilds);
  const bool issued = h.issued, overflow = h.overflow;
  h.active = false;
  if (rc) { cudaStreamSynchronize(c->st_d2h); return rc; }
  if (!issued) return c->fail(XNB_ERR_INVALID, "xnb_step_host: internal error (positions never became final)");
  const size_t n_now = (size_t)c->n_inner;
  if (n_out) *n_out = (int64_t)n_now;
  if (capacity < 0 && n_now != n) { cudaStreamSynchronize(c->st_d2h); return c->fail(XNB_ERR_INVALID, "xnb_step_host: the number of particles changed (a particle left a non-periodic domain)"); }
  if (overflow || n_now > h.capacity) { cudaStreamSynchronize(c->st_d2h); return c->fail(XNB_ERR_CAPACITY, "xnb_step_host_n: this rank now owns more particles than the caller's arrays hold (the device state is intact: enlarge and xnb_download_rvf)"); }
  // v and f are final after the sweep's epilogue
  for (int f = 0; f < 3; f++)
  {
    if (out_v && out_v[f] && n_now) CK(cudaMemcpyAsync(out_v[f], c->f64[c->cur][3 + f].p, n_now * 8, cudaMemcpyDeviceToHost, st));
    if (out_f && out_f[f] && n_now) CK(cudaMemcpyAsync(out_f[f], c->f64[c->cur][6 + f].p, n_now * 8, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaEventRecord(c->ev_d2h_done, c->st_d2h));
  CK(cudaStreamWaitEvent(st, c->ev_d2h_done, 0));
  CK(cudaStreamSynchronize(st));
  if (rebuilt_out) *rebuilt_out = rebuilds;
  return XNB_OK;
}

extern "C" {
int xnb_energy_virial(xnb_ctx* c, double eps, double sig, double rcut, double* epot, double virial[6], double* ekin, void* stream)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  if (!c->have_nbh) return c->fail(XNB_ERR_INVALID, "no neighbour list");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = c->n_inner;
  const unsigned nb2 = nblk(n, 256);
  unsigned nb = 0; double* evp = nullptr;
  // MODE 0 accumulates into f: run it on a zeroed scratch copy of f so the state is untouched
  DBuf<double>& scratch = c->ev_scratch; CK(scratch.ensure((size_t)c->n_total * 3 + 16, 0, 1.1));
  DBuf<double>& ekp = c->ev_ekin; CK(ekp.ensure((size_t)nb2 + 16, 0, 1.1));
  ParticlesP A = c->P(c->cur);
  CK(cudaMemsetAsync(scratch.p, 0, (size_t)c->n_total * 3 * 8, st));
  if (n)
  {
    int rc = launch_force<0, true>(c, false, make_lj(eps, sig, rcut), 0.0, scratch.p, scratch.p + c->n_total, scratch.p + 2 * c->n_total, &evp, &nb, st); if (rc) return rc;
    LAUNCH(k_ekin, nb2, 256, st, (int)n, A.vx, A.vy, A.vz, A.type, c->mass.p, ekp.p);
  }
  if (n == 0) { if (epot) *epot = 0; if (virial) for (int q = 0; q < 6; q++) virial[q] = 0; if (ekin) *ekin = 0; return XNB_OK; }
  std::vector<double> h((size_t)nb * 7 + nb2, 0.0);
  CK(cudaMemcpyAsync(h.data(), evp, (size_t)nb * 7 * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h.data() + (size_t)nb * 7, ekp.p, (size_t)nb2 * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  auto ksum = [&](size_t first, size_t count, size_t stride) { double s = 0., cc = 0.; for (size_t q = 0; q < count; q++) { const double x = h[first + q * stride]; const double t = s + x; cc += (std::fabs(s) >= std::fabs(x)) ? (s - t) + x : (x - t) + s; s = t; } return s + cc; };
  if (epot) *epot = ksum(0, nb, 7);
  if (virial) for (int q = 0; q < 6; q++) virial[q] = ksum((size_t)q + 1, nb, 7);
  if (ekin) *ekin = ksum((size_t)nb * 7, nb2, 1);
  return XNB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// views / downloads
// ---------------------------------------------------------------------------------------------------------------------
int xnb_view_chunk_neighbors(xnb_ctx* c, const uint16_t* const** d_cell_stream, const uint32_t** d_bytes, uint32_t* max_neighbors)
{
  if (!c || !c->have_nbh) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  { int rc = ensure_ghost_lists(c, nullptr); if (rc) return rc; }      // lists of the ghost cells are built on demand
  if (d_cell_stream) *d_cell_stream = (const uint16_t* const*)c->cell_stream.p;
  if (d_bytes) *d_bytes = c->cell_stream_bytes.p;
  if (max_neighbors) *max_neighbors = c->max_neighbors;
  return XNB_OK;
}

int64_t xnb_stream_pool_u16(const xnb_ctx* c) { return c ? c->pool_used : 0; }

int xnb_get_streams(xnb_ctx* c, uint32_t* size_u16, uint16_t* data)
{
  if (!c || !c->have_nbh) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  { int rc = ensure_ghost_lists(c, nullptr); if (rc) return rc; }      // lists of the ghost cells are built on demand
  CK(cudaDeviceSynchronize());
  const size_t nc = (size_t)c->g.n_cells;
  std::vector<uint32_t> sz(nc); std::vector<unsigned long long> off(nc);
  CK(cudaMemcpy(sz.data(), c->stream_size.p, nc * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(off.data(), c->stream_off.p, nc * 8, cudaMemcpyDeviceToHost));
  if (size_u16) memcpy(size_u16, sz.data(), nc * 4);
  if (data)
  {
    // the streams sit at stream_off[] in the pool (compact after the two-pass build, fixed slots after the tiled one)
    size_t extent = 0;
    for (size_t q = 0; q < nc; q++) if (sz[q]) extent = std::max(extent, (size_t)off[q] + sz[q]);
    std::vector<uint16_t> pool(extent);
    if (extent) CK(cudaMemcpy(pool.data(), c->pool.p, extent * 2, cudaMemcpyDeviceToHost));
    size_t o = 0;
    for (size_t q = 0; q < nc; q++) { if (sz[q]) memcpy(data + o, pool.data() + off[q], (size_t)sz[q] * 2); o += sz[q]; }
  }
  return XNB_OK;
}

int64_t xnb_get_amr(xnb_ctx* c, int64_t* sgs, uint32_t* sgc)
{
  if (!c) return -1;
  cudaDeviceSynchronize();
  const size_t nc = (size_t)c->g.n_cells;
  if (sgs)
  {
    std::vector<unsigned long long> h(nc + 1);
    if (cudaMemcpy(h.data(), c->sub_grid_start.p, (nc + 1) * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    for (size_t q = 0; q <= nc; q++) sgs[q] = (int64_t)h[q];
  }
  if (sgc && c->n_sub_grid_cells && c->max_side > 1)
    if (cudaMemcpy(sgc, c->sub_grid_cells.p, (size_t)c->n_sub_grid_cells * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return c->n_sub_grid_cells;
}

int xnb_get_backup(xnb_ctx* c, uint32_t* out)
{
  if (!c || !out) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  if (c->n_inner) CK(cudaMemcpy(out, c->backup.p, (size_t)c->n_inner * 12, cudaMemcpyDeviceToHost));
  return XNB_OK;
}

int xnb_view_particles(xnb_ctx* c, xnb_particle_view* out)
{
  if (!c || !out) return XNB_ERR_INVALID;
  int rc = ensure_grid(c); if (rc) return rc;
  ParticlesP A = c->P(c->cur);
  out->n_inner = c->n_inner; out->n_total = c->n_total; out->n_cells = c->g.n_cells;
  out->rx = A.rx; out->ry = A.ry; out->rz = A.rz; out->vx = A.vx; out->vy = A.vy; out->vz = A.vz; out->fx = A.fx; out->fy = A.fy; out->fz = A.fz;
  out->id = reinterpret_cast<uint64_t*>(A.id); out->type = A.type;
  out->cell_start = c->cell_start.p; out->cell_count = c->cell_count.p; out->particle_cell = c->atom_cell[c->cur_ac].p;
  return XNB_OK;
}

int64_t xnb_device_allocations(void) { return g_device_allocs; }
int64_t xnb_rebuild_count(const xnb_ctx* c) { return c ? c->rebuilds : 0; }
int64_t xnb_kernel_launches(const xnb_ctx* c) { return c ? c->launches : 0; }

int xnb_timing_enable(xnb_ctx* c, int on)
{
  if (!c) return XNB_ERR_INVALID;
  c->timing = on != 0;
  return XNB_OK;
}

int xnb_timing_read(xnb_ctx* c, double ms[XNB_T_COUNT], int64_t scopes[XNB_T_COUNT], int reset)
{
  if (!c) return XNB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  int rc = t_collect(c); if (rc) return rc;
  for (int cat = 0; cat < XNB_T_COUNT; cat++)
  {
    if (ms) ms[cat] = c->tpool[cat].ms;
    if (scopes) scopes[cat] = c->tpool[cat].n;
    if (reset) { c->tpool[cat].ms = 0; c->tpool[cat].n = 0; c->tpool[cat].dropped = 0; }
  }
  return XNB_OK;
}

int xnb_measure_dfma_peak(int device, double* tflops)
{
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return XNB_ERR_NO_DEVICE;
  if (cudaSetDevice(device) != cudaSuccess) return XNB_ERR_CUDA;
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, device);
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  double* out = nullptr;
  if (cudaMalloc(&out, (size_t)blocks * threads * 8) != cudaSuccess) return XNB_ERR_CUDA;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 5; rep++)
  {
    cudaEventRecord(e0);
    k_dfma_probe<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8.0 * iters * (double)blocks * threads;
    if (rep > 0 && ms > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  if (tflops) *tflops = best;
  return cudaGetLastError() == cudaSuccess ? XNB_OK : XNB_ERR_CUDA;
}

} // extern "C"
