/*
 * xnb_hotpath.h -- C-ABI of the B200-native exaNBody hot path (libxnb_hotpath.so).
 *
 * One xnb_ctx = one sub-domain on one GPU (the reference's "one MPI rank": a Grid + Domain + GridChunkNeighbors +
 * AmrGrid + PositionBackupData + GhostCommunicationScheme living in operator slots).  The ctx owns every device
 * buffer; callers borrow views.  All compute entry points enqueue hand-written sm_100a kernels on the caller's
 * cudaStream_t (passed as void*, NULL = default stream) and do not synchronise unless stated.  There is NO CPU
 * fallback: every entry point returns XNB_ERR_NO_DEVICE when no CUDA device is usable.
 *
 * Entry points that BLOCK THE HOST (they wait on the stream for small results that size the next launch; marked
 * "[host sync]" below): xnb_move_particles, xnb_rebuild_amr, xnb_ghost_comm_scheme, xnb_chunk_neighbors,
 * xnb_load_balance_rcb, xnb_read_displ_over / xnb_particle_displ_over, xnb_run_steps (one event wait per step, after the
 * fast path has been enqueued), xnb_step_host, xnb_first_iteration, xnb_energy_virial, and every xnb_get_* / xnb_download_*.
 * The others only enqueue.  Device buffers grow geometrically on demand (xnb_device_allocations() counts the calls).
 *
 * Return value: 0 = ok, otherwise an XNB_ERR_* code; xnb_last_error() gives the message
 * (the reference aborts through fatal_error(); the C++ shim in exanbody_b200/host maps non-zero to the same abort).
 *
 * Each entry point cites the reference operator / function it replaces (paths relative to the reference tree).
 */
#ifndef XNB_HOTPATH_H
#define XNB_HOTPATH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XNB_OK 0
#define XNB_ERR_NO_DEVICE 1
#define XNB_ERR_INVALID 2
#define XNB_ERR_CUDA 3
#define XNB_ERR_CAPACITY 4     /* u16 stream limits: > 65535 atoms in a cell, neighbour cell offset beyond +-15 */
#define XNB_ERR_LOST_PARTICLE 5 /* particle left a non-periodic domain (reference: otb_particles never re-homed) */
#define XNB_ERR_NCCL 6

typedef struct xnb_ctx xnb_ctx;

/* ---- life cycle -------------------------------------------------------------------------------------------- */
int  xnb_create(xnb_ctx** out, int cuda_device);        /* replaces onika init_cuda + slot construction          */
void xnb_destroy(xnb_ctx* ctx);
const char* xnb_last_error(const xnb_ctx* ctx);         /* ctx may be NULL: last error of a failed xnb_create     */
const char* xnb_version(void);

/* ---- configuration ----------------------------------------------------------------------------------------- */
/* op `domain` : src/core/include/exanb/core/domain.h:36-129 (bounds, cell_size, grid_dims, periodic; identity xform) */
int xnb_set_domain(xnb_ctx*, const double bounds_min[3], const double bounds_max[3], double cell_size,
                   const int64_t grid_dims[3], const int32_t periodic[3]);
/* op `init_rcb_grid` : src/grid_cell_particles/init_rcb_grid.cpp:37-84 + src/core/lib/simple_block_rcb.cpp:27-59.
   Chooses this rank's block of domain cells by recursive bisection; nranks == 1 => whole domain.                */
int xnb_init_rcb_grid(xnb_ctx*, int rank, int nranks);
/* op `nbh_dist` : src/particle_neighbors/nbh_dist.cpp:47-71.  nbh_dist = rcut_max + rcut_inc, max_displ = rcut_inc/2,
   ghost_dist = rcut_max + rcut_inc (identity xform).                                                            */
int xnb_set_nbh_dist(xnb_ctx*, double rcut_max, double rcut_inc);
/* per-type scalar `mass` used by divide_force_by_type_scalar: src/compute/vec3_typescalar_op.cu:71-122          */
int xnb_set_type_mass(xnb_ctx*, const double* mass_per_type, int n_types);
/* rebuild_amr sub_grid_density (data/config/update-particles.msp:9), default 6.5                                */
int xnb_set_sub_grid_density(xnb_ctx*, double density);
/* attach an NCCL communicator (ncclComm_t) for the multi-GPU halo; NULL = single GPU                            */
int xnb_set_nccl_comm(xnb_ctx*, void* nccl_comm);
/* ncclGetUniqueId / ncclCommInitRank through the NCCL already loaded in the process (128-byte id)               */
int xnb_nccl_unique_id(uint8_t id_out[128]);
int xnb_nccl_init_rank(xnb_ctx*, const uint8_t id[128], int rank, int nranks);

/* ---- particles in / out (HOST pointers; id/type may be NULL) ----------------------------------------------- */
/* replaces filling Grid cells (src/core/include/exanb/core/grid.h:57-688).  Particles may be in any order and
   anywhere in the domain: only those located in this rank's block are kept (as `lattice` does,
   generate_particle_lattice.h:289-321).  Synchronous.                                                           */
int xnb_set_particles(xnb_ctx*, int64_t n, const double* rx, const double* ry, const double* rz,
                      const double* vx, const double* vy, const double* vz, const uint64_t* id, const uint8_t* type);
int64_t xnb_num_inner(const xnb_ctx*);     /* particles of inner cells                                           */
int64_t xnb_num_total(const xnb_ctx*);     /* inner + ghost                                                      */
/* flat copy-out in device order [0,n_inner) inner particles sorted by cell, [n_inner,n_total) ghosts. Synchronous */
int xnb_get_particles(xnb_ctx*, int64_t first, int64_t n, double* rx, double* ry, double* rz,
                      double* vx, double* vy, double* vz, double* fx, double* fy, double* fz,
                      uint64_t* id, uint8_t* type, uint32_t* cell);
/* end-to-end helper: async H2D of r,v for the inner particles in current device order (pinned host memory)      */
int xnb_upload_rv(xnb_ctx*, const double* rx, const double* ry, const double* rz,
                  const double* vx, const double* vy, const double* vz, void* stream);
int xnb_download_rvf(xnb_ctx*, double* rx, double* ry, double* rz, double* vx, double* vy, double* vz,
                     double* fx, double* fy, double* fz, uint64_t* id, void* stream);

/* ---- grid description ------------------------------------------------------------------------------------- */
typedef struct xnb_grid_info
{
  int64_t dims[3];        /* local grid incl. ghost layers (Grid::dimension)      */
  int64_t offset[3];      /* Grid::offset : location of local cell (0,0,0) in the domain grid */
  int64_t ghost_layers;   /* Grid::ghost_layers, grid.h:110                        */
  int64_t n_cells;
  int64_t block_start[3]; /* inner block in domain cells                           */
  int64_t block_end[3];
} xnb_grid_info;
int xnb_get_grid_info(const xnb_ctx*, xnb_grid_info* out);
/* layout of the pair sweep's own copy of the neighbour lists ("compiled lists", derived from the GridChunkNeighbors
   streams after every xnb_chunk_neighbors; no reference counterpart: the reference sweeps the streams directly,
   src/compute/include/exanb/compute/compute_cell_particle_pairs_impl_default.h:143-179).  compiled = 0: no tile shape
   fits shared memory and the sweep reads the streams.                                                             */
typedef struct xnb_sweep_info
{
  int32_t compiled;       /* 1 if the sweep runs over compiled lists                                      */
  int32_t ghost;          /* lists compiled for ghost cells too (lennard_jones_force ghost=true)       */
  int64_t tile[3];        /* cells per tile (= per thread block)                                       */
  int64_t threads;        /* threads per block                                                         */
  int64_t blocks;
  int64_t smem_bytes;     /* staged positions of a tile's halo box                                     */
  int64_t rows;           /* 256-byte rows of the compiled lists (32 lanes x 4 candidates)             */
  int64_t candidates;     /* list entries of all swept particles (sum of neighbour counts)             */
  int64_t interior_tiles; /* tiles whose halo box holds no ghost cell: swept while the halo exchange   */
  int64_t boundary_tiles; /*   of xnb_run_steps is in flight (several ranks); the others wait for it   */
} xnb_sweep_info;
int xnb_get_sweep_info(const xnb_ctx*, xnb_sweep_info* out);
/* per-cell particle count and start index into the flat arrays (the per-cell SoA views of CellParticles)        */
int xnb_get_cells(xnb_ctx*, uint32_t* cell_start /* n_cells */, uint32_t* cell_count /* n_cells */);
/* DEVICE view of the grid: what the reference hands its operators as cells[c][field] (src/core/include/exanb/core/grid.h:70-71,
   678: per-cell SoA pointers).  Field f of particle p of cell c is f[cell_start[c] + p], p < cell_count[c]; inner particles
   are [0, n_inner) sorted by cell, ghosts follow.  All pointers are device pointers owned by the ctx; they stay valid until
   the next xnb_move_particles / xnb_rebuild_amr / xnb_ghost_comm_scheme (binning gathers into the other buffer, ghost creation
   may grow the arrays) and the next xnb_run_steps of more than one step (its sweeps write the next positions into the other position
   buffer and the two trade places): take the view again after those.  Does not synchronise: work enqueued on the caller's stream is
   ordered with the kernels that produce these arrays.                                                                */
typedef struct xnb_particle_view
{
  int64_t n_inner, n_total, n_cells;
  double *rx, *ry, *rz, *vx, *vy, *vz, *fx, *fy, *fz;
  uint64_t* id;
  uint8_t* type;
  const uint32_t* cell_start;      /* n_cells */
  const uint32_t* cell_count;      /* n_cells */
  const uint32_t* particle_cell;   /* n_total: local cell index of every particle */
} xnb_particle_view;
int xnb_view_particles(xnb_ctx*, xnb_particle_view* out);

/* ---- operators of the hot path (asynchronous on `stream`) --------------------------------------------------- */
/* op `move_particles` : src/grid_cell_particles/include/exanb/grid_cell_particles/move_particles_across_cells.h:78-235
   periodic wrap + cell location + re-binning of all inner particles (K1).  Multi-GPU: also performs the
   `migrate_cell_particles` hand-off of particles that left the block (src/mpi/migrate_cell_particles.cpp:101-143). */
int xnb_move_particles(xnb_ctx*, void* stream);   /* [host sync] */
/* op `rebuild_amr` : src/amr/rebuild_amr.cpp:35-62, amr_grid_algorithm.h:66-78,112-434 (in-cell sub-grid sort + tables) */
int xnb_rebuild_amr(xnb_ctx*, void* stream);   /* [host sync] */
/* op `backup_r` : src/io/backup_r.cpp:36-78, src/core/include/exanb/core/backup_r.h:31-51                       */
int xnb_backup_r(xnb_ctx*, void* stream);
/* op `ghost_comm_scheme` : src/mpi/update_ghosts_comm_scheme.cpp:84-489 (who sends which particles of which cell) */
int xnb_ghost_comm_scheme(xnb_ctx*, void* stream);   /* [host sync] */
/* ops `ghost_update_all` / `ghost_update_r` : src/mpi/update_ghosts.cu:45-64, include/exanb/mpi/grid_update_ghosts.h:63-202 */
int xnb_ghost_update_all(xnb_ctx*, void* stream);
int xnb_ghost_update_r(xnb_ctx*, void* stream);
/* how the halo travels between ranks (the role of the reference's MPI_Isend / MPI_Irecv per partner,
   update_ghosts_comm_manager.h:390,436): 1 = NVLink peer-memory mailboxes -- the pack kernel stores straight into the partner's
   HBM and raises a flag, the unpack kernel waits on it, and the per-step displacement all-reduce (particle_displ_over.cu:174) is
   a peer-store kernel too, so the step path has no host-side communication call; 0 = one ncclSend / ncclRecv per partner (ranks on
   different nodes, IPC not permitted, or XNB_GHOST_NCCL=1).  Decided collectively at the first xnb_ghost_comm_scheme.          */
int xnb_ghost_transport(const xnb_ctx*);
/* ops `amr_grid_pairs` + `chunk_neighbors` : src/particle_neighbors/chunk_neighbors.cpp:48-74,
   include/exanb/particle_neighbors/chunk_neighbors_execute.h:40-423 (K2).  Config = build_particle_offset, chunk 1. */
int xnb_chunk_neighbors(xnb_ctx*, void* stream);   /* [host sync] (an overflowing build is re-run with more room) */
/* op `zero_particle_force` : src/compute/zero_particle_force.cu:15-53                                           */
int xnb_zero_particle_force(xnb_ctx*, int ghost, void* stream);
/* The functor the pair sweeps are instantiated with (the functor concept: compute_pair_traits.h:24-74, restated in
   exanbody_b200/csrc/xnb_pair_functor.cuh).  XNB_FUNCTOR_LJ (default): LennardJonesForceFunctor restated with one
   reciprocal and a four-candidate hook.  XNB_FUNCTOR_LJ_REFERENCE_FORM: lj_compute_energy + the buffer-less operator()
   exactly as written in lennard_jones.cu:46-56,106-124 (sqrt, two divisions), driven through the generic buffer-less
   call, one candidate at a time -- the route any further functor of the concept takes.  Same pair sets; forces agree
   to rounding (<= 3e-13 relative).                                                                                */
#define XNB_FUNCTOR_LJ 0
#define XNB_FUNCTOR_LJ_REFERENCE_FORM 1
int xnb_set_pair_functor(xnb_ctx*, int functor);
/* op `lennard_jones_force` : contribs/md/lennard_jones/lennard_jones.cu:171-215 through
   compute_cell_particle_pairs (src/compute/include/exanb/compute/compute_cell_particle_pairs.h:122-189).
   ACCUMULATES into fx,fy,fz like the reference (call xnb_zero_particle_force first).                            */
int xnb_lennard_jones_force(xnb_ctx*, double epsilon, double sigma, double rcut, int ghost, void* stream);
/* op `gravitational_force` : contribs/pi/gravitational_force.cu:161-217 (functor :48-132, traits :144-150) -- a second
   functor of the concept, one that reads a PER-NEIGHBOUR field (the neighbour's type, for its mass; the reference passes
   SimpleNbhComputeBuffer<FieldSet<type>>, :175).  It runs through the general pair sweep (csrc/xnb_pair_generic.cuh:
   compute_cell_particle_pairs_impl_default.h:87-239 for any functor of the concept, reference-format streams, cells[c][field][p]
   views).  buffer_form = 0: the buffer-less call `func(dr, d2, type_a, fx, fy, fz, cells, cell_b, p_b, weight)`
   (impl_default.h:199-204); buffer_form = 1: the ComputePairBuffer2 call `func(n, buf, type_a, fx, fy, fz, cells)`
   (impl_default.h:213-222, compute_pair_buffer.h:150-243; capacity 512 neighbours inside the cut, more is an error).
   Masses: xnb_set_type_mass.  ACCUMULATES into fx,fy,fz.  ghost must be 0.                                         */
int xnb_gravitational_force(xnb_ctx*, double G, double rcut, int ghost, int buffer_form, void* stream);
/* op `average_neighbors_scalar` : src/compute/average_neighbors.cu:102-215 (functor :60-100, traits :104-113) -- a third kind of
   functor of the concept: a per-neighbour SCALAR FIELD and a PARTICLE CONTEXT (HasParticleContextStart / HasParticleContext /
   HasParticleContextStop: func(ctx, cells, cell_a, p_a, Start{}), func(ctx, dr, d2, cells, cell_b, p_b, w), func(..., Stop{});
   impl_default.h:152,199-204,224).  avg_field[a] = sum_b w(d) nbh_field[b] / sum_b w(d) over the listed neighbours with
   0 < d <= rcut, w(d) = a0 + a1 d + a2 d^2 + a3 d^3 (weight_function, NULL = {1,0,0,0}); nbh_field = XNB_FIELD_*.  The result is
   a ctx-owned generic real field, one double per inner particle in the current particle order.                         */
#define XNB_FIELD_RX 0
#define XNB_FIELD_RY 1
#define XNB_FIELD_RZ 2
#define XNB_FIELD_VX 3
#define XNB_FIELD_VY 4
#define XNB_FIELD_VZ 5
#define XNB_FIELD_FX 6
#define XNB_FIELD_FY 7
#define XNB_FIELD_FZ 8
#define XNB_FIELD_ID 9
#define XNB_FIELD_TYPE 10
int xnb_average_neighbors(xnb_ctx*, double rcut, const double weight_function[4], int nbh_field, void* stream);
int xnb_get_generic_field(xnb_ctx*, double* out /* n_inner */);   /* [host sync] */
/* ---- Newton-3 path ---------------------------------------------------------------------- */
/* ChunkNeighborsConfig::half_symmetric / skip_ghosts of the chunk_neighbors operator (chunk_neighbors_config.h:35-36),
   i.e. NeighborFilterHalfSymGhost (neighbor_filter_func.h:36-52): half_symmetric keeps b only if cell_b < cell_a or
   (same cell and p_b < p_a); skip_ghosts drops neighbours that live in ghost cells.  Takes effect at the next
   xnb_chunk_neighbors; the lists stay in the GridChunkNeighbors format (bit-identical to the reference's filter).  */
int xnb_set_chunk_neighbors_config(xnb_ctx*, int half_symmetric, int skip_ghosts);
/* compute_cell_particle_pairs<Symmetric> with ComputePairOptionalLocks<true> (impl_default.h:143-239,
   compute_pair_optional_args.h:152-161) and the LJ functor applied once per pair: f_a += de*dr, f_b -= de*dr
   (accumulates: zero the forces, ghosts included, first).  Needs half_symmetric lists.                            */
int xnb_lennard_jones_force_symmetric(xnb_ctx*, double epsilon, double sigma, double rcut, void* stream);
/* replaces op `update_force_from_ghost` (mpi/update_force_from_ghost.cu:44, update_from_ghosts.h:152-,
   update_from_ghost_functors.h:36-120, UpdateValueAdd): the force of every ghost is added to the particle it is an
   image of; between ranks the ghost exchange runs backwards (NCCL send/recv of fx,fy,fz).                         */
int xnb_update_force_from_ghost(xnb_ctx*, void* stream);

/* op `divide_force_by_type_scalar: mass` : src/compute/vec3_typescalar_op.cu:118-122                            */
int xnb_divide_force_by_mass(xnb_ctx*, void* stream);
/* ops `push_f_v_r` / `push_f_v` : src/defbox/include/exanb/defbox/push_vec3_2nd_order.h:29-118, push_vec3_1st_order.h:29-107 */
int xnb_push_f_v_r(xnb_ctx*, double dt, double dt_scale, void* stream);
int xnb_push_f_v(xnb_ctx*, double dt, double dt_scale, void* stream);
/* op `particle_displ_over` : src/mpi/particle_displ_over.cu:39-178.  Synchronises; *count_out = global count
   (all ranks) of particles displaced >= max_displ since backup_r.                                               */
int xnb_particle_displ_over(xnb_ctx*, uint64_t* count_out, void* stream);

/* ---- fused forms (same results, fewer HBM passes) ---------------------------------------------------------- */
/* verlet_first_half + trigger_move_particles (numerical-scheme.msp:13-15, update-particles.msp:1-6) in one kernel (K4);
   the displacement count stays on the device until xnb_read_displ_over().                                       */
int xnb_verlet_first_half(xnb_ctx*, double dt, void* stream);
int xnb_read_displ_over(xnb_ctx*, uint64_t* count_out, void* stream);   /* [host sync] */
/* compute_all_forces_energy of the LJ deck (input_lj_Ni.msp:88-92) + verlet_second_half in one kernel (K3):
   f = LJ(r) (written, not accumulated), a = f/m[type], v += a*dt/2.  dt_half_kick = 0 skips the kick.           */
int xnb_force_and_second_half(xnb_ctx*, double epsilon, double sigma, double rcut, double dt_half_kick, void* stream);
/* whole `numerical_scheme` loop (numerical-scheme.msp:21-25 + check_and_update_particles): nsteps iterations.
   Synchronises once per step to read the rebuild trigger (the reference does an MPI_Allreduce there).           */
int xnb_run_steps(xnb_ctx*, int nsteps, double dt, double epsilon, double sigma, double rcut, void* stream, int* rebuilds_out);   /* [host sync] once per step */
/* one iteration of the same loop for a caller whose particles live in HOST memory, as the reference's Grid does
   (core/grid.h:57-688, host / managed allocations): uploads r,v of the inner particles (current device order; NULL
   = keep the device copy), runs one step exactly as xnb_run_steps(1) does, and returns r,v,f (NULL = not wanted).
   Positions go back on a second stream as soon as they are final (after the first half, or after binning when
   the step rebuilds), overlapping the pair sweep; v and f follow the sweep.  out_id is written when the step
   rebuilt (the particle order changed: *rebuilt_out = 1) or when id_always != 0.  Pinned host buffers make the
   copies asynchronous.  Synchronises `stream` before returning.  Single sub-domain only (nranks == 1).          */
int xnb_step_host(xnb_ctx*, double dt, double epsilon, double sigma, double rcut,
                  const double* const in_r[3], const double* const in_v[3],
                  double* const out_r[3], double* const out_v[3], double* const out_f[3], uint64_t* out_id, int id_always,
                  void* stream, int* rebuilt_out);
/* the same with SEVERAL sub-domains (collective, like xnb_run_steps): when the step rebuilds, particles migrate between the ranks
   (mpi/migrate_cell_particles.cpp:101-143) and this rank then owns *n_out particles, possibly more or fewer than it uploaded.  The out
   arrays hold `capacity` elements each; r, v, f and ids of the rank's particles after the step are written (ids whenever the step
   rebuilt); the next call uploads *n_out elements.  XNB_ERR_CAPACITY if *n_out > capacity: the device state is intact, enlarge the
   arrays and fetch it with xnb_download_rvf.                                                                                   */
int xnb_step_host_n(xnb_ctx*, double dt, double epsilon, double sigma, double rcut,
                    const double* const in_r[3], const double* const in_v[3],
                    double* const out_r[3], double* const out_v[3], double* const out_f[3], uint64_t* out_id,
                    int64_t capacity, int64_t* n_out, void* stream, int* rebuilt_out);   /* [host sync] */
/* init_particles + first force (update-particles.msp:55-60, compute-loop.msp:1-7)                               */
int xnb_first_iteration(xnb_ctx*, double epsilon, double sigma, double rcut, void* stream);

/* ---- oracle-defined observables (unpinned by the reference) ---------------------------------- */
/* E = 1/2 sum_ij e_ij (un-shifted lj_compute_energy, lennard_jones.cu:46-56), W = -1/2 sum dr (x) f, inner atoms;
   ekin = sum 1/2 m v^2.  Synchronous, local to this rank.                                                       */
int xnb_energy_virial(xnb_ctx*, double epsilon, double sigma, double rcut, double* epot, double virial[6], double* ekin, void* stream);   /* [host sync] */

/* ---- views / downloads of derived data -------------------------------------------------------------------- */
/* GridChunkNeighbors (src/particle_neighbors/include/exanb/particle_neighbors/chunk_neighbors.h:40-120):
   device array of per-cell stream pointers (= GridChunkNeighborsData) and per-cell sizes in BYTES.              */
int xnb_view_chunk_neighbors(xnb_ctx*, const uint16_t* const** d_cell_stream, const uint32_t** d_cell_stream_size,
                             uint32_t* max_neighbors);
int64_t xnb_stream_pool_u16(const xnb_ctx*);    /* total u16 words in use (incl. 16-byte alignment padding)      */
/* host copy: per-cell size in u16 (without padding) and the streams concatenated in cell order. Synchronous.    */
int xnb_get_streams(xnb_ctx*, uint32_t* size_u16 /* n_cells */, uint16_t* data /* sum(size) or NULL */);
/* AmrGrid tables (src/amr/include/exanb/amr/amr_grid.h:31-61) */
int64_t xnb_get_amr(xnb_ctx*, int64_t* sub_grid_start /* n_cells+1 or NULL */, uint32_t* sub_grid_cells /* or NULL */);
/* PositionBackupData: 3 x u32 per inner particle in flat order */
int xnb_get_backup(xnb_ctx*, uint32_t* out /* 3*n_inner */);
/* counters */
int64_t xnb_rebuild_count(const xnb_ctx*);
int64_t xnb_kernel_launches(const xnb_ctx*);    /* number of kernels this ctx has launched so far                */
int64_t xnb_device_allocations(void);           /* cudaMalloc calls of the library so far (process-wide): constant over steady-state steps */
/* device time per kernel group since the last reset: CUDA event pairs recorded on the launching stream around each
   group, never synchronising inside the timed region; xnb_timing_read waits for the recorded events and sums them. */
#define XNB_T_FORCE 0        /* pair sweep (K3)                                  */
#define XNB_T_NBH 1          /* chunk_neighbors (K2)                             */
#define XNB_T_FIRST_HALF 2   /* verlet_first_half + displacement test (K4)       */
#define XNB_T_BIN 3          /* move_particles + rebuild_amr + backup_r (K1)     */
#define XNB_T_GHOST_SCHEME 4 /* ghost_comm_scheme + ghost_update_all             */
#define XNB_T_GHOST_UPDATE 5 /* ghost_update_r (K6/K7 + NCCL)                    */
#define XNB_T_COUNT 6
int xnb_timing_enable(xnb_ctx*, int on);
int xnb_timing_read(xnb_ctx*, double ms[XNB_T_COUNT], int64_t scopes[XNB_T_COUNT], int reset);

/* ---- host-side input operators of the LJ decks (not on the timed path) ------------------------------------ */
/* ops `lattice` (structure FCC) + `gaussian_noise_r` with deterministic_noise:
   src/grid_cell_particles/include/exanb/grid_cell_particles/generate_particle_lattice.h:247-388,
   src/compute/include/exanb/compute/gaussian_noise.h:61-80,150-161.  vel_sigma / spheres are the synthetic-config
   extensions of SURVEY.md 8d.  Returns the particle count (negative = capacity too small), domain-cell order, ids from 1. */
typedef struct xnb_lattice_cfg
{
  double bounds_min[3], bounds_max[3], cell_size;
  int64_t grid_dims[3];
  double lattice_a, noise_sigma, vel_sigma;
  int32_t n_spheres; double sphere_rmin, sphere_rmax, drift_speed;
} xnb_lattice_cfg;
int64_t xnb_host_lattice_fcc(const xnb_lattice_cfg* cfg, int64_t capacity, double* rx, double* ry, double* rz,
                             double* vx, double* vy, double* vz, uint64_t* id, uint8_t* type);

/* ---- static decomposition, host only (no CUDA): what every rank derives for itself and its partners -------------- */
/* src/core/lib/simple_block_rcb.cpp:27-59 (via init_rcb_grid.cpp:65-77): block [start,end) of `rank` among `nranks` */
int xnb_host_rcb_block(const int64_t grid_dims[3], int nranks, int rank, int64_t start[3], int64_t end[3]);
/* op `amr_grid_pairs` : max_distance_sub_cell_pairs (src/amr/lib/amr_grid_algorithm.cpp:102-218) -> AmrSubCellPairCache
   (src/amr/include/exanb/amr/amr_grid_algorithm.h:439-453), host only.  For every resolution pair (res_b outer, res_a <= res_b
   inner) and every neighbour-cell offset (k, j, i in [0, layers], layers = ceil(max_dist / cell_size)): the (sub-cell a, sub-cell b)
   pairs, coded (k << 10) | (j << 5) | i each, whose boxes are at most max_dist apart.  list_offsets: n_lists + 1 entries
   (n_lists = max_res (max_res + 1) / 2 * (layers + 1)^3); pairs: a, b interleaved.  Returns the number of u16 words (call with
   NULLs to size the buffers), -1 on bad arguments.  The product's own builds prune with measured bounding boxes instead
   (csrc/xnb_nbh_big.cuh); the cache is produced for operators that consume it.                                        */
int64_t xnb_host_amr_sub_cell_pairs(int max_res, double cell_size, double max_dist, uint64_t* list_offsets, uint16_t* pairs);
/* op `simple_cost_model` (src/mpi/include/exanb/mpi/simple_cost_model.h:67-146), arithmetic only: cost of a cell from its
   particle count, p = N / cell_size^3, cost = coefs[0] p^3 + coefs[1] p^2 + coefs[2] p + coefs[3] (the reference's default
   coefficients are {0, 0, 1, 0}).  Feed it the inner-cell counts of xnb_get_cells; ghost cells carry no cost (:103).  */
int xnb_host_simple_cost_model(int64_t n_cells, const uint32_t* cell_count, double cell_size, const double coefs[4],
                               double* cell_costs);
/* host half of op `load_balance_rcb` (src/mpi/load_balance_rcb.cpp:228-452,510-545, the path without Zoltan; SURVEY.md 8f
   rank 1): cost-weighted recursive bisection of the domain cell grid.  cell_costs = the all-reduced cost of every domain
   cell, index (k*dj + j)*di + i (CellCosts, e.g. simple_cost_model).  Every rank calls it with the same costs and obtains
   its own block; *block_cost (may be NULL) = the cost inside it, from which lb_inbalance = (max - avg)/avg follows.
   xnb_load_balance_rcb (below) applies the result to a live xnb_ctx.                                               */
int xnb_host_load_balance_rcb(const int64_t grid_dims[3], const double* cell_costs, int nranks, int rank,
                              int64_t start[3], int64_t end[3], double* block_cost);
/* op `load_balance_rcb` on a live context, COLLECTIVE (src/mpi/load_balance_rcb.cpp:51-601 without Zoltan + simple_cost_model.h:67-146
   + migrate_cell_particles.cpp:101-143): per-cell costs on the device from the current cell counts
   (coefs as xnb_host_simple_cost_model, NULL = the reference's defaults {0,0,1,0}), summed over the ranks (ncclAllReduce),
   cost-weighted recursive bisection (every rank derives the same table), then the new block replaces the old one and the
   migration hand-off of xnb_move_particles carries every particle to its new owner (any rank).  lb_inbalance = (max - avg)/avg
   of the block costs before / after (either may be NULL).  SYNCHRONISES the host.  Continue with the rebuild chain
   (xnb_rebuild_amr ... xnb_chunk_neighbors); forces and velocities travel with the particles.                        */
int xnb_load_balance_rcb(xnb_ctx*, const double coefs[4], double* lb_inbalance_before, double* lb_inbalance_after, void* stream);
/* GridBlock [start, end) of `rank` in domain cells (this rank's own table: identical on every rank) */
int xnb_get_block(const xnb_ctx*, int rank, int64_t start[3], int64_t end[3]);
/* src/mpi/update_ghosts_comm_scheme.cpp:168-196,429-443: the cells rank `from` sends to rank `to` (sender local cell,
   receiver local ghost cell, GhostBoundaryModifier flags, ghosts_comm_scheme.h:46-81), in the reference's order.
   Returns the item count (arrays are filled when capacity suffices); -1 on invalid arguments.                     */
int64_t xnb_host_ghost_items(const int64_t grid_dims[3], const int32_t periodic[3], int ghost_layers, int nranks, int from, int to,
                             int64_t capacity, uint32_t* src_cell, uint32_t* dst_cell, uint32_t* flags);

/* stand-alone FP64 FMA peak probe (DFMA/s) used for the FP64 roofline denominator                               */
int xnb_measure_dfma_peak(int cuda_device, double* tflops_out);

#ifdef __cplusplus
}
#endif
#endif
