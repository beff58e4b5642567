/*
 * xnb_oracle.h -- C interface of the CPU ORACLE for the exaNBody LJ hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  It is a from-scratch CPU restatement
 * (C++17 + OpenMP, compiled with -ffp-contract=off) of the reference algorithms on the
 * hot path (binning -> AMR sub-grid sort -> ghosts -> chunk neighbour streams -> LJ pair
 * sweep -> velocity-Verlet + displacement trigger).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (exanbody_b200/, include/xnb_hotpath.h) never links or calls it.
 *
 * Parity pinning: the composition is pinned against the reference's own golden file
 * contribs/microStamp/samples/benchmark_lj_snap/check_values_lj_Ni.dat (1e-5 on r, v, a
 * after the deck's 100 steps) -- see tests/test_oracle_kat.py.  Bit-level pair-set order of
 * operations (norm2 contraction), potential energy and virial are NOT pinned by any
 * reference test ("parity unpinned" for those) and are defined here.
 *
 * Every function in xnb_oracle.cpp cites the reference file:line it follows
 * (paths relative to the reference tree root).
 */
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct xo_config
{
  /* domain (core/domain.h:36-129) */
  double bounds_min[3];
  double bounds_max[3];
  double cell_size;
  int64_t grid_dims[3];
  int32_t periodic[3];
  /* FCC lattice + noise (lattice_generator.h:162-170, generate_particle_lattice.h:247-321, gaussian_noise.h) */
  double lattice_a;          /* cubic lattice constant                                  */
  double noise_sigma;        /* gaussian_noise_r sigma (0 = none)                       */
  double vel_sigma;          /* synthetic configs only: gaussian velocities (0 = none)  */
  int32_t n_spheres;         /* C5: keep atoms only inside n_spheres random spheres (0 = all) */
  double sphere_rmin, sphere_rmax, drift_speed;
  /* LJ (lennard_jones.cu:40-56) */
  double epsilon, sigma, rcut;
  /* global */
  double rcut_inc;           /* Verlet skin (nbh_dist.cpp:47)                           */
  double dt;
  double mass;               /* per-type scalar 'mass' (vec3_typescalar_op.cu)           */
  double sub_grid_density;   /* rebuild_amr (update-particles.msp:9) default 6.5         */
  int32_t max_neighbors;     /* compute buffer capacity; reference default 256           */
  int32_t serial_order;      /* 1: binning reproduces the reference's single-thread in-cell order */
} xo_config;

typedef struct xo_sim xo_sim;

xo_sim* xo_create(const xo_config* cfg);
void    xo_destroy(xo_sim*);
const char* xo_last_error(void);

/* input_data + init_particles + first force (main-config.msp:33-49, compute-loop.msp:1-7) */
int xo_init(xo_sim*);              /* = xo_generate + xo_first_iteration */
int xo_generate(xo_sim*);          /* lattice + gaussian_noise_r (+ synthetic velocities); grid without ghost layers */
int xo_first_iteration(xo_sim*);   /* move_particles + parallel_update_particles + first force */
/* n iterations of numerical_scheme (numerical-scheme.msp:21-25); returns #rebuilds done */
int xo_run(xo_sim*, int nsteps);

/* single operators, for operator-level parity tests */
int xo_move_particles(xo_sim*);           /* move_particles_across_cells.h:78-235 */
int xo_update_particles_full(xo_sim*);    /* update-particles.msp:62-68 (without move_particles) */
int xo_ghost_update_r(xo_sim*);           /* update_ghosts.cu:46 */
int xo_build_neighbors(xo_sim*);          /* amr_grid_pairs + chunk_neighbors */
int xo_compute_force(xo_sim*);            /* zero_particle_force{ghost} + lennard_jones_force + divide by mass */
/* ChunkNeighborsConfig::half_symmetric / skip_ghosts (chunk_neighbors_config.h:35-36, neighbor_filter_func.h:36-52); takes effect at the next build */
void xo_set_nbh_config(xo_sim*, int half_symmetric, int skip_ghosts);
/* Newton-3 sweep over half_symmetric lists + update_force_from_ghost + divide by mass */
int xo_compute_force_symmetric(xo_sim*);
/* zero_particle_force{ghost:true}; gravitational_force (contribs/pi/gravitational_force.cu): ADDS G ma mb / r^2 pair forces, masses by type */
int xo_zero_force(xo_sim*);
int xo_gravitational_force(xo_sim*, double G, double rcut, const double* type_mass, int n_types);
/* average_neighbors_scalar (src/compute/average_neighbors.cu): out[all particles, cell order]; field 0..8 = r v f, 9 = id, 10 = type */
int xo_average_neighbors(xo_sim*, double rcut, const double weight_function[4], int nbh_field, double* out);
int xo_push_f_v_r(xo_sim*);               /* push_vec3_2nd_order.h */
int xo_push_f_v(xo_sim*, double dt_scale);/* push_vec3_1st_order.h */
int64_t xo_displ_over(xo_sim*);           /* particle_displ_over.cu: count of atoms over threshold */

/* grid description (local grid including ghost layers) */
void    xo_grid_info(const xo_sim*, int64_t dims[3], int64_t offset[3], int64_t* ghost_layers, int64_t* n_cells);
int64_t xo_total_particles(const xo_sim*);     /* inner + ghost */
int64_t xo_inner_particles(const xo_sim*);
void    xo_cell_counts(const xo_sim*, int32_t* counts /* n_cells */);
/* flat copy-out in cell order (all cells incl. ghosts); any pointer may be NULL */
void    xo_get_particles(const xo_sim*, double* rx, double* ry, double* rz,
                         double* vx, double* vy, double* vz,
                         double* fx, double* fy, double* fz, uint64_t* id, uint8_t* type);
/* overwrite the whole grid content (cell counts + flat arrays in cell order), e.g. with the GPU's in-cell order */
int     xo_set_particles(xo_sim*, const int32_t* counts, const double* rx, const double* ry, const double* rz,
                         const double* vx, const double* vy, const double* vz,
                         const double* fx, const double* fy, const double* fz, const uint64_t* id, const uint8_t* type);
/* AMR tables (amr_grid.h:31-61) */
int64_t xo_amr_tables(const xo_sim*, int64_t* sub_grid_start /* n_cells+1 or NULL */, uint32_t* sub_grid_cells /* or NULL */);
/* AmrSubCellPairCache of the last amr_grid_pairs (amr_grid_algorithm.cpp:102-218): lists in (res_b, res_a, offset k, j, i) order; returns the
   number of u16 words, *max_res = the cache's resolution; list_offsets has n_lists + 1 entries; any pointer may be NULL */
int64_t xo_amr_pair_cache(const xo_sim*, int64_t* max_res, uint64_t* list_offsets, uint16_t* pairs);
/* backup_r (3 x u32 per inner atom, cell order over ALL cells, ghost cells contribute 0 entries) */
int64_t xo_get_backup(const xo_sim*, uint32_t* out /* or NULL */);

/* neighbour streams: GridChunkNeighbors layout (chunk_neighbors.h:42-120) */
int64_t xo_stream_total_u16(const xo_sim*);
void    xo_stream_sizes(const xo_sim*, uint32_t* size_u16 /* n_cells */);
void    xo_stream_data(const xo_sim*, uint16_t* out /* concatenated in cell order */);
int64_t xo_max_neighbors(const xo_sim*);
/* neighbour pair set of inner-cell atoms as (id_a,id_b) sorted; returns #pairs, writes if out != NULL */
int64_t xo_pairs(const xo_sim*, uint64_t* out /* 2 per pair */);
/* restated verify_chunk_neighbors / chunk_neighbors_stream_check invariants; 0 = ok */
int     xo_check_streams(const xo_sim*);

/* oracle-defined extras (unpinned by the reference): E = 1/2 sum e_ij, W = -1/2 sum dr (x) f_ij, inner atoms, rcut of cfg */
void    xo_energy_virial(const xo_sim*, double* epot, double virial[6] /* xx yy zz xy xz yz */, double* ekin);

int64_t xo_rebuild_count(const xo_sim*);
int     xo_num_threads(void);
void    xo_set_num_threads(int n);   /* omp_set_num_threads (torchrun exports OMP_NUM_THREADS=1 to its workers) */

#ifdef __cplusplus
}
#endif
