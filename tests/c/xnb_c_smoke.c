/* Plain C99 caller of the C-ABI: dlopen()s libxnb_hotpath.so, resolves the entry points by name and runs an LJ deck through
 * them -- what a binding in any language does.  No CUDA, C++ or torch types on this side of the boundary.
 *   usage: xnb_c_smoke <path to libxnb_hotpath.so>
 * prints "ok <atoms> <rebuilds> <max |sum f| / sum |f|>" and exits 0 on success. */
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "xnb_hotpath.h"

#define SYM(name) name##_t p_##name; *(void**)(&p_##name) = dlsym(h, #name); if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }
typedef int (*xnb_create_t)(xnb_ctx**, int);
typedef void (*xnb_destroy_t)(xnb_ctx*);
typedef const char* (*xnb_last_error_t)(const xnb_ctx*);
typedef int (*xnb_set_domain_t)(xnb_ctx*, const double*, const double*, double, const int64_t*, const int32_t*);
typedef int (*xnb_set_nbh_dist_t)(xnb_ctx*, double, double);
typedef int (*xnb_set_type_mass_t)(xnb_ctx*, const double*, int);
typedef int (*xnb_set_particles_t)(xnb_ctx*, int64_t, const double*, const double*, const double*, const double*, const double*, const double*, const uint64_t*, const uint8_t*);
typedef int64_t (*xnb_num_inner_t)(const xnb_ctx*);
typedef int (*xnb_get_particles_t)(xnb_ctx*, int64_t, int64_t, double*, double*, double*, double*, double*, double*, double*, double*, double*, uint64_t*, uint8_t*, uint32_t*);
typedef int (*xnb_first_iteration_t)(xnb_ctx*, double, double, double, void*);
typedef int (*xnb_run_steps_t)(xnb_ctx*, int, double, double, double, double, void*, int*);
typedef int64_t (*xnb_host_lattice_fcc_t)(const xnb_lattice_cfg*, int64_t, double*, double*, double*, double*, double*, double*, uint64_t*, uint8_t*);

int main(int argc, char** argv)
{
  if (argc < 2) { fprintf(stderr, "usage: %s libxnb_hotpath.so\n", argv[0]); return 2; }
  void* h = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  SYM(xnb_create) SYM(xnb_destroy) SYM(xnb_last_error) SYM(xnb_set_domain) SYM(xnb_set_nbh_dist) SYM(xnb_set_type_mass) SYM(xnb_set_particles)
  SYM(xnb_num_inner) SYM(xnb_get_particles) SYM(xnb_first_iteration) SYM(xnb_run_steps) SYM(xnb_host_lattice_fcc)

  /* reduced-unit LJ, FCC rho* = 0.8442, 8^3 unit cells, cell = 2 lattice constants */
  const double a = cbrt(4.0 / 0.8442);
  xnb_lattice_cfg cfg;
  int d;
  for (d = 0; d < 3; d++) { cfg.bounds_min[d] = 0.0; cfg.bounds_max[d] = 8.0 * a; cfg.grid_dims[d] = 4; }
  cfg.cell_size = 2.0 * a; cfg.lattice_a = a; cfg.noise_sigma = 0.02; cfg.vel_sigma = 1.2;
  cfg.n_spheres = 0; cfg.sphere_rmin = cfg.sphere_rmax = cfg.drift_speed = 0.0;
  const int64_t cap = 4 * 8 * 8 * 8;
  double* r = (double*)malloc(sizeof(double) * 9 * cap);
  uint64_t* id = (uint64_t*)malloc(sizeof(uint64_t) * cap);
  uint8_t* ty = (uint8_t*)malloc(cap);
  const int64_t n = p_xnb_host_lattice_fcc(&cfg, cap, r, r + cap, r + 2 * cap, r + 3 * cap, r + 4 * cap, r + 5 * cap, id, ty);
  if (n != cap) { fprintf(stderr, "lattice: %lld atoms\n", (long long)n); return 1; }

  xnb_ctx* c = NULL;
  int rc = p_xnb_create(&c, 0);
  if (rc) { fprintf(stderr, "xnb_create: %d %s\n", rc, p_xnb_last_error(NULL)); return rc == XNB_ERR_NO_DEVICE ? 77 : 1; }
  const int32_t periodic[3] = {1, 1, 1};
  const double mass = 1.0;
#define CALL(x) do { rc = (x); if (rc) { fprintf(stderr, "%s: %d %s\n", #x, rc, p_xnb_last_error(c)); return 1; } } while (0)
  CALL(p_xnb_set_domain(c, cfg.bounds_min, cfg.bounds_max, cfg.cell_size, cfg.grid_dims, periodic));
  CALL(p_xnb_set_nbh_dist(c, 2.5, 0.3));
  CALL(p_xnb_set_type_mass(c, &mass, 1));
  CALL(p_xnb_set_particles(c, n, r, r + cap, r + 2 * cap, r + 3 * cap, r + 4 * cap, r + 5 * cap, id, ty));
  CALL(p_xnb_first_iteration(c, 1.0, 1.0, 2.5, NULL));
  int rebuilds = 0;
  CALL(p_xnb_run_steps(c, 25, 0.005, 1.0, 1.0, 2.5, NULL, &rebuilds));
  if (p_xnb_num_inner(c) != n) { fprintf(stderr, "atoms lost\n"); return 1; }
  CALL(p_xnb_get_particles(c, 0, n, r, r + cap, r + 2 * cap, NULL, NULL, NULL, r + 6 * cap, r + 7 * cap, r + 8 * cap, id, NULL, NULL));
  double s[3] = {0, 0, 0}, sa = 0;
  int64_t i;
  for (i = 0; i < n; i++)
    for (d = 0; d < 3; d++) { const double f = r[(6 + d) * cap + i]; if (!isfinite(f)) { fprintf(stderr, "non-finite force\n"); return 1; } s[d] += f; sa += fabs(f); }
  const double rel = fmax(fabs(s[0]), fmax(fabs(s[1]), fabs(s[2]))) / sa;
  p_xnb_destroy(c);
  printf("ok %lld %d %.3e\n", (long long)n, rebuilds, rel);
  return (rel < 1e-11 && rebuilds > 0) ? 0 : 1;
}
