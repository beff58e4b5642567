"""ctypes binding of the CPU oracle (oracle/xnb_oracle.h).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  Nothing under exanbody_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libxnb_oracle.so")


class XoConfig(C.Structure):
    _fields_ = [
        ("bounds_min", C.c_double * 3), ("bounds_max", C.c_double * 3), ("cell_size", C.c_double),
        ("grid_dims", C.c_int64 * 3), ("periodic", C.c_int32 * 3),
        ("lattice_a", C.c_double), ("noise_sigma", C.c_double), ("vel_sigma", C.c_double),
        ("n_spheres", C.c_int32), ("sphere_rmin", C.c_double), ("sphere_rmax", C.c_double), ("drift_speed", C.c_double),
        ("epsilon", C.c_double), ("sigma", C.c_double), ("rcut", C.c_double),
        ("rcut_inc", C.c_double), ("dt", C.c_double), ("mass", C.c_double), ("sub_grid_density", C.c_double),
        ("max_neighbors", C.c_int32), ("serial_order", C.c_int32),
    ]


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("xnb_oracle.cpp", "xnb_oracle.h", "Makefile")]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(f) > os.path.getmtime(_LIB) for f in src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        P = C.c_void_p
        L.xo_create.restype = P; L.xo_create.argtypes = [C.POINTER(XoConfig)]
        L.xo_destroy.argtypes = [P]
        L.xo_last_error.restype = C.c_char_p
        for f in ("xo_init", "xo_generate", "xo_first_iteration", "xo_move_particles", "xo_update_particles_full", "xo_ghost_update_r", "xo_build_neighbors",
                  "xo_compute_force", "xo_compute_force_symmetric", "xo_push_f_v_r", "xo_check_streams", "xo_zero_force"):
            getattr(L, f).argtypes = [P]; getattr(L, f).restype = C.c_int
        L.xo_run.argtypes = [P, C.c_int]; L.xo_run.restype = C.c_int
        L.xo_amr_pair_cache.argtypes = [P, P, P, P]; L.xo_amr_pair_cache.restype = C.c_int64
        L.xo_average_neighbors.argtypes = [P, C.c_double, P, C.c_int, P]; L.xo_average_neighbors.restype = C.c_int
        L.xo_gravitational_force.argtypes = [P, C.c_double, C.c_double, P, C.c_int]; L.xo_gravitational_force.restype = C.c_int
        L.xo_push_f_v.argtypes = [P, C.c_double]; L.xo_push_f_v.restype = C.c_int
        for f in ("xo_displ_over", "xo_total_particles", "xo_inner_particles", "xo_stream_total_u16", "xo_max_neighbors",
                  "xo_rebuild_count"):
            getattr(L, f).argtypes = [P]; getattr(L, f).restype = C.c_int64
        L.xo_set_nbh_config.argtypes = [P, C.c_int, C.c_int]; L.xo_set_nbh_config.restype = None
        L.xo_grid_info.argtypes = [P, P, P, P, P]
        L.xo_cell_counts.argtypes = [P, P]
        L.xo_get_particles.argtypes = [P] * 12
        L.xo_set_particles.argtypes = [P] * 13; L.xo_set_particles.restype = C.c_int
        L.xo_amr_tables.argtypes = [P, P, P]; L.xo_amr_tables.restype = C.c_int64
        L.xo_get_backup.argtypes = [P, P]; L.xo_get_backup.restype = C.c_int64
        L.xo_stream_sizes.argtypes = [P, P]; L.xo_stream_data.argtypes = [P, P]
        L.xo_pairs.argtypes = [P, P]; L.xo_pairs.restype = C.c_int64
        L.xo_energy_virial.argtypes = [P, P, P, P]
        L.xo_num_threads.restype = C.c_int
        L.xo_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_config(bounds_max, cell_size, grid_dims, lattice_a, epsilon, sigma, rcut, rcut_inc, dt, mass=1.0, noise_sigma=0.0,
                vel_sigma=0.0, bounds_min=(0., 0., 0.), periodic=(1, 1, 1), max_neighbors=1024, sub_grid_density=6.5,
                n_spheres=0, sphere_rmin=0., sphere_rmax=0., drift_speed=0., serial_order=1):
    c = XoConfig()
    c.bounds_min[:] = bounds_min; c.bounds_max[:] = bounds_max; c.cell_size = cell_size
    c.grid_dims[:] = grid_dims; c.periodic[:] = periodic
    c.lattice_a = lattice_a; c.noise_sigma = noise_sigma; c.vel_sigma = vel_sigma
    c.n_spheres = n_spheres; c.sphere_rmin = sphere_rmin; c.sphere_rmax = sphere_rmax; c.drift_speed = drift_speed
    c.epsilon = epsilon; c.sigma = sigma; c.rcut = rcut; c.rcut_inc = rcut_inc; c.dt = dt; c.mass = mass
    c.sub_grid_density = sub_grid_density; c.max_neighbors = max_neighbors; c.serial_order = serial_order
    return c


class Oracle:
    """One oracle simulation (single rank, periodic images through self-partner ghosts)."""

    def __init__(self, cfg):
        self.L = lib()
        self.cfg = cfg
        self.h = self.L.xo_create(C.byref(cfg))

    def close(self):
        if self.h:
            self.L.xo_destroy(self.h); self.h = None

    def __del__(self):
        self.close()

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("oracle: " + self.L.xo_last_error().decode())

    def init(self): self._chk(self.L.xo_init(self.h))
    def generate(self): self._chk(self.L.xo_generate(self.h))
    def first_iteration(self): self._chk(self.L.xo_first_iteration(self.h))

    def run(self, n):
        r = self.L.xo_run(self.h, n)
        if r < 0:
            raise RuntimeError("oracle: " + self.L.xo_last_error().decode())
        return r

    def move_particles(self): self._chk(self.L.xo_move_particles(self.h))
    def update_particles_full(self): self._chk(self.L.xo_update_particles_full(self.h))
    def ghost_update_r(self): self._chk(self.L.xo_ghost_update_r(self.h))
    def build_neighbors(self): self._chk(self.L.xo_build_neighbors(self.h))
    def compute_force(self): self._chk(self.L.xo_compute_force(self.h))
    def compute_force_symmetric(self): self._chk(self.L.xo_compute_force_symmetric(self.h))
    def amr_pair_cache(self):
        """(max_res, list_offsets, pairs) of the AmrSubCellPairCache built by the last build_neighbors / update_particles_full"""
        mr = C.c_int64()
        n = self.L.xo_amr_pair_cache(self.h, C.addressof(mr), None, None)
        layers_lists = None
        # number of lists: walk once with a generous offsets buffer
        off = np.zeros(1 + 32 * 33 // 2 * 64, np.uint64); data = np.zeros(max(int(n), 1), np.uint16)
        self.L.xo_amr_pair_cache(self.h, None, off.ctypes.data_as(C.c_void_p), data.ctypes.data_as(C.c_void_p))
        return int(mr.value), off, data[:int(n)]
    def zero_force(self): self._chk(self.L.xo_zero_force(self.h))
    FIELDS = {"rx": 0, "ry": 1, "rz": 2, "vx": 3, "vy": 4, "vz": 5, "fx": 6, "fy": 7, "fz": 8, "id": 9, "type": 10}
    def average_neighbors(self, rcut, nbh_field, weight_function=(1.0, 0.0, 0.0, 0.0)):
        w = np.ascontiguousarray(list(weight_function) + [0.0] * (4 - len(weight_function)), np.float64)
        out = np.zeros(self.n_total())
        self._chk(self.L.xo_average_neighbors(self.h, float(rcut), _p(w), self.FIELDS[nbh_field], _p(out)))
        return out
    def gravitational_force(self, G, rcut, type_mass):
        m = np.ascontiguousarray(type_mass, np.float64)
        self._chk(self.L.xo_gravitational_force(self.h, float(G), float(rcut), m.ctypes.data_as(C.c_void_p), len(m)))

    def set_nbh_config(self, half_symmetric=False, skip_ghosts=False):
        self.L.xo_set_nbh_config(self.h, int(bool(half_symmetric)), int(bool(skip_ghosts)))
    def push_f_v_r(self): self._chk(self.L.xo_push_f_v_r(self.h))
    def push_f_v(self, s): self._chk(self.L.xo_push_f_v(self.h, s))
    def displ_over(self): return self.L.xo_displ_over(self.h)
    def check_streams(self):
        rc = self.L.xo_check_streams(self.h)
        return rc, self.L.xo_last_error().decode() if rc else ""
    def rebuild_count(self): return self.L.xo_rebuild_count(self.h)
    def max_neighbors(self): return self.L.xo_max_neighbors(self.h)
    def stream_total_u16(self): return self.L.xo_stream_total_u16(self.h)

    def grid_info(self):
        d = np.zeros(3, np.int64); o = np.zeros(3, np.int64); gl = C.c_int64(); nc = C.c_int64()
        self.L.xo_grid_info(self.h, _p(d), _p(o), C.addressof(gl), C.addressof(nc))
        return dict(dims=d, offset=o, ghost_layers=gl.value, n_cells=nc.value)

    def n_total(self): return self.L.xo_total_particles(self.h)
    def n_inner(self): return self.L.xo_inner_particles(self.h)

    def cell_counts(self):
        c = np.zeros(self.grid_info()["n_cells"], np.int32)
        self.L.xo_cell_counts(self.h, _p(c)); return c

    def particles(self):
        n = self.n_total()
        out = {k: np.zeros(n, np.float64) for k in ("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz")}
        out["id"] = np.zeros(n, np.uint64); out["type"] = np.zeros(n, np.uint8)
        self.L.xo_get_particles(self.h, *[_p(out[k]) for k in ("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz", "id", "type")])
        return out

    def set_particles(self, counts, p):
        counts = np.ascontiguousarray(counts, np.int32)
        arrs = [np.ascontiguousarray(p[k], np.float64) if k in p else None for k in ("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz")]
        ids = np.ascontiguousarray(p["id"], np.uint64) if "id" in p else None
        ty = np.ascontiguousarray(p["type"], np.uint8) if "type" in p else None
        self._chk(self.L.xo_set_particles(self.h, _p(counts), *[_p(a) for a in arrs], _p(ids), _p(ty)))

    def inner_mask(self):
        """bool mask over flat (cell-ordered) particles: True for particles of inner (non ghost) cells"""
        gi = self.grid_info(); d = gi["dims"]; gl = gi["ghost_layers"]
        k, j, i = np.meshgrid(np.arange(d[2]), np.arange(d[1]), np.arange(d[0]), indexing="ij")
        inner = ((i >= gl) & (i < d[0] - gl) & (j >= gl) & (j < d[1] - gl) & (k >= gl) & (k < d[2] - gl)).ravel()
        return np.repeat(inner, self.cell_counts())

    def amr_tables(self):
        nc = self.grid_info()["n_cells"]
        sgs = np.zeros(nc + 1, np.int64)
        n = self.L.xo_amr_tables(self.h, _p(sgs), None)
        sgc = np.zeros(n, np.uint32)
        self.L.xo_amr_tables(self.h, None, _p(sgc))
        return sgs, sgc

    def backup(self):
        n = self.L.xo_get_backup(self.h, None)
        b = np.zeros(n, np.uint32); self.L.xo_get_backup(self.h, _p(b)); return b

    def streams(self):
        nc = self.grid_info()["n_cells"]
        sz = np.zeros(nc, np.uint32); self.L.xo_stream_sizes(self.h, _p(sz))
        data = np.zeros(self.L.xo_stream_total_u16(self.h), np.uint16); self.L.xo_stream_data(self.h, _p(data))
        return sz, data

    def pairs(self):
        n = self.L.xo_pairs(self.h, None)
        out = np.zeros((n, 2), np.uint64); self.L.xo_pairs(self.h, _p(out)); return out

    def energy_virial(self):
        e = C.c_double(); k = C.c_double(); w = np.zeros(6)
        self.L.xo_energy_virial(self.h, C.addressof(e), _p(w), C.addressof(k))
        return e.value, w, k.value


def num_threads():
    return lib().xo_num_threads()


def set_num_threads(n):
    lib().xo_set_num_threads(int(n))
