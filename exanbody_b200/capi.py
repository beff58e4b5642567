"""ctypes binding of the C-ABI declared in include/xnb_hotpath.h.

This is the same stub a maintainer of the reference would write to call the library from Python (INTEGRATION.md shows the
C++ one).  It computes nothing itself: every method forwards to libxnb_hotpath.so, and loading fails loudly when the
library is absent (no CPU fallback)."""
import ctypes as C
import os
import numpy as np

from . import buildlib as _build

XNB_OK = 0
ERRORS = {1: "XNB_ERR_NO_DEVICE", 2: "XNB_ERR_INVALID", 3: "XNB_ERR_CUDA", 4: "XNB_ERR_CAPACITY", 5: "XNB_ERR_LOST_PARTICLE", 6: "XNB_ERR_NCCL"}

# every symbol include/xnb_hotpath.h declares (tests/test_capi_symbols.py checks header <-> library <-> this list)
SYMBOLS = [
    "xnb_create", "xnb_destroy", "xnb_last_error", "xnb_version", "xnb_set_domain", "xnb_init_rcb_grid", "xnb_set_nbh_dist",
    "xnb_set_type_mass", "xnb_set_sub_grid_density", "xnb_set_nccl_comm", "xnb_nccl_unique_id", "xnb_nccl_init_rank",
    "xnb_set_particles", "xnb_num_inner", "xnb_num_total", "xnb_get_particles", "xnb_upload_rv", "xnb_download_rvf",
    "xnb_get_grid_info", "xnb_get_sweep_info", "xnb_get_cells", "xnb_view_particles", "xnb_device_allocations", "xnb_move_particles", "xnb_rebuild_amr", "xnb_backup_r", "xnb_ghost_comm_scheme",
    "xnb_ghost_update_all", "xnb_ghost_update_r", "xnb_ghost_transport", "xnb_chunk_neighbors", "xnb_zero_particle_force", "xnb_set_pair_functor", "xnb_lennard_jones_force", "xnb_gravitational_force", "xnb_average_neighbors", "xnb_get_generic_field", "xnb_load_balance_rcb", "xnb_get_block", "xnb_host_amr_sub_cell_pairs",
    "xnb_divide_force_by_mass", "xnb_set_chunk_neighbors_config", "xnb_lennard_jones_force_symmetric", "xnb_update_force_from_ghost", "xnb_push_f_v_r", "xnb_push_f_v", "xnb_particle_displ_over", "xnb_verlet_first_half",
    "xnb_read_displ_over", "xnb_force_and_second_half", "xnb_run_steps", "xnb_step_host", "xnb_step_host_n", "xnb_first_iteration", "xnb_energy_virial",
    "xnb_view_chunk_neighbors", "xnb_stream_pool_u16", "xnb_get_streams", "xnb_get_amr", "xnb_get_backup",
    "xnb_rebuild_count", "xnb_kernel_launches", "xnb_timing_enable", "xnb_timing_read", "xnb_measure_dfma_peak",
    "xnb_host_lattice_fcc", "xnb_host_rcb_block", "xnb_host_load_balance_rcb", "xnb_host_simple_cost_model", "xnb_host_ghost_items",
]


class XnbGridInfo(C.Structure):
    _fields_ = [("dims", C.c_int64 * 3), ("offset", C.c_int64 * 3), ("ghost_layers", C.c_int64), ("n_cells", C.c_int64),
                ("block_start", C.c_int64 * 3), ("block_end", C.c_int64 * 3)]


class XnbSweepInfo(C.Structure):
    _fields_ = [("compiled", C.c_int32), ("ghost", C.c_int32), ("tile", C.c_int64 * 3), ("threads", C.c_int64), ("blocks", C.c_int64),
                ("smem_bytes", C.c_int64), ("rows", C.c_int64), ("candidates", C.c_int64), ("interior_tiles", C.c_int64), ("boundary_tiles", C.c_int64)]


class XnbParticleView(C.Structure):
    _fields_ = [("n_inner", C.c_int64), ("n_total", C.c_int64), ("n_cells", C.c_int64)] + [(k, C.c_void_p) for k in
                ("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz", "id", "type", "cell_start", "cell_count", "particle_cell")]


class XnbLatticeCfg(C.Structure):
    _fields_ = [("bounds_min", C.c_double * 3), ("bounds_max", C.c_double * 3), ("cell_size", C.c_double), ("grid_dims", C.c_int64 * 3),
                ("lattice_a", C.c_double), ("noise_sigma", C.c_double), ("vel_sigma", C.c_double),
                ("n_spheres", C.c_int32), ("sphere_rmin", C.c_double), ("sphere_rmax", C.c_double), ("drift_speed", C.c_double)]


class XnbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (ERRORS.get(code, code), msg))
        self.code = code


_lib = None


def library_path():
    # XNB_HOTPATH_LIB: development knob (kernel experiments built beside the product library); default = the in-tree build
    return os.environ.get("XNB_HOTPATH_LIB") or _build.LIB


def load():
    """dlopen libxnb_hotpath.so (built in-tree by exanbody_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ImportError("libxnb_hotpath.so is not built (%s): run `python -m exanbody_b200.buildlib`; there is no CPU fallback" % path)
    L = C.CDLL(path, mode=C.RTLD_GLOBAL)
    P, I, D, I64 = C.c_void_p, C.c_int, C.c_double, C.c_int64
    sig = {
        "xnb_create": (I, [C.POINTER(P), I]), "xnb_destroy": (None, [P]), "xnb_last_error": (C.c_char_p, [P]), "xnb_version": (C.c_char_p, []),
        "xnb_set_domain": (I, [P, P, P, D, P, P]), "xnb_init_rcb_grid": (I, [P, I, I]), "xnb_set_nbh_dist": (I, [P, D, D]),
        "xnb_set_type_mass": (I, [P, P, I]), "xnb_set_sub_grid_density": (I, [P, D]), "xnb_set_nccl_comm": (I, [P, P]),
        "xnb_nccl_unique_id": (I, [P]), "xnb_nccl_init_rank": (I, [P, P, I, I]),
        "xnb_set_particles": (I, [P, I64] + [P] * 8), "xnb_num_inner": (I64, [P]), "xnb_num_total": (I64, [P]),
        "xnb_get_particles": (I, [P, I64, I64] + [P] * 12), "xnb_upload_rv": (I, [P] * 8), "xnb_download_rvf": (I, [P] * 12),
        "xnb_get_grid_info": (I, [P, C.POINTER(XnbGridInfo)]), "xnb_get_sweep_info": (I, [P, C.POINTER(XnbSweepInfo)]), "xnb_get_cells": (I, [P, P, P]),
        "xnb_view_particles": (I, [P, C.POINTER(XnbParticleView)]), "xnb_device_allocations": (I64, []),
        "xnb_move_particles": (I, [P, P]), "xnb_rebuild_amr": (I, [P, P]), "xnb_backup_r": (I, [P, P]), "xnb_ghost_comm_scheme": (I, [P, P]),
        "xnb_ghost_update_all": (I, [P, P]), "xnb_ghost_update_r": (I, [P, P]), "xnb_chunk_neighbors": (I, [P, P]),
        "xnb_zero_particle_force": (I, [P, I, P]), "xnb_set_pair_functor": (I, [P, I]), "xnb_lennard_jones_force": (I, [P, D, D, D, I, P]), "xnb_gravitational_force": (I, [P, D, D, I, I, P]), "xnb_average_neighbors": (I, [P, D, P, I, P]), "xnb_ghost_transport": (I, [P]), "xnb_get_generic_field": (I, [P, P]), "xnb_load_balance_rcb": (I, [P, P, P, P, P]), "xnb_get_block": (I, [P, I, P, P]), "xnb_host_amr_sub_cell_pairs": (I64, [I, D, D, P, P]), "xnb_divide_force_by_mass": (I, [P, P]),
        "xnb_set_chunk_neighbors_config": (I, [P, I, I]), "xnb_lennard_jones_force_symmetric": (I, [P, D, D, D, P]), "xnb_update_force_from_ghost": (I, [P, P]),
        "xnb_push_f_v_r": (I, [P, D, D, P]), "xnb_push_f_v": (I, [P, D, D, P]), "xnb_particle_displ_over": (I, [P, P, P]),
        "xnb_verlet_first_half": (I, [P, D, P]), "xnb_read_displ_over": (I, [P, P, P]), "xnb_force_and_second_half": (I, [P, D, D, D, D, P]),
        "xnb_run_steps": (I, [P, I, D, D, D, D, P, P]), "xnb_step_host": (I, [P, D, D, D, D, P, P, P, P, P, P, I, P, P]), "xnb_step_host_n": (I, [P, D, D, D, D, P, P, P, P, P, P, I64, P, P, P]), "xnb_first_iteration": (I, [P, D, D, D, P]),
        "xnb_energy_virial": (I, [P, D, D, D, P, P, P, P]),
        "xnb_view_chunk_neighbors": (I, [P, P, P, P]), "xnb_stream_pool_u16": (I64, [P]), "xnb_get_streams": (I, [P, P, P]),
        "xnb_get_amr": (I64, [P, P, P]), "xnb_get_backup": (I, [P, P]), "xnb_rebuild_count": (I64, [P]), "xnb_kernel_launches": (I64, [P]),
        "xnb_host_lattice_fcc": (I64, [C.POINTER(XnbLatticeCfg), I64] + [P] * 8),
        "xnb_host_rcb_block": (I, [P, I, I, P, P]), "xnb_host_load_balance_rcb": (I, [P, P, I, I, P, P, P]),
        "xnb_host_simple_cost_model": (I, [I64, P, D, P, P]), "xnb_host_ghost_items": (I64, [P, P, I, I, I, I, I64, P, P, P]),
        "xnb_timing_enable": (I, [P, I]), "xnb_timing_read": (I, [P, P, P, I]), "xnb_measure_dfma_peak": (I, [I, P]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))   # raw address (e.g. pinned torch tensor data_ptr)


class Context:
    """One sub-domain on one GPU (xnb_ctx)."""

    def __init__(self, device=0):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.xnb_create(C.byref(h), device)
        if rc != XNB_OK:
            raise XnbError(rc, self.L.xnb_last_error(None).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.xnb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != XNB_OK:
            raise XnbError(rc, self.L.xnb_last_error(self.h).decode())

    # ---- configuration
    def set_domain(self, bounds_min, bounds_max, cell_size, grid_dims, periodic=(1, 1, 1)):
        a = np.asarray(bounds_min, np.float64); b = np.asarray(bounds_max, np.float64)
        d = np.asarray(grid_dims, np.int64); p = np.asarray(periodic, np.int32)
        self._ck(self.L.xnb_set_domain(self.h, _p(a), _p(b), float(cell_size), _p(d), _p(p)))

    def init_rcb_grid(self, rank=0, nranks=1): self._ck(self.L.xnb_init_rcb_grid(self.h, rank, nranks))
    def set_nbh_dist(self, rcut_max, rcut_inc): self._ck(self.L.xnb_set_nbh_dist(self.h, float(rcut_max), float(rcut_inc)))

    def set_type_mass(self, masses):
        m = np.ascontiguousarray(masses, np.float64)
        self._ck(self.L.xnb_set_type_mass(self.h, _p(m), len(m)))

    def set_sub_grid_density(self, d): self._ck(self.L.xnb_set_sub_grid_density(self.h, float(d)))

    def nccl_unique_id(self):
        b = np.zeros(128, np.uint8)
        rc = self.L.xnb_nccl_unique_id(_p(b))
        if rc != XNB_OK:
            raise XnbError(rc, "ncclGetUniqueId failed")
        return b

    def nccl_init_rank(self, uid, rank, nranks):
        b = np.ascontiguousarray(uid, np.uint8)
        self._ck(self.L.xnb_nccl_init_rank(self.h, _p(b), rank, nranks))

    # ---- particles
    def set_particles(self, rx, ry, rz, vx=None, vy=None, vz=None, id=None, type=None):
        f = lambda a: None if a is None else np.ascontiguousarray(a, np.float64)
        rx, ry, rz, vx, vy, vz = map(f, (rx, ry, rz, vx, vy, vz))
        ids = None if id is None else np.ascontiguousarray(id, np.uint64)
        ty = None if type is None else np.ascontiguousarray(type, np.uint8)
        self._ck(self.L.xnb_set_particles(self.h, len(rx), _p(rx), _p(ry), _p(rz), _p(vx), _p(vy), _p(vz), _p(ids), _p(ty)))

    @property
    def n_inner(self): return self.L.xnb_num_inner(self.h)
    @property
    def n_total(self): return self.L.xnb_num_total(self.h)

    def get_particles(self, first=0, n=None, fields=("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz", "id", "type", "cell")):
        if n is None:
            n = self.n_total - first
        names = ("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz", "id", "type", "cell")
        dt = dict(id=np.uint64, type=np.uint8, cell=np.uint32)
        out = {k: np.zeros(n, dt.get(k, np.float64)) for k in names if k in fields}
        self._ck(self.L.xnb_get_particles(self.h, first, n, *[_p(out.get(k)) for k in names]))
        return out

    def upload_rv(self, rx, ry, rz, vx, vy, vz, stream=None):
        self._ck(self.L.xnb_upload_rv(self.h, _p(rx), _p(ry), _p(rz), _p(vx), _p(vy), _p(vz), _p(stream)))

    def download_rvf(self, rx, ry, rz, vx, vy, vz, fx, fy, fz, id=None, stream=None):
        self._ck(self.L.xnb_download_rvf(self.h, _p(rx), _p(ry), _p(rz), _p(vx), _p(vy), _p(vz), _p(fx), _p(fy), _p(fz), _p(id), _p(stream)))

    def grid_info(self):
        gi = XnbGridInfo()
        self._ck(self.L.xnb_get_grid_info(self.h, C.byref(gi)))
        return dict(dims=np.array(gi.dims[:]), offset=np.array(gi.offset[:]), ghost_layers=gi.ghost_layers, n_cells=gi.n_cells,
                    block_start=np.array(gi.block_start[:]), block_end=np.array(gi.block_end[:]))

    def sweep_info(self):
        si = XnbSweepInfo()
        self._ck(self.L.xnb_get_sweep_info(self.h, C.byref(si)))
        return dict(compiled=bool(si.compiled), ghost=bool(si.ghost), tile=tuple(si.tile[:]), threads=si.threads, blocks=si.blocks,
                    smem_bytes=si.smem_bytes, rows=si.rows, candidates=si.candidates, interior_tiles=si.interior_tiles, boundary_tiles=si.boundary_tiles)

    def view_particles(self):
        """device pointers of the SoA particle arrays and of the per-cell tables (xnb_particle_view); no copy, no synchronisation"""
        v = XnbParticleView()
        self._ck(self.L.xnb_view_particles(self.h, C.byref(v)))
        return {k: getattr(v, k) for k, _ in XnbParticleView._fields_}

    def device_allocations(self): return int(self.L.xnb_device_allocations())

    def cells(self):
        nc = self.grid_info()["n_cells"]
        s = np.zeros(nc, np.uint32); c = np.zeros(nc, np.uint32)
        self._ck(self.L.xnb_get_cells(self.h, _p(s), _p(c)))
        return s, c

    # ---- operators
    def move_particles(self, stream=None): self._ck(self.L.xnb_move_particles(self.h, _p(stream)))
    def rebuild_amr(self, stream=None): self._ck(self.L.xnb_rebuild_amr(self.h, _p(stream)))
    def backup_r(self, stream=None): self._ck(self.L.xnb_backup_r(self.h, _p(stream)))
    def ghost_comm_scheme(self, stream=None): self._ck(self.L.xnb_ghost_comm_scheme(self.h, _p(stream)))
    def ghost_update_all(self, stream=None): self._ck(self.L.xnb_ghost_update_all(self.h, _p(stream)))
    def ghost_update_r(self, stream=None): self._ck(self.L.xnb_ghost_update_r(self.h, _p(stream)))
    def chunk_neighbors(self, stream=None): self._ck(self.L.xnb_chunk_neighbors(self.h, _p(stream)))
    def zero_particle_force(self, ghost=True, stream=None): self._ck(self.L.xnb_zero_particle_force(self.h, int(ghost), _p(stream)))
    def set_pair_functor(self, functor):
        """0 = restated LJ functor (default), 1 = the reference's literal LJ form through the generic buffer-less call"""
        self._ck(self.L.xnb_set_pair_functor(self.h, int(functor)))
    def lennard_jones_force(self, epsilon, sigma, rcut, ghost=False, stream=None):
        self._ck(self.L.xnb_lennard_jones_force(self.h, epsilon, sigma, rcut, int(ghost), _p(stream)))
    def ghost_transport(self):
        """'peer' (NVLink peer-memory mailboxes) or 'nccl' (send / recv per partner)"""
        return "peer" if self.L.xnb_ghost_transport(self.h) == 1 else "nccl"

    FIELDS = {"rx": 0, "ry": 1, "rz": 2, "vx": 3, "vy": 4, "vz": 5, "fx": 6, "fy": 7, "fz": 8, "id": 9, "type": 10}

    def average_neighbors(self, rcut, nbh_field, weight_function=None, stream=None):
        """op average_neighbors_scalar (src/compute/average_neighbors.cu): returns the averaged field of the inner particles (current order)"""
        w = None if weight_function is None else np.ascontiguousarray(list(weight_function) + [0.0] * (4 - len(weight_function)), np.float64)
        self._ck(self.L.xnb_average_neighbors(self.h, rcut, _p(w), self.FIELDS[nbh_field], _p(stream)))
        out = np.zeros(self.n_inner)
        self._ck(self.L.xnb_get_generic_field(self.h, _p(out)))
        return out

    def gravitational_force(self, G, rcut, ghost=False, buffer_form=False, stream=None):
        """op gravitational_force (contribs/pi/gravitational_force.cu) through the general pair sweep"""
        self._ck(self.L.xnb_gravitational_force(self.h, G, rcut, int(ghost), int(buffer_form), _p(stream)))
    def divide_force_by_mass(self, stream=None): self._ck(self.L.xnb_divide_force_by_mass(self.h, _p(stream)))

    def set_chunk_neighbors_config(self, half_symmetric=False, skip_ghosts=False):
        self._ck(self.L.xnb_set_chunk_neighbors_config(self.h, int(bool(half_symmetric)), int(bool(skip_ghosts))))

    def lennard_jones_force_symmetric(self, epsilon, sigma, rcut, stream=None):
        self._ck(self.L.xnb_lennard_jones_force_symmetric(self.h, epsilon, sigma, rcut, _p(stream)))

    def update_force_from_ghost(self, stream=None): self._ck(self.L.xnb_update_force_from_ghost(self.h, _p(stream)))
    def push_f_v_r(self, dt, dt_scale=1.0, stream=None): self._ck(self.L.xnb_push_f_v_r(self.h, dt, dt_scale, _p(stream)))
    def push_f_v(self, dt, dt_scale=0.5, stream=None): self._ck(self.L.xnb_push_f_v(self.h, dt, dt_scale, _p(stream)))

    def particle_displ_over(self, stream=None):
        v = C.c_uint64()
        self._ck(self.L.xnb_particle_displ_over(self.h, C.addressof(v), _p(stream)))
        return v.value

    def verlet_first_half(self, dt, stream=None): self._ck(self.L.xnb_verlet_first_half(self.h, dt, _p(stream)))

    def read_displ_over(self, stream=None):
        v = C.c_uint64()
        self._ck(self.L.xnb_read_displ_over(self.h, C.addressof(v), _p(stream)))
        return v.value

    def force_and_second_half(self, epsilon, sigma, rcut, dt_half_kick, stream=None):
        self._ck(self.L.xnb_force_and_second_half(self.h, epsilon, sigma, rcut, dt_half_kick, _p(stream)))

    def run_steps(self, nsteps, dt, epsilon, sigma, rcut, stream=None):
        r = C.c_int()
        self._ck(self.L.xnb_run_steps(self.h, nsteps, dt, epsilon, sigma, rcut, _p(stream), C.addressof(r)))
        return r.value

    def step_host(self, dt, epsilon, sigma, rcut, in_r=None, in_v=None, out_r=None, out_v=None, out_f=None, out_id=None, id_always=False, stream=None):
        """xnb_step_host: triples of host addresses (ints / arrays) or None; returns the number of rebuilds (0 or 1)"""
        def trip(t):
            if t is None:
                return None
            a = (C.c_void_p * 3)(*[_p(x) for x in t])
            return a
        keep = [trip(t) for t in (in_r, in_v, out_r, out_v, out_f)]
        r = C.c_int()
        self._ck(self.L.xnb_step_host(self.h, dt, epsilon, sigma, rcut, *[C.cast(k, C.c_void_p) if k is not None else None for k in keep],
                                      _p(out_id), int(bool(id_always)), _p(stream), C.addressof(r)))
        return r.value

    def step_host_n(self, dt, epsilon, sigma, rcut, capacity, in_r=None, in_v=None, out_r=None, out_v=None, out_f=None, out_id=None, stream=None):
        """xnb_step_host_n (several sub-domains): returns (rebuilds, particles this rank owns after the step)"""
        def trip(t):
            return None if t is None else (C.c_void_p * 3)(*[_p(x) for x in t])
        keep = [trip(t) for t in (in_r, in_v, out_r, out_v, out_f)]
        r = C.c_int(); n = C.c_int64()
        self._ck(self.L.xnb_step_host_n(self.h, dt, epsilon, sigma, rcut, *[C.cast(k, C.c_void_p) if k is not None else None for k in keep],
                                        _p(out_id), int(capacity), C.addressof(n), _p(stream), C.addressof(r)))
        return r.value, n.value

    def first_iteration(self, epsilon, sigma, rcut, stream=None): self._ck(self.L.xnb_first_iteration(self.h, epsilon, sigma, rcut, _p(stream)))

    def update_particles_full(self, stream=None):
        """parallel_update_particles of data/config/update-particles.msp:47-53 (after move_particles)"""
        self.rebuild_amr(stream); self.backup_r(stream); self.ghost_comm_scheme(stream); self.ghost_update_all(stream); self.chunk_neighbors(stream)

    def load_balance_rcb(self, coefs=None, stream=None):
        """op load_balance_rcb on the live context (collective): returns (lb_inbalance before, after); continue with update_particles_full()"""
        b = C.c_double(); a = C.c_double()
        k = None if coefs is None else np.ascontiguousarray(coefs, np.float64)
        self._ck(self.L.xnb_load_balance_rcb(self.h, _p(k), C.addressof(b), C.addressof(a), _p(stream)))
        return b.value, a.value

    def block(self, rank):
        s = np.zeros(3, np.int64); e = np.zeros(3, np.int64)
        self._ck(self.L.xnb_get_block(self.h, int(rank), _p(s), _p(e)))
        return s, e

    def energy_virial(self, epsilon, sigma, rcut, stream=None):
        e = C.c_double(); k = C.c_double(); w = np.zeros(6)
        self._ck(self.L.xnb_energy_virial(self.h, epsilon, sigma, rcut, C.addressof(e), _p(w), C.addressof(k), _p(stream)))
        return e.value, w, k.value

    # ---- derived data
    def streams(self):
        nc = self.grid_info()["n_cells"]
        sz = np.zeros(nc, np.uint32)
        self._ck(self.L.xnb_get_streams(self.h, _p(sz), None))
        data = np.zeros(int(sz.sum()), np.uint16)
        self._ck(self.L.xnb_get_streams(self.h, _p(sz), _p(data)))
        return sz, data

    def stream_sizes(self):
        """per-cell stream size in u16 words (GridChunkNeighbors::m_cell_stream_size / 2)"""
        sz = np.zeros(self.grid_info()["n_cells"], np.uint32)
        self._ck(self.L.xnb_get_streams(self.h, _p(sz), None))
        return sz

    def view_chunk_neighbors(self):
        ps = C.c_void_p(); pb = C.c_void_p(); mx = C.c_uint32()
        self._ck(self.L.xnb_view_chunk_neighbors(self.h, C.addressof(ps), C.addressof(pb), C.addressof(mx)))
        return ps.value, pb.value, mx.value

    def stream_pool_u16(self): return self.L.xnb_stream_pool_u16(self.h)

    def amr_tables(self):
        nc = self.grid_info()["n_cells"]
        sgs = np.zeros(nc + 1, np.int64)
        n = self.L.xnb_get_amr(self.h, _p(sgs), None)
        sgc = np.zeros(max(n, 0), np.uint32)
        if n > 0:
            self.L.xnb_get_amr(self.h, None, _p(sgc))
        return sgs, sgc

    def backup(self):
        b = np.zeros(3 * self.n_inner, np.uint32)
        self._ck(self.L.xnb_get_backup(self.h, _p(b)))
        return b

    def rebuild_count(self): return self.L.xnb_rebuild_count(self.h)
    def kernel_launches(self): return self.L.xnb_kernel_launches(self.h)
    def timing_enable(self, on=True): self._ck(self.L.xnb_timing_enable(self.h, int(on)))

    TIMING_SCOPES = ("force", "nbh", "first_half", "bin", "ghost_scheme", "ghost_update")

    def timing_read(self, reset=True):
        """device ms and scope count per kernel group (XNB_T_* of the header) since the last reset"""
        ms = np.zeros(len(self.TIMING_SCOPES)); n = np.zeros(len(self.TIMING_SCOPES), np.int64)
        self._ck(self.L.xnb_timing_read(self.h, _p(ms), _p(n), int(reset)))
        return {k: dict(ms=float(ms[i]), n=int(n[i])) for i, k in enumerate(self.TIMING_SCOPES)}


def measure_dfma_peak(device=0):
    L = load()
    v = C.c_double()
    rc = L.xnb_measure_dfma_peak(device, C.addressof(v))
    if rc != XNB_OK:
        raise XnbError(rc, "DFMA probe failed")
    return v.value


def lattice_fcc(bounds_max, cell_size, grid_dims, lattice_a, noise_sigma=0.0, vel_sigma=0.0, bounds_min=(0., 0., 0.),
                n_spheres=0, sphere_rmin=0., sphere_rmax=0., drift_speed=0.):
    """ops `lattice` + `gaussian_noise_r` (host side, deterministic): returns dict of numpy arrays in domain-cell order"""
    L = load()
    cfg = XnbLatticeCfg()
    cfg.bounds_min[:] = bounds_min; cfg.bounds_max[:] = bounds_max; cfg.cell_size = cell_size; cfg.grid_dims[:] = grid_dims
    cfg.lattice_a = lattice_a; cfg.noise_sigma = noise_sigma; cfg.vel_sigma = vel_sigma
    cfg.n_spheres = n_spheres; cfg.sphere_rmin = sphere_rmin; cfg.sphere_rmax = sphere_rmax; cfg.drift_speed = drift_speed
    cap = 4
    for d in range(3):
        cap *= int(np.ceil((bounds_max[d] - bounds_min[d]) / lattice_a)) + 1
    out = {k: np.zeros(cap, np.float64) for k in ("rx", "ry", "rz", "vx", "vy", "vz")}
    out["id"] = np.zeros(cap, np.uint64); out["type"] = np.zeros(cap, np.uint8)
    n = L.xnb_host_lattice_fcc(C.byref(cfg), cap, *[_p(out[k]) for k in ("rx", "ry", "rz", "vx", "vy", "vz", "id", "type")])
    if n < 0:
        raise XnbError(4, "lattice capacity")
    return {k: v[:n].copy() for k, v in out.items()}


def rcb_block(grid_dims, nranks, rank):
    """op init_rcb_grid (host only): this rank's block [start, end) of the domain cell grid"""
    L = load()
    gd = np.ascontiguousarray(grid_dims, np.int64); s = np.zeros(3, np.int64); e = np.zeros(3, np.int64)
    rc = L.xnb_host_rcb_block(_p(gd), nranks, rank, _p(s), _p(e))
    if rc:
        raise XnbError(rc, "xnb_host_rcb_block")
    return s, e


def simple_cost_model(cell_count, cell_size, coefs=(0.0, 0.0, 1.0, 0.0)):
    """op simple_cost_model, arithmetic only: per-cell cost from the particle counts"""
    L = load()
    cnt = np.ascontiguousarray(cell_count, np.uint32); co = np.ascontiguousarray(coefs, np.float64); out = np.zeros(cnt.size, np.float64)
    rc = L.xnb_host_simple_cost_model(cnt.size, _p(cnt), float(cell_size), _p(co), _p(out))
    if rc:
        raise XnbError(rc, "xnb_host_simple_cost_model")
    return out


def load_balance_rcb(grid_dims, cell_costs, nranks, rank):
    """op load_balance_rcb, host half: this rank's block of the cost-weighted recursive bisection and the cost inside it;
    cell_costs = all-reduced cost per domain cell, index (k*dj + j)*di + i"""
    L = load()
    gd = np.ascontiguousarray(grid_dims, np.int64); s = np.zeros(3, np.int64); e = np.zeros(3, np.int64)
    cc = np.ascontiguousarray(cell_costs, np.float64)
    if cc.size != int(gd[0] * gd[1] * gd[2]):
        raise ValueError("cell_costs must hold one value per domain cell")
    cost = C.c_double()
    rc = L.xnb_host_load_balance_rcb(_p(gd), _p(cc), nranks, rank, _p(s), _p(e), C.addressof(cost))
    if rc:
        raise XnbError(rc, "xnb_host_load_balance_rcb")
    return s, e, cost.value


def ghost_items(grid_dims, periodic, ghost_layers, nranks, src_rank, dst_rank):
    """op ghost_comm_scheme, static part (host only): (sender cell, receiver ghost cell, flags) of every cell src_rank sends to dst_rank"""
    L = load()
    gd = np.ascontiguousarray(grid_dims, np.int64); per = np.ascontiguousarray(periodic, np.int32)
    n = L.xnb_host_ghost_items(_p(gd), _p(per), ghost_layers, nranks, src_rank, dst_rank, 0, None, None, None)
    if n < 0:
        raise XnbError(2, "xnb_host_ghost_items")
    a = np.zeros(n, np.uint32); b = np.zeros(n, np.uint32); f = np.zeros(n, np.uint32)
    L.xnb_host_ghost_items(_p(gd), _p(per), ghost_layers, nranks, src_rank, dst_rank, n, _p(a), _p(b), _p(f))
    return a, b, f


def amr_sub_cell_pairs(max_res, cell_size, max_dist):
    """op amr_grid_pairs (host): (list_offsets, pairs) of the AmrSubCellPairCache"""
    L = load()
    layers = int(np.ceil(max_dist / cell_size))
    n_lists = max_res * (max_res + 1) // 2 * (layers + 1) ** 3
    n = L.xnb_host_amr_sub_cell_pairs(int(max_res), float(cell_size), float(max_dist), None, None)
    if n < 0:
        raise ValueError("amr_sub_cell_pairs: bad arguments")
    off = np.zeros(n_lists + 1, np.uint64); data = np.zeros(max(int(n), 1), np.uint16)
    L.xnb_host_amr_sub_cell_pairs(int(max_res), float(cell_size), float(max_dist), _p(off), _p(data))
    return off, data[:int(n)]
