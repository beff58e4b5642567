"""The drop-in boundary exercised the way a foreign caller sees it: a plain C99 program that dlopen()s the library,
the device view of the grid, and the promise that steady-state steps do not allocate device memory."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import lj_reduced_kwargs
import parity_util as U

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_program_runs_a_deck(tmp_path):
    from exanbody_b200 import buildlib
    exe = tmp_path / "xnb_c_smoke"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-D_GNU_SOURCE", "-I", os.path.join(ROOT, "include"),
                           "-o", str(exe), os.path.join(ROOT, "tests", "c", "xnb_c_smoke.c"), "-ldl", "-lm"])
    out = subprocess.run([str(exe), buildlib.LIB], capture_output=True, text=True)
    assert out.returncode == 0, (out.stdout, out.stderr)
    tok = out.stdout.split()
    assert tok[0] == "ok" and int(tok[1]) == 2048 and int(tok[2]) > 0


def _d2h(ptr, n, dtype):
    """cudaMemcpy of n elements from a raw device pointer (the runtime torch already loaded)"""
    import torch  # noqa: F401  (loads libcudart)
    rt = None
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            rt = ctypes.CDLL(name); break
        except OSError:
            continue
    assert rt is not None
    out = np.empty(n, dtype)
    rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    assert rt.cudaDeviceSynchronize() == 0
    assert rt.cudaMemcpy(out.ctypes.data, ctypes.c_void_p(ptr), out.nbytes, 2) == 0
    return out


def test_device_view_of_the_grid():
    """xnb_view_particles: device SoA pointers + per-cell tables = the cells[c][field] view of grid.h:70-71 without a host copy"""
    kw = lj_reduced_kwargs(ncell_units=8, cell_units=2)
    ctx = U.make_ctx(kw)
    ctx.first_iteration(kw["epsilon"], kw["sigma"], kw["rcut"])
    ctx.run_steps(12, kw["dt"], kw["epsilon"], kw["sigma"], kw["rcut"])
    v = ctx.view_particles()
    p = ctx.get_particles()
    start, count = ctx.cells()
    assert v["n_inner"] == ctx.n_inner and v["n_total"] == ctx.n_total and v["n_cells"] == len(count)
    for k in ("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz"):
        assert np.array_equal(_d2h(v[k], v["n_total"], np.float64), p[k]), k
    assert np.array_equal(_d2h(v["id"], v["n_total"], np.uint64), p["id"])
    assert np.array_equal(_d2h(v["type"], v["n_total"], np.uint8), p["type"])
    assert np.array_equal(_d2h(v["cell_start"], v["n_cells"], np.uint32), start)
    assert np.array_equal(_d2h(v["cell_count"], v["n_cells"], np.uint32), count)
    cell = _d2h(v["particle_cell"], v["n_total"], np.uint32)
    # every particle lies in the slice of its cell
    assert np.all(np.arange(v["n_total"]) >= start[cell]) and np.all(np.arange(v["n_total"]) < start[cell] + count[cell])


def test_steady_state_steps_do_not_allocate():
    """after the first rebuilds have sized the buffers, steps (rebuilds included) run without cudaMalloc / cudaFree"""
    kw = lj_reduced_kwargs(ncell_units=12, cell_units=2)
    ctx = U.make_ctx(kw)
    eps, sig, rc, dt = kw["epsilon"], kw["sigma"], kw["rcut"], kw["dt"]
    ctx.first_iteration(eps, sig, rc)
    assert ctx.run_steps(30, dt, eps, sig, rc) >= 2
    a0 = ctx.device_allocations()
    assert ctx.run_steps(60, dt, eps, sig, rc) >= 4
    assert ctx.device_allocations() == a0
