#!/usr/bin/env python
"""Newton-3 path against the full-list path on one workload of bench.py (default C2), operator by operator, CUDA-event times.
  python scripts/bench_newton3.py [C2|C4|...]       -> one JSON line (profiles/*_newton3_*.json)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                      # noqa: E402
import bench                      # noqa: E402
from exanbody_b200 import capi    # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C2"
    kw, desc = bench.workload(name, 1)
    eps, sig, rc, dt = kw["epsilon"], kw["sigma"], kw["rcut"], kw["dt"]
    torch.cuda.set_device(0)
    ctx = capi.Context(0)
    ctx.set_domain((0., 0., 0.), kw["bounds_max"], kw["cell_size"], kw["grid_dims"], (1, 1, 1))
    ctx.init_rcb_grid(0, 1); ctx.set_nbh_dist(rc, kw["rcut_inc"]); ctx.set_type_mass([kw["mass"]])
    inp = capi.lattice_fcc(kw["bounds_max"], kw["cell_size"], kw["grid_dims"], kw["lattice_a"], noise_sigma=kw["noise_sigma"], vel_sigma=kw["vel_sigma"])
    ctx.set_particles(inp["rx"], inp["ry"], inp["rz"], inp["vx"], inp["vy"], inp["vz"], inp["id"], inp["type"])
    st = torch.cuda.current_stream(); sh = st.cuda_stream
    ctx.first_iteration(eps, sig, rc, sh)
    ctx.run_steps(10, dt, eps, sig, rc, sh)
    n = ctx.n_inner

    def timed(fn, reps):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            fn()
        e1.record(st); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    out = {"workload": desc, "atoms": n}
    ctx.set_chunk_neighbors_config(); out["full_build_ms"] = timed(lambda: ctx.chunk_neighbors(sh), 3)
    out["full_stream_u16_per_atom"] = float(ctx.stream_sizes().sum()) / n

    def full():
        ctx.zero_particle_force(True, sh); ctx.lennard_jones_force(eps, sig, rc, False, sh); ctx.divide_force_by_mass(sh)
    out["full_force_ms"] = timed(full, 10)
    f_full = ctx.get_particles(0, n, fields=("fx", "fy", "fz"))
    ctx.set_chunk_neighbors_config(half_symmetric=True); out["half_build_ms"] = timed(lambda: ctx.chunk_neighbors(sh), 3)
    out["half_stream_u16_per_atom"] = float(ctx.stream_sizes().sum()) / n

    def sym():
        ctx.zero_particle_force(True, sh); ctx.lennard_jones_force_symmetric(eps, sig, rc, sh); ctx.update_force_from_ghost(sh); ctx.divide_force_by_mass(sh)
    out["newton3_force_ms"] = timed(sym, 10)
    f_sym = ctx.get_particles(0, n, fields=("fx", "fy", "fz"))
    out["newton3_sweep_only_ms"] = timed(lambda: ctx.lennard_jones_force_symmetric(eps, sig, rc, sh), 10)      # (accumulates: timing only)
    import numpy as np
    d = np.sqrt(sum((f_sym[k] - f_full[k]) ** 2 for k in ("fx", "fy", "fz"))); nrm = np.sqrt(sum(f_full[k] ** 2 for k in ("fx", "fy", "fz")))
    out["max_rel_force_diff"] = float((d / np.maximum(nrm, np.sqrt((nrm ** 2).mean()))).max())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
