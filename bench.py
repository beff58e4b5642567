#!/usr/bin/env python
"""bench.py -- atom-timesteps/s of the LJ hot path (binning + chunk neighbour build + ghost update + force + integrate).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload auto|C1|C2|C3|C4|C5]

One "step" = one MD time step (numerical_scheme of data/config/numerical-scheme.msp:21-25) of the named workload, neighbour
rebuilds included whenever the displacement trigger fires.  N=1 runs C2 (2 048 000 atoms); N>1 runs C3 (4 000 000 atoms per
GPU, spatial decomposition, halo over NVLink peer memory) under torchrun.  Prints ONE JSON line (rank 0).

  value     device-resident rate: atoms(all ranks) * K / max-over-ranks CUDA-event time of the K steps
  e2e       the same steps driven through the C-ABI with HOST buffers (one xnb_step_host call per step): r,v go up from pinned
            host memory, the step runs, r,v,f come back (ids too on the steps that rebuild) -- copies inside the timed region;
            with several ranks xnb_step_host_n on every rank (atoms migrate at rebuilds: the call returns the rank's new particle count)
  roofline  the dominant kernel (pair sweep k_lj_sweep_cl): algorithmic bytes per launch / its mean CUDA-event duration
  cpu_baseline  the CPU oracle (oracle/, the OpenMP restatement of the reference) on a bounded sample of the same workload

--impl reference times the reference's CPU implementation of the path.  The reference itself cannot be built here (it needs
onika, yaml-cpp and MPI, none of which is in the image: DESIGN.md), so that arm is the oracle port with every host thread.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

A_FCC = (4.0 / 0.8442) ** (1.0 / 3.0)      # rho* = 0.8442


def workload(name, nranks=1):
    """SURVEY.md 8(d) restated as concrete inputs (reduced LJ units except C1)."""
    def lj(units, cell_units=2, rcut=2.5, skin=0.3, noise=0.02, vel_sigma=1.2, **kw):
        units = tuple(units) if isinstance(units, (tuple, list)) else (units,) * 3
        d = dict(bounds_max=tuple(A_FCC * u for u in units), cell_size=A_FCC * cell_units, grid_dims=tuple(u // cell_units for u in units),
                 lattice_a=A_FCC, epsilon=1.0, sigma=1.0, rcut=rcut, rcut_inc=skin, dt=0.005, mass=1.0, noise_sigma=noise,
                 vel_sigma=vel_sigma, max_neighbors=1024)
        d.update(kw)
        return d
    if name == "C1":
        ev = 1.6021892e-19 / (1.66053904e-27 * 1e-20 / 1e-24)
        return dict(bounds_max=(139.2,) * 3, cell_size=13.92, grid_dims=(10,) * 3, lattice_a=3.48, epsilon=0.3729 * ev, sigma=2.2808,
                    rcut=2.5 * 2.2808, rcut_inc=2.0, dt=2e-3, mass=58.693, noise_sigma=0.05, vel_sigma=0.0, max_neighbors=1024), \
            "C1: Ni LJ FCC 40^3 unit cells (256000 atoms), rc=2.5 sigma, skin 2.0 A"
    if name == "C2":
        return lj(80), "C2: LJ FCC 80^3 unit cells = 2048000 atoms fp64, rc=2.5, skin=0.3, dt=0.005, T*=1.44 (binning + chunk neighbour build + force + integrate)"
    if name == "C3":
        mult = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[nranks]
        return lj(tuple(100 * m for m in mult)), "C3: LJ FCC weak scaling, 100^3 unit cells = 4000000 atoms per GPU, rc=2.5, skin=0.3, RCB blocks %dx%dx%d, NCCL halo" % mult
    if name == "C4":
        return lj(48, cell_units=4, rcut=5.0, noise=0.1), "C4: dense LJ rc=5 sigma, 48^3 unit cells = 442368 atoms, cell = 4a"
    if name == "C5":
        return lj(80, n_spheres=64, sphere_rmin=6.0, sphere_rmax=14.0, drift_speed=2.0), "C5: clusters + voids in an 80^3 unit cell box"
    raise SystemExit("unknown workload " + name)


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe): NVML polled every 5 ms from a
    thread (the timed C call releases the GIL); nvidia-smi -lms as the fallback when NVML cannot be loaded."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None
        self.thread = None
        self.stop_flag = False
        self.sm, self.mx, self.reasons = [], [], set()

    def _poll(self, nv, h):
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.gpu
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.thread = threading.Thread(target=self._poll, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if self.sm:
                out.update(sm_mhz=float(np.median(self.sm)), sm_max_mhz=float(max(self.mx)), reasons=sorted(self.reasons), samples=len(self.sm), source="nvml")
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm), source="nvidia-smi")
        return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel, workload_name="C2"):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture OF THE SAME WORKLOAD (null when there is none)"""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            e = json.load(open(p)).get(kernel)
            if e and e.get("workload", "C2") != workload_name:
                return None
            return float(e["dram_bytes_per_launch"]) if e else None      # bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)
        except Exception:
            return None
    return None


# ---------------------------------------------------------------------------------------------------------------------
def make_oracle_config(kw):
    from oracle import oracle as O
    return O.make_config(**kw)


def run_cpu_oracle(kw, nsteps, budget_s, keep=False):
    """the CPU oracle on the same workload: init untimed, then up to nsteps steps (stops early when budget_s is spent).
    keep: return the oracle (state after the steps) for the parity check instead of closing it."""
    from oracle import oracle as O
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm uses every host core the container exposes
    if "XNB_KEEP_OMP" not in os.environ:
        O.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    o = O.Oracle(make_oracle_config(kw))
    o.generate(); o.first_iteration()
    n = o.n_inner()
    done, rebuilds, t0 = 0, 0, time.perf_counter()
    chunk = 1
    while done < nsteps:
        k = min(chunk, nsteps - done)
        rebuilds += o.run(k); done += k
        el = time.perf_counter() - t0
        if el > budget_s:
            break
    el = time.perf_counter() - t0
    res = dict(atoms=n, steps=done, seconds=el, rebuilds=rebuilds, rate=n * done / el, threads=O.num_threads())
    if keep:
        res["oracle"] = o
    else:
        o.close()
    return res


def parity_against_oracle(o, steps, rebuilds_oracle, kw, device):
    """untimed tail: the SAME input stepped the same number of steps by the CUDA path (fresh context) and by the CPU oracle `o`
    (already stepped): atoms, rebuild count, positions, forces (|df| <= 1e-10 max(|f|, f_rms) is the single-step bar; after
    `steps` chaotic steps the bound reported is what it is), and the neighbour streams of the final configuration byte for byte
    (the oracle rebuilds its lists on the GPU's in-cell order first, as tests/test_gpu_parity.py does)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity_util as U
    eps, sig, rc, dt = kw["epsilon"], kw["sigma"], kw["rcut"], kw["dt"]
    ctx = U.make_ctx(kw, device=device)
    ctx.first_iteration(eps, sig, rc)
    rb = ctx.run_steps(steps, dt, eps, sig, rc)
    po = U.by_id(o.particles(), o.inner_mask()); pg = U.by_id(ctx.get_particles(0, ctx.n_inner))
    out = {"steps": steps, "atoms_equal": bool(np.array_equal(po["id"], pg["id"])), "rebuilds": [int(rb), int(rebuilds_oracle)],
           "rebuilds_equal": int(rb) == int(rebuilds_oracle)}
    if out["atoms_equal"]:
        L = np.array(kw["bounds_max"])
        dr = U.vec(pg, ("rx", "ry", "rz")) - U.vec(po, ("rx", "ry", "rz")); dr -= L * np.round(dr / L)
        out["max_position_error_over_cell"] = float(np.abs(dr).max() / kw["cell_size"])
        out["max_force_error"] = float(U.force_error(U.vec(pg, ("fx", "fy", "fz")), U.vec(po, ("fx", "fy", "fz"))))
        # single-step force parity on IDENTICAL positions: hand the GPU's state to the oracle and let it recompute lists and forces
        pcell, cnt = U.gpu_particles_cell_order(ctx)
        o.set_particles(cnt, pcell); o.build_neighbors(); o.compute_force()
        ctx.chunk_neighbors()               # lists of the FINAL configuration on both sides (the step loop's lists date from its last rebuild)
        sz_o, data_o = o.streams(); sz_g, data_g = ctx.streams()
        out["streams_equal"] = bool(np.array_equal(sz_o, sz_g) and np.array_equal(data_o, data_g))
        out["stream_words"] = int(np.asarray(sz_g, np.int64).sum())
        ctx.zero_particle_force(True); ctx.lennard_jones_force(eps, sig, rc); ctx.divide_force_by_mass()
        po2 = U.by_id(o.particles(), o.inner_mask()); pg2 = U.by_id(ctx.get_particles(0, ctx.n_inner))
        out["max_force_error_same_positions"] = float(U.force_error(U.vec(pg2, ("fx", "fy", "fz")), U.vec(po2, ("fx", "fy", "fz"))))
        out["ok"] = bool(out["rebuilds_equal"] and out["streams_equal"] and out["max_force_error_same_positions"] < 1e-10)
    else:
        out["ok"] = False
    ctx.close()
    return out


def parity_nranks(rank, world, local, dist):
    """untimed, before the timed region of an N > 1 run: a small lattice case (12x8x8 cells, 24576 atoms, 20 steps) stepped by the N
    ranks (spatial decomposition, NCCL halo, migration) and by the CPU ORACLE on rank 0: same atoms, same rebuild count, positions and
    forces within tolerance.  (tests/test_gpu_multi.py compares N ranks with 1 rank of the CUDA path; this is the independent checker.)"""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity_util as U
    from conftest import lj_reduced_kwargs
    kw = dict(lj_reduced_kwargs(ncell_units=8, cell_units=2), bounds_max=tuple(A_FCC * n for n in (24, 16, 16)), grid_dims=(12, 8, 8))
    eps, sig, rc, dt = kw["epsilon"], kw["sigma"], kw["rcut"], kw["dt"]
    steps = 20
    inp = U.generate_input(kw)
    ctx = U.make_ctx(kw, rank=rank, nranks=world, device=local, particles=inp)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.from_numpy(ctx.nccl_unique_id().copy())
    uid = uid.cuda(); dist.broadcast(uid, 0); ctx.nccl_init_rank(uid.cpu().numpy(), rank, world)
    ctx.first_iteration(eps, sig, rc)
    rb = ctx.run_steps(steps, dt, eps, sig, rc)
    mine = ctx.get_particles(0, ctx.n_inner)
    everyone = [None] * world
    dist.all_gather_object(everyone, {k: mine[k] for k in ("id", "rx", "ry", "rz", "fx", "fy", "fz")})
    ctx.close()
    out = None
    if rank == 0:
        o = U.make_oracle(kw)
        o.generate(); o.first_iteration()
        rb_o = o.run(steps)
        po = U.by_id(o.particles(), o.inner_mask())
        pg = U.by_id({k: np.concatenate([e[k] for e in everyone]) for k in everyone[0]})
        out = {"case": "LJ 24x16x16 unit cells, %d atoms, %d steps" % (len(po["id"]), steps), "ranks": world, "atoms_per_rank": [int(len(e["id"])) for e in everyone],
               "atoms_equal": bool(np.array_equal(po["id"], pg["id"])), "rebuilds": [int(rb), int(rb_o)]}
        if out["atoms_equal"]:
            L = np.array(kw["bounds_max"])
            dr = U.vec(pg, ("rx", "ry", "rz")) - U.vec(po, ("rx", "ry", "rz")); dr -= L * np.round(dr / L)
            out["max_position_error_over_cell"] = float(np.abs(dr).max() / kw["cell_size"])
            out["max_force_error"] = float(U.force_error(U.vec(pg, ("fx", "fy", "fz")), U.vec(po, ("fx", "fy", "fz"))))
            out["ok"] = bool(rb == rb_o and out["max_position_error_over_cell"] < 1e-10 and out["max_force_error"] < 1e-8)
        else:
            out["ok"] = False
        o.close()
    dist.barrier()
    return out


def weak_base_run(local, steps, warmup):
    """rank 0 alone, after the timed region of an N > 1 run: the SAME per-GPU workload (C3 block, 100^3 unit cells = 4 M atoms) on ONE GPU
    (periodic self-images instead of NCCL partners): the like-for-like denominator of the weak-scaling efficiency"""
    import torch
    from exanbody_b200 import capi
    kw, desc = workload("C3", 1)
    eps, sig, rc, dt = kw["epsilon"], kw["sigma"], kw["rcut"], kw["dt"]
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity_util as U
    ctx = U.make_ctx(kw, device=local)
    sh = torch.cuda.current_stream().cuda_stream
    ctx.first_iteration(eps, sig, rc, sh)
    ctx.run_steps(warmup, dt, eps, sig, rc, sh)
    ctx.timing_enable(True); ctx.timing_read(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(); rb = ctx.run_steps(steps, dt, eps, sig, rc, sh); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tim = ctx.timing_read(reset=True); ctx.timing_enable(False)
    n = ctx.n_inner
    ctx.close()
    return {"workload": desc, "atoms": int(n), "steps": steps, "rebuilds": int(rb), "value": n * steps / (ms * 1e-3), "ms_per_step": ms / steps,
            "breakdown_ms_per_step": {k: v["ms"] / steps for k, v in tim.items()}}


def short_run(name, device, steps=12, warmup=3):
    """a short single-GPU run of another BASELINE workload (C1, C4, C5) after the timed region: value, rebuilds, kernel breakdown"""
    import torch
    from exanbody_b200 import capi
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity_util as U
    kw, desc = workload(name)
    eps, sig, rc, dt = kw["epsilon"], kw["sigma"], kw["rcut"], kw["dt"]
    ctx = U.make_ctx(kw, device=device)
    sh = torch.cuda.current_stream().cuda_stream
    ctx.first_iteration(eps, sig, rc, sh)
    ctx.run_steps(warmup, dt, eps, sig, rc, sh)
    ctx.timing_enable(True); ctx.timing_read(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(); rb = ctx.run_steps(steps, dt, eps, sig, rc, sh); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tim = ctx.timing_read(reset=True); ctx.timing_enable(False)
    n = ctx.n_inner
    si = ctx.sweep_info()
    sz = ctx.stream_sizes()
    gi = ctx.grid_info(); d, gl = gi["dims"], gi["ghost_layers"]
    k, j, i = np.meshgrid(np.arange(d[2]), np.arange(d[1]), np.arange(d[0]), indexing="ij")
    inner = ((i >= gl) & (i < d[0] - gl) & (j >= gl) & (j < d[1] - gl) & (k >= gl) & (k < d[2] - gl)).ravel()
    S = 2.0 * float(sz[inner].sum()) / max(n, 1)
    peak, _ = measured_peaks()
    value = n * steps / (ms * 1e-3)
    fk = tim["force"]["ms"] / max(tim["force"]["n"], 1)
    nb = tim["nbh"]["ms"] / max(tim["nbh"]["n"], 1)
    out = {"workload": desc, "atoms": int(n), "steps": steps, "rebuilds": int(rb), "value": value, "ms_per_step": ms / steps,
           "sweep_kernel_ms": fk, "nbh_build_ms": nb, "compiled_lists": bool(si["compiled"]), "tile_cells": list(si["tile"]),
           "list_entries_per_atom": si["candidates"] / max(n, 1) if si["compiled"] else None,
           "whole_step_hbm_frac": (253.0 + S) * value / 1e9 / peak}
    ctx.close()
    if si["compiled"] and fk > 0:
        # F_alg = 8 N_list + 19 N_cut (+ 30 integrate) flop per atom-step, N_cut from N_list by the volume ratio; against the
        # DFMA rate measured now (C4 is the FP64-bound configuration: rc = 5 sigma, 530 list entries per atom)
        try:
            dfma = capi.measure_dfma_peak(device)
            n_list = si["candidates"] / max(n, 1)
            f_sweep = 8.0 * n_list + 19.0 * n_list * (rc / (rc + kw["rcut_inc"])) ** 3
            out["sweep_fp64_frac"] = f_sweep * n / (fk * 1e-3) / 1e12 / dfma
            out["whole_step_fp64_frac"] = (f_sweep + 30.0) * value / 1e12 / dfma
        except Exception:
            pass
    return out



def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name = args.workload if args.workload != "auto" else ("C2" if args.gpus == 1 else "C3")
    kw, desc = workload(name, args.gpus)
    # warmup steps are part of the bounded budget; the oracle is deterministic so they only warm caches
    budget = float(os.environ.get("XNB_REF_BUDGET_S", "150"))
    r = run_cpu_oracle(kw, args.steps, budget)
    line = {
        "impl": "reference", "metric": "atom-timesteps/s (LJ neighbor+force+ghost)", "value": r["rate"], "unit": "atom-timesteps/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": 0, "ms_per_step": 1e3 * r["seconds"] / max(r["steps"], 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "atoms": r["atoms"], "rebuilds": r["rebuilds"], "l2": "inputs larger than L2"},
        "cpu_baseline": {"value": r["rate"], "unit": "atom-timesteps/s", "cores": r["threads"], "kind": "port",
                         "sample": "CPU oracle (OpenMP restatement of the reference; the reference needs onika/yaml-cpp/MPI and cannot be built), "
                                   "%d steps of the full workload after untimed init, %d rebuilds" % (r["steps"], r["rebuilds"])},
        "e2e": {"value": r["rate"], "unit": "atom-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------------------
def b200_arm(args):
    import torch
    from exanbody_b200 import capi
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    name = args.workload if args.workload != "auto" else ("C2" if world == 1 else "C3")
    kw, desc = workload(name, world)
    eps, sig, rc, dt = kw["epsilon"], kw["sigma"], kw["rcut"], kw["dt"]
    par_n = None
    if world > 1 and not args.no_parity:
        try:
            par_n = parity_nranks(rank, world, local, dist)
        except Exception as ex:
            par_n = {"ok": False, "error": repr(ex)[:300]}

    ctx = capi.Context(local)
    ctx.set_domain((0., 0., 0.), kw["bounds_max"], kw["cell_size"], kw["grid_dims"], (1, 1, 1))
    ctx.init_rcb_grid(rank, world)
    ctx.set_nbh_dist(rc, kw["rcut_inc"])
    ctx.set_type_mass([kw["mass"]])
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.from_numpy(ctx.nccl_unique_id().copy())
        uid = uid.cuda(); dist.broadcast(uid, 0); ctx.nccl_init_rank(uid.cpu().numpy(), rank, world)
    gi = ctx.grid_info()
    inp = capi.lattice_fcc(kw["bounds_max"], kw["cell_size"], kw["grid_dims"], kw["lattice_a"], noise_sigma=kw["noise_sigma"], vel_sigma=kw["vel_sigma"],
                           n_spheres=kw.get("n_spheres", 0), sphere_rmin=kw.get("sphere_rmin", 0.), sphere_rmax=kw.get("sphere_rmax", 0.),
                           drift_speed=kw.get("drift_speed", 0.))      # whole domain; xnb_set_particles keeps this rank's block
    ctx.set_particles(inp["rx"], inp["ry"], inp["rz"], inp["vx"], inp["vy"], inp["vz"], inp["id"], inp["type"])
    del inp
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ctx.first_iteration(eps, sig, rc, sh)
    n_atoms_local = ctx.n_inner
    n_atoms = int(round(sum_over_ranks(float(n_atoms_local))))

    # ---- device-resident rate ---------------------------------------------------------------------------------------
    ctx.run_steps(args.warmup, dt, eps, sig, rc, sh)
    ctx.timing_enable(True); ctx.timing_read(reset=True)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    l0 = ctx.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    rebuilds = ctx.run_steps(args.steps, dt, eps, sig, rc, sh)
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    launches = ctx.kernel_launches() - l0
    tim = ctx.timing_read(reset=True)
    ctx.timing_enable(False)
    n_atoms_after = int(round(sum_over_ranks(float(ctx.n_inner))))
    assert n_atoms_after == n_atoms, "atoms lost: %d -> %d" % (n_atoms, n_atoms_after)
    value = n_atoms * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (pair sweep) -----------------------------------------------------------------
    sz = ctx.stream_sizes()
    d, gl = gi["dims"], gi["ghost_layers"]
    k, j, i = np.meshgrid(np.arange(d[2]), np.arange(d[1]), np.arange(d[0]), indexing="ij")
    inner = ((i >= gl) & (i < d[0] - gl) & (j >= gl) & (j < d[1] - gl) & (k >= gl) & (k < d[2] - gl)).ravel()
    S = 2.0 * float(sz[inner].sum()) / max(n_atoms_local, 1)          # stream bytes per inner atom (reference u16 format)
    force_bytes = 97.0 + S                                            # R r 24 + R stream S + W f 24 + R v 24 + R type 1 + W v 24
    # inside xnb_run_steps the sweep of step k also performs the first half of step k + 1 (k_lj_sweep_cl MODE 2): + W r' 24 + R backup 12
    # on those launches (all steps but the last of the call; the stand-alone first-half kernel then does not run)
    fused_frac = max(0.0, 1.0 - tim["first_half"]["n"] / float(max(args.steps, 1)))
    force_bytes += 36.0 * fused_frac
    step_bytes = 253.0 + S                                            # minimal fused traffic of a steady step
    peak, peak_src = measured_peaks()
    fk_ms = tim["force"]["ms"] / max(tim["force"]["n"], 1)
    achieved = force_bytes * n_atoms_local / (fk_ms * 1e-3) / 1e9 if fk_ms > 0 else 0.0
    si = ctx.sweep_info()
    kname = "k_lj_sweep_cl" if si["compiled"] else "k_lj_sweep"
    n_list = si["candidates"] / max(n_atoms_local, 1) if si["compiled"] else None        # list entries per atom (78 on the perfect lattice)
    roofline = {"bound": "hbm", "kernel": kname + " (pair sweep + fused second half kick)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(kname, name), "traffic_unit": "DRAM bytes per launch (ncu --set full, profiles/ncu_traffic.json)",
                "algorithmic_bytes_per_launch": force_bytes * n_atoms_local, "peak_source": peak_src,
                "algorithmic_bytes_per_atom": force_bytes, "stream_bytes_per_atom": S, "kernel_ms": fk_ms, "fused_next_first_half_fraction": fused_frac,
                "whole_step": {"algorithmic_bytes_per_atom": step_bytes, "achieved": step_bytes * value / world / 1e9, "frac": step_bytes * value / world / 1e9 / peak},
                "sweep": {"compiled_lists": si["compiled"], "tile_cells": si["tile"], "threads": si["threads"], "blocks": si["blocks"], "smem_bytes": si["smem_bytes"],
                          "list_entries_per_atom": n_list, "compiled_list_bytes_per_atom": (si["rows"] * 256.0 / max(n_atoms_local, 1)) if si["compiled"] else None}}
    breakdown = {k2: (v["ms"] / args.steps) for k2, v in tim.items()}
    dom = max(breakdown, key=lambda k2: breakdown[k2]) if breakdown else None
    roofline["dominant_by_time"] = {"group": dom, "ms_per_step": breakdown.get(dom), "share_of_step": (breakdown.get(dom, 0.0) / (ms / args.steps)) if ms > 0 else None,
                                    "groups": "force = pair sweep (+ second half kick, + the next step's first half when fused), nbh = chunk_neighbors (k_nbh_bits), first_half = verlet_first_half + displacement test, "
                                              "bin = move_particles + rebuild_amr + backup_r, ghost_scheme / ghost_update = halo"}

    # ---- end to end through the C-ABI with host buffers -----------------------------------------------------------------
    n = ctx.n_inner
    hb = {k2: torch.empty(n, dtype=torch.float64).pin_memory() for k2 in ("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz")}
    hid = torch.empty(n, dtype=torch.int64).pin_memory()
    ptr = {k2: v.data_ptr() for k2, v in hb.items()}
    ctx.download_rvf(ptr["rx"], ptr["ry"], ptr["rz"], ptr["vx"], ptr["vy"], ptr["vz"], ptr["fx"], ptr["fy"], ptr["fz"], hid.data_ptr(), sh)

    e2e_rebuilds = [0]

    def e2e_step():
        # particles live in host memory (as the reference's Grid does): one C-ABI call uploads r,v, runs the step and returns r,v,f
        # (positions travel back while the sweep runs; ids only when the step rebuilt, i.e. when the particle order changed)
        e2e_rebuilds[0] += ctx.step_host(dt, eps, sig, rc, in_r=(ptr["rx"], ptr["ry"], ptr["rz"]), in_v=(ptr["vx"], ptr["vy"], ptr["vz"]),
                                         out_r=(ptr["rx"], ptr["ry"], ptr["rz"]), out_v=(ptr["vx"], ptr["vy"], ptr["vz"]),
                                         out_f=(ptr["fx"], ptr["fy"], ptr["fz"]), out_id=hid.data_ptr(), stream=sh)
        assert ctx.n_inner == n

    e2e = None
    if world > 1 and not args.no_e2e:
        # several ranks: xnb_step_host_n -- atoms migrate at a rebuild, so the host arrays have room to spare and the call says how many
        # particles the rank owns afterwards; every step still uploads r, v and brings back r, v, f (+ ids when the step rebuilt)
        cap = int(n * 1.25) + 4096
        hbn = {k2: torch.empty(cap, dtype=torch.float64).pin_memory() for k2 in ("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz")}
        hidn = torch.empty(cap, dtype=torch.int64).pin_memory()
        pn = {k2: v.data_ptr() for k2, v in hbn.items()}
        ctx.download_rvf(pn["rx"], pn["ry"], pn["rz"], pn["vx"], pn["vy"], pn["vz"], pn["fx"], pn["fy"], pn["fz"], hidn.data_ptr(), sh)
        state = {"n": int(n), "rebuilds": 0, "h2d": 0, "d2h": 0}

        def e2e_step_n():
            n_in = state["n"]
            rbs, n_new = ctx.step_host_n(dt, eps, sig, rc, cap, in_r=(pn["rx"], pn["ry"], pn["rz"]), in_v=(pn["vx"], pn["vy"], pn["vz"]),
                                         out_r=(pn["rx"], pn["ry"], pn["rz"]), out_v=(pn["vx"], pn["vy"], pn["vz"]),
                                         out_f=(pn["fx"], pn["fy"], pn["fz"]), out_id=hidn.data_ptr(), stream=sh)
            state["n"] = n_new; state["rebuilds"] += rbs; state["h2d"] += 48 * n_in; state["d2h"] += 72 * n_new + 8 * n_new * rbs

        e2e_steps = max(3, min(args.steps, 20))
        for _ in range(3):
            e2e_step_n()
        barrier()
        state.update(rebuilds=0, h2d=0, d2h=0)
        e0.record(stream)
        for _ in range(e2e_steps):
            e2e_step_n()
        e1.record(stream)
        barrier()
        ems = max_over_ranks(e0.elapsed_time(e1))
        h2d_all = sum_over_ranks(float(state["h2d"])); d2h_all = sum_over_ranks(float(state["d2h"]))
        e2e = {"value": n_atoms * e2e_steps / (ems * 1e-3), "unit": "atom-timesteps/s", "h2d_bytes_per_step": h2d_all / e2e_steps,
               "d2h_bytes_per_step": d2h_all / e2e_steps, "steps": e2e_steps, "rebuilds": state["rebuilds"], "ms_per_step": ems / e2e_steps,
               "what": "per step one xnb_step_host_n call on every rank: r,v of the rank's atoms from pinned host memory -> one step (halo over peer "
                       "memory, migration at rebuilds) -> r,v,f back (ids too on the steps that rebuild); bytes summed over the ranks, time = max over ranks"}
    if world == 1 and not args.no_e2e:       # (N = 1: xnb_step_host, the arrays keep their size)
        e2e_steps = max(3, min(args.steps, 20))
        for _ in range(3):
            e2e_step()
        barrier()
        e2e_rebuilds[0] = 0
        e0.record(stream)
        for _ in range(e2e_steps):
            e2e_step()
        e1.record(stream)
        barrier()
        ems = e0.elapsed_time(e1)
        e2e = {"value": n_atoms * e2e_steps / (ems * 1e-3), "unit": "atom-timesteps/s", "h2d_bytes_per_step": 48 * n,
               "d2h_bytes_per_step": 72 * n + 8 * n * e2e_rebuilds[0] / e2e_steps, "steps": e2e_steps, "rebuilds": e2e_rebuilds[0], "ms_per_step": ems / e2e_steps,
               "what": "per step one xnb_step_host call: r,v from pinned host memory -> one step -> r,v,f back (ids too on the steps that rebuild: "
                       "the particle order changes only there); d2h_bytes_per_step is the mean over the timed steps"}

    # ---- CPU baseline (bounded sample, rank 0, N=1 only) ----------------------------------------------------------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = run_cpu_oracle(kw, 20, float(os.environ.get("XNB_CPU_BUDGET_S", "25")), keep=not args.no_parity)
        cpu = {"value": r["rate"], "unit": "atom-timesteps/s", "cores": r["threads"], "kind": "port",
               "sample": "CPU oracle (OpenMP port of the reference path), same %d-atom input, %d steps (%d rebuilds) after untimed init" % (r["atoms"], r["steps"], r["rebuilds"])}
        if "oracle" in r:
            # parity where the number is quoted: same input, same number of steps, CUDA path against the oracle (untimed)
            try:
                parity = parity_against_oracle(r["oracle"], r["steps"], r["rebuilds"], kw, local)
            except Exception as ex:                                        # never lose the bench line over the checker
                parity = {"ok": False, "error": repr(ex)[:300]}
            r["oracle"].close()
    extra = None
    if rank == 0 and world == 1 and not args.no_extra:
        extra = {}
        for wl in ("C1", "C4", "C5"):
            if wl == name:
                continue
            try:
                extra[wl] = short_run(wl, local)
            except Exception as ex:
                extra[wl] = {"error": repr(ex)[:300]}

    # ---- dynamic re-partition, on request: load_balance_rcb on the live contexts, then the same K steps again
    lbal = None
    if world > 1 and args.load_balance:
        atoms_before = [None] * world
        dist.all_gather_object(atoms_before, int(ctx.n_inner))
        b, a = ctx.load_balance_rcb(stream=sh)
        ctx.update_particles_full(sh)
        ctx.run_steps(args.warmup, dt, eps, sig, rc, sh)
        barrier()
        e0.record(stream); rb2 = ctx.run_steps(args.steps, dt, eps, sig, rc, sh); e1.record(stream)
        barrier()
        ms2 = max_over_ranks(e0.elapsed_time(e1))
        atoms_after = [None] * world
        dist.all_gather_object(atoms_after, int(ctx.n_inner))
        lbal = {"lb_inbalance": [b, a], "ms_per_step_before": ms / args.steps, "ms_per_step_after": ms2 / args.steps, "rebuilds_after": int(rb2),
                "atoms_per_rank_before": atoms_before, "atoms_per_rank_after": atoms_after, "value_after": sum(atoms_after) * args.steps / (ms2 * 1e-3)}
    weak_base = None
    if world > 1 and name == "C3" and not args.no_extra:
        if rank == 0:
            try:
                weak_base = weak_base_run(local, args.steps, args.warmup)
                weak_base["efficiency_vs_weak_base"] = value / (world * weak_base["value"])
            except Exception as ex:
                weak_base = {"error": repr(ex)[:300]}
        barrier()
    if rank == 0:
        try:
            dfma = capi.measure_dfma_peak(local)
        except Exception:
            dfma = None
        if dfma and n_list:
            # F_alg = 8 N_list + 19 N_cut + 30 flop per atom-step; N_cut estimated from N_list by the volume ratio
            # (rc / (rc + skin))^3 (54 of 78 on the perfect lattice).  The slower of the two ceilings bounds the path.
            n_cut = n_list * (rc / (rc + kw["rcut_inc"])) ** 3
            f_sweep = 8.0 * n_list + 19.0 * n_cut
            roofline["fp64"] = {"peak_tflops": dfma, "peak_source": "k_dfma_probe measured in this run", "flop_per_atom_sweep": f_sweep,
                                "achieved_tflops": f_sweep * n_atoms_local / (fk_ms * 1e-3) / 1e12 if fk_ms > 0 else 0.0,
                                "frac": (f_sweep * n_atoms_local / (fk_ms * 1e-3) / 1e12 / dfma) if fk_ms > 0 else 0.0,
                                "whole_step_ceiling_atom_steps_per_s": {"hbm": peak * 1e9 / step_bytes, "fp64": dfma * 1e12 / (f_sweep + 30.0)}}
        line = {
            "metric": "atom-timesteps/s (LJ neighbor+force+ghost)", "value": value, "unit": "atom-timesteps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "atoms": n_atoms, "rebuilds": rebuilds, "l2": "inputs larger than L2 (state + neighbour streams >> 126 MB), no flush",
                       "timing": "CUDA events on the launching stream, max over ranks",
                       "ghost_transport": (ctx.ghost_transport() if world > 1 else None)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "breakdown_ms_per_step": breakdown, "fp64_dfma_peak_tflops": dfma, "parity": parity, "extra_workloads": extra, "parity_nranks": par_n, "weak_base": weak_base, "load_balance": lbal,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed comparison of the CUDA path with the CPU oracle on the benchmark input")
    ap.add_argument("--load-balance", action="store_true", help="N > 1: after the timed region re-partition with load_balance_rcb and time the same steps again")
    ap.add_argument("--no-extra", action="store_true", help="skip the short runs of the other BASELINE workloads (C1, C4, C5)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)
    return b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
