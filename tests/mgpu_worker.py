"""Worker of tests/test_gpu_multi.py, launched by torch.distributed.run with one rank per GPU.
Every rank owns one block of the spatial decomposition (init_rcb_grid) and runs the LJ loop with the NCCL halo
(ghost_comm_scheme / ghost_update_all / ghost_update_r / migrate hand-off); rank 0 also runs the same input on a single
rank and compares: identical atoms, rebuild count, and bit-identical r, v, f (the in-cell order is by particle id, so the
neighbour streams and the summation order do not depend on the decomposition)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from conftest import lj_reduced_kwargs          # noqa: E402
import parity_util as U                          # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    nsteps = int(os.environ.get("XNB_MGPU_STEPS", "40"))
    cases = {
        "lj_12x8x8": dict(lj_reduced_kwargs(ncell_units=8, cell_units=2), bounds_max=tuple(((4.0 / 0.8442) ** (1 / 3.)) * n for n in (24, 16, 16)), grid_dims=(12, 8, 8)),
        "lj_voids": lj_reduced_kwargs(ncell_units=16, cell_units=2, n_spheres=6, sphere_rmin=3.0, sphere_rmax=6.0, drift_speed=1.5),
    }
    # same lattice case again with small sweep tiles: every rank then has interior tiles, i.e. the halo exchange of
    # xnb_run_steps runs on its own stream while they are swept (asserted below through xnb_get_sweep_info)
    cases["lj_12x8x8_overlap"] = cases["lj_12x8x8"]
    for name, kw in cases.items():
        if name.endswith("_overlap"):
            os.environ["XNB_CL_TILE"] = "2,1,1"
        else:
            os.environ.pop("XNB_CL_TILE", None)
        inp = U.generate_input(kw)
        ctx = U.make_ctx(kw, rank=rank, nranks=world, device=local, particles=inp)
        uid = [ctx.nccl_unique_id().copy() if rank == 0 else None]
        dist.broadcast_object_list(uid, 0)
        ctx.nccl_init_rank(uid[0], rank, world)
        eps, sig, rc, dt = kw["epsilon"], kw["sigma"], kw["rcut"], kw["dt"]
        ctx.first_iteration(eps, sig, rc)
        rb = ctx.run_steps(nsteps, dt, eps, sig, rc)
        transport = ctx.ghost_transport()
        expect = os.environ.get("XNB_EXPECT_TRANSPORT")
        assert expect is None or transport == expect, "halo transport is %s, expected %s" % (transport, expect)
        lb = None
        if name == "lj_voids":
            # load_balance_rcb on the live contexts (device cost model -> all-reduce -> cost-weighted RCB -> new blocks
            # -> migration to the new owners), then the rebuild chain, then more steps.  Clusters + voids: the static blocks are badly
            # balanced on purpose.  The single-rank reference below performs the same rebuild at the same step.
            n_before = ctx.n_inner
            lb = ctx.load_balance_rcb()
            ctx.update_particles_full()
            blocks = [tuple(map(tuple, ctx.block(r))) for r in range(world)]
            counts = [None] * world
            dist.all_gather_object(counts, (int(n_before), int(ctx.n_inner)))
            rb += ctx.run_steps(nsteps // 2, dt, eps, sig, rc)
        if name.endswith("_overlap"):
            si = ctx.sweep_info()
            assert si["compiled"] and si["interior_tiles"] > 0 and si["boundary_tiles"] > 0, si
        mine = ctx.get_particles(0, ctx.n_inner)
        everyone = [None] * world
        dist.all_gather_object(everyone, {k: mine[k] for k in ("id", "rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz")})
        if name == "lj_12x8x8":
            # host-resident particles with several ranks (xnb_step_host_n): every step uploads r, v from host arrays and brings r, v, f
            # (and the ids after a rebuild: atoms migrate between the ranks) back; same trajectory as the device-resident loop
            ctx2 = U.make_ctx(kw, rank=rank, nranks=world, device=local, particles=inp)
            uid2 = [ctx2.nccl_unique_id().copy() if rank == 0 else None]
            dist.broadcast_object_list(uid2, 0)
            ctx2.nccl_init_rank(uid2[0], rank, world)
            ctx2.first_iteration(eps, sig, rc)
            cap = int(ctx2.n_inner * 1.5) + 1024
            hb = {k: np.zeros(cap) for k in ("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz")}
            hid = np.zeros(cap, np.uint64)
            ctx2.download_rvf(hb["rx"], hb["ry"], hb["rz"], hb["vx"], hb["vy"], hb["vz"], hb["fx"], hb["fy"], hb["fz"], hid)
            n2, rb2, counts2 = ctx2.n_inner, 0, set()
            for _ in range(nsteps):
                r_, n2 = ctx2.step_host_n(dt, eps, sig, rc, cap, in_r=(hb["rx"], hb["ry"], hb["rz"]), in_v=(hb["vx"], hb["vy"], hb["vz"]),
                                          out_r=(hb["rx"], hb["ry"], hb["rz"]), out_v=(hb["vx"], hb["vy"], hb["vz"]), out_f=(hb["fx"], hb["fy"], hb["fz"]), out_id=hid)
                rb2 += r_; counts2.add(n2)
            assert rb2 == rb and n2 == ctx2.n_inner == ctx.n_inner, (rb2, rb, n2, ctx.n_inner)
            assert np.array_equal(hid[:n2], mine["id"]), "host-resident stepping: ids differ from the device-resident run"
            for k in ("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz"):
                assert np.array_equal(hb[k][:n2], mine[k]), "host-resident stepping: %s differs from the device-resident run" % k
            if rank == 0:
                print("step_host_n ok: %d ranks, %d steps, %d rebuilds, particle counts seen on rank 0: %s" % (world, nsteps, rb2, sorted(counts2)), flush=True)
            ctx2.close()
        # Newton-3 path on the final configuration: half_symmetric lists, symmetric sweep, and the ghost
        # exchange run backwards (update_force_from_ghost over NCCL) must reproduce the full-list forces (= a, mass 1 here)
        ctx.set_chunk_neighbors_config(half_symmetric=True); ctx.chunk_neighbors()
        ctx.zero_particle_force(ghost=True); ctx.lennard_jones_force_symmetric(eps, sig, rc); ctx.update_force_from_ghost(); ctx.divide_force_by_mass()
        sym = ctx.get_particles(0, ctx.n_inner)
        assert np.array_equal(sym["id"], mine["id"])
        if len(sym["id"]):
            err = U.force_error(U.vec(sym, ("fx", "fy", "fz")), U.vec(mine, ("fx", "fy", "fz")))
            assert err < 1e-10, "%s rank %d: Newton-3 forces differ from the full-list forces (%g)" % (name, rank, err)
        if rank == 0:
            ref = U.make_ctx(kw, device=local, particles=inp)
            ref.first_iteration(eps, sig, rc)
            rb_ref = ref.run_steps(nsteps, dt, eps, sig, rc)
            if lb is not None:
                ref.move_particles(); ref.update_particles_full()
                rb_ref += ref.run_steps(nsteps // 2, dt, eps, sig, rc)
                nb = [c[0] for c in counts]; na = [c[1] for c in counts]
                imb = lambda v: (max(v) - sum(v) / len(v)) / (sum(v) / len(v))
                assert sum(nb) == sum(na), "load_balance_rcb lost atoms: %s -> %s" % (nb, na)
                assert lb[1] <= lb[0] + 1e-12 and imb(na) <= imb(nb) + 1e-12, (lb, nb, na)
                print("load_balance_rcb: lb_inbalance %.3f -> %.3f, atoms per rank %s -> %s, blocks %s" % (lb[0], lb[1], nb, na, blocks), flush=True)
            pr = U.by_id(ref.get_particles(0, ref.n_inner))
            allp = U.by_id({k: np.concatenate([e[k] for e in everyone]) for k in everyone[0]})
            assert np.array_equal(allp["id"], pr["id"]), "%s: atoms lost or duplicated across ranks" % name
            assert rb == rb_ref and rb > 0, (name, rb, rb_ref)
            if name != "lj_voids":                 # clusters and voids: a rank may legitimately own no atom at all (covered on purpose)
                assert min(len(e["id"]) for e in everyone) > 0
            for k in ("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz"):
                assert np.array_equal(allp[k], pr[k]), "%s: %s differs between %d ranks and 1 rank (max %g)" % (name, k, world, np.abs(allp[k] - pr[k]).max())
            print("mgpu parity ok: %s, %d ranks, transport %s, %d atoms, %d steps, %d rebuilds, atoms per rank %s" % (name, world, transport, len(pr["id"]), nsteps, rb, [len(e["id"]) for e in everyone]), flush=True)
            ref.close()
        ctx.close()
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
