"""Worker of tests/test_gpu_multi.py::test_an_error_on_one_rank_ends_the_step_on_every_rank (2 ranks): a non periodic box, one atom
of rank 1 is shot out of the domain.  Rank 1's binning raises DERR_LOST_PARTICLE; rank 0 must return an error from the same call (the
count matrix of the migration carries every rank's error word) instead of waiting for rank 1 in the exchange that follows."""
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
from exanbody_b200 import capi                   # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    kw = dict(lj_reduced_kwargs(ncell_units=8, cell_units=2), bounds_max=tuple(((4.0 / 0.8442) ** (1 / 3.)) * n for n in (24, 16, 16)),
              grid_dims=(12, 8, 8), periodic=(0, 0, 0))
    inp = U.generate_input(kw)
    # the atom with the largest x (it lives on the last rank) leaves through the +x face within a few steps
    # (everybody else stands still, so that no thermal atom of the surface leaves through a face on the other rank)
    for f in ("vx", "vy", "vz"):
        inp[f][:] = 0.0
    k = int(np.argmax(inp["rx"]))
    inp["vx"][k] = 400.0
    ctx = U.make_ctx(kw, rank=rank, nranks=world, device=local, particles=inp)
    uid = [ctx.nccl_unique_id().copy() if rank == 0 else None]
    dist.broadcast_object_list(uid, 0)
    ctx.nccl_init_rank(uid[0], rank, world)
    eps, sig, rc, dt = kw["epsilon"], kw["sigma"], kw["rcut"], kw["dt"]
    ctx.first_iteration(eps, sig, rc)
    try:
        ctx.run_steps(20, dt, eps, sig, rc)
    except capi.XnbError as ex:
        print("rank %d: %s" % (rank, ex), flush=True)
        msg = str(ex)
    else:
        raise AssertionError("rank %d: run_steps returned although an atom left the non periodic domain" % rank)
    msgs = [None] * world
    dist.all_gather_object(msgs, msg)
    if rank == 0:
        assert sum("left a non periodic domain" in m for m in msgs) == 1 and sum("reported a device error" in m for m in msgs) == world - 1, msgs
        print("collective error ok: %s" % msgs, flush=True)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
