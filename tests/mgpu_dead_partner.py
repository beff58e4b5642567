"""Worker of tests/test_gpu_multi.py::test_a_silent_partner_is_an_error_not_a_hang (2 ranks, peer-memory halo): rank 1 stops stepping,
rank 0 keeps going -- its wait for rank 1's displacement count / halo slab must end in an error (XNB_PEER_TIMEOUT_MS), not in a hang."""
import os
import sys
import time

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
    kw = dict(lj_reduced_kwargs(ncell_units=8, cell_units=2), bounds_max=tuple(((4.0 / 0.8442) ** (1 / 3.)) * n for n in (24, 16, 16)), grid_dims=(12, 8, 8))
    inp = U.generate_input(kw)
    ctx = U.make_ctx(kw, rank=rank, nranks=world, device=local, particles=inp)
    uid = [ctx.nccl_unique_id().copy() if rank == 0 else None]
    dist.broadcast_object_list(uid, 0)
    ctx.nccl_init_rank(uid[0], rank, world)
    eps, sig, rc, dt = kw["epsilon"], kw["sigma"], kw["rcut"], kw["dt"]
    ctx.first_iteration(eps, sig, rc)
    ctx.run_steps(5, dt, eps, sig, rc)
    assert ctx.ghost_transport() == "peer", ctx.ghost_transport()
    dist.barrier()
    if rank == 0:
        t0 = time.time()
        try:
            ctx.run_steps(5, dt, eps, sig, rc)
        except capi.XnbError as ex:
            took = time.time() - t0
            assert "did not arrive" in str(ex), str(ex)
            assert took < 15.0, took
            print("silent partner reported after %.2f s: %s" % (took, ex), flush=True)
        else:
            raise AssertionError("run_steps returned although the partner had stopped")
    else:
        time.sleep(4.0)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
