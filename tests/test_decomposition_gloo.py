"""Host-side logic of the multi-GPU path, exercised on CPU with real multi-process ranks (gloo, world_size 2 and 4):
each rank derives its block and its ghost send/receive item lists for itself, then the ranks exchange them and check
that they agree -- what A will send to B is exactly what B expects from A (update_ghosts_comm_scheme.cpp:257-303 does
this with MPI_Alltoall + Isend/Irecv at run time)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, grid_dims, periodic, gl, q):
    try:
        from exanbody_b200 import capi
        os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        s, e = capi.rcb_block(grid_dims, world, rank)
        mine = {"block": (s.tolist(), e.tolist()), "send": {}, "recv": {}}
        for p in range(world):
            a, b, f = capi.ghost_items(grid_dims, periodic, gl, world, rank, p)        # what I send to p
            mine["send"][p] = (a.tolist(), b.tolist(), f.tolist())
            a, b, f = capi.ghost_items(grid_dims, periodic, gl, world, p, rank)        # what I expect from p
            mine["recv"][p] = (a.tolist(), b.tolist(), f.tolist())
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        # 1. blocks tile the domain exactly
        owner = np.full(tuple(grid_dims[::-1]), -1)
        for r, o in enumerate(everyone):
            (s0, e0) = o["block"]
            assert np.all(owner[s0[2]:e0[2], s0[1]:e0[1], s0[0]:e0[0]] == -1)
            owner[s0[2]:e0[2], s0[1]:e0[1], s0[0]:e0[0]] = r
        assert (owner >= 0).all()
        # 2. my receive lists are the partners' send lists, item by item and in the same order
        for p in range(world):
            assert everyone[p]["send"][rank] == mine["recv"][p], (rank, p)
        # 3. every ghost cell of my local grid is filled exactly once, and only ghost cells are
        ldims = [e[d] - s[d] + 2 * gl for d in range(3)]
        n_local = ldims[0] * ldims[1] * ldims[2]
        hits = np.zeros(n_local, np.int64)
        for p in range(world):
            np.add.at(hits, np.asarray(mine["recv"][p][1], np.int64), 1)
        k, j, i = np.meshgrid(np.arange(ldims[2]), np.arange(ldims[1]), np.arange(ldims[0]), indexing="ij")
        ghost = ~((i >= gl) & (i < ldims[0] - gl) & (j >= gl) & (j < ldims[1] - gl) & (k >= gl) & (k < ldims[2] - gl))
        if all(periodic):
            assert np.array_equal(hits.reshape(ghost.shape), ghost.astype(np.int64))
        else:
            assert (hits.reshape(ghost.shape)[~ghost] == 0).all() and hits.max() <= 1
        # 4. sender cells are inner cells of the sender
        for p in range(world):
            src = np.asarray(mine["send"][p][0], np.int64)
            if len(src):
                assert not ghost.ravel()[src].any()
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as ex:      # noqa: BLE001
        q.put((rank, "FAIL %r" % (ex,)))
        raise


@pytest.mark.parametrize("world,grid_dims,periodic,gl", [
    (2, (8, 6, 4), (1, 1, 1), 1),
    (2, (6, 6, 6), (1, 0, 1), 2),
    (4, (8, 8, 4), (1, 1, 1), 1),
])
def test_ranks_agree_on_blocks_and_ghost_items(world, grid_dims, periodic, gl):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, grid_dims, periodic, gl, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert all(r[1] == "ok" for r in res), res
    assert all(p.exitcode == 0 for p in procs)


def test_rcb_matches_reference_rule():
    """simple_block_rcb.cpp:27-59: split the longest axis (ties: i before j before k) at the middle, recursively"""
    from exanbody_b200 import capi
    s, e = capi.rcb_block((40, 40, 40), 8, 0); assert (s.tolist(), e.tolist()) == ([0, 0, 0], [20, 20, 20])
    s, e = capi.rcb_block((40, 40, 40), 8, 7); assert (s.tolist(), e.tolist()) == ([20, 20, 20], [40, 40, 40])
    s, e = capi.rcb_block((40, 40, 40), 2, 1); assert (s.tolist(), e.tolist()) == ([20, 0, 0], [40, 40, 40])
    s, e = capi.rcb_block((100, 50, 50), 4, 1); assert (s.tolist(), e.tolist()) == ([25, 0, 0], [50, 50, 50])
    s, e = capi.rcb_block((10, 10, 10), 3, 2); assert (s.tolist(), e.tolist()) == ([5, 5, 0], [10, 10, 10])
    s, e = capi.rcb_block((7, 3, 3), 1, 0); assert (s.tolist(), e.tolist()) == ([0, 0, 0], [7, 3, 3])
