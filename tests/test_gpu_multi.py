"""Multi-GPU parity: the spatially decomposed run (one rank per GPU, NCCL halo exchange) against the single-GPU run."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,transport", [(2, "peer"), (2, "nccl"), (4, "peer"), (8, "peer")])
def test_decomposed_run_is_bit_identical_to_single_gpu(world, transport):
    """transport: how the halo travels -- NVLink peer-memory mailboxes (xnb_peer_halo.cuh, the default on one node) or one
    ncclSend / ncclRecv per partner (XNB_GHOST_NCCL=1); the worker asserts which one ran"""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world + (16 if transport == "nccl" else 0)), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    env = dict(os.environ, XNB_EXPECT_TRANSPORT=transport)
    env.pop("XNB_GHOST_NCCL", None)
    if transport == "nccl":
        env["XNB_GHOST_NCCL"] = "1"
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("mgpu parity ok") == 3, r.stdout[-2000:]
    assert r.stdout.count("transport %s" % transport) == 3, r.stdout[-2000:]


def test_a_silent_partner_is_an_error_not_a_hang():
    """peer-memory halo: the waits inside k_peer_allsum / k_ghost_pull give up after XNB_PEER_TIMEOUT_MS (default 20 s) and the step
    loop returns XNB_ERR_NCCL before it would enter the rebuild's NCCL exchanges"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_dead_partner.py")]
    env = dict(os.environ, XNB_PEER_TIMEOUT_MS="500")
    env.pop("XNB_GHOST_NCCL", None)
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "silent partner reported" in r.stdout, r.stdout[-2000:]


def test_an_error_on_one_rank_ends_the_step_on_every_rank():
    """a particle leaves a non periodic domain on one rank: that rank reports it, and the other rank returns from the same call with an
    error too (the migration's count matrix carries the error words) -- the reference's fatal_error aborts every rank"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29534", os.path.join(ROOT, "tests", "mgpu_lost_particle.py")]
    env = dict(os.environ, XNB_PEER_TIMEOUT_MS="2000")
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "collective error ok" in r.stdout, r.stdout[-2000:]


def test_two_contexts_on_two_devices_in_one_process():
    """one process may hold sub-domains on several GPUs (kernel attributes and the current device are handled per context):
    the same input stepped alternately on cuda:0 and cuda:1 gives bit-identical results"""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import lj_reduced_kwargs
    import parity_util as U
    kw = lj_reduced_kwargs(ncell_units=8, cell_units=2)
    eps, sig, rc, dt = kw["epsilon"], kw["sigma"], kw["rcut"], kw["dt"]
    inp = U.generate_input(kw)
    ctxs = [U.make_ctx(kw, device=d, particles=inp) for d in (0, 1)]
    for c in ctxs:
        c.first_iteration(eps, sig, rc)
    rb = [0, 0]
    for _ in range(15):
        for q, c in enumerate(ctxs):
            rb[q] += c.run_steps(2, dt, eps, sig, rc)
    assert rb[0] == rb[1] and rb[0] > 0
    pa, pb = (c.get_particles(0, c.n_inner) for c in ctxs)
    for k in ("id", "rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz"):
        assert np.array_equal(pa[k], pb[k]), k
    ea, eb = (c.energy_virial(eps, sig, rc) for c in ctxs)
    assert ea[0] == eb[0]
