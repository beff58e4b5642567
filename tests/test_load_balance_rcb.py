"""Host half of op load_balance_rcb (src/mpi/load_balance_rcb.cpp:228-452,510-545, the path without Zoltan):
cost-weighted recursive bisection of the domain cell grid.  CPU only: properties, an independent numpy restatement, and a gloo run in
which every rank contributes the costs of its own cells, all-reduces them (the reference's MPI_Allreduce, :270) and derives its block."""
import multiprocessing as mp
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist

from exanbody_b200 import capi


def clustered_costs(dims, seed=3):
    """cell cost = particles^2-like weight of a few dense clusters on an almost empty background (C5 style)"""
    rng = np.random.default_rng(seed)
    k, j, i = np.meshgrid(np.arange(dims[2]), np.arange(dims[1]), np.arange(dims[0]), indexing="ij")
    c = np.full(i.shape, 0.01)
    for _ in range(5):
        ctr = rng.uniform(0, 1, 3) * np.array(dims); rad = rng.uniform(1.5, 4.0)
        c += 40.0 * np.exp(-((i - ctr[0]) ** 2 + (j - ctr[1]) ** 2 + (k - ctr[2]) ** 2) / (2 * rad * rad))
    return c.ravel()


def numpy_rcb(dims, costs, nparts, part):
    """independent restatement: same rules, written with numpy reductions"""
    c3 = np.asarray(costs).reshape(dims[2], dims[1], dims[0])
    s = [0, 0, 0]; e = list(dims)
    group, r = nparts, part
    while group > 1 and all(e[d] > s[d] for d in range(3)):
        blk = c3[s[2]:e[2], s[1]:e[1], s[0]:e[0]]
        d = [e[a] - s[a] for a in range(3)]
        prof = [blk.sum(axis=(0, 1)), blk.sum(axis=(0, 2)), blk.sum(axis=(1, 2))]
        left, right = group // 2, group - group // 2
        cand = []
        for a in range(3):
            v = prof[a]
            cl = np.concatenate([[0.0], np.cumsum(v)[:-1]]); cr = v.sum() - cl
            wb = np.maximum(cl / left, cr / right); wb[0] = v.sum() / right
            best = v.sum() / right; pos = 0
            # sequential accumulation as the reference does (floating-point order matters for ties)
            sl, sr = 0.0, float(sum(v.tolist()))
            for p in range(1, len(v)):
                sl += v[p - 1]; sr -= v[p - 1]
                w = max(sl / left, sr / right)
                if w < best:
                    best, pos = w, p
            cand.append(dict(pos=pos, wb=best, surf=d[(a + 1) % 3] * d[(a + 2) % 3], axis=a, valid=d[a] >= 2 and 0 < pos < d[a]))
        valid = [x for x in cand if x["valid"]]
        side = r >= left
        if valid:
            best = valid[0]
            for x in valid[1:]:
                mx = max(x["wb"], best["wb"]); mn = min(x["wb"], best["wb"])
                better = x["surf"] < best["surf"] if (mx == 0 or mn / mx > 0.95) else x["wb"] < best["wb"]
                if better:
                    best = x
            a, cut = best["axis"], s[best["axis"]] + best["pos"]
        else:
            a = 0 if d[0] >= d[1] and d[0] >= d[2] else (1 if d[1] >= d[0] and d[1] >= d[2] else 2)
            cut = s[a] + d[a] // 2
        if side:
            s[a] = cut; r -= left; group = right
        else:
            e[a] = cut; group = left
    return s, e


def all_blocks(dims, costs, n):
    return [capi.load_balance_rcb(dims, costs, n, r) for r in range(n)]


@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 8])
@pytest.mark.parametrize("dims", [(12, 12, 12), (20, 8, 6), (5, 3, 9)])
def test_blocks_tile_the_domain_and_balance(dims, n):
    costs = clustered_costs(dims)
    blocks = all_blocks(dims, costs, n)
    owner = np.full(dims[::-1], -1)
    for r, (s, e, cost) in enumerate(blocks):
        assert np.all(owner[s[2]:e[2], s[1]:e[1], s[0]:e[0]] == -1) and all(e[d] > s[d] for d in range(3))
        owner[s[2]:e[2], s[1]:e[1], s[0]:e[0]] = r
        assert abs(cost - costs.reshape(dims[::-1])[s[2]:e[2], s[1]:e[1], s[0]:e[0]].sum()) <= 1e-9 * costs.sum()
    assert (owner >= 0).all()
    bc = np.array([b[2] for b in blocks])
    assert abs(bc.sum() - costs.sum()) <= 1e-9 * costs.sum()
    if n in (2, 4, 8) and dims == (12, 12, 12):
        # better balanced than the static, cost-blind bisection (init_rcb_grid) on clustered costs
        c3 = costs.reshape(dims[::-1])
        static = []
        for r in range(n):
            s, e = capi.rcb_block(dims, n, r)
            static.append(c3[s[2]:e[2], s[1]:e[1], s[0]:e[0]].sum())
        static = np.array(static)
        imb = lambda x: (x.max() - x.mean()) / x.mean()        # lb_inbalance (load_balance_rcb.cpp:443-452)
        assert imb(bc) < imb(static)


@pytest.mark.parametrize("n", [2, 3, 4, 6, 8])
def test_matches_the_numpy_restatement(n):
    for dims, seed in (((12, 12, 12), 3), ((16, 10, 6), 5), ((7, 7, 7), 9)):
        costs = clustered_costs(dims, seed)
        for r in range(n):
            s, e, _ = capi.load_balance_rcb(dims, costs, n, r)
            s2, e2 = numpy_rcb(dims, costs, n, r)
            assert s.tolist() == s2 and e.tolist() == e2, (dims, n, r)


def test_uniform_and_zero_costs():
    dims = (8, 8, 8)
    # uniform costs: halves, quarters, eighths of the domain
    blocks = all_blocks(dims, np.ones(512), 8)
    assert sorted((e - s).tolist() for s, e, _ in blocks) == [[4, 4, 4]] * 8 and all(abs(c - 64.0) < 1e-12 for _, _, c in blocks)
    # zero costs: no cut balances anything -> the longest axis is halved regardless of costs (:389-412), as simple_block_rcb does
    for r in range(4):
        s, e, c = capi.load_balance_rcb((8, 6, 4), np.zeros(192), 4, r)
        s2, e2 = capi.rcb_block((8, 6, 4), 4, r)
        assert s.tolist() == s2.tolist() and e.tolist() == e2.tolist() and c == 0.0
    with pytest.raises(ValueError):
        capi.load_balance_rcb(dims, np.ones(10), 2, 0)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, dims, q):
    try:
        import torch
        os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        full = clustered_costs(dims).reshape(dims[::-1])
        # every rank knows the costs of the cells of its CURRENT (static) block only
        s, e = capi.rcb_block(dims, world, rank)
        mine = np.zeros_like(full); mine[s[2]:e[2], s[1]:e[1], s[0]:e[0]] = full[s[2]:e[2], s[1]:e[1], s[0]:e[0]]
        t = torch.from_numpy(mine.ravel().copy()); dist.all_reduce(t)                   # MPI_Allreduce(SUM) of load_balance_rcb.cpp:270
        assert np.allclose(t.numpy(), full.ravel(), rtol=0, atol=1e-12)
        ns, ne, cost = capi.load_balance_rcb(dims, t.numpy(), world, rank)
        everyone = [None] * world
        dist.all_gather_object(everyone, (ns.tolist(), ne.tolist(), cost))
        owner = np.full(dims[::-1], -1)
        for r, (s0, e0, _) in enumerate(everyone):
            assert np.all(owner[s0[2]:e0[2], s0[1]:e0[1], s0[0]:e0[0]] == -1)
            owner[s0[2]:e0[2], s0[1]:e0[1], s0[0]:e0[0]] = r
        assert (owner >= 0).all()
        costs = np.array([c for _, _, c in everyone])
        assert abs(costs.sum() - full.sum()) <= 1e-9 * full.sum()
        dist.barrier(); dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as ex:      # noqa: BLE001
        q.put((rank, "FAIL %r" % (ex,)))
        raise


@pytest.mark.parametrize("world", [2, 4])
def test_ranks_derive_a_consistent_partition(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue(); port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, (12, 10, 8), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(r[1] == "ok" for r in res), res


def test_simple_cost_model_and_the_chain_to_blocks():
    """simple_cost_model.h:143-145: cost = d1 p + d2 p^2 + d3 p^3 + cc with p = N / cell volume; chained into the bisection, a dense
    slab is shared evenly where the static, cost-blind blocks leave one rank with most of it"""
    cnt = np.array([0, 1, 32, 256], np.uint32)
    cs = 2.0
    p = cnt / cs ** 3
    assert np.array_equal(capi.simple_cost_model(cnt, cs), p)                                   # default coefficients {0, 0, 1, 0}
    co = (0.5, 0.25, 2.0, 3.0)
    assert np.allclose(capi.simple_cost_model(cnt, cs, co), p * 2.0 + p * p * 0.25 + p ** 3 * 0.5 + 3.0, rtol=1e-15, atol=0)
    with pytest.raises(capi.XnbError):
        capi.simple_cost_model(cnt, 0.0)
    dims = (16, 4, 4)
    counts = np.full(dims[::-1], 2, np.uint32); counts[:, :, :4] = 60                             # a dense slab at low i
    costs = capi.simple_cost_model(counts.ravel(), cs)
    blocks = [capi.load_balance_rcb(dims, costs, 4, r) for r in range(4)]
    bc = np.array([c for _, _, c in blocks])
    assert (bc.max() - bc.mean()) / bc.mean() < 0.05                                            # cuts across the slab balance it exactly
    c3 = costs.reshape(dims[::-1]); static = []
    for r in range(4):
        s, e = capi.rcb_block(dims, 4, r)
        static.append(c3[s[2]:e[2], s[1]:e[1], s[0]:e[0]].sum())
    static = np.array(static)
    assert (static.max() - static.mean()) / static.mean() > 0.5                                 # the cost-blind blocks do not


def test_bad_costs_and_empty_blocks_are_errors():
    """the reference aborts with 'Assigned grid block is empty' (load_balance_rcb.cpp:410-411,457); NaN / negative costs would poison the bisection"""
    import pytest
    def n_errors(dims, n):
        bad = 0
        for r in range(n):
            try:
                capi.load_balance_rcb(dims, np.ones(int(np.prod(dims))), n, r)
            except Exception:
                bad += 1
        return bad
    assert n_errors((1, 1, 1), 2) == 1              # more ranks than cells: the rank left without a cell gets an error, not an empty block
    assert n_errors((3, 3, 3), 64) >= 64 - 27
    assert n_errors((4, 4, 4), 8) == 0
    bad = np.ones(64); bad[5] = np.nan
    with pytest.raises(Exception):
        capi.load_balance_rcb((4, 4, 4), bad, 2, 0)
    bad[5] = -1.0
    with pytest.raises(Exception):
        capi.load_balance_rcb((4, 4, 4), bad, 2, 0)
