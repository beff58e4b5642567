"""Shared helpers of the parity tests: build the same simulation in the CPU oracle and in the CUDA library (through the
C-ABI binding) and compare them.  The oracle is the checker only."""
import numpy as np

from oracle import oracle as O
from exanbody_b200 import capi

FIELDS = ("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz", "id", "type")


def make_oracle(kw):
    return O.Oracle(O.make_config(**kw))


def make_ctx(kw, rank=0, nranks=1, device=0, particles=None):
    """configure an xnb_ctx like the deck does: domain, init_rcb_grid, nbh_dist, masses, then lattice + noise input"""
    ctx = capi.Context(device)
    ctx.set_domain(kw.get("bounds_min", (0., 0., 0.)), kw["bounds_max"], kw["cell_size"], kw["grid_dims"], kw.get("periodic", (1, 1, 1)))
    ctx.init_rcb_grid(rank, nranks)
    ctx.set_nbh_dist(kw["rcut"], kw["rcut_inc"])
    ctx.set_type_mass([kw.get("mass", 1.0)])
    ctx.set_sub_grid_density(kw.get("sub_grid_density", 6.5))
    if particles is None:
        particles = generate_input(kw)
    ctx.set_particles(particles["rx"], particles["ry"], particles["rz"], particles["vx"], particles["vy"], particles["vz"],
                      particles["id"], particles["type"])
    return ctx


def generate_input(kw):
    return capi.lattice_fcc(kw["bounds_max"], kw["cell_size"], kw["grid_dims"], kw["lattice_a"], noise_sigma=kw.get("noise_sigma", 0.),
                            vel_sigma=kw.get("vel_sigma", 0.), bounds_min=kw.get("bounds_min", (0., 0., 0.)),
                            n_spheres=kw.get("n_spheres", 0), sphere_rmin=kw.get("sphere_rmin", 0.), sphere_rmax=kw.get("sphere_rmax", 0.),
                            drift_speed=kw.get("drift_speed", 0.))


def gpu_cell_order(ctx):
    """index array that reorders the GPU's flat particle arrays ([inner by cell][ghosts by item]) into cell order"""
    start, count = ctx.cells()
    total = int(count.sum())
    first = np.cumsum(count) - count
    return (np.repeat(start.astype(np.int64), count) + (np.arange(total) - np.repeat(first.astype(np.int64), count))), count


def gpu_particles_cell_order(ctx):
    order, count = gpu_cell_order(ctx)
    p = ctx.get_particles()
    return {k: p[k][order] for k in p}, count.astype(np.int32)


def by_id(p, mask=None):
    """dict of arrays sorted by particle id (optionally restricted by mask)"""
    ids = p["id"] if mask is None else p["id"][mask]
    o = np.argsort(ids, kind="stable")
    return {k: (v if mask is None else v[mask])[o] for k, v in p.items()}


def force_error(fa, fb):
    """per-particle |df| relative to max(|f|, f_rms)  (SURVEY.md 8c tolerance definition)"""
    d = np.sqrt(((fa - fb) ** 2).sum(axis=1))
    n = np.sqrt((fb ** 2).sum(axis=1))
    rms = np.sqrt((n ** 2).mean()) if len(n) else 1.0
    return (d / np.maximum(n, rms)).max() if len(d) else 0.0


def vec(p, names):
    return np.stack([p[n] for n in names], axis=1)


def decode_pairs(counts, ids, sizes, data, dims, inner_cells):
    """pure-python decoder of GridChunkNeighbors streams (cell-ordered) into a sorted (id_a,id_b) array; small cases only"""
    cstart = np.concatenate([[0], np.cumsum(np.asarray(counts, np.int64))])
    sstart = np.concatenate([[0], np.cumsum(np.asarray(sizes, np.int64))])
    out = []
    di, dj = int(dims[0]), int(dims[1])
    for ca in inner_cells:
        n = int(counts[ca])
        if n == 0:
            continue
        st = data[sstart[ca]:sstart[ca + 1]]
        assert st[0] == 1 and st[1] == 0
        off = st[:2 * (n + 1)].view(np.uint32)
        lst = st[2 * (n + 1):]
        for pa in range(n):
            q = int(off[pa]) - 1
            groups = int(lst[q]); q += 1
            for _ in range(groups):
                enc = int(lst[q]); cnt = int(lst[q + 1]); q += 2
                ri, rj, rk = (enc & 31) - 16, ((enc >> 5) & 31) - 16, ((enc >> 10) & 31) - 16
                cb = ca + (rk * dj + rj) * di + ri
                for t in range(cnt):
                    out.append((ids[cstart[ca] + pa], ids[cstart[cb] + int(lst[q + t])]))
                q += cnt
    a = np.array(out, dtype=np.uint64).reshape(-1, 2)
    return a[np.lexsort((a[:, 1], a[:, 0]))]
