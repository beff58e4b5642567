#!/usr/bin/env python
"""Design study (numpy only, no GPU, no oracle): how many shared-memory gathers the pair-merged list layout of DESIGN.md 3.1 would
need under different partner choices and group formations, on a C2-like configuration (FCC rho* = 0.8442, cells of 2a = 32 atoms,
list radius 2.8 sigma, thermal disorder imitated by Gaussian displacements).

  python scripts/study_pair_merge.py [unit cells per axis = 16] [displacement sigma = 0.15]

Prints, per strategy: mean union length per pair, padded rows (4 entries) per warp, and gathers per (particle, list entry) relative
to the unpaired layout (one gather per padded entry)."""
import sys
import numpy as np

rng = np.random.default_rng(1)
nu = int(sys.argv[1]) if len(sys.argv) > 1 else 16
disp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.15
a = (4.0 / 0.8442) ** (1.0 / 3.0)
L = nu * a
basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]])
ijk = np.stack(np.meshgrid(np.arange(nu), np.arange(nu), np.arange(nu), indexing="ij"), -1).reshape(-1, 1, 3)
r = ((ijk + basis[None]).reshape(-1, 3) * a + rng.normal(0, disp, (nu ** 3 * 4, 3))) % L
n = len(r)
cs = 2 * a; nc = nu // 2
cell = np.floor(r / cs).astype(int) % nc
cid = (cell[:, 2] * nc + cell[:, 1]) * nc + cell[:, 0]
R = 2.8

# neighbour lists by cell binning (periodic)
order = np.argsort(cid, kind="stable"); start = np.searchsorted(cid[order], np.arange(nc ** 3 + 1))
lists = [None] * n
for c in range(nc ** 3):
    ci, cj, ck = c % nc, (c // nc) % nc, c // (nc * nc)
    mine = order[start[c]:start[c + 1]]
    cand = np.concatenate([order[start[cc]:start[cc + 1]] for cc in
                           {(((ck + dk) % nc) * nc + (cj + dj) % nc) * nc + (ci + di) % nc for dk in (-1, 0, 1) for dj in (-1, 0, 1) for di in (-1, 0, 1)}])
    d = r[mine][:, None, :] - r[cand][None, :, :]; d -= L * np.round(d / L)
    d2 = (d ** 2).sum(-1)
    for q, p in enumerate(mine):
        m = (d2[q] <= R * R) & (d2[q] > 0)
        lists[p] = np.sort(cand[m])
ln = np.array([len(x) for x in lists])
print("atoms %d, cells %d (%.1f per cell), list entries per atom %.1f (min %d max %d)" % (n, nc ** 3, n / nc ** 3, ln.mean(), ln.min(), ln.max()))

# tiles of 4x2x2 cells, as the sweep uses
tiles = {}
for c in range(nc ** 3):
    ci, cj, ck = c % nc, (c // nc) % nc, c // (nc * nc)
    tiles.setdefault((ci // 4, cj // 2, ck // 2), []).append(c)


def rows_unpaired():
    tot = 0
    for cells in tiles.values():
        parts = np.concatenate([order[start[c]:start[c + 1]] for c in cells])
        for g in range(0, len(parts), 32):
            tot += -(-ln[parts[g:g + 32]].max() // 4)
    return tot


base_rows = rows_unpaired()
print("unpaired: %.2f rows per 32 particles, padded entries per particle %.1f" % (base_rows / (n / 32), base_rows * 128 / n))


def union_len(p, q):
    return len(np.union1d(lists[p], lists[q]))


def morton(parts):
    rel = (r[parts] / cs) % 1.0
    b = np.minimum((rel * 4).astype(int), 3)
    key = np.zeros(len(parts), int)
    for bit in range(2):
        for d in range(3):
            key |= ((b[:, d] >> bit) & 1) << (3 * bit + d)
    return parts[np.argsort(key, kind="stable")]


def pairs_morton(parts):
    s = morton(parts)
    return [(s[k], s[k + 1] if k + 1 < len(s) else -1) for k in range(0, len(s), 2)]


def pairs_greedy(parts):
    """closest pair first (periodic distance inside the cell is plain distance)"""
    left = list(parts); out = []
    pos = r[parts]; d = pos[:, None, :] - pos[None, :, :]; d -= L * np.round(d / L)
    d2 = (d ** 2).sum(-1); np.fill_diagonal(d2, 1e9)
    idx = {p: k for k, p in enumerate(parts)}
    alive = np.ones(len(parts), bool)
    while alive.sum() >= 2:
        sub = np.where(alive)[0]
        m = d2[np.ix_(sub, sub)]
        k = np.unravel_index(np.argmin(m), m.shape)
        p, q = sub[k[0]], sub[k[1]]
        out.append((parts[p], parts[q])); alive[p] = alive[q] = False
    if alive.any():
        out.append((parts[np.where(alive)[0][0]], -1))
    return out


def study(name, pair_fn, balance):
    tot_rows = 0; unions = []; npairs = 0
    for cells in tiles.values():
        pairs = []
        for c in cells:
            pairs += pair_fn(order[start[c]:start[c + 1]])
        ul = np.array([union_len(p, q) if q >= 0 else len(lists[p]) for p, q in pairs])
        unions += ul.tolist(); npairs += len(pairs)
        if balance:
            ul = np.sort(ul)[::-1]
        for g in range(0, len(ul), 32):
            tot_rows += -(-ul[g:g + 32].max() // 4)
    unions = np.array(unions)
    gathers = tot_rows * 128.0            # one gather per padded union entry
    print("%-34s union per pair %.1f (%.2f lists), rows per warp %.1f, gathers per (particle, padded unpaired entry) %.2f, evaluations x%.2f"
          % (name, unions.mean(), unions.mean() / ln.mean(), tot_rows / (npairs / 32), gathers / (base_rows * 128.0), 2.0 * gathers / (base_rows * 128.0)))


study("Morton-consecutive, groups as stored", pairs_morton, False)
study("Morton-consecutive, length-balanced", pairs_morton, True)
study("closest-first matching, as stored", pairs_greedy, False)
study("closest-first matching, balanced", pairs_greedy, True)


def rows_unpaired_sorted():
    """the unpaired layout with a tile's particles ordered by list length before groups of 32 are formed"""
    tot = 0
    for cells in tiles.values():
        parts = np.concatenate([order[start[c]:start[c + 1]] for c in cells])
        l = np.sort(ln[parts])[::-1]
        for g in range(0, len(l), 32):
            tot += -(-l[g:g + 32].max() // 4)
    return tot


sr = rows_unpaired_sorted()
ideal = sum(-(-int(ln[np.concatenate([order[start[c]:start[c + 1]] for c in cells])].sum()) // 128) for cells in tiles.values())
print("unpaired, groups formed by list length: padded entries per particle %.1f (as stored %.1f, no padding at all %.1f)" % (sr * 128 / n, base_rows * 128 / n, ideal * 128 / n))
