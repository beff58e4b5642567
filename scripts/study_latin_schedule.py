"""Model: shared-memory wavefronts of the sweep's position gathers when every lane orders its own list entries so that, at step s,
lane L prefers a candidate whose staged index is congruent to (L + s) mod 16 (a per-lane 'Latin' schedule: the lanes of a half-warp
then tend to hit different 8-byte bank pairs, the lanes of a quarter-warp different 16-byte bank quads).
LDS.128 over {x,y}: a quarter-warp (8 lanes) per wavefront when their 16-byte slots (index mod 8) differ -> wavefronts = sum over the
four quarters of the largest multiplicity.  LDS.64 over z: a half-warp per wavefront, 8-byte slots (index mod 16)."""
import numpy as np

rng = np.random.default_rng(1)


def wavefronts(idx):
    """idx: (steps, 32) staged indices; returns mean wavefronts per step for LDS.128 and LDS.64"""
    w128 = 0.0; w64 = 0.0
    for row in idx:
        for q in range(4):
            w128 += np.bincount(row[8 * q:8 * q + 8] % 8, minlength=8).max()
        for h in range(2):
            w64 += np.bincount(row[16 * h:16 * h + 16] % 16, minlength=16).max()
    return w128 / len(idx), w64 / len(idx)


def schedule(lst, lane):
    """greedy per-lane order: at step s take an entry of residue (lane + s) mod 16 if there is one, else of the fullest bucket"""
    buckets = [[] for _ in range(16)]
    for v in lst:
        buckets[v % 16].append(v)
    out = []
    for s in range(len(lst)):
        r = (lane + s) % 16
        if not buckets[r]:
            # prefer a residue that keeps mod 8 right, else the fullest
            r2 = (r + 8) % 16
            r = r2 if buckets[r2] else max(range(16), key=lambda b: len(buckets[b]))
        out.append(buckets[r].pop())
    return out


def run(n_groups=200, nlist=78, spread=8, halo=3000):
    base = []; lat = []
    for _ in range(n_groups):
        lens = np.clip(rng.normal(nlist, spread, 32).astype(int), 40, 120)
        L = lens.max()
        lists = [sorted(rng.choice(halo, n, replace=False).tolist()) for n in lens]
        self_idx = rng.integers(0, halo - 32) + np.arange(32)
        a = np.empty((L, 32), np.int64); b = np.empty((L, 32), np.int64)
        for lane in range(32):
            pad = [int(self_idx[lane])] * (L - lens[lane])
            a[:, lane] = lists[lane] + pad
            b[:, lane] = schedule(lists[lane], lane) + pad
        base.append(wavefronts(a)); lat.append(wavefronts(b))
    base = np.mean(base, 0); lat = np.mean(lat, 0)
    print("ascending order : LDS.128 %.2f  LDS.64 %.2f  total %.2f wavefronts per candidate-step" % (base[0], base[1], base.sum()))
    print("latin schedule  : LDS.128 %.2f  LDS.64 %.2f  total %.2f" % (lat[0], lat[1], lat.sum()))


if __name__ == "__main__":
    run()
