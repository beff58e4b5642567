"""Oracle self-checks on CPU: restated reference invariants (verify_chunk_neighbors.cpp:89-130,
chunk_neighbors_stream_check.h:52-99, debug_total_force.cpp:51-93) and physics properties, on every test configuration."""
import numpy as np
import pytest

from conftest import ni_deck_kwargs, lj_reduced_kwargs
from oracle import oracle as O

CASES = {
    "ni2": ni_deck_kwargs(cells=2),
    "lj2k": lj_reduced_kwargs(ncell_units=8, cell_units=2),
    "lj_gap2": lj_reduced_kwargs(ncell_units=6, cell_units=1),
    "lj_voids": lj_reduced_kwargs(ncell_units=12, cell_units=2, n_spheres=5, sphere_rmin=2.5, sphere_rmax=5.0, drift_speed=1.0),
}


@pytest.mark.parametrize("case", list(CASES))
def test_streams_and_pairs(case):
    o = O.Oracle(O.make_config(**CASES[case]))
    o.init()
    rc, msg = o.check_streams()
    assert rc == 0, msg
    pairs = o.pairs()
    # the full (non symmetric) list is symmetric as a set: (a,b) listed <=> (b,a) listed
    fwd = set(map(tuple, pairs.tolist())); assert len(fwd) == len(pairs)
    assert all((b, a) in fwd for a, b in fwd)
    # brute force reference on ids: minimum-image distance <= rc + skin (no particle sits within 1 ulp of the radius here)
    p = o.particles(); m = o.inner_mask()
    kw = CASES[case]
    L = np.array(kw["bounds_max"]); r = np.stack([p["rx"][m], p["ry"][m], p["rz"][m]], 1); ids = p["id"][m]
    if len(ids) <= 3000:
        d = r[:, None, :] - r[None, :, :]; d -= L * np.round(d / L)
        d2 = (d ** 2).sum(-1)
        nd = kw["rcut"] + kw["rcut_inc"]
        ia, ib = np.nonzero((d2 <= nd * nd) & (d2 > 0))
        brute = set(zip(ids[ia].tolist(), ids[ib].tolist()))
        assert brute == fwd


@pytest.mark.parametrize("case", list(CASES))
def test_total_force_and_energy_conservation(case):
    kw = CASES[case]
    o = O.Oracle(O.make_config(**kw))
    o.init()
    p = o.particles(); m = o.inner_mask()
    f = np.stack([p["fx"][m], p["fy"][m], p["fz"][m]], 1)
    assert np.abs(f.sum(0)).max() <= 1e-11 * max(np.abs(f).sum(), 1.0)
    e0, w0, k0 = o.energy_virial()
    o.run(40)
    e1, w1, k1 = o.energy_virial()
    # velocity-Verlet: total energy drifts only through the unshifted cut-off and dt^2 errors
    assert abs((e1 + k1) - (e0 + k0)) <= 2e-3 * max(abs(e0), abs(k0), 1.0)
    p = o.particles(); m = o.inner_mask()
    mom = np.stack([p["vx"][m], p["vy"][m], p["vz"][m]], 1).sum(0)
    assert np.abs(mom).max() <= 1e-9 * max(1.0, np.abs(p["vx"][m]).sum())


def test_backup_codec_roundtrip():
    """encode_double_u32 / restore_u32_double (core/backup_r.h:31-51): |restore(encode(x)) - x| <= cell/2^33"""
    kw = CASES["lj2k"]
    o = O.Oracle(O.make_config(**kw)); o.init()
    gi = o.grid_info(); d = gi["dims"]; gl = gi["ghost_layers"]; cs = kw["cell_size"]
    p = o.particles(); m = o.inner_mask(); cnt = o.cell_counts()
    cell = np.repeat(np.arange(len(cnt)), cnt)[m]
    ci = cell % d[0]; cj = (cell // d[0]) % d[1]; ck = cell // (d[0] * d[1])
    b = o.backup().reshape(-1, 3)
    for ax, (c, name) in enumerate(((ci, "rx"), (cj, "ry"), (ck, "rz"))):
        org = (gi["offset"][ax] + c) * cs
        back = org + b[:, ax].astype(np.float64) * cs / 2.0 ** 32
        assert np.abs(back - p[name][m]).max() <= cs / 2.0 ** 33 * 1.0001
    assert o.displ_over() == 0


@pytest.mark.parametrize("case", list(CASES))
def test_half_symmetric_lists_and_newton3_sweep(case):
    """ChunkNeighborsConfig::half_symmetric / skip_ghosts (chunk_neighbors_config.h:35-36, neighbor_filter_func.h:36-52) and the
    symmetric sweep + update_force_from_ghost: the half list holds each unordered pair once, and the Newton-3 forces equal the
    full-list forces to rounding"""
    kw = CASES[case]
    o = O.Oracle(O.make_config(**kw)); o.init()
    p = o.particles(); m = o.inner_mask()
    f_full = np.stack([p["fx"][m], p["fy"][m], p["fz"][m]], 1)
    n_full = o.stream_total_u16()
    full = o.pairs()
    o.set_nbh_config(half_symmetric=True); o.build_neighbors()
    rc, msg = o.check_streams(); assert rc == 0, msg
    half = o.pairs()
    assert o.stream_total_u16() < n_full
    # pairs() lists (id_a, id_b) for inner a: an inner-inner pair appears once, an inner-ghost pair once on at most one side
    hs = set(map(tuple, half.tolist())); fs = set(map(tuple, full.tolist()))
    assert hs <= fs and all(((a, b) in hs) or ((b, a) in hs) for a, b in fs)
    o.compute_force_symmetric()
    p = o.particles()
    f_sym = np.stack([p["fx"][m], p["fy"][m], p["fz"][m]], 1)
    scale = max(np.sqrt((f_full ** 2).sum(1).mean()), 1e-300)
    assert np.abs(f_sym - f_full).max() <= 1e-12 * max(scale, np.abs(f_full).max())
    # skip_ghosts: no listed neighbour lives in a ghost cell -> fewer entries than the full list whenever there are ghosts
    o.set_nbh_config(skip_ghosts=True); o.build_neighbors()
    rc, msg = o.check_streams(); assert rc == 0, msg
    assert o.stream_total_u16() < n_full
    o.set_nbh_config(); o.build_neighbors()
    assert o.stream_total_u16() == n_full


@pytest.mark.parametrize("field,weights", [("vx", (1.0,)), ("id", (1.0, 0.0, -0.05)), ("rz", (2.0, -0.3, 0.0, 0.01))])
@pytest.mark.parametrize("case", ["lj2k", "lj_voids"])
def test_average_neighbors_against_brute_force(case, field, weights):
    """average_neighbors_scalar (src/compute/average_neighbors.cu:38-100): the reference ships no test for it; the restatement is pinned
    by an all-pairs evaluation with minimum-image distances (position fields: of the listed image, i.e. unwrapped relative to a)"""
    kw = CASES[case]
    o = O.Oracle(O.make_config(**kw))
    o.init()
    rcut = kw["rcut"]
    w4 = tuple(weights) + (0.0,) * (4 - len(weights))
    avg = o.average_neighbors(rcut, field, w4)
    p = o.particles(); m = o.inner_mask()
    L = np.array(kw["bounds_max"]); r = np.stack([p["rx"][m], p["ry"][m], p["rz"][m]], 1)
    val = p[field][m].astype(np.float64); got = avg[m]
    assert np.all(avg[~m] == 0.0)
    n = len(val)
    ref = np.zeros(n)
    for a in range(n):
        d = r - r[a]; sh = L * np.round(d / L); d -= sh
        d2 = (d ** 2).sum(1)
        sel = (d2 > 0) & (d2 <= rcut * rcut)
        dd = np.sqrt(d2[sel])
        w = w4[0] + w4[1] * dd + w4[2] * d2[sel] + w4[3] * d2[sel] * dd
        v = val[sel] - (sh[sel, 2] if field == "rz" else 0.0)       # a ghost image carries the shifted coordinate
        ref[a] = (w * v).sum() / w.sum() if w.sum() > 0 else (w * v).sum()
    scale = max(np.abs(ref).max(), 1e-300)
    assert np.abs(ref - got).max() <= 1e-11 * scale
    o.close()
