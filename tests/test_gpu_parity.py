"""GPU parity tests proper: the CUDA path, called through the C-ABI, against the CPU oracle on the same seeded inputs.

Bars (SURVEY.md 8c / BASELINE.json north_star): cell membership, AMR tables, backup codes and neighbour streams bit-exact;
neighbour pair sets (by id) identical; forces / positions / velocities within 1e-10 relative (|df| <= 1e-10*max(|f|, f_rms))."""
import numpy as np
import pytest

from conftest import ni_deck_kwargs, lj_reduced_kwargs
import parity_util as U

pytestmark = pytest.mark.gpu

TOL = 1e-10

CASES = {
    # verbatim reference deck: 16384 atoms, 256 per cell -> AMR side 3, 1 ghost layer
    "ni16k": ni_deck_kwargs(),
    # reduced LJ, 32 atoms per cell (AMR side 1), liquid-like velocities
    "lj2k": lj_reduced_kwargs(ncell_units=8, cell_units=2),
    # cell smaller than the list radius: 2 neighbour layers, 2 ghost layers
    "lj_gap2": lj_reduced_kwargs(ncell_units=6, cell_units=1),
    # non cubic domain
    "lj_slab": dict(lj_reduced_kwargs(ncell_units=8, cell_units=2), bounds_max=tuple(((4.0 / 0.8442) ** (1 / 3.)) * n for n in (8, 12, 16)), grid_dims=(4, 6, 8)),
    # clusters + voids (C5 style): empty cells, ragged occupancy
    "lj_voids": lj_reduced_kwargs(ncell_units=12, cell_units=2, n_spheres=5, sphere_rmin=2.5, sphere_rmax=5.0, drift_speed=1.0),
    # dense regime: cell = 4a (256 atoms, AMR side 3), rc = 5 sigma needs 1024-neighbour buffers in the oracle
    "lj_dense": lj_reduced_kwargs(ncell_units=12, cell_units=4, rcut=5.0, noise=0.1),
}


def setup_pair(kw):
    o = U.make_oracle(kw)
    o.generate()
    inp = U.generate_input(kw)
    # the product's host-side lattice/noise operators and the oracle's must agree bit for bit
    po = o.particles()
    for k in ("rx", "ry", "rz", "vx", "vy", "vz", "id"):
        assert np.array_equal(po[k], inp[k]), k
    ctx = U.make_ctx(kw, particles=inp)
    return o, ctx


@pytest.mark.parametrize("nbh_kernels", ["tiled", "untiled"])
@pytest.mark.parametrize("case", list(CASES))
def test_rebuild_pipeline_bit_exact(case, nbh_kernels, monkeypatch):
    # "tiled": k_nbh_bits (fp32 classification into accept bits + exact fp64 decision inside the band; streams of the ghost cells
    # built on demand); "untiled": the per-particle all-fp64 kernels kept as the fallback for tiles that do not fit shared memory
    if nbh_kernels == "untiled":
        monkeypatch.setenv("XNB_NBH_UNTILED", "1")
    else:
        monkeypatch.delenv("XNB_NBH_UNTILED", raising=False)
    kw = CASES[case]
    o, ctx = setup_pair(kw)
    o.move_particles(); o.update_particles_full()
    ctx.move_particles(); ctx.update_particles_full()
    gi_o, gi = o.grid_info(), ctx.grid_info()
    assert np.array_equal(gi_o["dims"], gi["dims"]) and np.array_equal(gi_o["offset"], gi["offset"]) and gi_o["ghost_layers"] == gi["ghost_layers"]
    # ---- binning + ghosts: same particles in every cell (inner and ghost)
    pg, cnt_g = U.gpu_particles_cell_order(ctx)
    po, cnt_o = o.particles(), o.cell_counts()
    assert np.array_equal(cnt_g, cnt_o)
    assert ctx.n_inner == o.n_inner() and ctx.n_total == o.n_total()
    cell_of = np.repeat(np.arange(len(cnt_o)), cnt_o)
    key_o = np.lexsort((po["id"], cell_of)); key_g = np.lexsort((pg["id"], cell_of))
    assert np.array_equal(po["id"][key_o], pg["id"][key_g])
    for k in ("rx", "ry", "rz", "vx", "vy", "vz"):           # positions incl. the periodic shift of ghosts: bit exact
        assert np.array_equal(po[k][key_o], pg[k][key_g]), k
    # ---- AMR tables (cumulative sub-cell offsets do not depend on the order inside a sub-cell)
    sgs_o, sgc_o = o.amr_tables(); sgs_g, sgc_g = ctx.amr_tables()
    assert np.array_equal(sgs_o, sgs_g) and np.array_equal(sgc_o, sgc_g)
    # ---- backup_r codes, matched by id
    inner = o.inner_mask()
    bo = o.backup().reshape(-1, 3); bg = ctx.backup().reshape(-1, 3)
    ids_g_inner = ctx.get_particles(0, ctx.n_inner, fields=("id",))["id"]
    assert np.array_equal(bo[np.argsort(po["id"][inner])], bg[np.argsort(ids_g_inner)])
    # ---- neighbour streams: byte-equal once the oracle is given the GPU's in-cell order
    pairs_ref = o.pairs()
    o.set_particles(cnt_g, pg)
    o.build_neighbors()
    rc, msg = o.check_streams(); assert rc == 0, msg
    sz_o, data_o = o.streams(); sz_g, data_g = ctx.streams()
    assert np.array_equal(sz_o, sz_g)
    assert np.array_equal(data_o, data_g)
    assert ctx.view_chunk_neighbors()[2] == o.max_neighbors()
    # ---- and the pair set (by id) equals the one of the oracle's own, independently ordered run
    assert np.array_equal(o.pairs(), pairs_ref)


def boundary_case():
    """hand-made input that sits ON the decision boundaries: an exact FCC lattice with a = 1 (coordinates are multiples of
    0.5, so every d2 is exact) and nbh_dist = 1.0, i.e. the whole second shell has d2 == max_dist^2 exactly (kept: <=);
    some atoms nudged by +-1e-13 / +-3e-8 (just inside / outside), and a few coincident duplicates (d2 == 0: never
    neighbours, chunk_neighbors_execute.h:225-227)."""
    n = 8
    kw = dict(bounds_max=(8.0,) * 3, cell_size=2.0, grid_dims=(4,) * 3, lattice_a=1.0, epsilon=1.0, sigma=0.5, rcut=0.75, rcut_inc=0.25,
              dt=1e-4, mass=1.0, noise_sigma=0.0, vel_sigma=0.0, max_neighbors=1024)
    basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]])
    ijk = np.stack(np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij"), -1).reshape(-1, 1, 3)
    r = (ijk + basis[None]).reshape(-1, 3).astype(np.float64) + 0.25          # +0.25 keeps atoms off the cell faces
    rng = np.random.default_rng(7)
    k = rng.choice(len(r), 400, replace=False)
    r[k[:100], 0] += 1e-13; r[k[100:200], 1] -= 1e-13; r[k[200:300], 2] += 3e-8; r[k[300:400], 0] -= 3e-8
    dup = r[rng.choice(len(r), 12, replace=False)].copy()                        # coincident distinct particles
    r = np.concatenate([r, dup])
    m = len(r)
    p = dict(rx=r[:, 0].copy(), ry=r[:, 1].copy(), rz=r[:, 2].copy(), vx=np.zeros(m), vy=np.zeros(m), vz=np.zeros(m),
             id=np.arange(1, m + 1, dtype=np.uint64), type=np.zeros(m, np.uint8))
    return kw, p


@pytest.mark.parametrize("nbh_kernels", ["tiled", "untiled"])
def test_neighbour_decisions_on_the_boundary(nbh_kernels, monkeypatch):
    if nbh_kernels == "untiled":
        monkeypatch.setenv("XNB_NBH_UNTILED", "1")
    else:
        monkeypatch.delenv("XNB_NBH_UNTILED", raising=False)
    kw, p = boundary_case()
    o = U.make_oracle(kw); o.generate()
    counts = np.zeros(o.grid_info()["n_cells"], np.int32); counts[0] = len(p["id"])
    o.set_particles(counts, p)
    ctx = U.make_ctx(kw, particles=p)
    o.move_particles(); o.update_particles_full()
    ctx.move_particles(); ctx.update_particles_full()
    pairs_ref = o.pairs()
    # second shell (6 pairs at d2 == max_dist^2) must be present for unperturbed atoms: 12 + 6 neighbours
    assert len(pairs_ref) > 17 * 2000
    pg, cnt_g = U.gpu_particles_cell_order(ctx)
    assert np.array_equal(cnt_g, o.cell_counts())
    o.set_particles(cnt_g, pg)
    o.build_neighbors()
    sz_o, data_o = o.streams(); sz_g, data_g = ctx.streams()
    assert np.array_equal(sz_o, sz_g) and np.array_equal(data_o, data_g)
    assert np.array_equal(o.pairs(), pairs_ref)
    # coincident particles are not neighbours of each other
    ids = pairs_ref.astype(np.int64)
    pos = {int(i): (x, y, z) for i, x, y, z in zip(p["id"], p["rx"], p["ry"], p["rz"])}
    n_dup = len(p["id"]) - 12
    for a, b in ids[ids[:, 0] > n_dup]:
        assert pos[int(a)] != pos[int(b)]


@pytest.mark.parametrize("case", ["ni16k", "lj2k", "lj_gap2", "lj_voids"])
def test_stream_decoding_independent_of_oracle(case):
    """decode the GPU streams with a third, pure-python reader and compare the pair set with the oracle's"""
    kw = CASES[case]
    if case == "ni16k":
        kw = ni_deck_kwargs(cells=2)
    o, ctx = setup_pair(kw)
    o.move_particles(); o.update_particles_full()
    ctx.move_particles(); ctx.update_particles_full()
    pg, cnt = U.gpu_particles_cell_order(ctx)
    sz, data = ctx.streams()
    gi = ctx.grid_info(); d = gi["dims"]; gl = gi["ghost_layers"]
    inner_cells = [(k * d[1] + j) * d[0] + i for k in range(gl, d[2] - gl) for j in range(gl, d[1] - gl) for i in range(gl, d[0] - gl)]
    pairs = U.decode_pairs(cnt, pg["id"], sz, data, d, inner_cells)
    assert np.array_equal(pairs, o.pairs())


@pytest.mark.parametrize("functor", ["lj", "lj_reference_form"])
@pytest.mark.parametrize("sweep", ["compiled", "streams"])
@pytest.mark.parametrize("case", list(CASES))
def test_force_parity(case, sweep, functor, monkeypatch):
    """sweep = compiled: k_lj_sweep_cl over the compiled lists (the default); streams: k_lj_sweep reading the
    GridChunkNeighbors streams directly (the fallback when no tile shape fits).
    functor = lj: the restated Lennard-Jones functor (four-candidate hook); lj_reference_form: lj_compute_energy + the
    buffer-less operator() as written in lennard_jones.cu:46-56,106-124, driven through the generic buffer-less call of
    the functor concept (xnb_pair_functor.cuh) -- the route any other functor would take"""
    if sweep == "streams":
        monkeypatch.setenv("XNB_SWEEP_STREAMS", "1")
    kw = CASES[case]
    o, ctx = setup_pair(kw)
    ctx.set_pair_functor(1 if functor == "lj_reference_form" else 0)
    o.first_iteration()
    ctx.first_iteration(kw["epsilon"], kw["sigma"], kw["rcut"])
    po = U.by_id(o.particles(), o.inner_mask())
    pg = U.by_id(ctx.get_particles(0, ctx.n_inner))
    assert np.array_equal(po["id"], pg["id"])
    err = U.force_error(U.vec(pg, ("fx", "fy", "fz")), U.vec(po, ("fx", "fy", "fz")))
    assert err < TOL, err
    # sum of forces vanishes (debug_total_force.cpp): m*a summed over all atoms
    ftot = np.abs(U.vec(pg, ("fx", "fy", "fz")).sum(axis=0)).max()
    fscale = np.abs(U.vec(pg, ("fx", "fy", "fz"))).sum()
    assert ftot <= 1e-11 * max(fscale, 1.0)
    # energy / virial (oracle-defined, unpinned by the reference)
    e_o, w_o, k_o = o.energy_virial()
    e_g, w_g, k_g = ctx.energy_virial(kw["epsilon"], kw["sigma"], kw["rcut"])
    assert abs(e_g - e_o) <= TOL * abs(e_o)
    assert np.abs(w_g - w_o).max() <= TOL * np.abs(w_o).max()
    assert abs(k_g - k_o) <= TOL * max(abs(k_o), 1e-300)


@pytest.mark.parametrize("tile", ["", "1,1,1", "2,2,2", "4,4,1", "3,3,2", "8,2,2"])
@pytest.mark.parametrize("case", ["ni16k", "lj2k", "lj_gap2", "lj_voids", "lj_dense"])
def test_compiled_lists_sweep_equals_stream_sweep(case, tile, monkeypatch):
    """the compiled lists hold the same candidates in the same order as the streams: forces, energy and virial of the two
    sweeps are bit-identical per atom, whatever the tile shape (also with ghost cells swept)"""
    kw = CASES[case]
    eps, sig, rc = kw["epsilon"], kw["sigma"], kw["rcut"]
    res = []
    for streams in (True, False):
        if streams:
            monkeypatch.setenv("XNB_SWEEP_STREAMS", "1")
        else:
            monkeypatch.delenv("XNB_SWEEP_STREAMS")
            if tile:
                monkeypatch.setenv("XNB_CL_TILE", tile)
        _, ctx = setup_pair(kw)
        ctx.first_iteration(eps, sig, rc)
        p = ctx.get_particles()
        ev = ctx.energy_virial(eps, sig, rc)
        ctx.zero_particle_force(True); ctx.lennard_jones_force(eps, sig, rc, ghost=True)
        pg = ctx.get_particles()
        res.append((p, ev, pg))
    (pa, eva, pga), (pb, evb, pgb) = res
    for k in ("fx", "fy", "fz"):
        assert np.array_equal(pa[k], pb[k]), k
        assert np.array_equal(pga[k], pgb[k]), k
    # energy / virial: per-tile partial sums, so only the summation order differs
    assert abs(eva[0] - evb[0]) <= 1e-12 * abs(eva[0]) and np.abs(eva[1] - evb[1]).max() <= 1e-12 * np.abs(eva[1]).max()


@pytest.mark.parametrize("case", ["ni16k", "lj2k", "lj_gap2", "lj_voids"])
def test_unfused_operator_sequence_equals_fused(case):
    """the reference's operator-by-operator sequence (zero_particle_force, lennard_jones_force, divide_force_by_type_scalar,
    push_f_v_r, push_f_v, particle_displ_over) gives the same numbers as the fused kernels"""
    kw = CASES[case]
    _, a = setup_pair(kw)
    _, b = setup_pair(kw)
    eps, sig, rc, dt = kw["epsilon"], kw["sigma"], kw["rcut"], kw["dt"]
    a.first_iteration(eps, sig, rc)
    b.move_particles(); b.update_particles_full()
    b.zero_particle_force(True); b.lennard_jones_force(eps, sig, rc); b.divide_force_by_mass()
    pa, pb = a.get_particles(), b.get_particles()
    for k in ("fx", "fy", "fz"):
        assert np.array_equal(pa[k], pb[k]), k
    a.verlet_first_half(dt); over_a = a.read_displ_over()
    b.push_f_v_r(dt, 1.0); b.push_f_v(dt, 0.5); over_b = b.particle_displ_over()
    assert over_a == over_b
    a.ghost_update_r(); b.ghost_update_r()
    a.force_and_second_half(eps, sig, rc, 0.5 * dt)
    b.zero_particle_force(True); b.lennard_jones_force(eps, sig, rc); b.divide_force_by_mass(); b.push_f_v(dt, 0.5)
    pa, pb = a.get_particles(), b.get_particles()
    for k in ("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz"):
        assert np.array_equal(pa[k], pb[k]), k


@pytest.mark.parametrize("case,nsteps", [("ni16k", 100), ("lj2k", 60), ("lj_gap2", 40), ("lj_voids", 60), ("lj_dense", 12)])
def test_trajectory_parity(case, nsteps):
    kw = CASES[case]
    o, ctx = setup_pair(kw)
    o.first_iteration()
    ctx.first_iteration(kw["epsilon"], kw["sigma"], kw["rcut"])
    rb_o = o.run(nsteps)
    rb_g = ctx.run_steps(nsteps, kw["dt"], kw["epsilon"], kw["sigma"], kw["rcut"])
    assert rb_g == rb_o                       # rebuild frequency parity (u32-quantised trigger)
    if case != "ni16k":
        assert rb_o > 0                       # make sure the rebuild path was exercised
    po = U.by_id(o.particles(), o.inner_mask())
    pg = U.by_id(ctx.get_particles(0, ctx.n_inner))
    assert np.array_equal(po["id"], pg["id"])
    L = np.array(kw["bounds_max"]) - np.array(kw.get("bounds_min", (0., 0., 0.)))
    dr = U.vec(pg, ("rx", "ry", "rz")) - U.vec(po, ("rx", "ry", "rz"))
    dr -= L * np.round(dr / L)
    scale_r = kw["cell_size"]
    assert np.abs(dr).max() <= 1e-9 * scale_r, np.abs(dr).max()
    vo = U.vec(po, ("vx", "vy", "vz")); vg = U.vec(pg, ("vx", "vy", "vz"))
    vrms = max(np.sqrt((vo ** 2).sum(axis=1).mean()), 1e-300)
    assert np.abs(vg - vo).max() <= 1e-8 * vrms, np.abs(vg - vo).max() / vrms
    err = U.force_error(U.vec(pg, ("fx", "fy", "fz")), U.vec(po, ("fx", "fy", "fz")))
    assert err < 1e-7, err                    # chaotic amplification of 1e-16 rounding over the run; single-step bar is 1e-10 above


def test_speculative_fast_path_changes_nothing(monkeypatch):
    """xnb_run_steps enqueues ghost_update_r + the sweep before the host has read the displacement count (the launches are void
    when a rebuild is due): same rebuild steps and bit-identical state as the strictly sequential loop"""
    kw = CASES["lj2k"]
    eps, sig, rc, dt = kw["epsilon"], kw["sigma"], kw["rcut"], kw["dt"]
    out = []
    for seq in (False, True):
        if seq:
            monkeypatch.setenv("XNB_NO_SPECULATION", "1")
        _, ctx = setup_pair(kw)
        ctx.first_iteration(eps, sig, rc)
        rb = [ctx.run_steps(1, dt, eps, sig, rc) for _ in range(30)]
        out.append((rb, ctx.get_particles(0, ctx.n_inner)))
    (rb_a, pa), (rb_b, pb) = out
    assert rb_a == rb_b and sum(rb_a) > 0
    for k in ("id", "rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz"):
        assert np.array_equal(pa[k], pb[k]), k


@pytest.mark.parametrize("case", ["lj2k", "lj_voids", "ni16k"])
def test_in_cell_sort_kernels_agree(case, monkeypatch):
    """binning sorts every cell by particle id with one warp per cell (cells of up to 64 particles: keys in registers; fuller cells
    through global scratch) or one block per cell: same particle order, hence bit-identical runs"""
    kw = CASES[case]
    eps, sig, rc, dt = kw["epsilon"], kw["sigma"], kw["rcut"], kw["dt"]
    out = []
    for var in ("XNB_CELLSORT_WARP", "XNB_CELLSORT_BLOCK"):
        monkeypatch.delenv("XNB_CELLSORT_WARP", raising=False); monkeypatch.delenv("XNB_CELLSORT_BLOCK", raising=False)
        monkeypatch.setenv(var, "1")
        _, ctx = setup_pair(kw)
        ctx.first_iteration(eps, sig, rc)
        order0 = ctx.get_particles(0, ctx.n_inner, fields=("id",))["id"].copy()
        rb = ctx.run_steps(25, dt, eps, sig, rc)
        out.append((order0, rb, ctx.get_particles(0, ctx.n_inner)))
        ctx.close()
    assert np.array_equal(out[0][0], out[1][0]) and out[0][1] == out[1][1]
    for k in ("id", "rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz"):
        assert np.array_equal(out[0][2][k], out[1][2][k]), k


@pytest.mark.parametrize("case", ["lj2k", "lj_voids", "lj_gap2", "ni16k"])
def test_fused_next_first_half_changes_nothing(case, monkeypatch):
    """inside one xnb_run_steps call the sweep of step k also performs the first half of step k + 1 (k_lj_sweep_cl MODE 2: positions
    into the other buffer, displacement count into the other counter slot).  Same rebuild steps and bit-identical state as the loop
    with a stand-alone first-half kernel and as thirty single-step calls, also across two consecutive calls"""
    kw = CASES[case]
    eps, sig, rc, dt = kw["epsilon"], kw["sigma"], kw["rcut"], kw["dt"]
    out = []
    for mode in ("fused", "plain", "single"):
        if mode == "plain":
            monkeypatch.setenv("XNB_NO_FUSED_FIRST_HALF", "1")
        else:
            monkeypatch.delenv("XNB_NO_FUSED_FIRST_HALF", raising=False)
        _, ctx = setup_pair(kw)
        ctx.first_iteration(eps, sig, rc)
        l0 = ctx.kernel_launches()
        if mode == "single":
            rb = sum(ctx.run_steps(1, dt, eps, sig, rc) for _ in range(30))
        else:
            rb = ctx.run_steps(23, dt, eps, sig, rc) + ctx.run_steps(7, dt, eps, sig, rc)
        out.append((rb, ctx.get_particles(0, ctx.n_inner), ctx.kernel_launches() - l0, ctx.energy_virial(eps, sig, rc)))
        ctx.close()
    assert out[0][0] == out[1][0] == out[2][0] and (out[0][0] > 0 or case == "ni16k")      # (the Ni deck does not rebuild within 30 steps)
    for other in (1, 2):
        for k in ("id", "rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz"):
            assert np.array_equal(out[0][1][k], out[other][1][k]), (k, other)
        assert all(np.array_equal(np.asarray(a), np.asarray(b)) for a, b in zip(out[0][3], out[other][3]))
    if case != "ni16k":                      # (large cells: the plane-staged sweep has no fused form)
        assert out[0][2] < out[1][2]         # the fused loop launches fewer kernels


@pytest.mark.parametrize("case", ["ni16k", "lj2k", "lj_gap2", "lj_voids"])
def test_newton3_path(case):
    """chunk_neighbors with ChunkNeighborsConfig::half_symmetric / skip_ghosts (neighbor_filter_func.h:36-52)
    builds byte-identical GridChunkNeighbors streams; zero_particle_force{ghost} + the symmetric LJ sweep (each pair once, f_b -=
    via FP64 atomics where the reference takes per-cell locks) + update_force_from_ghost + divide by mass gives the oracle's
    Newton-3 forces, which are the full-list forces"""
    kw = CASES[case]
    eps, sig, rc = kw["epsilon"], kw["sigma"], kw["rcut"]
    o, ctx = setup_pair(kw)
    o.move_particles(); o.update_particles_full()
    ctx.move_particles(); ctx.update_particles_full()
    # full-list forces of the standard path
    ctx.zero_particle_force(ghost=True); ctx.lennard_jones_force(eps, sig, rc); ctx.divide_force_by_mass()
    f_full = U.vec(U.by_id(ctx.get_particles(0, ctx.n_inner)), ("fx", "fy", "fz"))
    n_full = int(ctx.streams()[0].sum())
    # ---- Newton-3 forces
    o.set_nbh_config(half_symmetric=True); o.build_neighbors(); o.compute_force_symmetric()
    po = U.by_id(o.particles(), o.inner_mask())
    ctx.set_chunk_neighbors_config(half_symmetric=True)
    with pytest.raises(Exception):
        ctx.lennard_jones_force_symmetric(eps, sig, rc)            # the full lists are void now
    ctx.chunk_neighbors()
    assert int(ctx.streams()[0].sum()) < n_full
    with pytest.raises(Exception):
        ctx.lennard_jones_force(eps, sig, rc)                      # ... and the full-list sweep refuses half lists
    for functor in (0, 1):
        ctx.set_pair_functor(functor)
        ctx.zero_particle_force(ghost=True); ctx.lennard_jones_force_symmetric(eps, sig, rc); ctx.update_force_from_ghost(); ctx.divide_force_by_mass()
        pg = U.by_id(ctx.get_particles(0, ctx.n_inner))
        assert np.array_equal(po["id"], pg["id"])
        f_sym = U.vec(pg, ("fx", "fy", "fz"))
        assert U.force_error(f_sym, U.vec(po, ("fx", "fy", "fz"))) < TOL
        assert U.force_error(f_sym, f_full) < TOL
    ctx.set_pair_functor(0)
    # ---- streams of the three filtered configurations, byte for byte, with the oracle in the GPU's in-cell order
    pg_all, cnt_g = U.gpu_particles_cell_order(ctx)
    o.set_particles(cnt_g, pg_all)
    for half, skip in ((True, False), (False, True), (True, True), (False, False)):
        o.set_nbh_config(half_symmetric=half, skip_ghosts=skip); o.build_neighbors()
        rc_, msg = o.check_streams(); assert rc_ == 0, msg
        ctx.set_chunk_neighbors_config(half_symmetric=half, skip_ghosts=skip); ctx.chunk_neighbors()
        sz_o, data_o = o.streams(); sz_g, data_g = ctx.streams()
        assert np.array_equal(sz_o, sz_g), (half, skip)
        assert np.array_equal(data_o, data_g), (half, skip)
        assert ctx.view_chunk_neighbors()[2] == o.max_neighbors()
    assert int(ctx.streams()[0].sum()) == n_full
    # the symmetric sweep refuses full lists
    with pytest.raises(Exception):
        ctx.lennard_jones_force_symmetric(eps, sig, rc)


@pytest.mark.parametrize("sequential", [False, True])
def test_step_host_equals_run_steps(sequential, monkeypatch):
    """xnb_step_host (particles resident in host memory, positions copied back while the sweep runs) = upload + xnb_run_steps(1)
    + download, bit for bit, on steps with and without a rebuild; ids are rewritten exactly when the step rebuilt"""
    import torch
    if sequential:
        monkeypatch.setenv("XNB_NO_SPECULATION", "1")
    kw = CASES["lj2k"]
    eps, sig, rc, dt = kw["epsilon"], kw["sigma"], kw["rcut"], kw["dt"]
    _, ca = setup_pair(kw)
    _, cb = setup_pair(kw)
    ca.first_iteration(eps, sig, rc); cb.first_iteration(eps, sig, rc)
    n = ca.n_inner
    names = ("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz")
    hb = {k: torch.zeros(n, dtype=torch.float64).pin_memory() for k in names}
    hid = torch.zeros(n, dtype=torch.int64).pin_memory()
    p0 = cb.get_particles(0, n)
    for k in names:
        hb[k].numpy()[:] = p0[k]
    hid.numpy()[:] = p0["id"].astype(np.int64)
    P = {k: v.data_ptr() for k, v in hb.items()}
    total = 0
    for it in range(30):
        rb_a = ca.run_steps(1, dt, eps, sig, rc)
        id_before = hid.numpy().copy()
        rb_b = cb.step_host(dt, eps, sig, rc, in_r=(P["rx"], P["ry"], P["rz"]), in_v=(P["vx"], P["vy"], P["vz"]),
                            out_r=(P["rx"], P["ry"], P["rz"]), out_v=(P["vx"], P["vy"], P["vz"]), out_f=(P["fx"], P["fy"], P["fz"]),
                            out_id=hid.data_ptr())
        assert rb_a == rb_b, it
        total += rb_b
        pa = ca.get_particles(0, n)
        for k in names:
            assert np.array_equal(pa[k], hb[k].numpy()), (it, k)
        assert np.array_equal(pa["id"].astype(np.int64), hid.numpy()), it
        if not rb_b:
            assert np.array_equal(id_before, hid.numpy())
    assert total > 0
    # only part of the outputs wanted, inputs kept on the device
    rb_a = ca.run_steps(1, dt, eps, sig, rc)
    hb["fx"].zero_(); hid.zero_()
    rb_b = cb.step_host(dt, eps, sig, rc, out_f=(P["fx"], None, None), out_id=hid.data_ptr(), id_always=True)
    pa = ca.get_particles(0, n)
    assert rb_a == rb_b and np.array_equal(pa["fx"], hb["fx"].numpy()) and np.array_equal(pa["id"].astype(np.int64), hid.numpy())


def test_sweep_info():
    kw = CASES["lj2k"]
    _, ctx = setup_pair(kw)
    assert not ctx.sweep_info()["compiled"]                 # no lists yet
    ctx.first_iteration(kw["epsilon"], kw["sigma"], kw["rcut"])
    si = ctx.sweep_info()
    assert si["compiled"] and not si["ghost"]
    assert si["blocks"] == si["interior_tiles"] + si["boundary_tiles"] and si["threads"] % 32 == 0
    # list entries of the inner particles = stream words minus group counters, (code, count) headers and the offset tables
    pairs = sum(len(v) for v in ctx_pairs(ctx))
    assert si["candidates"] == pairs
    assert si["rows"] * 128 >= si["candidates"]


def ctx_pairs(ctx):
    """per inner particle neighbour lists decoded from the GridChunkNeighbors streams (pure python reader)"""
    gi = ctx.grid_info(); d, gl = gi["dims"], gi["ghost_layers"]
    start, cnt = ctx.cells()
    sz, data = ctx.streams()
    off = np.concatenate([[0], np.cumsum(np.asarray(sz, np.int64))])
    lists = []
    for c in range(gi["n_cells"]):
        i, j, k = c % d[0], (c // d[0]) % d[1], c // (d[0] * d[1])
        if not (gl <= i < d[0] - gl and gl <= j < d[1] - gl and gl <= k < d[2] - gl) or cnt[c] == 0:
            continue
        s = data[off[c]:off[c] + sz[c]]
        n = int(cnt[c])
        table = s[:2 * (n + 1)].view(np.uint32)
        body = s[2 * (n + 1):]
        for p in range(n):
            w = body[table[p] - 1:table[p + 1] - 1]
            g, q, cand = int(w[0]), 1, []
            for _ in range(g):
                m = int(w[q + 1]); cand.extend(w[q + 2:q + 2 + m].tolist()); q += 2 + m
            lists.append(cand)
    return lists


def test_reference_golden_file_through_the_cuda_path():
    """the reference's own regression deck, run end to end on the GPU: check_values_lj_Ni.dat within the deck's 1e-5"""
    from test_oracle_kat import compare_with_golden
    kw = ni_deck_kwargs()
    ctx = U.make_ctx(kw)
    ctx.first_iteration(kw["epsilon"], kw["sigma"], kw["rcut"])
    ctx.run_steps(100, kw["dt"], kw["epsilon"], kw["sigma"], kw["rcut"])
    p = ctx.get_particles(0, ctx.n_inner)
    (re, ae, ve), (rl2, al2, vl2) = compare_with_golden(p, np.ones(len(p["id"]), bool), 55.68)
    assert max(re, ae, ve, rl2, al2, vl2) < 1e-5, (re, ae, ve)
    assert re < 1e-10 and ae < 1e-6 and ve < 1e-8, (re, ae, ve)


def test_error_behaviour():
    from exanbody_b200 import capi
    kw = CASES["lj2k"]
    ctx = U.make_ctx(kw)
    with pytest.raises(capi.XnbError) as e:
        ctx.lennard_jones_force(1.0, 1.0, 2.5)          # no neighbour list yet
    assert e.value.code == 2
    c2 = capi.Context(0)
    with pytest.raises(capi.XnbError):
        c2.set_domain((0, 0, 0), (10, 10, 10), 3.0, (4, 4, 4))     # bounds do not match grid (check_domain)
    c2.set_domain((0, 0, 0), (4, 4, 4), 1.0, (4, 4, 4))
    c2.set_nbh_dist(20.0, 1.0)                                        # 21 cell layers: beyond the +-15 of the 5-bit codec
    with pytest.raises(capi.XnbError) as e:
        c2.grid_info()
    assert e.value.code == 4
    # non periodic domain: a particle that leaves is reported, not silently dropped
    kw2 = dict(kw, periodic=(0, 0, 0))
    inp = U.generate_input(kw2)
    inp["rx"][0] = -1.0
    c3 = U.make_ctx(kw2, particles=inp)
    with pytest.raises(capi.XnbError) as e:
        c3.move_particles()
    assert e.value.code == 5
    # empty input is fine
    c4 = capi.Context(0)
    c4.set_domain((0, 0, 0), (8, 8, 8), 2.0, (4, 4, 4)); c4.set_nbh_dist(1.5, 0.3)
    c4.set_particles(np.zeros(0), np.zeros(0), np.zeros(0))
    c4.first_iteration(1.0, 1.0, 1.5)
    assert c4.n_total == 0 and c4.run_steps(3, 0.005, 1.0, 1.0, 1.5) == 0


@pytest.mark.parametrize("buffer_form", [False, True])
@pytest.mark.parametrize("case", ["lj2k", "lj_gap2", "lj_voids", "ni16k"])
def test_gravitational_force_second_functor(case, buffer_form):
    """a second functor of the concept, one that reads a per-neighbour field -- gravitational_force
    (contribs/pi/gravitational_force.cu): masses by particle TYPE of the central particle and of the neighbour, through both call forms
    (buffer-less / ComputePairBuffer2) of the general pair sweep, against the oracle's restatement on the same lists.  Three types
    with different masses; forces within 1e-10, the two call forms within rounding of each other."""
    kw = CASES[case]
    inp = U.generate_input(kw)
    inp["type"] = (inp["id"] % 3).astype(np.uint8)
    type_mass = np.array([1.0, 2.5, 0.125])
    G, rcut = 0.75, kw["rcut"]
    o = U.make_oracle(kw)
    o.generate()
    ctx = U.make_ctx(kw, particles=inp)
    ctx.set_type_mass(type_mass)
    ctx.move_particles(); ctx.update_particles_full()
    # the oracle takes the GPU's grid content (in-cell order and types included), builds its own lists and sweeps them
    o.move_particles(); o.update_particles_full()
    pcell, cnt = U.gpu_particles_cell_order(ctx)
    assert set(np.unique(pcell["type"])) == {0, 1, 2}
    o.set_particles(cnt, pcell); o.build_neighbors()
    o.zero_force(); o.gravitational_force(G, rcut, type_mass)
    ctx.zero_particle_force(True); ctx.gravitational_force(G, rcut, buffer_form=buffer_form)
    po = U.by_id(o.particles(), o.inner_mask()); pg = U.by_id(ctx.get_particles(0, ctx.n_inner))
    assert np.array_equal(po["id"], pg["id"]) and np.array_equal(po["type"], pg["type"])
    fo, fg = U.vec(po, ("fx", "fy", "fz")), U.vec(pg, ("fx", "fy", "fz"))
    assert np.abs(fo).max() > 0
    assert U.force_error(fg, fo) <= TOL
    # ACCUMULATES like the reference operator: a second call doubles the forces
    ctx.gravitational_force(G, rcut, buffer_form=not buffer_form)
    pg2 = U.by_id(ctx.get_particles(0, ctx.n_inner))
    assert U.force_error(U.vec(pg2, ("fx", "fy", "fz")), 2.0 * fo) <= TOL


@pytest.mark.parametrize("field,weights", [("vx", None), ("id", (1.0, 0.0, -0.05, 0.0)), ("type", (2.0, -0.3, 0.0, 0.01)), ("rz", (0.0, 1.0))])
@pytest.mark.parametrize("case", ["lj2k", "lj_voids"])
def test_average_neighbors_third_functor(case, field, weights):
    """SURVEY 8(f) rank 3, the remaining call form of the functor concept: a PARTICLE CONTEXT (Start / per pair / Stop) and a per-neighbour
    scalar field -- average_neighbors_scalar (src/compute/average_neighbors.cu:38-175): avg[a] = sum w(d) field[b] / sum w(d) over the
    listed neighbours within rcut, w a cubic in d.  Against the oracle's restatement on the same lists (same summation order: equal to
    rounding of the final division), and -- to pin the oracle itself -- against a brute-force evaluation over ALL particle pairs."""
    from scipy.spatial import cKDTree
    kw = CASES[case]
    inp = U.generate_input(kw)
    inp["type"] = (inp["id"] % 3).astype(np.uint8)
    rcut = kw["rcut"]
    o = U.make_oracle(kw)
    o.generate()
    ctx = U.make_ctx(kw, particles=inp)
    ctx.move_particles(); ctx.update_particles_full()
    o.move_particles(); o.update_particles_full()
    pcell, cnt = U.gpu_particles_cell_order(ctx)
    o.set_particles(cnt, pcell); o.build_neighbors()
    w4 = (1.0, 0.0, 0.0, 0.0) if weights is None else tuple(weights) + (0.0,) * (4 - len(weights))
    avg_o = o.average_neighbors(rcut, field, w4)
    avg_g = ctx.average_neighbors(rcut, field, weights)
    allo = o.particles(); mask = o.inner_mask()
    po = U.by_id(dict(allo, avg=avg_o), mask)
    pg = U.by_id(dict(ctx.get_particles(0, ctx.n_inner), avg=avg_g))
    assert np.array_equal(po["id"], pg["id"])
    scale = max(np.abs(po["avg"]).max(), 1e-300)
    assert scale > 0 and np.abs(pg["avg"] - po["avg"]).max() <= 1e-13 * scale
    # brute force over every particle (ghost images stand for the periodic copies)
    pa = ctx.get_particles(0, ctx.n_total)
    xyz = np.stack([pa["rx"], pa["ry"], pa["rz"]], 1)
    val = pa[field].astype(np.float64)
    tree = cKDTree(xyz)
    ni = ctx.n_inner
    ref = np.zeros(ni)
    for a, nb in enumerate(tree.query_ball_point(xyz[:ni], rcut * (1 + 1e-9))):
        nb = np.asarray(nb, np.int64)
        d2 = ((xyz[nb] - xyz[a]) ** 2).sum(1)
        nb = nb[(d2 > 0) & (d2 <= rcut * rcut)]; d2 = ((xyz[nb] - xyz[a]) ** 2).sum(1)
        d = np.sqrt(d2)
        w = w4[0] + w4[1] * d + w4[2] * d2 + w4[3] * d2 * d
        ref[a] = (w * val[nb]).sum() / w.sum() if w.sum() > 0 else (w * val[nb]).sum()
    assert np.abs(ref - avg_g).max() <= 1e-10 * scale
    o.close(); ctx.close()


@pytest.mark.parametrize("name,steps", [("C5", 40), ("C4", 3), ("C1", 6)])
def test_full_size_workloads_against_the_oracle(name, steps):
    """parity where the numbers are quoted (BASELINE.json configs at their stated size, the inputs bench.py times): the CUDA path and
    the CPU oracle step the same input the same number of steps -- same atoms, SAME REBUILD COUNT (the C5 bar of SURVEY 8d), forces on
    identical positions within 1e-10, neighbour streams of the final configuration byte for byte.  Exercises what only large inputs
    reach: capacity regrowth of the tiled builds, the large-cell build (C1, C4: 256 atoms per cell), ragged occupancy (C5)."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    kw, _ = bench.workload(name)
    r = bench.run_cpu_oracle(kw, steps, 600.0, keep=True)
    assert r["steps"] == steps
    out = bench.parity_against_oracle(r["oracle"], steps, r["rebuilds"], kw, 0)
    r["oracle"].close()
    assert out["atoms_equal"] and out["rebuilds_equal"], out
    assert out["streams_equal"], out
    assert out["max_force_error_same_positions"] < TOL, out
    assert out["ok"], out
    if name in ("C4", "C5"):
        assert r["rebuilds"] > 0


@pytest.mark.parametrize("planes", ["XNB_CL_PLANES", "XNB_CL_NO_PLANES"])
@pytest.mark.parametrize("case", ["ni16k", "lj_dense"])
def test_large_cell_sweeps_equal_the_stream_sweep(case, planes, monkeypatch):
    """large cells (256 per cell): k_nbh_big emits the compiled rows either in one segment (swept by k_lj_sweep_cl, whole halo staged) or in
    one segment per z-plane of halo cells (k_lj_sweep_pl, a third of the halo staged at a time).  Both hold the stream's candidates in
    the stream's order: trajectories (several rebuilds), forces, energy and virial bit-identical to the sweep over the streams."""
    kw = CASES[case]
    eps, sig, rc, dt = kw["epsilon"], kw["sigma"], kw["rcut"], kw["dt"]
    res = []
    for streams in (True, False):
        monkeypatch.delenv("XNB_SWEEP_STREAMS", raising=False); monkeypatch.delenv("XNB_CL_PLANES", raising=False); monkeypatch.delenv("XNB_CL_NO_PLANES", raising=False)
        monkeypatch.setenv("XNB_SWEEP_STREAMS" if streams else planes, "1")
        _, ctx = setup_pair(kw)
        ctx.first_iteration(eps, sig, rc)
        rb = ctx.run_steps(12, dt, eps, sig, rc)
        si = ctx.sweep_info()
        assert bool(si["compiled"]) == (not streams)
        res.append((ctx.get_particles(0, ctx.n_inner), ctx.energy_virial(eps, sig, rc), rb))
    (pa, eva, ra), (pb, evb, rbb) = res
    assert ra == rbb
    for k in ("id", "rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz"):
        assert np.array_equal(pa[k], pb[k]), k
    assert abs(eva[0] - evb[0]) <= 1e-12 * abs(eva[0]) and np.abs(eva[1] - evb[1]).max() <= 1e-12 * np.abs(eva[1]).max()
