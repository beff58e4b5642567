"""How the unit constants in tests/conftest.py were identified (run from the repo root; needs only the oracle).

The error against check_values_lj_Ni.dat is linear in the relative change of epsilon/mass and vanishes at +7.86e-6
relative to CODATA-2018, which is e = 1.6021892e-19 C (CODATA 1973) over u = 1.66053904e-27 kg (CODATA 2014)."""
import os, sys
import numpy as np
import yaml
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O

gold = yaml.safe_load(open(os.path.join(os.path.dirname(__file__), "check_values_lj_Ni.dat")))["values"]


def run(ev_int):
    o = O.Oracle(O.make_config(bounds_max=(55.68,) * 3, cell_size=13.92, grid_dims=(4, 4, 4), lattice_a=3.48,
                               epsilon=0.3729 * ev_int, sigma=2.2808, rcut=4.1, rcut_inc=2.0, dt=2e-3, mass=58.693,
                               noise_sigma=0.05, max_neighbors=256))
    o.init(); o.run(100)
    p = o.particles(); m = o.inner_mask()
    idx = {int(i): k for k, i in enumerate(p["id"]) if m[k]}
    re = ae = ve = 0.0
    for row in gold:
        ref = [float.fromhex(x) for x in row[1:]]; k = idx[int(row[0])]
        d = np.array([p["rx"][k], p["ry"][k], p["rz"][k]]) - ref[0:3]; d -= 55.68 * np.round(d / 55.68)
        a = np.array([p["fx"][k], p["fy"][k], p["fz"][k]]) - ref[3:6]
        v = np.array([p["vx"][k], p["vy"][k], p["vz"][k]]) - ref[6:9]
        re = max(re, np.linalg.norm(d)); ae = max(ae, np.linalg.norm(a)); ve = max(ve, np.linalg.norm(v))
    return re, ae, ve


if __name__ == "__main__":
    base = 1.602176634e-19 / 1.66053906660e-23
    for d in (-1e-5, 0.0, 3e-6, 1e-5, 3e-5):
        print("CODATA2018 * (1%+.0e)" % d, run(base * (1 + d)))
    for e, u in ((1.6021892e-19, 1.66053906660e-27), (1.6021892e-19, 1.66053904e-27)):
        print("e=%g u=%g" % (e, u), run(e / (u * 1e4)))
