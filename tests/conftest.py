import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


_HAVE_GPU = None


def have_gpu():
    """one probe per session through the C-ABI itself: xnb_create answers XNB_ERR_NO_DEVICE on a box without a GPU"""
    global _HAVE_GPU
    if _HAVE_GPU is None:
        try:
            from exanbody_b200 import capi
            capi.Context(0).close()
            _HAVE_GPU = True
        except Exception:
            _HAVE_GPU = False
    return _HAVE_GPU


def pytest_collection_modifyitems(config, items):
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if gpu_items and not have_gpu():
        skip = pytest.mark.skip(reason="no CUDA device (the hot path has no CPU fallback)")
        for it in gpu_items:
            it.add_marker(skip)


# Unit constants recovered from the reference's golden file (tests/golden/README.md): with these the oracle reproduces
# check_values_lj_Ni.dat to 1e-10; onika (not in the reference tree) holds the originals.
ELEMENTARY_CHARGE = 1.6021892e-19   # C
ATOMIC_MASS = 1.66053904e-27        # kg
EV_INTERNAL = ELEMENTARY_CHARGE / (ATOMIC_MASS * 1e-20 / 1e-24)   # 1 eV in Da*ang^2/ps^2


def ni_deck_kwargs(cells=4):
    """contribs/microStamp/samples/benchmark_lj_snap/input_lj_Ni.msp (cells=4: verbatim 16384-atom deck)"""
    return dict(bounds_max=(13.92 * cells,) * 3, cell_size=13.92, grid_dims=(cells,) * 3, lattice_a=3.48,
                epsilon=0.3729 * EV_INTERNAL, sigma=2.2808, rcut=4.1, rcut_inc=2.0, dt=2e-3, mass=58.693,
                noise_sigma=0.05, max_neighbors=256)


def lj_reduced_kwargs(ncell_units=8, cell_units=2, rcut=2.5, skin=0.3, noise=0.02, vel_sigma=1.2, **kw):
    """synthetic reduced-unit LJ FCC config: rho*=0.8442, cell = cell_units lattice constants"""
    a = (4.0 / 0.8442) ** (1.0 / 3.0)
    nc = ncell_units // cell_units
    d = dict(bounds_max=(a * ncell_units,) * 3, cell_size=a * cell_units, grid_dims=(nc,) * 3, lattice_a=a,
             epsilon=1.0, sigma=1.0, rcut=rcut, rcut_inc=skin, dt=0.005, mass=1.0, noise_sigma=noise,
             vel_sigma=vel_sigma, max_neighbors=1024)
    d.update(kw)
    return d
