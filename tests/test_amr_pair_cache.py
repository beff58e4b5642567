"""op amr_grid_pairs: the AmrSubCellPairCache of amr/lib/amr_grid_algorithm.cpp:102-218 as the library's host function produces
it, against (1) the oracle's cache built inside its neighbour pipeline for the reference Ni deck and (2) an independent numpy
restatement, plus the structural facts the reference relies on (list count, codes, symmetry at offset 0)."""
import numpy as np
import pytest

from conftest import ni_deck_kwargs
from exanbody_b200 import capi
import parity_util as U


def numpy_lists(max_res, cs, md):
    layers = int(np.ceil(md / cs))
    out = []
    for rb in range(1, max_res + 1):
        for ra in range(1, rb + 1):
            sa, sb = cs / ra, cs / rb
            ga = np.stack(np.meshgrid(np.arange(ra), np.arange(ra), np.arange(ra), indexing="ij"), -1).reshape(-1, 3)[:, ::-1]   # rows ordered k, j, i; columns i, j, k
            gb = np.stack(np.meshgrid(np.arange(rb), np.arange(rb), np.arange(rb), indexing="ij"), -1).reshape(-1, 3)[:, ::-1]
            for ck in range(layers + 1):
                for cj in range(layers + 1):
                    for ci in range(layers + 1):
                        alo = ga * sa; ahi = (ga + 1) * sa
                        blo = np.array([ci, cj, ck]) * cs + gb * sb; bhi = np.array([ci, cj, ck]) * cs + (gb + 1) * sb
                        gap = np.maximum(0.0, np.maximum(blo[None] - ahi[:, None], alo[:, None] - bhi[None]))
                        d2 = gap[..., 0] ** 2 + gap[..., 1] ** 2 + gap[..., 2] ** 2
                        ia, ib = np.nonzero(d2 <= md * md)
                        code = lambda g: (g[:, 2] << 10) | (g[:, 1] << 5) | g[:, 0]
                        out.append(np.stack([code(ga[ia]), code(gb[ib])], 1).ravel().astype(np.uint16))
    return out


@pytest.mark.parametrize("max_res,cs,md", [(1, 3.0, 2.8), (3, 13.92, 6.1), (4, 2.0, 3.1)])
def test_matches_an_independent_restatement(max_res, cs, md):
    off, data = capi.amr_sub_cell_pairs(max_res, cs, md)
    ref = numpy_lists(max_res, cs, md)
    layers = int(np.ceil(md / cs))
    assert len(off) - 1 == len(ref) == max_res * (max_res + 1) // 2 * (layers + 1) ** 3
    for q, r in enumerate(ref):
        assert np.array_equal(data[int(off[q]):int(off[q + 1])], r), q
    # offset (0,0,0) with res_a == res_b: the relation is symmetric
    first = data[int(off[0]):int(off[1])].reshape(-1, 2)
    assert set(map(tuple, first)) == set(map(tuple, first[:, ::-1]))


def test_matches_the_oracle_cache_of_the_reference_deck():
    kw = ni_deck_kwargs()
    o = U.make_oracle(kw)
    o.generate(); o.move_particles(); o.update_particles_full()
    max_res, off_o, data_o = o.amr_pair_cache()
    assert max_res == 3                                   # 256 atoms per cell -> sub-grid side 3 (amr_grid_algorithm.h:66-78)
    off, data = capi.amr_sub_cell_pairs(max_res, kw["cell_size"], kw["rcut"] + kw["rcut_inc"])
    assert np.array_equal(off, off_o[:len(off)]) and np.array_equal(data, data_o)
    with pytest.raises(ValueError):
        capi.amr_sub_cell_pairs(0, 1.0, 1.0)
