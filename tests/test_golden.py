"""CPU: the oracle against golden vectors produced by the reference's OWN kernel on a B200
(tests/golden/make_golden.py).  This is what pins the oracle."""
import glob
import os

import numpy as np
import pytest

from oracle import ivfpq_oracle as O, c_oracle as CO

FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ivfpq_topk_*.npz")))


def test_golden_files_present():
    assert len(FILES) >= 4


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_oracle_matches_reference_kernel(path, impl):
    z = np.load(path)
    k = int(z["k"])
    args = (z["storage"], z["lut"], z["is_empty"], z["cell_start"], z["cell_size"], z["n_probe_list"], k)
    if impl == "numpy":
        v, a = O.ivfpq_topk(*args)
    else:
        if not CO.available():
            pytest.skip("oracle/liboracle_scan.so not built")
        v, a, _ = CO.ivfpq_topk(*args)
    assert np.array_equal(v, z["ref_values"])                 # fp32 sums in the reference's order: bit-identical
    tf = z["tie_free"]
    assert np.array_equal(a[tf], z["ref_address"][tf])        # same addresses, same order, wherever order is defined


# ---------------------------------------------------------------------------------------------------------------
# round 2: the residual scan and the placement kernels, pinned the same way
# ---------------------------------------------------------------------------------------------------------------
RES_FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "residual_*.npz")))
PLACEMENT = os.path.join(os.path.dirname(__file__), "golden", "placement.npz")


@pytest.mark.parametrize("path", RES_FILES, ids=[os.path.basename(p) for p in RES_FILES])
def test_residual_oracle_matches_reference_kernel(path):
    """oracle.ivfpq_topk_residual_precomputed == the reference's ivfpq_topk_residual_precomputed kernel
    (kernels/cuda/ivfpq_topk.cu:1039-1207) on a B200."""
    z = np.load(path)
    v, a = O.ivfpq_topk_residual_precomputed(z["storage"], z["part1"], z["part2"], z["cells"], z["base_sims"], z["is_empty"],
                                             z["cell_start"], z["cell_size"], z["n_probe_list"], int(z["k"]))
    assert np.array_equal(v, z["ref_values"])
    tf = z["tie_free"]
    assert np.array_equal(a[tf], z["ref_address"][tf])


def test_round2_golden_files_present():
    assert len(RES_FILES) >= 2 and os.path.exists(PLACEMENT)


def test_placement_oracle_matches_reference_kernels():
    """oracle.get_ioa == get_ioa.cu:8-47 and the oracle's 'ioa-th empty slot of the cell' rule (container_add)
    == get_write_address_v2.cu:9-41, both as run on a B200."""
    z = np.load(PLACEMENT)
    cells, is_empty, start, cap = z["cells"], z["is_empty"], z["cell_start"], z["cell_capacity"]
    ioa = O.get_ioa(cells)
    assert np.array_equal(ioa, z["ref_ioa"])
    w = np.array([start[c] + np.flatnonzero(is_empty[start[c]:start[c] + cap[c]] == 1)[i] for c, i in zip(cells, ioa)])
    assert np.array_equal(w, z["ref_write_address"])
