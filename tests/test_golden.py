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
