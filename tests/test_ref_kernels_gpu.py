"""Pin the oracle (and the product) against the REFERENCE'S OWN kernel, compiled unmodified from
/root/reference into oracle/_ref/ (oracle/build_ref.py) and executed here on the GPU."""
import numpy as np
import pytest
import torch

from oracle import ivfpq_oracle as O, build_state as B, ref_kernels as R
from helpers import make_index

pytestmark = pytest.mark.gpu


def _inputs(st, x):
    xx, sims, cells, npl = O.coarse_probe(st, x)
    lut = O.precompute_adc(xx, torch.from_numpy(st.pq_codebook), st.distance)
    cn = cells.numpy()
    return xx, lut, st.cell_start[cn], st.cell_size[cn], npl, cells


@pytest.mark.parametrize("M,d", [(8, 32), (16, 64), (64, 128), (120, 240)])
@pytest.mark.parametrize("kind", ["integer", "randn"])
def test_reference_kernel_vs_oracle_vs_ours(cuda_device, M, d, kind):
    if not R.available(M):
        pytest.skip("oracle/_ref not built")
    import torchpq_b200 as T
    k = 20
    if kind == "integer":
        st, queries = B.integer_state(d, M, 16, 4000, seed=M, lo=-12, hi=13)
        x = queries(96)
    else:
        torch.manual_seed(M)
        st = B.build_state(torch.randn(d, 4000), M, 16, vq_iters=2, pq_iters=1)
        x = torch.randn(d, 96)
    st.n_probe = 6
    xx, lut, cs, cz, npl, cells = _inputs(st, x)
    ov, oa = O.ivfpq_topk(st.storage, lut.numpy(), st.is_empty, cs, cz, npl.numpy(), k + 1)
    g = lambda a: torch.as_tensor(a).cuda()
    rv, ra = R.ivfpq_topk(g(st.storage), lut.cuda(), g(cs), g(cz), g(st.is_empty), npl.cuda(), k)
    rv, ra = rv.cpu().numpy(), ra.cpu().numpy()
    # (1) oracle == reference kernel: identical fp32 values (same m-ascending sum) on every row, and
    #     identical addresses IN ORDER on rows whose top-(k+1) scores are pairwise distinct.  (Among tied
    #     scores the reference's bitonic network returns duplicated addresses -- observed here on B200,
    #     DESIGN.md "reference quirks" -- so tied rows can only be compared by value.)
    assert np.array_equal(rv, ov[:, :k])
    ok = np.all(np.diff(ov[:, :k + 1], axis=1) != 0, axis=1)
    assert ok.sum() >= 8, f"only {ok.sum()} tie-free rows"
    assert np.array_equal(ra[ok], oa[ok, :k])
    # (2) ours (reference-layout op) == reference kernel, bit for bit on values
    v1, a1 = T.fn.IVFPQTopk(M).topk(g(st.storage), lut.cuda(), g(cs), g(cz), g(st.is_empty), npl.cuda(), k=k)
    assert np.array_equal(v1.cpu().numpy(), rv)
    assert np.array_equal(a1.cpu().numpy()[ok], ra[ok])
    # (3) ours (full search, scan layout) against the reference kernel: exact on integers, 1e-3 otherwise
    ix = make_index(st)
    v2, i2, a2 = ix.search_cells(xx.cuda().contiguous(), cells.cuda(), n_probe_list=npl.cuda(), k=k, return_address=True)
    v2, a2 = v2.cpu().numpy(), a2.cpu().numpy()
    if kind == "integer":
        assert np.array_equal(v2, rv)
    else:
        assert np.allclose(v2, rv, rtol=1e-3, atol=0)
    if kind == "integer":
        assert np.array_equal(a2[ok], ra[ok])
    else:   # fp32 sums in a different association order: near-ties may swap; demand >= 99.5 % agreement
        assert (a2[ok] == ra[ok]).mean() >= 0.995
