"""Residual IVFPQ (pq_use_residual=True): ours == oracle == the reference's own
`ivfpq_topk_residual_precomputed` kernel (oracle/_ref)."""
import numpy as np
import pytest
import torch

from oracle import ivfpq_oracle as O, build_state as B, ref_kernels as R
from helpers import assert_close_results

pytestmark = pytest.mark.gpu


def make_residual_index(st, device="cuda:0"):
    import torchpq_b200 as T
    ix = T.IVFPQIndex(st.d_vector, st.n_subvectors, st.n_cells, initial_size=1, device=device, pq_use_residual=True)
    return ix.load_state(st)


@pytest.mark.parametrize("M,d,smart,nq", [
    (8, 32, True, 64), (64, 128, False, 64), (16, 64, True, 64),     # query half of the LUT built in the CTA (d/M 4, 2, 4)
    (96, 192, True, 40), (120, 240, False, 40),                      # M > 64: staged query half, one table per CTA
    (8, 24, True, 40), (64, 512, True, 24), (120, 960, True, 16),    # d/M = 3, 8, 8
    (16, 64, False, 3), (96, 192, True, 2),                          # tiny batches: every query is cut into slices
])
def test_residual_search_matches_oracle_and_reference_kernel(cuda_device, M, d, smart, nq):
    torch.manual_seed(M)
    base = torch.randn(d, 5000)
    st = B.build_state_residual(base, M, 16, vq_iters=3, pq_iters=2 if M > 64 else 3)
    st.n_probe, st.use_smart_probing = 6, smart
    x = torch.randn(d, nq)
    k = 20
    ov, oi, oa = O.search_residual(st, x, k=k + 1, return_address=True)
    ix = make_residual_index(st)
    v, i, a = ix.search(x.cuda(), k=k, return_address=True)
    assert_close_results(v.cpu().numpy(), i.cpu().numpy(), ov[:, :k], oi[:, :k], rtol=1e-3, min_overlap=0.995)
    if R.available(M):
        xx, sims, cells, npl = O.coarse_probe(st, x)
        p1, p2 = O.residual_parts(xx, torch.from_numpy(st.vq_codebook), torch.from_numpy(st.pq_codebook))
        cn = cells.numpy()
        g = lambda t: torch.as_tensor(t).cuda()
        rv, ra = R.ivfpq_topk_residual_precomputed(g(st.storage), p1.cuda(), p2.cuda(), cells.cuda(), sims.cuda(),
                                                   g(st.cell_start[cn]), g(st.cell_size[cn]), g(st.is_empty), npl.cuda(), k)
        rv, ra = rv.cpu().numpy(), ra.cpu().numpy()
        assert np.array_equal(rv, ov[:, :k])                       # oracle == reference kernel, bit for bit
        tf = np.all(np.diff(ov, axis=1) != 0, axis=1)
        assert np.array_equal(ra[tf], oa[tf, :k])
        assert np.allclose(v.cpu().numpy(), rv, rtol=1e-3, atol=0)
        assert (a.cpu().numpy()[tf] == ra[tf]).mean() >= 0.995


def test_residual_train_add_search_and_decode(cuda_device):
    import torchpq_b200 as T
    import bench
    torch.manual_seed(0)
    base = torch.randn(32, 20000, device="cuda")
    ix = T.IVFPQIndex(32, 16, 64, initial_size=64, device="cuda:0", pq_use_residual=True)
    ix.train(base[:, :8000].contiguous())
    ix.add(base)
    ix.n_probe = 16
    x = torch.randn(32, 100, device="cuda")
    v, i = ix.search(x, k=10)
    ov, oi = O.search_residual(bench.to_oracle_state(ix), x.cpu(), k=10)
    assert_close_results(v.cpu().numpy(), i.cpu().numpy(), ov, oi, rtol=1e-3, min_overlap=0.995)
    codes, cells = ix.encode(base[:, :100].contiguous())
    rec = ix.decode((codes, cells))
    err_res = (rec - base[:, :100]).norm() / base[:, :100].norm()
    assert err_res < 0.6                                          # residual quantisation reconstructs the vectors


def test_residual_probe_list_zero_and_nan(cuda_device):
    """n_probe_list entries 0 and INT64_MIN (the NaN-derived value, IVFPQIndex.py:508-510): the residual kernel's loop is
    `for cCell < nProbe` with nProbe read as a 32-bit int (ivfpq_topk.cu:1080), so both scan nothing -> (-inf, -1)
    rows; other rows are unaffected."""
    torch.manual_seed(11)
    base = torch.randn(32, 4000)
    st = B.build_state_residual(base, 8, 16)
    st.n_probe, st.use_smart_probing = 6, False
    x = torch.randn(32, 12)
    xx, sims, cells, npl = O.coarse_probe(st, x)
    npl = npl.clone()
    npl[0], npl[1], npl[2], npl[3] = 0, O.INT64_MIN, 1, 7          # 7 > n_probe: clamped
    cn = cells.numpy()
    p1, p2 = O.residual_parts(xx, torch.from_numpy(st.vq_codebook), torch.from_numpy(st.pq_codebook))
    ov, oa = O.ivfpq_topk_residual_precomputed(st.storage, p1.numpy(), p2.numpy(), cn, sims.numpy(), st.is_empty,
                                               st.cell_start[cn], st.cell_size[cn], npl.numpy(), 15)
    ix = make_residual_index(st)
    v, i, a = ix.search_cells(x.cuda(), cells.cuda(), base_sims=sims.cuda(), n_probe_list=npl.cuda(), k=15, return_address=True)
    v, a = v.cpu().numpy(), a.cpu().numpy()
    assert np.isinf(v[:2]).all() and (a[:2] == -1).all() and (i[:2].cpu().numpy() == -1).all()
    assert np.isinf(ov[:2]).all()
    assert np.allclose(v[2:], ov[2:], rtol=1e-3, atol=0)
    tf = np.all(np.diff(ov[2:], axis=1) != 0, axis=1)
    assert (a[2:][tf] == oa[2:][tf]).mean() >= 0.995
