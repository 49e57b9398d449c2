"""GPU parity on BASELINE.json's configurations *as stated* ("What's weak" #1).

  config 1: IVFPQIndex d=128, M=8, n_cells=64 on 10k vectors, search k=10        (C1)
  config 4: cosine + M=120 + d=960 (d/M = 8) + n_probe=64 + k=100, at reduced N   (C4's kernel combination:
            128 KB LUT -> one 16-warp CTA per SM, in-CTA LUT build for d/M = 8, cosine pre-normalisation)
Each in two flavours: integer-valued state (every score exact -> values and addresses bit-identical to the oracle)
and randn (values within 1e-3 relative, id overlap >= 99.5 %).  The oracle is oracle/ivfpq_oracle.py (CPU).
"""
import numpy as np
import pytest
import torch

from oracle import ivfpq_oracle as O, build_state as B
from helpers import make_index, tie_free_rows, assert_close_results

pytestmark = pytest.mark.gpu


# ----------------------------------------------------------------------------- config 1
@pytest.mark.parametrize("n_probe,smart", [(1, True), (8, True), (8, False), (64, True)])
def test_config1_integer_exact(cuda_device, n_probe, smart):
    st, queries = B.integer_state(128, 8, 64, 10_000, seed=101)
    st.n_probe, st.use_smart_probing = n_probe, smart
    x = queries(200)
    ov, oi, oa = O.search(st, x, k=11, return_address=True)
    ix = make_index(st)
    v, i, a = ix.search(x.cuda(), k=10, return_address=True)
    v, i, a = v.cpu().numpy(), i.cpu().numpy(), a.cpu().numpy()
    assert np.array_equal(v, ov[:, :10])
    assert np.array_equal(a, oa[:, :10])
    ok = tie_free_rows(ov, 10)
    assert np.array_equal(i[ok], oi[ok, :10])


@pytest.mark.parametrize("n_probe", [1, 8])
def test_config1_randn(cuda_device, n_probe):
    torch.manual_seed(1)
    base = torch.randn(128, 10_000)
    st = B.build_state(base, 8, 64, vq_iters=5, pq_iters=3, n_train=10_000)
    st.n_probe = n_probe
    x = torch.randn(128, 500)
    truth = O.exact_topk(base, x, 10, "euclidean").numpy()
    for smart in (True, False):
        st.use_smart_probing = smart
        ov, oi = O.search(st, x, k=10)
        ix = make_index(st)
        v, i = ix.search(x.cuda(), k=10)
        v, i = v.cpu().numpy(), i.cpu().numpy()
        assert_close_results(v, i, ov, oi, rtol=1e-3, min_overlap=0.995)
        assert abs(O.recall_at_k(i, truth) - O.recall_at_k(oi, truth)) <= 1e-3          # recall@10 within 0.1 %


# ----------------------------------------------------------------------------- config 4 (reduced N)
def _unit_queries(d, nq, nnz, seed):
    """+-1 on exactly `nnz` = 4^n coordinates: the L2 norm is a power of two, so x / (|x| + 1e-9) is exact in fp32
    and every cosine LUT entry / coarse score / ADC sum is an exactly representable multiple of 1/|x|."""
    rng = np.random.default_rng(seed)
    x = np.zeros((d, nq), np.float32)
    for q in range(nq):
        idx = rng.choice(d, nnz, replace=False)
        x[idx, q] = rng.choice(np.array([-1.0, 1.0], np.float32), nnz)
    return torch.from_numpy(x)


@pytest.mark.parametrize("smart", [False, True])
def test_config4_integer_exact(cuda_device, smart):
    st, _ = B.integer_state(960, 120, 128, 30_000, seed=44, distance="cosine", lo=-2, hi=3)
    st.n_probe, st.use_smart_probing = 64, smart
    x = _unit_queries(960, 48, 256, seed=5)
    ov, oi, oa = O.search(st, x, k=101, return_address=True)
    ix = make_index(st)
    lay = ix.layout()
    assert lay.m_pad == 128                                      # 128 KB LUT: the 16-warp, one-CTA-per-SM launch shape
    v, i, a = ix.search(x.cuda(), k=100, return_address=True)
    v, i, a = v.cpu().numpy(), i.cpu().numpy(), a.cpu().numpy()
    assert np.array_equal(v, ov[:, :100])
    assert np.array_equal(a, oa[:, :100])
    ok = tie_free_rows(ov, 100)
    assert np.array_equal(i[ok], oi[ok, :100])


def test_config4_randn(cuda_device):
    torch.manual_seed(4)
    base = torch.randn(960, 12_000)
    st = B.build_state(base, 120, 128, distance="cosine", vq_iters=3, pq_iters=2, n_train=4000)
    st.n_probe = 64
    x = torch.randn(960, 100)
    truth = O.exact_topk(base, x, 100, "cosine").numpy()
    for smart in (True, False):
        st.use_smart_probing = smart
        ov, oi = O.search(st, x, k=100)
        ix = make_index(st)
        v, i = ix.search(x.cuda(), k=100)
        v, i = v.cpu().numpy(), i.cpu().numpy()
        assert_close_results(v, i, ov, oi, rtol=1e-3, min_overlap=0.995)
        assert abs(O.recall_at_k(i, truth) - O.recall_at_k(oi, truth)) <= 1e-3


def test_config4_small_batch_slices(cuda_device):
    """nq = 3: each query is cut into slices (several CTAs rebuild the same 128 KB LUT) and merged."""
    st, _ = B.integer_state(960, 120, 128, 30_000, seed=45, distance="cosine", lo=-2, hi=3)
    st.n_probe, st.use_smart_probing = 64, False
    x = _unit_queries(960, 3, 64, seed=6)
    ov, oi, oa = O.search(st, x, k=100, return_address=True)
    ix = make_index(st)
    v, i, a = ix.search(x.cuda(), k=100, return_address=True)
    assert np.array_equal(v.cpu().numpy(), ov) and np.array_equal(a.cpu().numpy(), oa)


# ----------------------------------------------------------------------------- container life cycle
def test_churn_keeps_capacity_bounded(cuda_device):
    """remove + add cycles at a constant live count must not grow the container (_cell_size is a
    high-water mark and free space is counted from live items, so holes are refilled without expansion)."""
    import torchpq_b200 as T
    torch.manual_seed(3)
    base = torch.randn(32, 4000, device="cuda")
    ix = T.IVFPQIndex(32, 8, 16, initial_size=64, device="cuda:0")
    ix.train(base[:, :2000].contiguous())
    ids = ix.add(base)
    cap0 = ix.capacity
    for cycle in range(6):
        gone = ids[cycle::2][:1500]
        ix.remove(ids=gone)
        ids_new = ix.add(torch.randn(32, gone.shape[0], device="cuda"))   # fresh vectors land in different cells
        ids = torch.cat([ids[~torch.isin(ids, gone)], ids_new])
        assert ix.n_items == 4000
        assert (ix._cell_size <= ix._cell_capacity).all()
    assert ix.capacity <= 4 * cap0, (cap0, ix.capacity)
    ix.n_probe = 8
    import bench
    x = torch.randn(32, 50, device="cuda")
    v, i = ix.search(x, k=10)
    ov, oi = O.search(bench.to_oracle_state(ix), x.cpu(), k=10)
    assert_close_results(v.cpu().numpy(), i.cpu().numpy(), ov, oi, rtol=1e-3, min_overlap=0.995)


def test_state_dict_round_trip_then_remove_and_add(cuda_device):
    """load_state_dict restores _max_id from _address2id: remove(ids) and add() work right after a load."""
    import torchpq_b200 as T
    torch.manual_seed(5)
    base = torch.randn(32, 3000, device="cuda")
    a = T.IVFPQIndex(32, 8, 16, initial_size=256, device="cuda:0")
    a.train(base[:, :2000].contiguous())
    ids = a.add(base)
    sd = {k: v.clone() for k, v in a.state_dict().items()}
    b = T.IVFPQIndex(32, 8, 16, initial_size=1, device="cuda:0")
    b.load_state_dict(sd)
    assert b.max_id == 2999
    b.remove(ids=ids[:100])
    assert b.n_items == 2900
    new_ids = b.add(base[:, :10].contiguous())
    assert new_ids.min().item() == 3000                               # ids continue after the loaded maximum
    assert (b.get_address_by_id(new_ids) >= 0).all()
    b.n_probe, a.n_probe = 4, 4
    x = torch.randn(32, 20, device="cuda")
    v, i = b.search(x, k=5)
    assert not torch.isin(i, ids[:100]).any()


def test_search_cells_rejects_out_of_range_cells(cuda_device):
    """Probe entries outside [0, n_cells) are empty segments, not out-of-bounds reads."""
    st, queries = B.integer_state(32, 8, 8, 2000, seed=2)
    st.n_probe, st.use_smart_probing = 4, False
    ix = make_index(st)
    x = queries(16).cuda()
    cells = torch.randint(0, 8, (16, 4), device="cuda")
    v0, i0 = ix.search_cells(x, cells, k=10)
    bad = torch.cat([cells, torch.full((16, 1), 99, device="cuda"), torch.full((16, 1), -3, device="cuda")], 1)
    v1, i1 = ix.search_cells(x, bad, k=10)
    assert torch.equal(v0, v1) and torch.equal(i0, i1)
