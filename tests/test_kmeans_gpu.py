"""GPU parity of tpq_max_sim / tpq_compute_centroids / tpq_pq_decode against the oracle and against the
reference's own kernels (oracle/_ref)."""
import numpy as np
import pytest
import torch

from oracle import kmeans_oracle as K, ref_kernels as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("l,d,n,k", [(3, 2, 1000, 256), (2, 16, 700, 40), (1, 64, 513, 256), (4, 8, 300, 300)])
@pytest.mark.parametrize("distance", ["euclidean", "cosine"])
def test_max_sim(cuda_device, l, d, n, k, distance):
    import torchpq_b200 as T
    torch.manual_seed(l * 7 + d)
    data, cent = torch.randn(l, d, n), torch.randn(l, d, k)
    osim, olab = K.max_sim(data.numpy(), cent.numpy(), distance)
    sim, lab = T.fn.max_sim(data.cuda(), cent.cuda(), distance)
    sim, lab = sim.cpu().numpy(), lab.cpu().numpy()
    assert np.array_equal(sim, osim)                       # same fmaf chain -> bit-identical
    assert np.array_equal(lab, olab)
    if R.aux_available():                                  # the reference's own max_sim_tn kernel
        rv, ri = R.max_sim_tn(data.cuda(), cent.cuda(), "euclidean" if distance == "euclidean" else "inner", dim=2)
        assert np.array_equal(rv.cpu().numpy(), osim)
        # the reference's label write races across its 128-wide centroid tiles (max_sim.cu:173-178):
        # compare where the winner is unique
        rl = ri.cpu().numpy()
        assert (rl == olab).mean() > (0.999 if k <= 128 else 0.99)   # k > 128: two racing tiles per row


def test_max_sim_integer_ties_lowest_index(cuda_device):
    import torchpq_b200 as T
    rng = np.random.default_rng(0)
    data = torch.from_numpy(rng.integers(-2, 3, (2, 4, 500)).astype(np.float32))
    cent = torch.from_numpy(rng.integers(-2, 3, (2, 4, 300)).astype(np.float32))   # many duplicate centroids
    osim, olab = K.max_sim(data.numpy(), cent.numpy())
    sim, lab = T.fn.max_sim(data.cuda(), cent.cuda())
    assert np.array_equal(lab.cpu().numpy(), olab) and np.array_equal(sim.cpu().numpy(), osim)


@pytest.mark.parametrize("l,d,n,k", [(3, 2, 5000, 256), (2, 16, 3000, 16), (1, 130, 20000, 256), (2, 8, 100, 1000)])
def test_compute_centroids(cuda_device, l, d, n, k):
    import torchpq_b200 as T
    torch.manual_seed(d)
    data = torch.randn(l, d, n)
    labels = torch.randint(0, max(1, k - 3), (l, n))       # the last clusters stay empty -> 0
    oc = K.compute_centroids(data.numpy(), labels.numpy(), k)
    c = T.fn.compute_centroids(data.cuda(), labels.cuda(), k).cpu().numpy()
    assert np.allclose(c, oc, rtol=1e-5, atol=1e-6)
    assert (c[:, :, k - 3:] == 0).all()
    if R.aux_available() and k in (16, 256):
        rc = R.compute_centroids(data.cuda(), labels.cuda(), k).cpu().numpy()
        assert np.allclose(rc, oc, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("M,dsub,n", [(8, 4, 1000), (64, 2, 4096), (3, 8, 17), (120, 8, 333)])
def test_pq_decode(cuda_device, M, dsub, n):
    import torchpq_b200 as T
    torch.manual_seed(M)
    cb = torch.randn(M, dsub, 256)
    code = torch.randint(0, 256, (M, n), dtype=torch.uint8)
    o = K.pq_decode(cb.numpy(), code.numpy())
    g = T.fn.pq_decode(cb.cuda(), code.cuda()).cpu().numpy()
    assert np.array_equal(g, o)
    if R.aux_available():
        assert np.array_equal(R.pq_decode(cb.cuda(), code.cuda()).cpu().numpy(), o)


def _exact_sim_of(data, cent, labels):
    """fp32 fmaf-chain similarity of every point to the centroid `labels` picks (oracle arithmetic)."""
    l, d, n = data.shape
    out = np.zeros((l, n), np.float32)
    for li in range(l):
        c = cent[li][:, labels[li]]                           # [d, n]
        acc = np.zeros(n, np.float32)
        for e in range(d):
            dif = (data[li, e] - c[e]).astype(np.float32)
            acc = K._fma32(-dif, dif, acc)
        out[li] = acc
    return out


@pytest.mark.parametrize("l,d,n,k", [(3, 64, 4096, 256), (2, 16, 1000, 40), (1, 8, 132, 256), (5, 32, 300, 100), (64, 64, 2048, 256)])
def test_max_sim_tensor_core(cuda_device, l, d, n, k):
    """TF32 tcgen05 assignment: labels equal the exact ones except between near-equidistant centroids, and the
    reported similarity is the EXACT fp32 value of the chosen centroid."""
    import torchpq_b200 as T
    torch.manual_seed(l + d + k)
    data, cent = torch.randn(l, d, n), torch.randn(l, d, k)
    osim, olab = K.max_sim(data.numpy(), cent.numpy())
    sim, lab = T.fn.max_sim(data.cuda(), cent.cuda(), exact=False)
    sim_fast, lab_fast = T.fn.max_sim(data.cuda(), cent.cuda(), exact=False, exact_values=False)
    torch.cuda.synchronize()
    assert torch.equal(lab, lab_fast)                                  # same labels, TF32-derived values
    assert torch.allclose(sim_fast, sim, rtol=5e-3, atol=5e-2)
    sim, lab = sim.cpu().numpy(), lab.cpu().numpy()
    assert lab.min() >= 0 and lab.max() < k
    agree = (lab == olab).mean()
    assert agree >= 0.985, f"label agreement {agree}"
    assert np.array_equal(sim, _exact_sim_of(data.numpy(), cent.numpy(), lab))
    # where the label differs the chosen centroid is (almost) as close as the best one
    bad = lab != olab
    if bad.any():
        rel = (osim[bad] - sim[bad]) / np.abs(osim[bad])
        assert rel.min() >= 0 and rel.max() < 5e-3


def test_fit_matches_oracle_loop(cuda_device):
    """build.fit = the reference's MultiKMeans.fit loop (clustering/MultiKMeans.py:415-453): from identical initial
    centroids on integer-valued data (sums of members are exact in fp32, so every Lloyd step is order-independent)
    the exact-assignment GPU loop and the CPU restatement agree bit for bit: centroids, labels, iteration count."""
    from torchpq_b200 import build
    rng = np.random.default_rng(3)
    data = torch.from_numpy(rng.integers(-8, 9, (3, 4, 3000)).astype(np.float32))
    c0 = data[:, :, rng.choice(3000, 16, replace=False)].clone()
    ocent, olab, oinertia, oit = K.fit(data.numpy(), c0.numpy(), max_iter=12, tol=1e-4)
    cent, lab, inertia, it = build.fit(data.cuda().contiguous(), 16, max_iter=12, tol=1e-4, centroids=c0.cuda(), exact=True)
    assert it == oit
    assert np.array_equal(lab.cpu().numpy(), olab)
    assert np.allclose(cent.cpu().numpy(), ocent, rtol=1e-6, atol=1e-6)
    assert abs(inertia - oinertia) <= 1e-4 * max(1.0, abs(oinertia))


def test_fit_stops_on_sum_of_squares_and_picks_best_redo(cuda_device):
    """error = sum((new - old)^2) <= tol stops the loop (not a mean shift); with n_redo > 1 the lowest-inertia run wins."""
    from torchpq_b200 import build
    torch.manual_seed(0)
    centers = torch.randn(1, 8, 4) * 20
    data = (centers[:, :, torch.randint(0, 4, (2000,))] + torch.randn(1, 8, 2000)).cuda().contiguous()
    cent, lab, inertia, it = build.fit(data, 4, max_iter=50, tol=1e-4, n_redo=1, seed=1)
    assert it < 50                                                   # converged by the tolerance, not by max_iter
    runs = [build.fit(data, 4, max_iter=50, tol=1e-4, n_redo=1, seed=s)[2] for s in (5, 6, 7)]
    best = build.fit(data, 4, max_iter=50, tol=1e-4, n_redo=3, seed=5)[2]
    assert abs(best - min(runs)) <= 1e-3 * abs(min(runs))
