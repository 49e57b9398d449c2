"""CPU restatement of the k-means / codec neighbours of the search path (TEST INFRASTRUCTURE).

    max_sim            kernels/cuda/max_sim.cu:182-309 (thread_nseuclidean :78-98 / thread_matmul, reduce_dim_2 :152-180)
                       as called by MultiKMeans.get_labels (clustering/MultiKMeans.py:314-333)
    compute_centroids  kernels/cuda/compute_centroids.cu:9-86
    pq_decode          kernels/cuda/pq_decode.cu:7-53 ; PQCodec._decode_cpu codec/PQCodec.py:95-111
"""
import numpy as np


def _fma32(a, b, c):
    """fp32 fused multiply-add emulated through float64 (exact product, one rounding to fp32)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def max_sim(data, cent, distance="euclidean", chunk=4096):
    """data [l, d, n], cent [l, d, k] -> (maxsims [l, n] f32, labels [l, n] i64).
    Similarity accumulated feature by feature with fmaf(-dif, dif, acc) (euclidean) or fmaf(a, b, acc)
    (inner), e ascending, as the kernel does; arg-max = lowest index among equal maxima."""
    data, cent = np.asarray(data, np.float32), np.asarray(cent, np.float32)
    l, d, n = data.shape
    k = cent.shape[2]
    sims = np.empty((l, n), np.float32)
    labels = np.empty((l, n), np.int64)
    for li in range(l):
        for s in range(0, n, chunk):
            x = data[li, :, s:s + chunk]                         # [d, m]
            acc = np.zeros((x.shape[1], k), np.float32)
            for e in range(d):
                a = x[e][:, None]
                b = cent[li, e][None, :]
                if distance == "euclidean":
                    dif = (a - b).astype(np.float32)
                    acc = _fma32(-dif, dif, acc)
                else:
                    acc = _fma32(np.broadcast_to(a, acc.shape), np.broadcast_to(b, acc.shape), acc)
            labels[li, s:s + chunk] = acc.argmax(axis=1)
            sims[li, s:s + chunk] = acc.max(axis=1)
    return sims, labels


def compute_centroids(data, labels, k):
    """data [l, d, n], labels [l, n] -> centroids [l, d, k] (member mean in float64, empty cluster -> 0)."""
    data = np.asarray(data, np.float32)
    l, d, n = data.shape
    out = np.zeros((l, d, k), np.float64)
    for li in range(l):
        cnt = np.bincount(labels[li], minlength=k).astype(np.float64)
        for e in range(d):
            out[li, e] = np.bincount(labels[li], weights=data[li, e].astype(np.float64), minlength=k)
        nz = cnt > 0
        out[li][:, nz] /= cnt[nz]
        out[li][:, ~nz] = 0
    return out.astype(np.float32)


def pq_decode(codebook, code):
    """codebook [M, dsub, 256], code [M, n] u8 -> [M*dsub, n] f32."""
    M, dsub, _ = codebook.shape
    n = code.shape[1]
    out = np.empty((M, dsub, n), np.float32)
    for m in range(M):
        out[m] = codebook[m][:, code[m]]
    return out.reshape(M * dsub, n)


def fit(data, centroids0, max_iter, tol=1e-4):
    """MultiKMeans.fit for one redo from given initial centroids, euclidean (clustering/MultiKMeans.py:431-441):
    labels = arg-max similarity, new centroids = member means, error = sum((new - old)^2), stop when error <= tol.
    -> (centroids, labels of the last assignment, inertia = mean(-maxsims), iterations)."""
    cent = np.asarray(centroids0, np.float32)
    k = cent.shape[2]
    n_iter = 0
    for _ in range(max_iter):
        sims, labels = max_sim(data, cent, "euclidean")
        new = compute_centroids(data, labels, k)
        diff = (new - cent).astype(np.float32)
        error = float((diff * diff).astype(np.float32).sum(dtype=np.float32))
        cent = new
        n_iter += 1
        if error <= tol:
            break
    return cent, labels, float((-sims).mean()), n_iter
