"""Shared helpers for the parity tests."""
import numpy as np
import torch

from oracle import ivfpq_oracle as O


def make_index(st, device="cuda:0"):
    import torchpq_b200 as T
    ix = T.IVFPQIndex(st.d_vector, st.n_subvectors, st.n_cells, initial_size=1, distance=st.distance, device=device)
    return ix.load_state(st)


def tie_free_rows(vals_k1: np.ndarray, k: int) -> np.ndarray:
    """Rows of an oracle top-(k+1) result that are tie-free at the k-th boundary (score k != score k+1),
    i.e. whose top-k id SET does not depend on any tie rule."""
    return vals_k1[:, k - 1] != vals_k1[:, k]


def assert_close_results(vals, ids, ovals, oids, rtol=1e-3, min_overlap=0.999):
    """fp32 parity: values within rtol (relative), id sets overlap >= min_overlap."""
    vals, ids, ovals, oids = map(np.asarray, (vals, ids, ovals, oids))
    finite = np.isfinite(ovals)
    assert np.array_equal(np.isfinite(vals), finite)
    denom = np.maximum(np.abs(ovals[finite]), 1e-6)
    rel = np.abs(vals[finite] - ovals[finite]) / denom
    assert rel.max(initial=0.0) <= rtol, f"max relative error {rel.max()}"
    nq, k = oids.shape
    hit = sum(np.intersect1d(ids[q][ids[q] >= 0], oids[q][oids[q] >= 0]).shape[0] for q in range(nq))
    tot = int((oids >= 0).sum())
    assert tot == 0 or hit / tot >= min_overlap, f"id overlap {hit / max(tot, 1)}"
