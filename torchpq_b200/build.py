"""Index build side (train / encode / add) -- the callers *before* the search hot path.

SURVEY.md section 8(f) rank 1-2 ("next" rows).  The two compute steps (assignment = tpq_max_sim,
centroid update = tpq_compute_centroids) are the library's CUDA kernels; the container bookkeeping
(index of appearance, expansion, write addresses) is expressed with PyTorch tensor ops on the index
device.  It is what builds the large synthetic indexes for the search benchmark; it is NOT the
measured path.  Semantics followed:

    IVFPQIndex.train        torchpq/index/IVFPQIndex.py:234-260  (VQ: <=15 Lloyd iters, PQ: <=25)
    KMeans / MultiKMeans    torchpq/clustering/KMeans.py:399-438, MultiKMeans.py:415-453
        assignment = arg-max of negative squared L2 (kernels/cuda/max_sim.cu:78-98)
        update     = member mean, empty cluster -> 0 (kernels/cuda/compute_centroids.cu:9-86)
    IVFPQIndex.add          torchpq/index/IVFPQIndex.py:316-364
    CellContainer.add       torchpq/container/CellContainer.py:313-367 (ioa, expand loop, write address
                            = ioa-th empty slot of the cell, get_write_address_v2.cu:9-41)
Unlike the reference (unseeded np.random.choice, MultiKMeans.py:277-283) initial centroids come
from a seeded torch.Generator: two runs start from the same centroids.  (Codebooks are reproducible to fp32
rounding, not bit for bit: the centroid update accumulates per-tile partial sums with float atomics.)
"""
from __future__ import annotations

import torch


def _assign(data: torch.Tensor, cent: torch.Tensor, exact: bool = True) -> torch.Tensor:
    """data [l, d, n], cent [l, d, k] -> labels [l, n] int64: tpq_max_sim.  exact=True: the reference's fp32
    arithmetic (used for encoding, so codes match the reference's); exact=False: TF32 tensor-core kernel where it
    applies (used inside the Lloyd iterations, where a label flip between near-equidistant centroids is harmless)."""
    from . import fn
    return fn.max_sim(data.contiguous(), cent.contiguous(), "euclidean", exact=exact, exact_values=False)[1]


def _init_centroids(data: torch.Tensor, k: int, seed: int) -> torch.Tensor:
    """MultiKMeans.initialize_centroids, init_mode="random" (MultiKMeans.py:271-283): k distinct data points, the same
    columns for every sub-quantizer -- from a seeded generator instead of the reference's unseeded np.random.choice."""
    l, d, n = data.shape
    gen = torch.Generator(device="cpu").manual_seed(seed)
    pick = torch.randperm(n, generator=gen)[:k].to(data.device)
    cent = data[:, :, pick].clone()
    if cent.shape[2] < k:                                              # fewer points than clusters
        cent = torch.cat([cent, torch.zeros(l, d, k - cent.shape[2], device=data.device)], dim=2)
    return cent


def fit(data: torch.Tensor, k: int, max_iter: int, tol: float = 1e-4, n_redo: int = 1, seed: int = 0,
        centroids: torch.Tensor | None = None, distance: str = "euclidean", exact: bool = False):
    """MultiKMeans.fit (clustering/MultiKMeans.py:415-453; KMeans.fit, KMeans.py:399-438, is the l = 1 case):
    l independent k-means over data [l, d, n].  Per redo: Lloyd iterations until `error = sum((new - old)^2) <= tol`
    (calculate_error, :143-149 -- a SUM of squares over all l*d*k entries, not a mean shift) or max_iter; the redo
    with the lowest inertia = mean(-maxsims) (:151-152) wins.  -> (centroids [l, d, k], labels [l, n], inertia, n_iter).
    As in the reference the returned labels are those of the last assignment, i.e. relative to the centroids BEFORE
    the final update.  exact=True: the reference's fp32 assignment arithmetic bit for bit (max_sim.cu:78-98);
    exact=False: TF32 tensor-core assignment where the shape allows (labels can differ between near-equidistant
    centroids only)."""
    from . import fn
    assert data.dim() == 3 and data.is_contiguous(), "use .contiguous()"                  # MultiKMeans.py:421
    assert distance in ("euclidean", "cosine")
    best = None
    best_inertia = 1e32
    for redo in range(n_redo):
        cent = centroids if (centroids is not None and redo == 0) else _init_centroids(data, k, seed + redo)
        n_iter = 0
        for _ in range(max_iter):
            if distance == "cosine":                                    # get_labels, MultiKMeans.py:314-333
                dn = data / (data.norm(dim=-2, keepdim=True) + 1e-8)
                cn = cent / (cent.norm(dim=-2, keepdim=True) + 1e-8)
                maxsims, labels = fn.max_sim(dn.contiguous(), cn.contiguous(), "cosine", exact=True)
            else:
                maxsims, labels = fn.max_sim(data, cent.contiguous(), "euclidean", exact=exact, exact_values=n_redo > 1)
            new = fn.compute_centroids(data, labels, k)
            error = (new - cent).pow(2).sum()
            cent = new
            n_iter += 1
            if float(error) <= tol:                                     # one host sync per iteration, as in the reference
                break
        inertia = float((-maxsims).mean())
        if inertia < best_inertia:
            best, best_inertia = (cent, labels, inertia, n_iter), inertia
    return best


def multi_kmeans(data: torch.Tensor, k: int, max_iter: int, tol: float = 1e-4, seed: int = 0,
                 distance: str = "euclidean") -> torch.Tensor:
    """l independent k-means over data [l, d, n] -> centroids [l, d, k] (fit with the reference's defaults)."""
    return fit(data, k, max_iter, tol=tol, n_redo=1, seed=seed, distance=distance)[0]


def train(index, x: torch.Tensor, seed: int = 0, vq_iters: int = 15, pq_iters: int = 25):
    """IVFPQIndex.train: fit the coarse quantizer and the M sub-quantizers on x [d, n]."""
    assert x.dim() == 2 and x.shape[0] == index.d_vector
    from . import fn
    if index.distance == "cosine":
        x = fn.normalize(x.contiguous())
    vq = multi_kmeans(x[None].contiguous(), index.n_cells, vq_iters, seed=seed)[0].contiguous()
    if getattr(index, "pq_use_residual", False):                       # IVFPQIndex.py:248-256: PQ learns x - vq(x)
        x = x - vq[:, _assign(x[None].contiguous(), vq[None])[0]]
    sub = x.reshape(index.n_subvectors, index.d_subvector, x.shape[1]).contiguous()
    pq = multi_kmeans(sub, 256, pq_iters, seed=seed + 1, distance=index.distance).contiguous()   # PQCodec.py:27-32
    index.vq_codec.set_codebook(vq)
    index.pq_codec.set_codebook(pq)
    index._state_changed()


def encode(index, x: torch.Tensor):
    """-> (cells [n] i64, codes [M, n] u8) for x [d, n] (already normalised for cosine)."""
    cells = _assign(x[None], index.vq_codec.codebook[None])[0]
    if getattr(index, "pq_use_residual", False):                       # IVFPQIndex.py:281-284
        x = x - index.vq_codec.codebook[:, cells]
    sub = x.reshape(index.n_subvectors, index.d_subvector, x.shape[1]).contiguous()
    codes = _assign(sub, index.pq_codec.codebook).to(torch.uint8)
    return cells, codes


def _ioa(cells: torch.Tensor) -> torch.Tensor:
    """index of appearance of each label among equal labels, input order (get_ioa.cu:8-47)."""
    n = cells.shape[0]
    order = torch.sort(cells, stable=True).indices
    sc = cells[order]
    first = torch.ones(n, dtype=torch.bool, device=cells.device)
    first[1:] = sc[1:] != sc[:-1]
    ar = torch.arange(n, device=cells.device)
    run_start = torch.cummax(torch.where(first, ar, torch.zeros_like(ar)), 0).values
    ioa = torch.empty(n, dtype=torch.long, device=cells.device)
    ioa[order] = ar - run_start
    return ioa


def _expand(index, cells: torch.Tensor):
    """CellContainer.expand (CellContainer.py:249-311), one rebuild for all listed cells."""
    dev = index._storage.device
    grow = torch.zeros(index.n_cells, dtype=torch.long, device=dev)
    grow[cells] = index._cell_capacity[cells] if index.expand_mode == "double" else index.expand_step_size
    shift = torch.cumsum(grow, 0) - grow                           # slots inserted before each cell
    old_start, old_cap = index._cell_start, index._cell_capacity
    new_capacity_total = index.capacity + int(grow.sum().item())
    # old address -> new address
    adr = torch.arange(index.capacity, device=dev)
    cell_of = torch.searchsorted(old_start, adr, right=True) - 1
    new_adr = adr + shift[cell_of]
    storage = torch.zeros(index._storage.shape[0], new_capacity_total, 4, dtype=torch.uint8, device=dev)
    storage[:, new_adr] = index._storage
    a2i = -torch.ones(new_capacity_total, dtype=torch.long, device=dev)
    a2i[new_adr] = index._address2id
    emp = torch.ones(new_capacity_total, dtype=torch.uint8, device=dev)
    emp[new_adr] = index._is_empty
    for name, val in (("_storage", storage), ("_address2id", a2i), ("_is_empty", emp)):
        delattr(index, name)
        index.register_buffer(name, val)
    index._cell_start += shift
    index._cell_capacity += grow


def container_add(index, codes: torch.Tensor, cells: torch.Tensor, ids=None, return_address=False):
    """CellContainer.add (CellContainer.py:313-367)."""
    assert codes.dtype == torch.uint8 and cells.dtype == torch.long
    assert codes.shape[0] == index.code_size and codes.shape[1] == cells.shape[0]
    dev = index._storage.device
    n = cells.shape[0]
    if ids is None:
        ids = torch.arange(n, device=dev, dtype=torch.long) + index.max_id + 1
    else:
        assert ids.dtype == torch.long and ids.shape[0] == n
        ids = ids.to(dev)
    ioa = _ioa(cells)
    while True:
        # free slots of a cell = capacity - LIVE items.  Without holes live == _cell_size and this is the reference's
        # `capacity - cell_size - (ioa + 1)` (CellContainer.py:338-341); with holes left by remove() it stops the cell
        # from expanding while it still has room (and _cell_size, a high-water mark here, from inflating).
        occupied = torch.cumsum((index._is_empty == 0).to(torch.long), 0)
        occupied = torch.cat([occupied.new_zeros(1), occupied])
        live = occupied[index._cell_start + index._cell_capacity] - occupied[index._cell_start]
        free = index._cell_capacity[cells] - live[cells] - (ioa + 1)
        need = cells[free < 0].unique()
        if need.shape[0] == 0:
            break
        _expand(index, need)
    empty_adr = torch.nonzero(index._is_empty == 1)[:, 0]           # sorted
    first_empty = torch.searchsorted(empty_adr, index._cell_start)   # index of each cell's first empty slot
    write = empty_adr[first_empty[cells] + ioa]
    M = index.code_size
    index._storage[:, write] = codes.reshape(M // 4, 4, n).transpose(1, 2)
    index._address2id[write] = ids
    if n:
        index._max_id = max(index._max_id, int(ids.max().item()))
    index._is_empty[write] = 0
    # _cell_size = extent the scan must cover = highest occupied slot + 1 (== old + count when the cell has no holes)
    extent = write - index._cell_start[cells] + 1
    index._cell_size.scatter_reduce_(0, cells, extent, reduce="amax", include_self=True)
    index._state_changed()
    return (ids, write) if return_address else ids


def add(index, x: torch.Tensor, ids=None, return_address=False, chunk: int = 1 << 20):
    """IVFPQIndex.add: encode x [d, n] and store it."""
    assert x.dim() == 2 and x.shape[0] == index.d_vector
    from . import fn
    cells_l, codes_l = [], []
    for s in range(0, x.shape[1], chunk):
        xc = x[:, s:s + chunk].contiguous()
        if index.distance == "cosine":
            xc = fn.normalize(xc)
        c, q = encode(index, xc)
        cells_l.append(c)
        codes_l.append(q)
    return container_add(index, torch.cat(codes_l, 1), torch.cat(cells_l, 0), ids, return_address)
