"""Cell-sharded multi-GPU search: one process per GPU, one NCCL all-gather per query batch.

The reference is single-GPU only (no NCCL / torch.distributed anywhere).
Partitioning: rank r scans the cells c with c % world == r; the two small
codebooks and the address->id table are replicated; every rank sees the full query batch, runs
the (deterministic, identical) coarse probe redundantly, scans only its own cells and emits its
local top-k as packed 64-bit (score, address) keys.  One ``all_gather_into_tensor`` of
[nq, k] keys (8 B per candidate) follows, then the same merge kernel that combines CTA slices
picks the global top-k -- the union of per-shard top-k's contains the global top-k, and keys
carry GLOBAL addresses, so sharded == unsharded bit for bit, ties included.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import fn


def merge_gathered(keys_all: torch.Tensor, address2id: torch.Tensor):
    """keys_all [world, nq, k] -> (values, ids, address); CUDA kernel tpq_merge_topk."""
    return fn.merge_topk(keys_all.contiguous(), address2id)


def _gather_rows(t: torch.Tensor, world: int, group):
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    return out


def sharded_search(index, x: torch.Tensor, k: int, group=None, return_address: bool = False,
                   split_coarse: bool = True):
    """``index`` holds the full reference state on every rank and has been given its shard with
    ``index.set_shard(rank, world)``.  Returns the same (values, ids[, address]) on every rank.

    split_coarse (default): the coarse probe -- the one stage that does not shrink with the shard -- is itself split
    by QUERIES: rank r probes queries [r*nq/world, (r+1)*nq/world) and a small all-gather (n_probe+1 int64 per query)
    hands every rank the full probe lists.  Same deterministic kernel on every rank, so the lists are bit-identical
    to the replicated computation.  split_coarse=False runs the whole coarse probe on every rank (one collective
    per batch in total instead of two)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world > 1 and split_coarse:
        nq, n_probe = x.shape[1], int(index.n_probe)
        xq = fn.normalize(x.contiguous()) if index.distance == "cosine" else x.contiguous()   # IVFPQIndex.py:474-475
        per = (nq + world - 1) // world
        lo = min(nq, rank * per)
        hi = min(nq, lo + per)
        mine = torch.zeros(per, 2 * n_probe + 1, dtype=torch.long, device=x.device)
        if hi > lo:
            sims, cells, npl = fn.coarse_probe(xq[:, lo:hi].contiguous(), index.vq_codec.codebook, n_probe,
                                               index.use_smart_probing, index.smart_probing_temperature)
            mine[:hi - lo, :n_probe] = cells
            mine[:hi - lo, n_probe] = npl
            mine[:hi - lo, n_probe + 1:] = sims.view(torch.int32).to(torch.long)      # fp32 bits, exact round trip
        allp = _gather_rows(mine, world, group)[:nq]
        base = allp[:, n_probe + 1:].to(torch.int32).view(torch.float32).contiguous()
        keys = index.search_cells(xq, allp[:, :n_probe].contiguous(), base_sims=base,
                                  n_probe_list=allp[:, n_probe].contiguous(), k=k, return_keys=True)[2]
    else:
        _, _, keys = index.search(x, k=k, return_keys=True)
    if world == 1:
        keys_all = keys[None]
    else:
        keys_all = torch.empty((world * keys.shape[0], keys.shape[1]), dtype=keys.dtype, device=keys.device)
        dist.all_gather_into_tensor(keys_all, keys, group=group)
        keys_all = keys_all.view(world, keys.shape[0], keys.shape[1])
    values, ids, address = merge_gathered(keys_all, index._address2id)
    return (values, ids, address) if return_address else (values, ids)


def broadcast_state(index, src: int = 0, group=None):
    """Replicate rank ``src``'s index buffers (after train/add) to every rank."""
    rank = dist.get_rank(group)
    dev = torch.device(index.device)
    names = ["_storage", "_is_empty", "_cell_start", "_cell_size", "_cell_capacity", "_address2id"]
    meta = [None]
    if rank == src:
        meta[0] = {n: (tuple(getattr(index, n).shape), getattr(index, n).dtype) for n in names}
        meta[0]["vq"] = tuple(index.vq_codec.codebook.shape)
        meta[0]["pq"] = tuple(index.pq_codec.codebook.shape)
        meta[0]["max_id"] = index._max_id
    dist.broadcast_object_list(meta, src=src, group=group)
    m = meta[0]
    for n in names:
        if rank != src:
            shape, dtype = m[n]
            delattr(index, n)
            index.register_buffer(n, torch.empty(shape, dtype=dtype, device=dev))
        dist.broadcast(getattr(index, n), src=src, group=group)
    vq = index.vq_codec.codebook if rank == src else torch.empty(m["vq"], dtype=torch.float32, device=dev)
    pq = index.pq_codec.codebook if rank == src else torch.empty(m["pq"], dtype=torch.float32, device=dev)
    dist.broadcast(vq, src=src, group=group)
    dist.broadcast(pq, src=src, group=group)
    if rank != src:
        index.vq_codec.set_codebook(vq)
        index.pq_codec.set_codebook(pq)
        index._max_id = m["max_id"]
    index._state_changed()
