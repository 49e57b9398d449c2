"""GPU, world_size 2, NCCL: torchpq_b200.dist.sharded_search == the unsharded search, bit for bit.
Skipped on boxes with fewer than two GPUs (the gloo tests cover the host-side logic on CPU)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    import torchpq_b200 as T
    from torchpq_b200 import dist as tdist
    from oracle import build_state as B
    torch.manual_seed(3)
    st = B.build_state(torch.randn(64, 6000), 16, 32, vq_iters=2, pq_iters=1)
    st.n_probe = 8
    ix = T.IVFPQIndex(64, 16, 32, initial_size=1, device=f"cuda:{rank}").load_state(st)
    x = torch.randn(64, 301, generator=torch.Generator().manual_seed(9)).to(f"cuda:{rank}")
    v0, i0, a0 = ix.search(x, k=25, return_address=True)
    ix.set_shard(rank, world)
    ok = True
    for split in (True, False):
        v, i, a = tdist.sharded_search(ix, x, 25, return_address=True, split_coarse=split)
        ok = ok and torch.equal(v, v0) and torch.equal(i, i0) and torch.equal(a, a0)
    # shard-only ranks (dist.distribute): rank 0 owns the full index, every rank ends up with its cells' scan layout
    # only (no _storage / _is_empty anywhere), for both grids a 2-rank world allows
    for grid in ((2, 1), (1, 2)):
        full = T.IVFPQIndex(64, 16, 32, initial_size=1, device=f"cuda:{rank}")
        if rank == 0:
            full.load_state(st)
        tdist.distribute(full, 0, grid=grid)
        ok = ok and full._storage is None and full._is_empty is None
        v, i, a = tdist.sharded_search(full, x, 25, return_address=True, grid=grid, split_coarse=False)
        ok = ok and torch.equal(v, v0) and torch.equal(i, i0) and torch.equal(a, a0)
        if grid == (2, 1):
            lay = full.layout()
            ok = ok and lay.n_blocks < ix.layout().n_blocks + 64 and full.resident_bytes() > 0
    # fused scan + exchange (P2P key push into every rank's symmetric-memory gather buffer) == the NCCL all-gather path,
    # on a batch large enough not to be sliced, several steps in a row (the two gather buffers alternate)
    xb = torch.randn(64, 1500, generator=torch.Generator().manual_seed(11)).to(f"cuda:{rank}")
    ref = tdist.sharded_search(full, xb, 25, return_address=True, grid=(1, 2), split_coarse=False, exchange="nccl")
    p2p = "unavailable"
    try:
        for grid, split in (((1, 2), False), ((2, 1), True), ((2, 1), False)):
            ix2 = T.IVFPQIndex(64, 16, 32, initial_size=1, device=f"cuda:{rank}")
            if rank == 0:
                ix2.load_state(st)
            tdist.distribute(ix2, 0, grid=grid)
            for _ in range(3):
                got = tdist.sharded_search(ix2, xb, 25, return_address=True, grid=grid, split_coarse=split, exchange="p2p")
                ok = ok and all(torch.equal(a_, b_) for a_, b_ in zip(got, ref))
        p2p = "ok"
    except AssertionError as e:
        p2p = f"unavailable: {e}"
    q.put((rank, bool(ok), p2p))
    dist.destroy_process_group()


def test_sharded_search_nccl():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 32500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    print("fused exchange:", [r[2] for r in res])
    assert sorted(r[:2] for r in res) == [(0, True), (1, True)]
    assert all(r[2] == "ok" for r in res), res
