"""CPU, world_size 2, gloo: the host-side logic of the cell-sharded search (torchpq_b200/dist.py):
shard = cells mod world, one all-gather of packed keys, merge == unsharded.  The per-shard scan is
played by the oracle (the product's scan needs a GPU); key packing restates common.cuh make_key."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pack_keys(vals, adr):
    b = vals.astype(np.float32).view(np.uint32)
    o = np.where(b & 0x80000000, ~b, b | 0x80000000).astype(np.uint64)
    key = (o << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - adr.astype(np.uint64))
    key[adr < 0] = 0
    return key


def unpack_keys(key):
    o = (key >> np.uint64(32)).astype(np.uint32)
    b = np.where(o & 0x80000000, o ^ 0x80000000, ~o).astype(np.uint32)
    v = b.view(np.float32).copy()
    a = (np.uint64(0xFFFFFFFF) - (key & np.uint64(0xFFFFFFFF))).astype(np.int64)
    v[key == 0] = -np.inf
    a[key == 0] = -1
    return v, a


def shard_search(st, x, k, rank, world):
    from oracle import ivfpq_oracle as O
    xx, sims, cells, npl = O.coarse_probe(st, x)
    lut = O.precompute_adc(xx, torch.from_numpy(st.pq_codebook), st.distance).numpy()
    cn = cells.numpy()
    size = st.cell_size[cn].copy()
    size[(cn % world) != rank] = 0                              # not my cell -> nothing to scan
    return O.ivfpq_topk(st.storage, lut, st.is_empty, st.cell_start[cn], size, npl.numpy(), k)


def worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ivfpq_oracle as O, build_state as B
    st, queries = B.integer_state(32, 8, 16, 3000, seed=2)
    st.n_probe, st.use_smart_probing = 6, True
    x, k = queries(20), 15
    v, a = shard_search(st, x, k, rank, world)
    keys = torch.from_numpy(pack_keys(v, a).view(np.int64))
    gathered = torch.empty((world * keys.shape[0], keys.shape[1]), dtype=torch.int64)
    dist.all_gather_into_tensor(gathered, keys)                 # the one collective of the path
    allk = gathered.view(world, keys.shape[0], -1).numpy().view(np.uint64).transpose(1, 0, 2).reshape(keys.shape[0], -1)
    top = np.sort(allk, axis=1)[:, ::-1][:, :k]
    mv, ma = unpack_keys(np.ascontiguousarray(top))
    fv, fi, fa = O.search(st, x, k=k, return_address=True)
    ok = np.array_equal(mv, fv) and np.array_equal(ma, fa)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_equals_unsharded_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_key_packing_roundtrip_and_order():
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.normal(size=100).astype(np.float32) * 1e3, np.array([0.0, -0.0 + 0.0, 1e-38, -1e-38], np.float32)])
    a = rng.integers(0, 2 ** 31, v.shape[0])
    k = pack_keys(v, a)
    v2, a2 = unpack_keys(k)
    assert np.array_equal(v2, v) and np.array_equal(a2, a)
    order = np.argsort(k)[::-1]
    expect = np.lexsort((a, -v.astype(np.float64)))
    assert np.array_equal(v[order], v[expect])


# ---------------------------------------------------------------------------------------------------------------
# torchpq_b200.dist.sharded_search itself (its collective plumbing: query-split coarse probe, packed all-gather of
# cells / n_probe_list / coarse similarities, key all-gather, merge), with the CUDA ops replaced by the oracle.
# ---------------------------------------------------------------------------------------------------------------
class _OracleIndex:
    """Stands in for IVFPQIndex on CPU: same attributes / methods sharded_search touches."""

    def __init__(self, st, rank, world):
        self.st, self.rank, self.world = st, rank, world
        self.n_probe, self.distance = st.n_probe, st.distance
        self.use_smart_probing, self.smart_probing_temperature = st.use_smart_probing, st.smart_probing_temperature
        self.vq_codec = type("C", (), {"codebook": torch.from_numpy(st.vq_codebook)})()
        self._address2id = torch.from_numpy(st.address2id)

    def search_cells(self, x, cells, base_sims=None, n_probe_list=None, k=1, return_keys=False):
        from oracle import ivfpq_oracle as O
        cn = cells.numpy()
        size = self.st.cell_size[cn].copy()
        size[(cn % self.world) != self.rank] = 0
        lut = O.precompute_adc(x, torch.from_numpy(self.st.pq_codebook), self.st.distance).numpy()
        v, a = O.ivfpq_topk(self.st.storage, lut, self.st.is_empty, self.st.cell_start[cn], size, n_probe_list.numpy(), k)
        return None, None, torch.from_numpy(pack_keys(v, a).view(np.int64))


def _cpu_coarse_probe(x, vq_codebook, n_probe, smart, temperature):
    from oracle import ivfpq_oracle as O
    sims = O.negative_squared_l2(x, vq_codebook).contiguous()
    s, c = O.topk_desc(sims, n_probe)
    npl = O.smart_probing(s, n_probe, temperature) if (smart and n_probe > 1) else torch.zeros(x.shape[1], dtype=torch.long) + n_probe
    return s, c, npl


def _cpu_merge(keys_all, address2id):
    k = keys_all.shape[2]
    allk = keys_all.numpy().view(np.uint64).transpose(1, 0, 2).reshape(keys_all.shape[1], -1)
    top = np.ascontiguousarray(np.sort(allk, axis=1)[:, ::-1][:, :k])
    v, a = unpack_keys(top)
    ids = np.where(a >= 0, address2id.numpy()[np.maximum(a, 0)], -1)
    return torch.from_numpy(v), torch.from_numpy(ids), torch.from_numpy(a)


def worker_sharded_search(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ivfpq_oracle as O, build_state as B
    from torchpq_b200 import dist as tdist
    tdist.fn.coarse_probe = _cpu_coarse_probe            # the only CUDA ops on this path
    tdist.fn.merge_topk = _cpu_merge
    st, queries = B.integer_state(32, 8, 16, 3000, seed=5)
    st.n_probe, st.use_smart_probing = 6, True
    x, k = queries(21), 12                                # 21 queries over 2 ranks: uneven split, padded slice
    fv, fi, fa = O.search(st, x, k=k, return_address=True)
    ok = True
    for split in (True, False):
        ix = _OracleIndex(st, rank, world)
        if not split:                                     # replicated coarse probe: search() = coarse + search_cells
            def search(xx, k=1, return_keys=False, _ix=ix):
                s, c, npl = _cpu_coarse_probe(xx, _ix.vq_codec.codebook, _ix.n_probe, True, 30.0)
                return _ix.search_cells(xx, c, base_sims=s, n_probe_list=npl, k=k, return_keys=True)
            ix.search = search
        v, ids, a = tdist.sharded_search(ix, x, k, return_address=True, split_coarse=split)
        ok = ok and np.array_equal(v.numpy(), fv) and np.array_equal(a.numpy(), fa) and np.array_equal(ids.numpy(), fi)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_search_plumbing_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=worker_sharded_search, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


# ---------------------------------------------------------------------------------------------------------------
# 2-D grid: world 4 = 2 cell shards x 2 query groups (dist.make_grid), split coarse probe inside each query group
# ---------------------------------------------------------------------------------------------------------------
def worker_grid(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ivfpq_oracle as O, build_state as B
    from torchpq_b200 import dist as tdist
    tdist.fn.coarse_probe = _cpu_coarse_probe
    tdist.fn.merge_topk = _cpu_merge
    st, queries = B.integer_state(32, 8, 16, 3000, seed=7)
    st.n_probe, st.use_smart_probing = 6, True
    x, k = queries(23), 9                                 # 23 queries over 2 groups: 12 + 11 (padded), then 6 + 6 per rank
    fv, fi, fa = O.search(st, x, k=k, return_address=True)
    ok = True
    for groups in (2, 1, 4):
        grid = tdist.make_grid(world, groups)
        shards = grid[0]
        cgs = [dist.new_group(list(range(g * shards, (g + 1) * shards))) for g in range(groups)]
        ix = _OracleIndex(st, rank % shards, shards)

        def search(xx, k=1, return_keys=False, _ix=ix):   # replicated coarse probe (used when shards == 1)
            s, c, npl = _cpu_coarse_probe(xx, _ix.vq_codec.codebook, _ix.n_probe, True, 30.0)
            return _ix.search_cells(xx, c, base_sims=s, n_probe_list=npl, k=k, return_keys=True)
        ix.search = search
        v, ids, a = tdist.sharded_search(ix, x, k, return_address=True, grid=grid, coarse_group=cgs[rank // shards])
        ok = ok and np.array_equal(v.numpy(), fv) and np.array_equal(a.numpy(), fa) and np.array_equal(ids.numpy(), fi)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_grid_sharded_search_gloo():
    world = 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=worker_grid, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]
