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
