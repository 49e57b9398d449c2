#!/usr/bin/env python
"""Headline benchmark: queries/sec at recall@100 of IVFPQ search (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c4|tiny]

A "step" is one ``IVFPQIndex.search`` call over one batch of ``--nq`` synthetic queries.
Default workload = BASELINE.json configs[2] (the config the metric is quoted on):
10M x 128 fp32 randn base, M=64, n_cells=4096, n_probe=32, k=100, index cell-sharded over N GPUs.
Prints ONE JSON line on rank 0 (see the task's bench contract for the fields).

--impl reference: the reference has no CPU path and cannot be imported without CuPy + a GPU
(torchpq/__init__.py:2-5), so the reference arm times the repo's CPU restatement of it
(oracle/: torch-CPU matmuls + the OpenMP C scan) on the host cores, on a bounded query sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

WORKLOADS = {
    # name: (N, d, M, n_cells, n_probe, k, distance, n_train, description)
    "c3":   (10_000_000, 128, 64, 4096, 32, 100, "euclidean", 1 << 20, "10Mx128 fp32 randn, M=64, n_cells=4096, nprobe=32, k=100"),
    "c2":   (1_000_000, 128, 64, 1024, 32, 100, "euclidean", 1 << 18, "1Mx128 fp32 randn, M=64, n_cells=1024, nprobe=32, k=100"),
    "c4":   (1_000_000, 960, 120, 1024, 64, 100, "cosine", 1 << 17, "1Mx960 fp32 randn (GIST-shaped), M=120, n_cells=1024, nprobe=64, cosine, k=100"),
    "tiny": (100_000, 128, 64, 256, 16, 100, "euclidean", 1 << 16, "100kx128 smoke workload"),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------- synthetic index (setup, not timed)
def gen_base(d, n, device, seed=1234, chunk=1 << 20):
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.empty(d, n, dtype=torch.float32, device=device)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        x[:, s:e] = torch.randn(d, e - s, generator=g, device=device)
    return x


def build_index(wl, device, vq_iters=4, pq_iters=4):
    """Train + add with the torch-op build side (torchpq_b200/build.py); returns (index, base)."""
    import torchpq_b200 as T
    from torchpq_b200 import build
    N, d, M, C, n_probe, k, distance, n_train, _ = wl
    base = gen_base(d, N, device)
    tmp = T.IVFPQIndex(d, M, C, initial_size=1, distance=distance, device=str(device))
    build.train(tmp, base[:, :n_train].contiguous(), seed=0, vq_iters=vq_iters, pq_iters=pq_iters)
    cells_l, codes_l = [], []
    for s in range(0, N, 1 << 20):
        xc = base[:, s:s + (1 << 20)].contiguous()
        if distance == "cosine":
            xc = T.fn.normalize(xc)
        c, q = build.encode(tmp, xc)
        cells_l.append(c); codes_l.append(q)
    cells, codes = torch.cat(cells_l), torch.cat(codes_l, 1)
    initial_size = int(torch.bincount(cells, minlength=C).max().item())     # no expansion, capacity = C * max cell
    index = T.IVFPQIndex(d, M, C, initial_size=initial_size, distance=distance, device=str(device))
    index.vq_codec.set_codebook(tmp.vq_codec.codebook)
    index.pq_codec.set_codebook(tmp.pq_codec.codebook)
    build.container_add(index, codes, cells)
    index.n_probe = n_probe
    return index, base


def build_state_torch(wl, device, vq_iters=4, pq_iters=4):
    """Reference arm only: the same synthetic index built with plain torch ops (no torchpq_b200 code at all) and
    returned as an oracle IndexState.  Seeded Lloyd with matmul arg-max assignment, placement = stable sort by cell
    (what CellContainer.add does for one add into a fresh container, CellContainer.py:334-353)."""
    from oracle import ivfpq_oracle as O
    N, d, M, C, n_probe, k, distance, n_train, _ = wl
    base = gen_base(d, N, device)
    if distance == "cosine":
        base = base / (base.norm(dim=0, keepdim=True) + 1e-9)

    def assign(data, cent, budget=1 << 29):                              # data [l,d,n], cent [l,d,k] -> [l,n]
        l, _, n = data.shape
        out = torch.empty(l, n, dtype=torch.long, device=device)
        c2 = (cent ** 2).sum(1)
        per = max(1, budget // max(1, l * cent.shape[2]))
        for s in range(0, n, per):
            sim = torch.baddbmm(-c2[:, None, :], data[:, :, s:s + per].transpose(1, 2), cent, alpha=2.0)
            out[:, s:s + per] = sim.argmax(2)
        return out

    def kmeans(data, kk, iters, seed):
        l, dd, n = data.shape
        g = torch.Generator(device="cpu").manual_seed(seed)
        cent = data[:, :, torch.randperm(n, generator=g)[:kk].to(device)].clone()
        ones = torch.ones(l, n, device=device)
        for _ in range(iters):
            lab = assign(data, cent)
            sums = torch.zeros(l, dd, kk, device=device).scatter_add_(2, lab[:, None, :].expand(l, dd, n), data)
            cnt = torch.zeros(l, kk, device=device).scatter_add_(1, lab, ones)
            cent = torch.where(cnt[:, None, :] > 0, sums / cnt.clamp(min=1)[:, None, :], torch.zeros((), device=device))
        return cent

    tr = base[:, :n_train].contiguous()
    vq = kmeans(tr[None], C, vq_iters, 0)[0].contiguous()
    pq = kmeans(tr.reshape(M, d // M, n_train).contiguous(), 256, pq_iters, 1).contiguous()
    cells_l, codes_l = [], []
    for s in range(0, N, 1 << 20):
        xc = base[:, s:s + (1 << 20)].contiguous()
        cells_l.append(assign(xc[None], vq[None])[0])
        codes_l.append(assign(xc.reshape(M, d // M, xc.shape[1]), pq).to(torch.uint8))
    cells, codes = torch.cat(cells_l), torch.cat(codes_l, 1)
    counts = torch.bincount(cells, minlength=C)
    initial = int(counts.max().item())
    order = torch.sort(cells, stable=True).indices
    starts = torch.arange(C, device=device) * initial
    first = torch.cumsum(counts, 0) - counts
    rank_in_cell = torch.empty(N, dtype=torch.long, device=device)
    rank_in_cell[order] = torch.arange(N, device=device) - first[cells[order]]
    adr = starts[cells] + rank_in_cell
    cap = C * initial
    storage = torch.zeros(M // 4, cap, 4, dtype=torch.uint8, device=device)
    storage[:, adr] = codes.reshape(M // 4, 4, N).transpose(1, 2)
    a2i = -torch.ones(cap, dtype=torch.long, device=device); a2i[adr] = torch.arange(N, device=device)
    emp = torch.ones(cap, dtype=torch.uint8, device=device); emp[adr] = 0
    c = lambda t: t.cpu().numpy()
    return O.IndexState(d_vector=d, n_subvectors=M, n_cells=C, distance=distance, vq_codebook=c(vq), pq_codebook=c(pq),
                        storage=c(storage), is_empty=c(emp), cell_start=c(starts), cell_size=c(counts),
                        cell_capacity=c(torch.zeros(C, dtype=torch.long, device=device) + initial), address2id=c(a2i),
                        max_id=N - 1, n_probe=n_probe)


def gen_queries(d, nq, n_batches, device, seed=4321):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return [torch.randn(d, nq, generator=g).pin_memory() if device.type == "cuda" else torch.randn(d, nq, generator=g)
            for _ in range(n_batches)]


def exact_truth(base, x, k, distance, chunk=1 << 20):
    """Exact top-k ids of x [d, nq] in base by brute force fp32 (ids = column index = add order)."""
    if distance == "cosine":
        base_n = None
        x = x / (x.norm(dim=0, keepdim=True) + 1e-9)
    best_v = torch.full((x.shape[1], k), -float("inf"), device=x.device)
    best_i = torch.zeros((x.shape[1], k), dtype=torch.long, device=x.device)
    x2 = (x ** 2).sum(0)[:, None]
    for s in range(0, base.shape[1], chunk):
        b = base[:, s:s + chunk]
        if distance == "cosine":
            sim = x.T @ (b / (b.norm(dim=0, keepdim=True) + 1e-9))
        else:
            sim = 2 * (x.T @ b) - x2 - (b ** 2).sum(0)[None, :]
        v, i = torch.topk(sim, k, dim=1)
        cat_v = torch.cat([best_v, v], 1); cat_i = torch.cat([best_i, i + s], 1)
        sel = torch.topk(cat_v, k, dim=1).indices
        best_v, best_i = torch.gather(cat_v, 1, sel), torch.gather(cat_i, 1, sel)
    return best_i


def recall(found, truth):
    found, truth = found.cpu().numpy(), truth.cpu().numpy()
    hit = sum(np.intersect1d(found[q], truth[q]).shape[0] for q in range(truth.shape[0]))
    return hit / float(truth.size)


# ----------------------------------------------------------------------------- oracle (CPU) leg
def to_oracle_state(index):
    from oracle import ivfpq_oracle as O
    return O.IndexState(
        d_vector=index.d_vector, n_subvectors=index.n_subvectors, n_cells=index.n_cells, distance=index.distance,
        vq_codebook=index.vq_codec.codebook.cpu().numpy(), pq_codebook=index.pq_codec.codebook.cpu().numpy(),
        storage=index._storage.cpu().numpy(), is_empty=index._is_empty.cpu().numpy(),
        cell_start=index._cell_start.cpu().numpy(), cell_size=index._cell_size.cpu().numpy(),
        cell_capacity=index._cell_capacity.cpu().numpy(), address2id=index._address2id.cpu().numpy(),
        max_id=index.max_id, n_probe=index.n_probe, use_smart_probing=index.use_smart_probing,
        smart_probing_temperature=index.smart_probing_temperature)


def oracle_search(st, x, k, threads):
    """oracle.search with the scan done by the OpenMP C restatement (same semantics, all host threads)."""
    from oracle import ivfpq_oracle as O, c_oracle as CO
    xx, sims, cells, npl = O.coarse_probe(st, x)
    lut = O.precompute_adc(xx, torch.from_numpy(st.pq_codebook), st.distance).numpy()
    cn = cells.numpy()
    vals, adr, used = CO.ivfpq_topk(st.storage, lut, st.is_empty, st.cell_start[cn], st.cell_size[cn], npl.numpy(), k, threads)
    return vals, O.get_id_by_address(st.address2id, adr), used


def time_oracle(st, xs, k, threads, target_s=12.0):
    """Time the CPU oracle on a bounded sample; returns (qps, n_queries, threads_used)."""
    n = 32
    t0 = time.perf_counter(); _, _, used = oracle_search(st, xs[:, :n].contiguous(), k, threads); dt = time.perf_counter() - t0
    n2 = int(min(xs.shape[1], max(n, n * target_s / max(dt, 1e-3))))
    t0 = time.perf_counter(); _, _, used = oracle_search(st, xs[:, :n2].contiguous(), k, threads); dt = time.perf_counter() - t0
    return n2 / dt, n2, used


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nme, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        self.f.close()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("TPQ_BENCH_WORKLOAD", "c3"), choices=list(WORKLOADS))
    ap.add_argument("--nq", type=int, default=10000)
    ap.add_argument("--no-smart", action="store_true", help="use_smart_probing=False (fixed work per query)")
    ap.add_argument("--cpu-sample", type=float, default=12.0, help="seconds of CPU-oracle work for cpu_baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    N, d, M, C, n_probe, k, distance, n_train, desc = wl
    have_cuda = torch.cuda.is_available()
    device = torch.device(f"cuda:{local_rank}" if have_cuda else "cpu")
    if have_cuda:
        torch.cuda.set_device(device)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)

    if args.impl == "reference":
        if rank != 0:
            return 0
        return run_reference(args, wl, device, threads)

    assert have_cuda, "bench.py --impl ours needs a CUDA device (no CPU fallback)"
    import torch.distributed as dist
    import torchpq_b200 as T
    from torchpq_b200 import dist as tdist
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    assert T._lib.lib.tpq_device_supported(local_rank), "libtpq_b200.so is built for sm_100a"

    t0 = time.time()
    if rank == 0:
        index, base = build_index(wl, device)
    else:
        index, base = T.IVFPQIndex(d, M, C, initial_size=1, distance=distance, device=str(device)), None
        index.n_probe = n_probe
    if world > 1:
        tdist.broadcast_state(index, 0)
    index.use_smart_probing = not args.no_smart
    index.set_shard(rank, world)
    lay = index.layout()
    torch.cuda.synchronize()
    log(f"[rank {rank}] index built in {time.time() - t0:.1f}s: capacity={index.capacity} blocks={lay.n_blocks}")

    nq = args.nq
    xs_host = gen_queries(d, nq, 4, device)
    xs_dev = [x.to(device) for x in xs_host]

    def step_device(i):
        x = xs_dev[i % len(xs_dev)]
        if world > 1:
            return tdist.sharded_search(index, x, k)
        return index.search(x, k=k)

    # end-to-end leg: host (pinned) queries in, host (pinned) results out, every step.  Copies run on their own
    # stream with double-buffered staging tensors, so step i's D2H and step i+1's H2D overlap step i+1's search --
    # all of it inside the timed region.
    copy_stream = torch.cuda.Stream(device)
    out_v = [torch.empty(nq, k, dtype=torch.float32).pin_memory() for _ in range(2)]
    out_i = [torch.empty(nq, k, dtype=torch.long).pin_memory() for _ in range(2)]
    x_in = [torch.empty(d, nq, dtype=torch.float32, device=device) for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    res = [None, None]

    def step_e2e(i):
        b = i & 1
        main = torch.cuda.current_stream(device)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[b])                       # search i-2 no longer reads x_in[b]
            x_in[b].copy_(xs_host[i % len(xs_host)], non_blocking=True)
            ev_in[b].record(copy_stream)
        main.wait_event(ev_in[b])
        if world > 1:
            v, ids = tdist.sharded_search(index, x_in[b], k)
        else:
            v, ids = index.search(x_in[b], k=k)
        ev_free[b].record(main)
        ev_done[b].record(main)
        res[b] = (v, ids)                                            # keep alive until the copy stream has read them
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_done[b])
            out_v[b].copy_(v, non_blocking=True); out_i[b].copy_(ids, non_blocking=True)
            v.record_stream(copy_stream); ids.record_stream(copy_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        torch.cuda.current_stream(device).wait_stream(copy_stream) if fn is step_e2e else None
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local_rank) if rank == 0 else None      # samples across warm-up and both timed regions
    for i in range(args.warmup):
        step_device(i); step_e2e(i)
    ms_dev = timed(step_device, args.steps)
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if sampler else None
    qps = nq * args.steps / (ms_dev / 1e3)
    qps_e2e = nq * args.steps / (ms_e2e / 1e3)

    # ---- roofline of the scan kernel: algorithmic bytes / CUDA-event duration of the scan launches
    T._lib.lib.tpq_profile_enable(1)
    barrier()
    reps = min(args.steps, 5)
    for i in range(reps):
        index.search(xs_dev[i % len(xs_dev)], k=k)
    import ctypes
    ms_scan, n_launch = ctypes.c_float(0), ctypes.c_int(0)
    T._lib.check(T._lib.lib.tpq_profile_scan_ms(ctypes.byref(ms_scan), ctypes.byref(n_launch)))
    T._lib.lib.tpq_profile_enable(0)
    alg_bytes = 0
    for i in range(reps):                                    # B_q = M * sum_{j<P_q} cell_size[cells[q,j]] + 12 k
        x = xs_dev[i % len(xs_dev)]
        _, cells, npl = T.fn.coarse_probe(x, index.vq_codec.codebook, n_probe, index.use_smart_probing,
                                          index.smart_probing_temperature)
        P = npl.clamp(1, n_probe)
        sizes = index._cell_size[cells]
        owned = (cells % world) == rank
        mask = (torch.arange(n_probe, device=device)[None, :] < P[:, None]) & owned
        alg_bytes += int((sizes * mask).sum().item()) * M + 12 * k * nq
    scan_ms_per_launch = ms_scan.value / max(1, n_launch.value)
    achieved = alg_bytes / max(1, n_launch.value) / (scan_ms_per_launch / 1e3) / 1e9 if n_launch.value else None
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic = None
    tp = os.path.join(ROOT, "profiles", "scan_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload)
        except Exception:
            traffic = None

    # sharded run: does the cell-sharded search (all ranks, NCCL) return exactly what the unsharded index returns?
    sharded_ok = None
    if world > 1:
        try:
            v_s, i_s = tdist.sharded_search(index, xs_dev[0], k)
            torch.cuda.synchronize()
            if rank == 0:
                index.set_shard(0, 1)
                v_u, i_u = index.search(xs_dev[0], k=k)
                sharded_ok = bool(torch.equal(v_s, v_u) and torch.equal(i_s, i_u))
                index.set_shard(rank, world)
        except Exception as e:                                       # report, never take the number down
            sharded_ok = f"check failed: {e!r}"
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- recall@100 (ours vs exact brute force) on 1000 queries; oracle on its CPU sample
    nr = min(1000, nq)
    xr = xs_dev[0][:, :nr].contiguous()
    truth = exact_truth(base, xr, k, distance)
    if world > 1:
        index.set_shard(0, 1)                                # full index on rank 0 for the recall / parity leg
    v_r, ids_r = index.search(xr, k=k)
    rec = recall(ids_r, truth)

    cpu = None
    try:
        st = to_oracle_state(index)
        qps_cpu, n_cpu, used = time_oracle(st, xs_host[0].clone(), k, threads, args.cpu_sample)
        nv = min(n_cpu, nr, 128)
        ov, oi, _ = oracle_search(st, xs_host[0][:, :nv].contiguous(), k, threads)
        rec_oracle = recall(torch.from_numpy(oi), truth[:nv])
        rec_ours_same = recall(ids_r[:nv], truth[:nv])
        same = float((np.sort(ids_r[:nv].cpu().numpy(), 1) == np.sort(oi, 1)).all(axis=1).mean())
        cpu = {"value": qps_cpu, "unit": "queries/s", "cores": used, "kind": "port",
               "sample": f"{n_cpu} of the {nq} queries of batch 0, full {args.workload} index, oracle (torch-CPU matmul + OpenMP C scan)",
               "recall_at_100": rec_oracle, "recall_at_100_ours_same_queries": rec_ours_same,
               "rows_with_identical_id_sets": same}
    except Exception as e:  # the CPU leg must not take the GPU number down with it
        cpu = {"value": None, "error": repr(e)}

    # our kernels per search: 2 x col_sqnorm, coarse_gemm, probe_select, ivfpq_scan, merge_topk
    # (+ normalize_columns for cosine, + lut_scan when d/M is not 1/2/4, + the cross-shard merge when sharded)
    launches_per_step = 6 + (1 if distance == "cosine" else 0) + (0 if (d // M) in (1, 2, 4) else 1) + (1 if world > 1 else 0)
    line = {
        "metric": "queries/sec @ recall@100, IVFPQ search", "value": qps, "unit": "queries/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8 codes / fp32 LUT+accumulate",
        "data": "synthetic (randn base seed 1234, randn queries seed 4321; codebooks trained on-device by seeded k-means)",
        "config": {"workload": f"{args.workload}: {desc}", "n_query_per_step": nq, "index_sharding": f"cells mod {world}",
                   "use_smart_probing": index.use_smart_probing,
                   "l2_policy": "inputs larger than L2 (code store >= 640 MB vs 126 MB L2; 4 rotating query batches)"},
        "recall_at_100": rec, "sharded_equals_unsharded": sharded_ok,
        "e2e": {"value": qps_e2e, "unit": "queries/s", "h2d_bytes_per_step": d * nq * 4, "d2h_bytes_per_step": nq * k * 12,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": {"kernel": "ivfpq_scan_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes / max(1, n_launch.value),
                     "scan_ms_per_launch": scan_ms_per_launch, "scan_share_of_step": scan_ms_per_launch / (ms_dev / args.steps)},
        "cpu_baseline": cpu, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_reference(args, wl, device, threads):
    """CPU arm: the oracle restatement on the host cores; each step = a bounded query sample."""
    N, d, M, C, n_probe, k, distance, n_train, desc = wl
    st = build_state_torch(wl, device)            # plain torch ops only; nothing of torchpq_b200 is imported on this arm
    st.use_smart_probing = not args.no_smart
    if device.type == "cuda":
        torch.cuda.empty_cache()
    xs = gen_queries(d, args.nq, 1, torch.device("cpu"))[0]
    # size the per-step sample so that (steps + warmup) fit in about two minutes
    qps0, n0, used = time_oracle(st, xs, k, threads, target_s=3.0)
    per_step = int(max(16, min(args.nq, qps0 * 120.0 / max(1, args.steps + args.warmup))))
    for i in range(args.warmup):
        oracle_search(st, xs[:, :per_step].contiguous(), k, threads)
    t0 = time.perf_counter()
    for i in range(args.steps):
        oracle_search(st, xs[:, :per_step].contiguous(), k, threads)
    dt = time.perf_counter() - t0
    qps = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": "queries/sec @ recall@100, IVFPQ search", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8 codes / fp32 LUT+accumulate",
        "data": "synthetic (same generator as the GPU arm)",
        "config": {"workload": f"{args.workload}: {desc}", "n_query_per_step": per_step,
                   "use_smart_probing": st.use_smart_probing},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": used, "kind": "port",
                         "sample": f"{per_step} queries per step against the full {args.workload} index (reference has no CPU path; oracle restatement)"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
