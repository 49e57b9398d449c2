#!/usr/bin/env python
"""Headline benchmark: queries/sec at recall@100 of IVFPQ search (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c4|c3c|tiny]

A "step" is one ``IVFPQIndex.search`` call over one batch of ``--nq`` synthetic queries.
Default workload = BASELINE.json configs[2] (the config the metric is quoted on):
10M x 128 fp32 randn base, M=64, n_cells=4096, n_probe=32, k=100, index cell-sharded over N GPUs.
Prints ONE JSON line on rank 0 (see the task's bench contract for the fields).  At N=1 with the default workload
the line carries a ``secondary`` block: the same measurement on configs 2 and 4, on a clustered variant of config 3
(meaningful recall, smart probing actually pruning), the C5 k-means kernels, the build side, and the reference's own
``ivfpq_topk`` GPU kernel (compiled unmodified for sm_100a into oracle/_ref) timed on the same index and queries.

--impl reference: the reference has no CPU path and cannot be imported without CuPy + a GPU
(torchpq/__init__.py:2-5), so the reference arm times the repo's CPU restatement of it
(oracle/: torch-CPU matmuls + the OpenMP C scan) on the host cores, on a bounded query sample.
"""
from __future__ import annotations

import argparse
import json
import math
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
    "c3c":  (10_000_000, 128, 64, 4096, 32, 100, "euclidean", 1 << 20, "10Mx128 fp32 clustered (256 Gaussian centres + unit noise), M=64, n_cells=4096, nprobe=32, k=100"),
    "tiny": (100_000, 128, 64, 256, 16, 100, "euclidean", 1 << 16, "100kx128 smoke workload"),
}
CLUSTERED = {"c3c"}
VQ_ITERS, PQ_ITERS = 15, 25          # the reference's training lengths (IVFPQIndex.py:63-71, PQCodec.py:27-32)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def host_threads():
    """CPU threads this process may actually use: the affinity mask capped by the cgroup cpu quota
    (os.cpu_count() ignores both and made the round-1 CPU baseline swing 8x between boxes)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    quota = None
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:                      # cgroup v2: "<quota> <period>" or "max <period>"
            q, p = f.read().split()[:2]
            if q != "max":
                quota = float(q) / float(p)
    except (OSError, ValueError):
        try:
            q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / p
        except (OSError, ValueError):
            pass
    if quota:
        n = max(1, min(n, int(math.floor(quota + 1e-9)) or 1))
    return n, {"affinity": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None,
               "os_cpu_count": os.cpu_count(), "cgroup_quota_cpus": quota}


# ----------------------------------------------------------------------------- synthetic index (setup, not timed)
def cluster_centres(d, n_cells, device, seed=99):
    """4 * sqrt(C) Gaussian centres (unit-variance coordinates); points = centre + unit noise."""
    g = torch.Generator(device=device).manual_seed(seed)
    return torch.randn(d, 4 * int(round(math.sqrt(n_cells))), generator=g, device=device)


def gen_base(d, n, device, seed=1234, chunk=1 << 20, centres=None):
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.empty(d, n, dtype=torch.float32, device=device)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        x[:, s:e] = torch.randn(d, e - s, generator=g, device=device)
        if centres is not None:
            pick = torch.randint(0, centres.shape[1], (e - s,), generator=g, device=device)
            x[:, s:e] += centres[:, pick]
    return x


def build_index(wl, device, vq_iters=VQ_ITERS, pq_iters=PQ_ITERS, clustered=False):
    """Train + add with the library's build side (torchpq_b200/build.py); returns (index, base)."""
    import torchpq_b200 as T
    from torchpq_b200 import build
    N, d, M, C, n_probe, k, distance, n_train, _ = wl
    centres = cluster_centres(d, C, device) if clustered else None
    base = gen_base(d, N, device, centres=centres)
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


def build_state_torch(wl, device, vq_iters=VQ_ITERS, pq_iters=PQ_ITERS, clustered=False):
    """Reference arm only: the same synthetic index built with plain torch ops (no torchpq_b200 code at all) and
    returned as an oracle IndexState.  Seeded Lloyd with matmul arg-max assignment and the reference's stopping rule
    (sum of squared centroid shifts <= 1e-4), placement = stable sort by cell (what CellContainer.add does for one add
    into a fresh container, CellContainer.py:334-353)."""
    from oracle import ivfpq_oracle as O
    N, d, M, C, n_probe, k, distance, n_train, _ = wl
    base = gen_base(d, N, device, centres=cluster_centres(d, C, device) if clustered else None)
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
            new = torch.where(cnt[:, None, :] > 0, sums / cnt.clamp(min=1)[:, None, :], torch.zeros((), device=device))
            err = float((new - cent).pow(2).sum())
            cent = new
            if err <= 1e-4:
                break
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


def gen_queries(d, nq, n_batches, device, seed=4321, centres=None):
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = []
    for _ in range(n_batches):
        x = torch.randn(d, nq, generator=g)
        if centres is not None:
            c = centres.cpu()
            x = x + c[:, torch.randint(0, c.shape[1], (nq,), generator=g)]
        out.append(x.pin_memory() if device.type == "cuda" else x)
    return out


def exact_truth(base, x, k, distance, chunk=1 << 20):
    """Exact top-k ids of x [d, nq] in base by brute force fp32 (ids = column index = add order)."""
    if distance == "cosine":
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
    """oracle.search with the scan done by the OpenMP C restatement (same semantics, all usable host threads)."""
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


def parity_vs_oracle(st, x_host, v_ours, ids_ours, truth, k, threads, n):
    """ours vs the CPU oracle on the first n queries: id sets, values (rtol 1e-3), recall of both."""
    ov, oi, _ = oracle_search(st, x_host[:, :n].contiguous(), k, threads)
    v, ids = v_ours[:n].cpu().numpy(), ids_ours[:n].cpu().numpy()
    same_sets = float((np.sort(ids, 1) == np.sort(oi, 1)).all(axis=1).mean())
    fin = np.isfinite(ov)
    rel = np.abs(v[fin] - ov[fin]) / np.maximum(np.abs(ov[fin]), 1e-6)
    out = {"n_queries": n, "rows_with_identical_id_sets": same_sets, "id_overlap": float(np.mean([
               np.intersect1d(ids[q], oi[q]).shape[0] / float(k) for q in range(n)])),
           "values_max_rel_err": float(rel.max(initial=0.0)), "values_within_1e-3": bool(rel.max(initial=0.0) <= 1e-3)}
    if truth is not None:
        out["recall_at_100_oracle"] = recall(torch.from_numpy(oi), truth[:n])
        out["recall_at_100_ours"] = recall(torch.from_numpy(ids), truth[:n])
    return out


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


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        try:
            return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- device measurements on one index
def cuda_time(fn, steps, warmup=3, device=None):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record(); torch.cuda.synchronize(device)
    return e0.elapsed_time(e1) / steps


def scan_profile_begin():
    """Bracket every scan-kernel launch from now on with CUDA events on its own stream (tpq_profile_*)."""
    import torchpq_b200 as T
    T._lib.lib.tpq_profile_enable(1)


def scan_profile_end():
    """-> (summed scan ms, launches) since scan_profile_begin()."""
    import ctypes
    import torchpq_b200 as T
    ms_scan, n_launch = ctypes.c_float(0), ctypes.c_int(0)
    T._lib.check(T._lib.lib.tpq_profile_scan_ms(ctypes.byref(ms_scan), ctypes.byref(n_launch)))
    T._lib.lib.tpq_profile_enable(0)
    return ms_scan.value, n_launch.value


def algorithmic_bytes(index, xs_dev, k, steps, rank=0, shards=1, query_slice=None):
    """Sum over the `steps` batches of the timed region (batch i = xs_dev[i % len]) of
    B_q = M * sum_{j < P_q} cell_size[cells[q, j]] over the cells this rank owns + 12 k  (SURVEY.md section 8d)."""
    import torchpq_b200 as T
    n_probe, M = int(index.n_probe), index.n_subvectors
    dev = xs_dev[0].device
    sel = (lambda x: x) if query_slice is None else (lambda x: x[:, query_slice[0]:query_slice[1]].contiguous())
    per_batch, probes = [], []
    for b in range(min(steps, len(xs_dev))):
        x = sel(xs_dev[b])
        if index.distance == "cosine":
            x = T.fn.normalize(x)
        _, cells, npl = T.fn.coarse_probe(x, index.vq_codec.codebook, n_probe, index.use_smart_probing,
                                          index.smart_probing_temperature)
        P = npl.clamp(1, n_probe)
        sizes = index._cell_size[cells]
        owned = (cells % shards) == rank
        mask = (torch.arange(n_probe, device=dev)[None, :] < P[:, None]) & owned
        per_batch.append(int((sizes * mask).sum().item()) * M + 12 * k * x.shape[1])
        probes.append(float(P.float().mean().item()))
    total = sum(per_batch[i % len(per_batch)] for i in range(steps))
    return total, sum(probes) / len(probes)


def scan_roofline(index, xs_dev, k, reps, rank=0, shards=1, query_slice=None):
    """Roofline of the scan kernel on this rank in a stand-alone loop of `reps` searches (secondary workloads)."""
    dev = xs_dev[0].device
    sel = (lambda x: x) if query_slice is None else (lambda x: x[:, query_slice[0]:query_slice[1]].contiguous())
    torch.cuda.synchronize(dev)
    scan_profile_begin()
    for i in range(reps):
        index.search(sel(xs_dev[i % len(xs_dev)]), k=k)
    ms, n_launch = scan_profile_end()
    alg_bytes, probes = algorithmic_bytes(index, xs_dev, k, reps, rank, shards, query_slice)
    n = max(1, n_launch)
    return {"scan_ms_per_launch": ms / n, "algorithmic_bytes_per_launch": alg_bytes / n,
            "achieved_gbs": (alg_bytes / n) / (ms / n / 1e3) / 1e9 if n_launch else None,
            "mean_probes_scanned": probes, "launches": n_launch}


def profiled_traffic(workload, world):
    """dram bytes per scan launch from an ncu --set full capture of THIS workload at N=1 (profiles/scan_traffic.json);
    null anywhere it was not profiled."""
    if world != 1:
        return None
    tp = os.path.join(ROOT, "profiles", "scan_traffic.json")
    try:
        return json.load(open(tp)).get(workload)
    except Exception:
        return None


def measure_single_gpu(name, index, base, device, nq, steps, threads, with_oracle_parity=True, centres=None):
    """Device-timed QPS + scan roofline + recall on one workload (secondary block)."""
    N, d, M, C, n_probe, k, distance, n_train, desc = WORKLOADS[name]
    xs_host = gen_queries(d, nq, 4, device, centres=centres)
    xs_dev = [x.to(device) for x in xs_host]
    ms = cuda_time(lambda i: index.search(xs_dev[i % 4], k=k), steps, device=device)
    rf = scan_roofline(index, xs_dev, k, min(steps, 4))
    peak, _ = hbm_peak()
    nr = min(1000, nq)
    xr = xs_dev[0][:, :nr].contiguous()
    truth = exact_truth(base, xr, k, distance)
    v_r, ids_r = index.search(xr, k=k)
    out = {"workload": f"{name}: {desc}", "n_query_per_step": nq, "queries_per_s": nq / ms * 1e3, "ms_per_step": ms,
           "recall_at_100": recall(ids_r, truth), "use_smart_probing": bool(index.use_smart_probing),
           "mean_probes_scanned": rf["mean_probes_scanned"],
           "roofline": {"kernel": "ivfpq_scan_kernel", "achieved": rf["achieved_gbs"], "peak": peak, "unit": "GB/s",
                        "frac": rf["achieved_gbs"] / peak if rf["achieved_gbs"] else None,
                        "algorithmic_bytes_per_launch": rf["algorithmic_bytes_per_launch"],
                        "scan_ms_per_launch": rf["scan_ms_per_launch"], "scan_share_of_step": rf["scan_ms_per_launch"] / ms,
                        "step_frac_of_peak": rf["algorithmic_bytes_per_launch"] / (ms / 1e3) / 1e9 / peak}}
    if with_oracle_parity:
        try:
            st = to_oracle_state(index)
            out["vs_oracle"] = parity_vs_oracle(st, xs_host[0], v_r, ids_r, truth, k, threads, nr)
        except Exception as e:
            out["vs_oracle"] = {"error": repr(e)}
    return out


def reference_gpu_kernel(index, xs_dev, k, steps=3):
    """The reference's own `ivfpq_topk` kernel (kernels/cuda/ivfpq_topk.cu:822-971, compiled unmodified for sm_100a
    with its own launch parameters into oracle/_ref, tpb=256) on the SAME index state and queries, fed the LUT and the
    gathered cell tables the reference's host code would feed it.  SURVEY section 2b's bar: beat this kernel."""
    try:
        from oracle import ref_kernels as R
        import torchpq_b200 as T
        M = index.n_subvectors
        if not R.available(M):
            return {"unavailable": f"oracle/_ref has no ivfpq_topk build for M={M}"}
        if index.capacity > (1 << 24):
            return {"unavailable": "capacity above 2^24: the reference kernel carries addresses in a float"}
        x = xs_dev[0]
        xq = T.fn.normalize(x) if index.distance == "cosine" else x
        _, cells, npl = T.fn.coarse_probe(xq, index.vq_codec.codebook, int(index.n_probe), index.use_smart_probing,
                                          index.smart_probing_temperature)
        lut = T.fn.precompute_adc(xq, index.pq_codec.codebook, index.distance)
        cs, cz = index._cell_start[cells].contiguous(), index._cell_size[cells].contiguous()
        run = lambda i: R.ivfpq_topk(index._storage, lut, cs, cz, index._is_empty, npl, k)
        ms = cuda_time(run, steps, warmup=1, device=x.device)
        P = npl.clamp(1, int(index.n_probe))
        mask = torch.arange(cells.shape[1], device=x.device)[None, :] < P[:, None]
        byts = int((cz * mask).sum().item()) * M + 12 * k * x.shape[1]
        rv, ra = run(0)
        v, _, a = index.search(x, k=k, return_address=True)
        agree = float(torch.isclose(rv, v, rtol=1e-4, atol=0).all(dim=1).float().mean().item())
        peak, _ = hbm_peak()
        return {"kernel": "torchpq ivfpq_topk (tpb=256), compiled for sm_100a from /root/reference sources", "ms_per_launch": ms,
                "queries_per_s_scan_only": x.shape[1] / ms * 1e3, "scan_gbs": byts / (ms / 1e3) / 1e9,
                "frac_of_hbm_peak": byts / (ms / 1e3) / 1e9 / peak, "rows_with_values_equal_to_ours_rtol_1e-4": agree,
                "note": "scan kernel only (LUT, coarse probe and gathers prepared outside the timed region)"}
    except Exception as e:
        return {"error": repr(e)}


def secondary_block(device, threads, c3_index, c3_xs_dev, k3, steps):
    """C2 / C4 / clustered-C3 search lines, C5 k-means kernels, build side, reference GPU kernel (N=1 only)."""
    import torchpq_b200 as T
    from torchpq_b200 import build
    sec = {}
    peak, _ = hbm_peak()
    t0 = time.time()
    sec["reference_gpu_kernel_c3"] = reference_gpu_kernel(c3_index, c3_xs_dev, k3)
    for name in ("c2", "c4", "c3c"):
        try:
            clustered = name in CLUSTERED
            wl = WORKLOADS[name]
            index, base = build_index(wl, device, clustered=clustered)
            centres = cluster_centres(wl[1], wl[3], device) if clustered else None
            sec[name] = measure_single_gpu(name, index, base, device, 10000, steps, threads, centres=centres)
            if name == "c3c":
                # smart probing at the reference's default temperature (30) never prunes on these distances; a user who
                # wants pruning lowers `smart_probing_temperature` -- same index, same queries, ours vs the CPU oracle again
                sweep = []
                xs_c = [x.to(device) for x in gen_queries(wl[1], 10000, 4, device, centres=centres)]
                truth_c = exact_truth(base, xs_c[0][:, :1000].contiguous(), wl[5], wl[6])
                for temp in (1.0, 0.3, 0.1):
                    index.smart_probing_temperature = temp
                    ms = cuda_time(lambda i: index.search(xs_c[i % 4], k=wl[5]), 5, device=device)
                    _, _, npl = T.fn.coarse_probe(xs_c[0], index.vq_codec.codebook, wl[4], True, temp)
                    ids_t = index.search(xs_c[0][:, :1000].contiguous(), k=wl[5])[1]
                    row = {"temperature": temp, "queries_per_s": 10000 / ms * 1e3,
                           "mean_n_probe_list": float(npl.clamp(1, wl[4]).float().mean().item()),
                           "recall_at_100": recall(ids_t, truth_c)}
                    try:
                        st_c = to_oracle_state(index)
                        _, oi, _ = oracle_search(st_c, xs_c[0][:, :200].cpu().contiguous(), wl[5], threads)
                        row["rows_with_identical_id_sets_vs_oracle_200q"] = float(
                            (np.sort(ids_t[:200].cpu().numpy(), 1) == np.sort(oi, 1)).all(axis=1).mean())
                    except Exception as e:
                        row["oracle_error"] = repr(e)
                    sweep.append(row)
                index.smart_probing_temperature = 30.0
                sec[name]["smart_probing_temperature_sweep"] = sweep
                sec[name]["note"] = ("queries drawn from the same Gaussian mixture as the base: recall is meaningful; "
                                     "smart probing prunes once the temperature is lowered from the reference default of 30")
            del index, base
            torch.cuda.empty_cache()
        except Exception as e:
            sec[name] = {"error": repr(e)}
    # ---- batch-size sweep on the C3 index (also report nq in {1, 100, 1000}); small batches are cut into
    #      slices so that every SM has work, the slices are merged by merge_topk_kernel
    try:
        sweep = {}
        for nq_s in (1, 100, 1000):
            xs_s = [x[:, :nq_s].contiguous() for x in c3_xs_dev]
            ms = cuda_time(lambda i: c3_index.search(xs_s[i % 4], k=k3), 20, device=device)
            sweep[str(nq_s)] = {"ms_per_batch": ms, "queries_per_s": nq_s / ms * 1e3}
        sec["c3_batch_size_sweep"] = sweep
    except Exception as e:
        sec["c3_batch_size_sweep"] = {"error": repr(e)}
    # ---- residual IVFPQ (pq_use_residual=True) on the C3 shape, same queries
    try:
        wl = WORKLOADS["c3"]
        N, d, M, C, n_probe = wl[0], wl[1], wl[2], wl[3], wl[4]
        base = gen_base(d, N, device)
        tmp = T.IVFPQIndex(d, M, C, initial_size=1, device=str(device), pq_use_residual=True)
        build.train(tmp, base[:, :wl[7]].contiguous(), seed=0, vq_iters=VQ_ITERS, pq_iters=PQ_ITERS)
        cl, co = [], []
        for s0 in range(0, N, 1 << 20):
            c, q = build.encode(tmp, base[:, s0:s0 + (1 << 20)].contiguous()); cl.append(c); co.append(q)
        cells, codes = torch.cat(cl), torch.cat(co, 1)
        rix = T.IVFPQIndex(d, M, C, initial_size=int(torch.bincount(cells, minlength=C).max().item()), device=str(device),
                           pq_use_residual=True)
        rix.vq_codec.set_codebook(tmp.vq_codec.codebook); rix.pq_codec.set_codebook(tmp.pq_codec.codebook)
        build.container_add(rix, codes, cells)
        rix.n_probe = n_probe
        ms = cuda_time(lambda i: rix.search(c3_xs_dev[i % 4], k=k3), steps, device=device)
        truth = exact_truth(base, c3_xs_dev[0][:, :1000].contiguous(), k3, "euclidean")
        sec["c3_residual"] = {"queries_per_s": 10000 / ms * 1e3, "ms_per_step": ms,
                              "recall_at_100": recall(rix.search(c3_xs_dev[0][:, :1000].contiguous(), k=k3)[1], truth),
                              "note": "pq_use_residual=True, precomputed per-cell LUT half (reference on T4/Sift1M: 0.60x its plain index)"}
        del base, tmp, rix, cells, codes
        torch.cuda.empty_cache()
    except Exception as e:
        sec["c3_residual"] = {"error": repr(e)}
    # ---- C5: MultiKMeans assignment + centroid update (BASELINE config 5)
    try:
        l, d, n, kk = 64, 64, 1_000_000, 256
        data = torch.randn(l, d, n, device=device)
        cent = data[:, :, :kk].contiguous()
        byts = 4 * l * d * n + 4 * l * d * kk + 12 * l * n
        ms = cuda_time(lambda i: T.fn.max_sim(data, cent, exact=False, exact_values=False), 3, warmup=1, device=device)
        sec["c5_assign"] = {"kernel": "assign_tc_kernel (tcgen05 TF32)", "ms": ms, "gbs": byts / ms / 1e6,
                            "frac_of_hbm_peak": byts / ms / 1e6 / peak, "tflops_gemm_form": 2.0 * l * n * d * kk / ms / 1e9}
        lab = T.fn.max_sim(data, cent, exact=False, exact_values=False)[1]
        byts_c = 4 * l * d * n + 8 * l * n + 4 * l * d * kk
        ms = cuda_time(lambda i: T.fn.compute_centroids(data, lab, kk), 3, warmup=1, device=device)
        sec["c5_compute_centroids"] = {"ms": ms, "gbs": byts_c / ms / 1e6, "frac_of_hbm_peak": byts_c / ms / 1e6 / peak}
        del data, lab
        torch.cuda.empty_cache()
    except Exception as e:
        sec["c5"] = {"error": repr(e)}
    # ---- build side: add 1M vectors (encode + place) into a fresh index with the C3 codebooks
    try:
        wl = WORKLOADS["c3"]
        xb = gen_base(wl[1], 1_000_000, device, seed=77)
        ix2 = T.IVFPQIndex(wl[1], wl[2], wl[3], initial_size=512, device=str(device))
        ix2.vq_codec.set_codebook(c3_index.vq_codec.codebook); ix2.pq_codec.set_codebook(c3_index.pq_codec.codebook)
        build.encode(ix2, xb[:, :1 << 16].contiguous())
        torch.cuda.synchronize(); t1 = time.perf_counter()
        cells, codes = build.encode(ix2, xb)
        torch.cuda.synchronize(); t_enc = time.perf_counter() - t1
        t1 = time.perf_counter()
        build.container_add(ix2, codes, cells)
        torch.cuda.synchronize(); t_place = time.perf_counter() - t1
        sec["add_1M"] = {"encode_s": t_enc, "place_s": t_place, "vectors_per_s": 1e6 / (t_enc + t_place),
                         "encode_input_gbs": xb.numel() * 4 / t_enc / 1e9}
        del xb, ix2
        torch.cuda.empty_cache()
    except Exception as e:
        sec["add_1M"] = {"error": repr(e)}
    sec["seconds"] = time.time() - t0
    return sec


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
    ap.add_argument("--query-groups", type=int, default=0,
                    help="multi-GPU grid: ranks = cell_shards x query_groups (0 = auto: 1 up to 4 GPUs, 2 at 8)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"],
                    help="multi-GPU key exchange: fused scan + P2P push into symmetric memory (auto/p2p) or NCCL all-gather")
    ap.add_argument("--pipeline", type=int, default=2,
                    help="streams the timed loops alternate over: step i+1's probe/scan overlaps step i's exchange, merge "
                         "and tail (1 = strictly serial steps)")
    ap.add_argument("--split-coarse", default="auto", choices=["auto", "on", "off"],
                    help="multi-GPU coarse probe: split by queries inside a query group + all-gather of the probe lists (on), or "
                         "replicated per query group (off); auto = on without query groups, off with them (measured at 8 GPUs, "
                         "4 x 2: on 6.02 M device / 4.67 M e2e q/s, off 5.84 M / 5.26 M -- the extra collective and its host-side "
                         "packing cost the end-to-end loop more than the replicated 5000-query probe costs the device)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary block (c2/c4/c3c/C5/build/reference kernel)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    clustered = args.workload in CLUSTERED
    N, d, M, C, n_probe, k, distance, n_train, desc = wl
    have_cuda = torch.cuda.is_available()
    device = torch.device(f"cuda:{local_rank}" if have_cuda else "cpu")
    if have_cuda:
        torch.cuda.set_device(device)
    threads, thread_info = host_threads()
    torch.set_num_threads(threads)
    groups = args.query_groups if args.query_groups > 0 else (2 if world >= 8 else 1)
    if world % groups:
        groups = 1
    shards = world // groups
    split_coarse = (groups == 1) if args.split_coarse == "auto" else (args.split_coarse == "on")
    config = {"workload": f"{args.workload}: {desc}", "n_query_per_step": args.nq,
              "index_sharding": f"cells mod {shards}" + (f" x {groups} query groups" if groups > 1 else ""),
              "use_smart_probing": not args.no_smart, "pipeline_streams": max(1, args.pipeline),
              "coarse_probe": ("replicated per query group" if not split_coarse else "split by queries inside a query group + all-gather of probe lists") if world > 1 else "single GPU",
              "exchange": ("fused scan + P2P key push (symmetric memory) + barrier" if args.exchange != "nccl" else "NCCL all-gather") if world > 1 else "none",
              "l2_policy": "inputs larger than L2 (code store >= 640 MB vs 126 MB L2; 4 rotating query batches)",
              "training": f"seeded k-means, <= {VQ_ITERS} (coarse) / {PQ_ITERS} (PQ) Lloyd iterations, tol 1e-4 (the reference's settings)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        return run_reference(args, wl, device, threads, thread_info, config, clustered)

    assert have_cuda, "bench.py --impl ours needs a CUDA device (no CPU fallback)"
    import torch.distributed as dist
    import torchpq_b200 as T
    from torchpq_b200 import dist as tdist
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    assert T._lib.lib.tpq_device_supported(local_rank), "libtpq_b200.so is built for sm_100a"
    grid = (shards, groups)
    coarse_group = None
    if world > 1 and groups > 1:
        for g in range(groups):                                       # every rank creates every group, keeps its own
            pg = dist.new_group(list(range(g * shards, (g + 1) * shards)))
            if rank // shards == g:
                coarse_group = pg

    nq = args.nq
    centres = cluster_centres(d, C, device) if clustered else None
    xs_host = gen_queries(d, nq, 4, device, centres=centres)
    xs_dev = [x.to(device) for x in xs_host]

    t0 = time.time()
    st_cpu, truth, unsharded0, full_bytes = None, None, None, None
    nr = min(1000, nq)
    if rank == 0:
        index, base = build_index(wl, device, clustered=clustered)
        index.use_smart_probing = not args.no_smart
        # everything that needs the FULL index happens here, before the index is cut into shards
        truth = exact_truth(base, xs_dev[0][:, :nr].contiguous(), k, distance)
        del base
        torch.cuda.empty_cache()
        index.layout()
        full_bytes = index.resident_bytes()                            # reference buffers + scan layout of all cells
        if world > 1:
            unsharded0 = index.search(xs_dev[0], k=k)
        try:
            st_cpu = to_oracle_state(index)
        except Exception as e:
            log(f"oracle state export failed: {e!r}")
    else:
        index = T.IVFPQIndex(d, M, C, initial_size=1, distance=distance, device=str(device))
        index.n_probe = n_probe
    index.use_smart_probing = not args.no_smart
    if world > 1:
        tdist.distribute(index, 0, grid=grid)
    lay = index.layout()
    torch.cuda.synchronize()
    resident = index.resident_bytes()
    log(f"[rank {rank}] index ready in {time.time() - t0:.1f}s: capacity={index.capacity} blocks={lay.n_blocks} resident={resident / 1e6:.0f} MB")

    def search(x):
        if world > 1:
            return tdist.sharded_search(index, x, k, grid=grid, coarse_group=coarse_group, exchange=args.exchange,
                                        split_coarse=split_coarse)
        return index.search(x, k=k)

    # Steps are independent batches, so consecutive steps alternate over `--pipeline` streams: step i+1's coarse probe and
    # the head of its scan run while step i's exchange barrier, merge and last scan wave finish (all inside the timed
    # region; the roofline leg below times strictly serial steps).
    n_pipe = max(1, args.pipeline)
    pipe = [torch.cuda.Stream(device) for _ in range(n_pipe)] if n_pipe > 1 else [torch.cuda.current_stream(device)]

    def step_device(i):
        with torch.cuda.stream(pipe[i % n_pipe]):
            return search(xs_dev[i % len(xs_dev)])

    def step_serial(i):
        return search(xs_dev[i % len(xs_dev)])

    # end-to-end leg: host (pinned) queries in, host (pinned) results out, every step.  H2D and D2H run on their own
    # streams with staging tensors per pipeline slot, so step i's D2H and step i+1's H2D overlap the searches -- all of it
    # inside the timed region.  (Round 1 used ONE copy stream: H2D i+1 then queued behind D2H i, which waits for search
    # i, so every step paid search + D2H + H2D in series -- 1.07 ms of the 2.88 ms e2e step at 8 GPUs.)
    h2d_stream, d2h_stream = torch.cuda.Stream(device), torch.cuda.Stream(device)
    n_slot = 2 * n_pipe
    out_v = [torch.empty(nq, k, dtype=torch.float32).pin_memory() for _ in range(n_slot)]
    out_i = [torch.empty(nq, k, dtype=torch.long).pin_memory() for _ in range(n_slot)]
    x_in = [torch.empty(d, nq, dtype=torch.float32, device=device) for _ in range(n_slot)]
    ev_in = [torch.cuda.Event() for _ in range(n_slot)]
    ev_done = [torch.cuda.Event() for _ in range(n_slot)]
    ev_free = [torch.cuda.Event() for _ in range(n_slot)]
    res = [None] * n_slot

    def step_e2e(i):
        b = i % n_slot
        main = pipe[i % n_pipe]
        with torch.cuda.stream(h2d_stream):
            h2d_stream.wait_event(ev_free[b])                        # the search that last used x_in[b] has read it
            x_in[b].copy_(xs_host[i % len(xs_host)], non_blocking=True)
            ev_in[b].record(h2d_stream)
        with torch.cuda.stream(main):
            main.wait_event(ev_in[b])
            v, ids = search(x_in[b])
            ev_free[b].record(main)
            ev_done[b].record(main)
        res[b] = (v, ids)                                            # keep alive until the copy stream has read them
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(ev_done[b])
            out_v[b].copy_(v, non_blocking=True); out_i[b].copy_(ids, non_blocking=True)
            v.record_stream(d2h_stream); ids.record_stream(d2h_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        cur = torch.cuda.current_stream(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in pipe + [h2d_stream, d2h_stream]:
            st.wait_event(e0)                                        # nothing of the timed work starts before e0
        for i in range(steps):
            fn(i)
        for st in pipe + [h2d_stream, d2h_stream]:
            cur.wait_stream(st)                                      # ... and all of it ends before e1
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local_rank) if rank == 0 else None      # samples across warm-up and both timed regions
    for i in range(max(args.warmup, 2 * n_pipe)):                   # every pipeline slot warmed (symmetric-memory set-up included)
        step_device(i); step_e2e(i); step_serial(i)
    scan_profile_begin()                                          # roofline leg: K steps strictly one after the other,
    ms_serial = timed(step_serial, args.steps)                    # CUDA events around every scan launch (tpq_profile_*)
    scan_ms_total, scan_launches = scan_profile_end()
    ms_dev = timed(step_device, args.steps)
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if sampler else None
    qps = nq * args.steps / (ms_dev / 1e3)
    qps_e2e = nq * args.steps / (ms_e2e / 1e3)

    # ---- roofline of the scan kernel on this rank (its own query group, its own cells)
    barrier()
    per_group = (nq + groups - 1) // groups
    qsl = None if groups == 1 else (min(nq, (rank // shards) * per_group), min(nq, (rank // shards + 1) * per_group))
    alg_total, mean_probes = algorithmic_bytes(index, xs_dev, k, max(1, scan_launches), rank=rank % shards, shards=shards, query_slice=qsl)
    nl = max(1, scan_launches)
    rf = {"scan_ms_per_launch": scan_ms_total / nl, "algorithmic_bytes_per_launch": alg_total / nl,
          "achieved_gbs": (alg_total / nl) / (scan_ms_total / nl / 1e3) / 1e9 if scan_launches else None,
          "mean_probes_scanned": mean_probes, "launches": scan_launches}
    peak, peak_src = hbm_peak()

    # per-rank resident bytes (max over ranks) -- "shard the index" means memory per rank falls with the shard count
    res_t = torch.tensor([float(resident)], device=device)
    if world > 1:
        dist.all_reduce(res_t, op=dist.ReduceOp.MAX)

    # sharded run: does the cell-sharded search (all ranks, NCCL) return exactly what the unsharded index returned?
    sharded_ok = None
    v_r = ids_r = None
    try:
        v_s, i_s = search(xs_dev[0])
        torch.cuda.synchronize()
        v_r, ids_r = v_s[:nr], i_s[:nr]
        if world > 1 and rank == 0:
            sharded_ok = bool(torch.equal(v_s, unsharded0[0]) and torch.equal(i_s, unsharded0[1]))
    except Exception as e:                                           # report, never take the number down
        sharded_ok = f"check failed: {e!r}"
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- recall@100 (ours vs exact brute force) on 1000 queries; CPU oracle: timing sample + parity on the same 1000
    rec = recall(ids_r, truth) if ids_r is not None else None
    cpu = None
    try:
        qps_cpu, n_cpu, used = time_oracle(st_cpu, xs_host[0].clone(), k, threads, args.cpu_sample)
        par = parity_vs_oracle(st_cpu, xs_host[0], v_r, ids_r, truth, k, threads, nr)
        cpu = {"value": qps_cpu, "unit": "queries/s", "cores": used, "kind": "port", "threads_source": thread_info,
               "sample": f"{n_cpu} of the {nq} queries of batch 0, full {args.workload} index, oracle (torch-CPU matmul + OpenMP C scan)",
               "recall_at_100": par.get("recall_at_100_oracle"), "recall_at_100_ours_same_queries": par.get("recall_at_100_ours"),
               "parity_queries": par["n_queries"], "rows_with_identical_id_sets": par["rows_with_identical_id_sets"],
               "id_overlap": par["id_overlap"], "values_max_rel_err": par["values_max_rel_err"],
               "values_within_1e-3": par["values_within_1e-3"]}
    except Exception as e:  # the CPU leg must not take the GPU number down with it
        cpu = {"value": None, "error": repr(e)}
    st_cpu = None

    secondary = None
    if world == 1 and args.workload == "c3" and not args.no_secondary:
        try:
            secondary = secondary_block(device, threads, index, xs_dev, k, min(args.steps, 10))
        except Exception as e:
            secondary = {"error": repr(e)}

    # our kernels per step of the `value` loop: 2 x col_sqnorm, coarse_gemm, probe_select, ivfpq_scan (+ normalize_columns for
    # cosine, + lut_scan when d/M is not 1/2/4/8), then one merge_topk on a single GPU; sharded: one merge_topk per query
    # group, plus the rank-local merge/decode when the exchange is the NCCL all-gather (the fused push skips it).  The
    # symmetric-memory barrier and NCCL kernels are torch's, not counted.
    launches_per_step = 5 + (1 if distance == "cosine" else 0) + (0 if (d // M) in (1, 2, 4, 8) else 1)
    launches_per_step += 1 if world == 1 else groups + (1 if args.exchange == "nccl" else 0)
    ach = rf["achieved_gbs"]
    line = {
        "metric": "queries/sec @ recall@100, IVFPQ search", "value": qps, "unit": "queries/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8 codes / fp32 LUT+accumulate",
        "data": "synthetic (" + ("Gaussian-mixture" if clustered else "randn") + " base seed 1234, queries seed 4321; codebooks trained on-device by seeded k-means)",
        "config": config,
        "recall_at_100": rec, "sharded_equals_unsharded": sharded_ok,
        "resident_bytes_per_rank_max": res_t.item(), "resident_bytes_full_index_one_gpu": full_bytes,
        "e2e": {"value": qps_e2e, "unit": "queries/s", "h2d_bytes_per_step": d * nq * 4, "d2h_bytes_per_step": nq * k * 12,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": {"kernel": "ivfpq_scan_kernel", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": (ach / peak) if ach else None, "traffic": profiled_traffic(args.workload, world), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": rf["algorithmic_bytes_per_launch"],
                     "scan_ms_per_launch": rf["scan_ms_per_launch"],
                     "serial_ms_per_step": ms_serial / args.steps,
                     "scan_share_of_step": rf["scan_ms_per_launch"] / (ms_serial / args.steps),
                     "mean_probes_scanned": rf["mean_probes_scanned"], "rank": 0, "launches_timed": rf["launches"],
                     "timed_in": "a third timed loop of the same K steps run strictly serially (CUDA events around every scan launch); "
                                 "`value` and `e2e` alternate steps over `config.pipeline_streams` streams"},
        "cpu_baseline": cpu, "clocks": clocks, "secondary": secondary,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_reference(args, wl, device, threads, thread_info, config, clustered):
    """CPU arm: the oracle restatement on the host cores; each step = a bounded query sample."""
    N, d, M, C, n_probe, k, distance, n_train, desc = wl
    st = build_state_torch(wl, device, clustered=clustered)   # plain torch ops only; nothing of torchpq_b200 is imported on this arm
    st.use_smart_probing = not args.no_smart
    if device.type == "cuda":
        torch.cuda.empty_cache()
    centres = cluster_centres(d, C, torch.device("cpu")) if clustered else None
    xs = gen_queries(d, args.nq, 1, torch.device("cpu"), centres=centres)[0]
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
        "config": config,
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": used, "kind": "port", "threads_source": thread_info,
                         "sample": f"{per_step} of the {args.nq} queries per step against the full {args.workload} index "
                                   "(reference has no CPU path; oracle restatement: torch-CPU matmuls + OpenMP C scan)",
                         "sample_queries_per_step": per_step},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
