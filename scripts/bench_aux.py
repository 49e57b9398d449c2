"""Secondary measurements on one B200 (run under gpurun; results -> gpurun_out/aux_bench.json):
C5 MultiKMeans assignment (GB/s vs HBM roofline), compute_centroids, pq_decode, and search QPS on C2 / C4.
CUDA events, 3 warm-ups, inputs larger than L2 where the workload allows."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchpq_b200 as T
import bench

dev = torch.device("cuda:0")
out = {}
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


which = sys.argv[1:] or ["c5", "centroids", "decode", "c2", "c4"]
if "c5" in which or "centroids" in which:
    l, d, n, k = 64, 64, 1_000_000, 256
    data = torch.randn(l, d, n, device=dev)
    cent = data[:, :, :k].contiguous()
    if "c5" in which:
        byts = 4 * l * d * n + 4 * l * d * k + 12 * l * n
        for exact, ev, name, kern in ((False, False, "c5_max_sim_tc", "assign_tc_kernel (tcgen05 TF32, TF32-derived maxsim)"),
                                      (False, True, "c5_max_sim_tc_exact_values", "assign_tc_kernel (tcgen05 TF32 + exact fp32 maxsim)"),
                                      (True, True, "c5_max_sim_exact", "max_sim_kernel (fp32 SIMT, exact)")):
            if exact and "c5fast" in which:
                continue
            ms = timeit(lambda: T.fn.max_sim(data, cent, exact=exact, exact_values=ev), reps=3, warm=1)
            out[name] = {"ms": ms, "algorithmic_GB": byts / 1e9, "GBps": byts / ms / 1e6, "frac_of_hbm_peak": byts / ms / 1e6 / peak,
                         "tflops_gemm_form": 2.0 * l * n * d * k / ms / 1e9, "kernel": kern}
            print(out[name], flush=True)
        lt = T.fn.max_sim(data[:2], cent[:2], exact=False)[1]
        le = T.fn.max_sim(data[:2], cent[:2], exact=True)[1]
        out["c5_label_agreement_tc_vs_exact"] = float((lt == le).float().mean())
    if "centroids" in which:
        lab = T.fn.max_sim(data, cent)[1]
        ms = timeit(lambda: T.fn.compute_centroids(data, lab, k), reps=3, warm=1)
        byts = 4 * l * d * n + 8 * l * n + 4 * l * d * k
        out["c5_compute_centroids"] = {"ms": ms, "algorithmic_GB": byts / 1e9, "GBps": byts / ms / 1e6, "frac_of_hbm_peak": byts / ms / 1e6 / peak}
        print(out["c5_compute_centroids"], flush=True)
    del data
if "decode" in which:
    M, dsub, n = 64, 2, 10_000_000
    cb = torch.randn(M, dsub, 256, device=dev)
    code = torch.randint(0, 256, (M, n), dtype=torch.uint8, device=dev)
    ms = timeit(lambda: T.fn.pq_decode(cb, code))
    byts = M * n + 4 * M * dsub * n
    out["pq_decode_10M"] = {"ms": ms, "algorithmic_GB": byts / 1e9, "GBps": byts / ms / 1e6, "frac_of_hbm_peak": byts / ms / 1e6 / peak}
    print(out["pq_decode_10M"], flush=True)
    del code
for name in ("c2", "c4"):
    if name not in which:
        continue
    wl = bench.WORKLOADS[name]
    index, base = bench.build_index(wl, dev)
    nq = 10000 if name == "c2" else 1000
    xs = [x.to(dev) for x in bench.gen_queries(wl[1], nq, 4, dev)]
    k = wl[5]
    T._lib.lib.tpq_profile_enable(0)
    i = [0]
    def step():
        index.search(xs[i[0] % 4], k=k); i[0] += 1
    ms = timeit(step, reps=10)
    truth = bench.exact_truth(base, xs[0][:, :500].contiguous(), k, wl[6])
    rec = bench.recall(index.search(xs[0][:, :500].contiguous(), k=k)[1], truth)
    out[f"{name}_search"] = {"nq": nq, "ms_per_batch": ms, "qps": nq / ms * 1e3, "recall_at_100": rec, "workload": wl[8]}
    print(out[f"{name}_search"], flush=True)
    del index, base
if "c3sweep" in which or "build" in which:
    wl = bench.WORKLOADS["c3"]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    index, base = bench.build_index(wl, dev)
    torch.cuda.synchronize(); t_build = time.perf_counter() - t0
    out["c3_build"] = {"seconds_total": t_build, "what": "gen 10M x 128 randn + train (4+4 Lloyd iters on 1M) + encode 10M + container_add",
                       "reference_T4_sift1m": "train 4.45 s + add 10.72 s for 1M (BASELINE.md)"}
    # add path alone: encode + place 1M fresh vectors into an empty index with the trained codebooks
    from torchpq_b200 import build
    ix2 = T.IVFPQIndex(wl[1], wl[2], wl[3], initial_size=4096, device="cuda:0")
    ix2.vq_codec.set_codebook(index.vq_codec.codebook); ix2.pq_codec.set_codebook(index.pq_codec.codebook)
    xb = base[:, :1_000_000].contiguous()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ix2.add(xb)
    torch.cuda.synchronize(); out["add_1M"] = {"seconds": time.perf_counter() - t0, "vectors_per_s": 1e6 / (time.perf_counter() - t0)}
    print(out["c3_build"], out["add_1M"], flush=True)
    del ix2, xb
    if "c3sweep" in which:
        k = wl[5]
        res = {}
        for nq in (1, 10, 100, 1000, 10000):
            xs = [x.to(dev) for x in bench.gen_queries(wl[1], nq, 4, dev)]
            i = [0]
            def step():
                index.search(xs[i[0] % 4], k=k); i[0] += 1
            ms = timeit(step, reps=20)
            res[nq] = {"ms_per_batch": ms, "qps": nq / ms * 1e3}
        out["c3_batch_size_sweep"] = res
        print(res, flush=True)
    del index, base
if "c3res" in which:
    # residual IVFPQ on the C3 shape (reference: 72k q/s vs 120k q/s plain on T4/Sift1M)
    wl = bench.WORKLOADS["c3"]
    N, d, M, Cn, n_probe, k = wl[0], wl[1], wl[2], wl[3], wl[4], wl[5]
    from torchpq_b200 import build
    base = bench.gen_base(d, N, dev)
    tmp = T.IVFPQIndex(d, M, Cn, initial_size=1, device="cuda:0", pq_use_residual=True)
    build.train(tmp, base[:, :wl[7]].contiguous(), seed=0, vq_iters=4, pq_iters=4)
    cl, co = [], []
    for s0 in range(0, N, 1 << 20):
        c, q = build.encode(tmp, base[:, s0:s0 + (1 << 20)].contiguous()); cl.append(c); co.append(q)
    cells, codes = torch.cat(cl), torch.cat(co, 1)
    index = T.IVFPQIndex(d, M, Cn, initial_size=int(torch.bincount(cells, minlength=Cn).max().item()), device="cuda:0", pq_use_residual=True)
    index.vq_codec.set_codebook(tmp.vq_codec.codebook); index.pq_codec.set_codebook(tmp.pq_codec.codebook)
    build.container_add(index, codes, cells)
    index.n_probe = n_probe
    xs = [x.to(dev) for x in bench.gen_queries(d, 10000, 4, dev)]
    i = [0]
    def step():
        index.search(xs[i[0] % 4], k=k); i[0] += 1
    ms = timeit(step, reps=10)
    truth = bench.exact_truth(base, xs[0][:, :500].contiguous(), k, "euclidean")
    rec = bench.recall(index.search(xs[0][:, :500].contiguous(), k=k)[1], truth)
    out["c3_residual_search"] = {"nq": 10000, "ms_per_batch": ms, "qps": 1e4 / ms * 1e3, "recall_at_100": rec}
    print(out["c3_residual_search"], flush=True)
    del index, base, tmp
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/aux_bench.json", "w"), indent=1)
