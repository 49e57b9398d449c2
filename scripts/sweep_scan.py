"""Experiment driver (GPU box): time the scan kernel under different launch shapes on one workload.
usage: python scripts/sweep_scan.py [workload] [cfg ...]"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import torchpq_b200 as T

wl_name = sys.argv[1] if len(sys.argv) > 1 else "c3"
shard = int(os.environ.get("SHARD_WORLD", "1"))          # emulate one rank of a cell-sharded index on this GPU
cfgs = sys.argv[2:] or ["8x2", "8x3", "12x2", "16x1", "4x4", "4x6"]
wl = bench.WORKLOADS[wl_name]
dev = torch.device("cuda:0")
index, base = bench.build_index(wl, dev)
del base
index.set_shard(0, shard)
k = wl[5]
xs = [x.to(dev) for x in bench.gen_queries(wl[1], 10000, 4, dev)]
lib = T._lib.lib
for smart in (True,):
    index.use_smart_probing = smart
    for cfg in cfgs:
        if cfg.startswith("boot"):
            os.environ["TPQ_BOOT_R"] = cfg[4:]; os.environ.pop("TPQ_SCAN_CFG", None)
        else:
            os.environ["TPQ_SCAN_CFG"] = cfg
        for i in range(2):
            index.search(xs[i], k=k)
        torch.cuda.synchronize()
        lib.tpq_profile_enable(1)
        t0 = time.perf_counter()
        for i in range(8):
            index.search(xs[i % 4], k=k)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 8
        ms, n = ctypes.c_float(0), ctypes.c_int(0)
        lib.tpq_profile_scan_ms(ctypes.byref(ms), ctypes.byref(n))
        lib.tpq_profile_enable(0)
        print(f"smart={smart} cfg={cfg:6s} scan {ms.value / n.value:8.3f} ms/launch   search {dt * 1e3:8.3f} ms/batch", flush=True)
