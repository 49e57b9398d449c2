"""CPU: small pieces of host logic that need no GPU."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_make_grid():
    from torchpq_b200.dist import make_grid
    assert make_grid(8) == (8, 1) and make_grid(8, 2) == (4, 2) and make_grid(8, 8) == (1, 8) and make_grid(1) == (1, 1)
    with pytest.raises(AssertionError):
        make_grid(8, 3)


def test_host_threads_respects_affinity_and_quota():
    import bench
    n, info = bench.host_threads()
    assert 1 <= n <= (info["affinity"] or os.cpu_count())
    if info["cgroup_quota_cpus"]:
        assert n <= max(1, int(info["cgroup_quota_cpus"] + 1e-9))


def test_workload_table_matches_baseline_configs():
    """bench.WORKLOADS c2 / c3 / c4 are BASELINE.json configs[1..3] (shape, M, n_cells, n_probe, k, distance)."""
    import json
    import bench
    cfg = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    c2, c3, c4 = bench.WORKLOADS["c2"], bench.WORKLOADS["c3"], bench.WORKLOADS["c4"]
    assert (c2[0], c2[1], c2[2], c2[3], c2[4], c2[5]) == (1_000_000, 128, 64, 1024, 32, 100) and "1M" in cfg[1] and "n_cells=1024" in cfg[1]
    assert (c3[0], c3[1], c3[2], c3[3], c3[4], c3[5]) == (10_000_000, 128, 64, 4096, 32, 100) and "10M" in cfg[2] and "n_cells=4096" in cfg[2]
    assert (c4[0], c4[1], c4[2], c4[3], c4[4], c4[6]) == (1_000_000, 960, 120, 1024, 64, "cosine") and "M=120" in cfg[3] and "nprobe=64" in cfg[3]
