mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_build_gpu.py tests/test_search_gpu.py tests/test_kmeans_gpu.py -m gpu -q --timeout=600 2>&1 | grep -E "Error|error|assert|passed|failed" | head -20
timeout 600 python scripts/bench_aux.py centroids 2>&1 | grep -v "^$" | tail -1
