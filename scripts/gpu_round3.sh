mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kmeans_gpu.py tests/test_search_gpu.py -m gpu -q 2>&1 | tail -3
timeout 600 python scripts/bench_aux.py centroids c5fast c5 2>&1 | grep -v "^$" | tail -4
# launch list of the default bench command (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"coarse_gemm|probe_select|ivfpq_scan|merge_topk|col_sqnorm|normalize_columns|lut_scan" -c 80 --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-secondary --cpu-sample 0.2 > gpurun_out/r02_launches_c3.log 2>&1
tail -2 gpurun_out/r02_launches_c3.log | cut -c1-300
bash scripts/prof_scan.sh r02_c3 c3 1
bash scripts/prof_scan.sh r02_c4 c4 1
bash scripts/prof_scan.sh r02_c3_shard8 c3 8
