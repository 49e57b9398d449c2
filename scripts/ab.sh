mkdir -p gpurun_out; rm -f gpurun_out/sweep_scan.jsonl
export TPQ_B200_LIB=$PWD/torchpq_b200/libtpq_b200_dbg.so
export TPQ_BOOT_MODE=1
timeout 600 python scripts/sweep_scan.py c4 1 2>&1 | tail -1
echo "== r01"
TPQ_B200_LIB=$PWD/torchpq_b200/libtpq_b200_r01dbg.so timeout 600 python scripts/sweep_scan.py c3 1 8 2>&1 | tail -2
TPQ_B200_LIB=$PWD/torchpq_b200/libtpq_b200_r01dbg.so timeout 600 python scripts/sweep_scan.py c4 1 2>&1 | tail -1
echo "== boot_r 1 (tree)"
TPQ_BOOT_R=1 timeout 600 python scripts/sweep_scan.py c3 1 8 2>&1 | tail -2
echo "== boot_r 2 (tree)"
timeout 600 python scripts/sweep_scan.py c3 1 8 2>&1 | tail -2
echo "== 8x2 at shard 8 / 8x3 at shard 1"
TPQ_SCAN_CFG=8x2 timeout 600 python scripts/sweep_scan.py c3 8 2>&1 | tail -1
TPQ_SCAN_CFG=8x3 timeout 600 python scripts/sweep_scan.py c3 1 2>&1 | tail -1
