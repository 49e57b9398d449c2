#!/bin/bash
# One GPU-box visit: parity tests, then the scan sweeps with the experiment build.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import cupy; print('cupy', cupy.__version__)" > gpurun_out/cupy_probe.txt 2>&1 || echo "cupy: not importable" >> gpurun_out/cupy_probe.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
rm -f gpurun_out/sweep_scan.jsonl
timeout 600 python scripts/sweep_scan.py c3 1 8 2>&1 | tail -2
export TPQ_B200_LIB=$PWD/torchpq_b200/libtpq_b200_dbg.so
timeout 600 python scripts/sweep_scan.py c3 1 8 2>&1 | tail -4
timeout 300 python scripts/sweep_scan.py c2 1 2>&1 | tail -2
timeout 600 python scripts/sweep_scan.py c4 1 2>&1 | tail -2
NQ=1000 timeout 600 python scripts/sweep_scan.py c4 1 2>&1 | tail -2
