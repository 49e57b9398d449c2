mkdir -p gpurun_out
run() { # n groups exchange tag
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $1 --steps 20 --warmup 3 --query-groups $2 --exchange $3 --cpu-sample 2 > gpurun_out/scale_$4.json 2> gpurun_out/scale_$4.err ) 2>&1 | grep real
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_$4.json").read().strip().splitlines()[-1])
    print("$4", "qps=%.0f e2e=%.0f ms=%.3f e2e_ms=%.3f scan_ms=%.3f frac=%.3f eq=%s resident=%.0fMB" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["scan_ms_per_launch"], d["roofline"]["frac"], d["sharded_equals_unsharded"], d["resident_bytes_per_rank_max"]/1e6))
except Exception as e:
    print("$4 failed", e); print(open("gpurun_out/scale_$4.err").read()[-1500:])
PY
}
run 8 2 p2p n8_4x2_p2p
run 8 2 nccl n8_4x2_nccl
run 8 1 p2p n8_8x1_p2p
run 4 1 p2p n4_4x1_p2p
run 2 1 p2p n2_p2p
