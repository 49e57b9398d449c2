mkdir -p gpurun_out
run() { # n tag extra-args...
  n=$1; tag=$2; shift 2
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 20 --warmup 3 --cpu-sample 2 "$@" > gpurun_out/scale_$tag.json 2> gpurun_out/scale_$tag.err ) 2>&1 | grep real
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "qps=%.0f e2e=%.0f ms=%.3f e2e_ms=%.3f scan_ms=%.3f frac=%.3f eq=%s resident=%.0fMB" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["scan_ms_per_launch"], d["roofline"]["frac"], d["sharded_equals_unsharded"], d["resident_bytes_per_rank_max"]/1e6))
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/scale_$tag.err").read()[-1500:])
PY
}
run 1 n1 --no-secondary
run 8 n8_default
run 8 n8_nosplit --no-split-coarse
run 4 n4_default
run 2 n2_default
