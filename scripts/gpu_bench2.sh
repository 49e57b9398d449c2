mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q 2>&1 | tail -3
( time timeout 1500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err ) 2>&1 | grep real
tail -3 gpurun_out/bench_n1.err
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err ) 2>&1 | grep real
tail -3 gpurun_out/bench_n2.err
