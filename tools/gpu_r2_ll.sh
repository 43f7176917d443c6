#!/bin/bash
# 2-GPU, final code of round 2: sharded NCCL equality test + bench N=2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_nccl_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2ll_nccl_test.log 2>&1; echo "nccl test rc=$?"; tail -n 3 gpurun_out/r2ll_nccl_test.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2ll_bench_n2.json 2> gpurun_out/r2ll_bench_n2.err
echo "bench rc=$?"
tail -c 3500 gpurun_out/r2ll_bench_n2.json
