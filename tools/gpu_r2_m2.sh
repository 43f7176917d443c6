#!/bin/bash
# 2 GPUs: N-GPU == 1-GPU with the real engine (NCCL), then the bench at N = 2
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_sharded_nccl_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2m_nccl_test.log 2>&1; echo "nccl test exit $?"; tail -n 6 gpurun_out/r2m_nccl_test.log
LEAN="EGR_BENCH_C5=0"
env $LEAN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2m_bench_n2.json 2> gpurun_out/r2m_bench_n2.err; echo "bench n2 exit $?"; cat gpurun_out/r2m_bench_n2.json | cut -c1-3000; tail -n 5 gpurun_out/r2m_bench_n2.err
