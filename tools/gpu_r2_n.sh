#!/bin/bash
# 2-GPU: e2e phase diagnostic + sharded equality + bench N=2
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/e2e_phases_nccl.py > gpurun_out/r2n_phases.log 2>&1
echo "phases rc=$?"
grep "^\[r" gpurun_out/r2n_phases.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2n_bench_n2.json 2> gpurun_out/r2n_bench_n2.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/r2n_bench_n2.json
nvidia-smi topo -m > gpurun_out/r2n_topo.txt 2>&1
