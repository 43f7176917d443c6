#!/bin/bash
# 8-GPU, final code of round 2: bench N=8 (c2 weak scaling + c3 strong scaling with phase times + c5 chain)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2mm_bench_n8.json 2> gpurun_out/r2mm_bench_n8.err
echo "bench rc=$?"
python3 - <<'PY'
import json
d=json.loads(open('gpurun_out/r2mm_bench_n8.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'])
print('c3',json.dumps(d['c3']))
print('c5',json.dumps(d['chain_c5']))
PY
tail -n 3 gpurun_out/r2mm_bench_n8.err
