#!/bin/bash
# 2-GPU bench with the default sub-batch of 48
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2pp_bench_n2.json 2> gpurun_out/r2pp_bench_n2.err
echo "bench rc=$?"
python3 - <<'PY'
import json
d=json.loads(open('gpurun_out/r2pp_bench_n2.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print('c3',d['c3']['value'],d['c3']['seconds'],d['c3']['phases_max_over_ranks'])
print('c5',d['chain_c5']['value'])
PY
