#!/bin/bash
# sub-batch size of the engine (EGREGORA_FLASHSR_BATCH): c3 through node.run() on one GPU at 16 / 24 / 32 / 48
mkdir -p gpurun_out
for b in 16 32 48; do
  EGREGORA_FLASHSR_BATCH=$b EGR_BENCH_CPU=0 EGR_BENCH_EAGER=0 EGR_BENCH_PATHB=0 EGR_BENCH_C5=0 EGR_BENCH_BATCHED=0 timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r2nn_b$b.json 2> gpurun_out/r2nn_b$b.err
  python3 - <<PY
import json
d=json.loads(open('gpurun_out/r2nn_b$b.json').read().strip().splitlines()[-1])
print('sub-batch $b: c3', d['c3']['value'], 'x RT', d['c3']['seconds'], 's model_ms', d['c3']['phases_max_over_ranks']['model_ms'])
PY
  timeout 200 python tools/section_times.py $b 4 2>/dev/null | tail -6
  nvidia-smi --query-gpu=memory.used --format=csv,noheader
done
