#!/bin/bash
mkdir -p gpurun_out
LEAN="EGR_BENCH_CPU=0 EGR_BENCH_PATHB=0 EGR_BENCH_C5=0 EGR_BENCH_EAGER=0"
for b in 8 12 16; do
  echo "=== EGREGORA_FLASHSR_BATCH=$b"
  env $LEAN EGREGORA_FLASHSR_BATCH=$b timeout 900 python bench.py > gpurun_out/r2l_bench_b$b.json 2> gpurun_out/r2l_bench_b$b.err; echo "exit $?"
  python - <<PY
import json
d=json.load(open("gpurun_out/r2l_bench_b$b.json"))
print("c2 ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], "unet ms", d["roofline"]["unet_region"]["ms"])
print("c3", d["c3"]["value"], d["c3"]["seconds"], d["c3"]["phases_max_over_ranks"])
print("batched", d["batched"])
PY
done
timeout 300 python tools/section_times.py 16 1 2>/dev/null | tail -7
