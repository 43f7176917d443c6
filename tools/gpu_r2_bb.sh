#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
  EGR_BENCH_C3=0 EGR_BENCH_C5=0 EGR_BENCH_CPU=0 EGR_BENCH_EAGER=0 EGR_BENCH_PATHB=0 EGR_BENCH_BATCHED=0 timeout 300 python bench.py --steps 10 --warmup 3 2>gpurun_out/r2bb_err_$i.txt | python tools/bench_brief.py
done
uptime
