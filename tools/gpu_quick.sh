#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-300} python -m pytest "$@" -q -m gpu -p no:cacheprovider -x > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n 4 gpurun_out/$name.log; }
run ops_tc tests/test_ops_gpu.py -k "conv or attention_gemm"
run flashsr tests/test_flashsr_gpu.py
timeout 600 python tools/gemm_probe.py "${PROBE:-}" 2>&1 | tee gpurun_out/probe.txt
EGR_BENCH_CPU=0 EGR_BENCH_PATHB=0 timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_quick.json 2> gpurun_out/bench.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'roof', d['roofline']['achieved'], d['roofline']['gemm_share_of_plan'], 'batched', d['batched'])
PY
tail -n 3 gpurun_out/bench.err
