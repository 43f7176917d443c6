#!/bin/bash
# One gpurun call: full-size property test, bench line, ncu launch list of one pass.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "cpus: $(nproc)" >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests/test_flashsr_gpu.py -q -m gpu -k full -p no:cacheprovider > gpurun_out/full.log 2>&1; echo "exit $?" >> gpurun_out/full.log; tail -n 5 gpurun_out/full.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 1 1 1 > gpurun_out/profile_step.log 2>&1; echo "ncu exit $?"; tail -n 3 gpurun_out/profile_step.log
