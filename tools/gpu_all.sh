#!/bin/bash
# tests + bench + launch list in one gpurun call
bash tools/gpu_tests.sh
EGR_BENCH_CPU=${EGR_BENCH_CPU:-0} timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 1 1 1 > gpurun_out/profile_step.log 2>&1; echo "ncu exit $?"; tail -n 3 gpurun_out/profile_step.log
python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/launches.summary.txt 2>&1; cat gpurun_out/launches.summary.txt
