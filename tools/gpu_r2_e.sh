#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mega_gpu.py tests/test_flashsr_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/r2e_tests.log 2>&1; echo "tests exit $?"; tail -n 6 gpurun_out/r2e_tests.log
timeout 300 python tools/mega_trace.py 1 1 > gpurun_out/r2e_trace_b1.txt 2>&1; cat gpurun_out/r2e_trace_b1.txt | grep -v Warning | tail -n 45
timeout 300 python tools/mega_trace.py 8 1 > gpurun_out/r2e_trace_b8.txt 2>&1; cat gpurun_out/r2e_trace_b8.txt | grep -v Warning | tail -n 45
timeout 300 python tools/section_times.py 1 1 2>/dev/null | tail -7
timeout 300 python tools/section_times.py 8 1 2>/dev/null | tail -7
