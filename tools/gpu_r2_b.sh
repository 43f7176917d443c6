#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_flashsr_gpu.py tests/test_noise.py tests/test_checkpoint_gpu.py tests/test_fatllama_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2b_new_tests.log 2>&1; echo "new tests exit $?"; tail -n 40 gpurun_out/r2b_new_tests.log
cat gpurun_out/parity_full.json
timeout 600 python tools/parity_diag.py 1 > gpurun_out/r2b_parity_diag.log 2>&1; echo "diag exit $?"; tail -n 75 gpurun_out/r2b_parity_diag.log
EGR_BENCH_VERBOSE=1 timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench exit $?"; cat gpurun_out/r2b_bench.json; tail -n 40 gpurun_out/r2b_bench.err
