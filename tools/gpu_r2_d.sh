#!/bin/bash
# first hardware run of the persistent UNet kernel: every step under its own timeout (a hang must not eat the box)
mkdir -p gpurun_out
rm -f gpurun_out/parity_full.json
timeout 600 python -m pytest tests/test_mega_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/r2d_mega.log 2>&1; echo "mega tests exit $?"; tail -n 30 gpurun_out/r2d_mega.log
timeout 900 python -m pytest tests/test_flashsr_gpu.py tests/test_ops_gpu.py tests/test_zz_fused_qkv_gpu.py tests/test_zz_chain_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2d_tests.log 2>&1; echo "tests exit $?"; tail -n 15 gpurun_out/r2d_tests.log
cat gpurun_out/parity_full.json
timeout 300 python tools/section_times.py 1 1 2>/dev/null | tail -7
timeout 300 python tools/section_times.py 8 1 2>/dev/null | tail -7
EGR_NO_MEGA=1 timeout 300 python tools/section_times.py 1 1 2>/dev/null | tail -7
EGR_BENCH_CPU=0 EGR_BENCH_PATHB=0 EGR_BENCH_C5=0 EGR_BENCH_EAGER=0 timeout 900 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench exit $?"; cat gpurun_out/r2d_bench.json; tail -n 5 gpurun_out/r2d_bench.err
