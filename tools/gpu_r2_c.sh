#!/bin/bash
# folded split-K + variance-preserving synthetic init + fused qkv/emb default: parity at size, bit-identity, bench
mkdir -p gpurun_out
rm -f gpurun_out/parity_full.json
timeout 1500 python -m pytest tests/test_flashsr_gpu.py tests/test_checkpoint_gpu.py tests/test_fatllama_gpu.py tests/test_zz_fused_qkv_gpu.py tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2c_tests.log 2>&1; echo "tests exit $?"; tail -n 30 gpurun_out/r2c_tests.log
cat gpurun_out/parity_full.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 600 python tools/parity_diag.py 1 > gpurun_out/r2c_parity_diag.log 2>&1; echo "diag exit $?"; head -n 3 gpurun_out/r2c_parity_diag.log | tail -n 1; tail -n 12 gpurun_out/r2c_parity_diag.log
EGR_BENCH_VERBOSE=1 timeout 900 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench exit $?"; cat gpurun_out/r2c_bench.json; tail -n 12 gpurun_out/r2c_bench.err
timeout 300 python tools/section_times.py 1 1 2>/dev/null | tail -7
timeout 300 python tools/section_times.py 8 1 2>/dev/null | tail -7
EGR_TC_NO_FOLD=1 timeout 300 python tools/section_times.py 8 1 2>/dev/null | tail -7
