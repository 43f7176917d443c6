#!/bin/bash
# last call of the round: default path (EGR_PDL unset) on the final library — smoke + op / persistent-kernel tests
mkdir -p gpurun_out
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2rr_smoke.log 2>&1; echo "smoke exit $?"; tail -n 1 gpurun_out/r2rr_smoke.log
timeout 120 python -m pytest tests/test_ops_gpu.py tests/test_mega_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2rr_ops.log 2>&1; echo "ops+mega exit $?"; tail -n 2 gpurun_out/r2rr_ops.log
