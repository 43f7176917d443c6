#!/bin/bash
# programmatic dependent launch (EGR_PDL=1): smoke, op + persistent-kernel tests, full-size FlashSR tests, section times on / off
mkdir -p gpurun_out
export EGR_PDL=1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2qq_smoke.log 2>&1; rc=$?; echo "smoke(PDL) exit $rc"; tail -n 1 gpurun_out/r2qq_smoke.log
[ $rc -ne 0 ] && exit 1
timeout 200 python tools/section_times.py 1 1 2>/dev/null | tail -6
EGR_PDL=0 timeout 200 python tools/section_times.py 1 1 2>/dev/null | tail -6
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_mega_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2qq_ops.log 2>&1; echo "ops+mega(PDL) exit $?"; tail -n 2 gpurun_out/r2qq_ops.log
timeout 400 python -m pytest tests/test_flashsr_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2qq_e2e.log 2>&1; echo "flashsr(PDL) exit $?"; tail -n 2 gpurun_out/r2qq_e2e.log
timeout 200 python tools/section_times.py 8 1 2>/dev/null | tail -6
