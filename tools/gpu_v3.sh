#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-300} python -m pytest "$@" -q -m gpu -p no:cacheprovider -x > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n 12 gpurun_out/$name.log; }
run ops_tc_basic tests/test_ops_gpu.py -k "conv2d_tc or conv1d_tc_persistent"
run ops_tc_rest tests/test_ops_gpu.py -k "stride2 or transposed or conv1d or conv_transpose or attention_gemm"
run flashsr_tiny tests/test_flashsr_gpu.py -k "tiny"
timeout 600 python tools/gemm_probe.py 2>&1 | tee gpurun_out/probe_v3.txt
timeout 120 python tools/gemm_trace.py "640->640 k3" 2>&1 | cut -c1-1500
timeout 120 python tools/gemm_trace.py "128->128 k3 d1 (512" 2>&1 | cut -c1-1500
