#!/bin/bash
# Runs the GPU test groups under separate timeouts so one hung kernel cannot eat the whole gpurun call.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-400} python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n 6 gpurun_out/$name.log; }
run wola tests/test_wola_gpu.py
run ops_simt tests/test_ops_gpu.py -k "simt or groupnorm or layernorm or attention_small or snake or stft"
run ops_tc_basic tests/test_ops_gpu.py -k "conv2d_tc or conv1d_tc_persistent"
run ops_tc_rest tests/test_ops_gpu.py -k "stride2 or transposed or conv1d or conv_transpose or attention_gemm"
run flashsr_tiny tests/test_flashsr_gpu.py -k "tiny"
run fft tests/test_fft_gpu.py
run fatllama tests/test_fatllama_gpu.py
run resample tests/test_resample.py
run dfn_mix tests/test_dfn_mix.py
run eval tests/test_eval_metrics.py
