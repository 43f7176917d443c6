#!/bin/bash
# round 2, session 2, seventh call: per-block epilogue stamps of the 48-channel vocoder convs (trace build)
mkdir -p gpurun_out
export EGREGORA_B200_LIB=$PWD/comfyui-egregora-audio-super-resolution_b200/libegregora_b200_trace.so
timeout 300 python tools/gemm_trace.py "conv1d 48->48 k3" 1 2>&1 | tail -8 | cut -c1-1200
timeout 300 python tools/gemm_trace.py "conv1d 48->48 k3" 8 2>&1 | tail -8 | cut -c1-1200
timeout 300 python tools/gemm_trace.py "conv1d 96->96 k3" 8 2>&1 | tail -8 | cut -c1-1200
timeout 300 python tools/gemm_trace.py "conv2d 128->128 k3 d1 (512, 256)" 1 2>&1 | tail -8 | cut -c1-1200
