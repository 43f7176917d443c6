#!/bin/bash
mkdir -p gpurun_out
P=$PWD/comfyui-egregora-audio-super-resolution_b200
EGREGORA_B200_LIB=$P/libegregora_b200_old.so timeout 300 python tools/op_times.py 1 > gpurun_out/r2s_ops_old.tsv 2>/dev/null
timeout 300 python tools/op_times.py 1 > gpurun_out/r2s_ops_new.tsv 2>/dev/null
EGR_TC_GMAX=2 timeout 300 python tools/op_times.py 1 > gpurun_out/r2s_ops_new_g2.tsv 2>/dev/null
wc -l gpurun_out/r2s_ops_*.tsv
