#!/bin/bash
mkdir -p gpurun_out
echo "=== default"; timeout 600 python tools/gemm_probe.py 2>&1 | tee gpurun_out/probe_default.txt
echo "=== NO_MT2"; EGR_TC_NO_MT2=1 timeout 600 python tools/gemm_probe.py 2>&1 | tee gpurun_out/probe_nomt2.txt
echo "=== NO_HALO"; EGR_TC_NO_HALO=1 timeout 600 python tools/gemm_probe.py conv1d 2>&1 | tee gpurun_out/probe_nohalo.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:gemm_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_small python tools/gemm_probe.py "640->640 k1" > gpurun_out/ncu_small.log 2>&1; echo "ncu small $?"
timeout 300 $NCU -k regex:gemm_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_big python tools/gemm_probe.py "128->128 k3 d1 (512" > gpurun_out/ncu_big.log 2>&1; echo "ncu big $?"
