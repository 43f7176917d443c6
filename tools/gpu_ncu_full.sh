#!/bin/bash
# ncu --set full captures of the top kernels (one GPU, short commands).  Reports land in gpurun_out/.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
# tcgen05 tap-GEMM: launch indices (among gemm_tc launches of one c2 pass) of representative shapes
for idx in 0 19 311 391 429; do
  timeout 300 $NCU -k regex:gemm_tc_kernel -s $idx -c 1 -f -o gpurun_out/prof_gemm_$idx python tools/profile_step.py 1 1 1 > gpurun_out/ncu_gemm_$idx.log 2>&1
  echo "gemm $idx exit $?"
done
timeout 300 $NCU -k regex:snake_aa_kernel -s 80 -c 1 -f -o gpurun_out/prof_snake python tools/profile_step.py 1 1 1 > gpurun_out/ncu_snake.log 2>&1; echo "snake exit $?"
timeout 300 $NCU -k regex:gn_ -s 4 -c 2 -f -o gpurun_out/prof_gn python tools/profile_step.py 1 1 1 > gpurun_out/ncu_gn.log 2>&1; echo "gn exit $?"
timeout 300 $NCU -k regex:fl_ -s 1 -c 2 -f -o gpurun_out/prof_fl python tools/profile_fatllama.py 3 > gpurun_out/ncu_fl.log 2>&1; echo "fl exit $?"
ls -la gpurun_out/*.ncu-rep
