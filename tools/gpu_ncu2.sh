#!/bin/bash
mkdir -p gpurun_out
# DRAM traffic + duration of every launch of one c2 pass (cheap metrics, all kernels)
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/traffic.csv python tools/profile_step.py 1 1 1 > gpurun_out/traffic.log 2>&1; echo "traffic exit $?"
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
# representative GEMM launches (plan order among gemm_tc launches): 0 = vae hi-res 128ch, 19 = vae mid 1024ch, 311 = vae dec, 391/429 = vocoder
for idx in 0 19 100 391 429; do
  timeout 300 $NCU -k regex:gemm_tc_kernel -s $idx -c 1 -f -o gpurun_out/prof2_gemm_$idx python tools/profile_step.py 1 1 1 > gpurun_out/ncu2_gemm_$idx.log 2>&1; echo "gemm $idx exit $?"
done
timeout 300 $NCU -k regex:snake_aa_kernel -s 80 -c 1 -f -o gpurun_out/prof2_snake python tools/profile_step.py 1 1 1 > /dev/null 2>&1; echo "snake $?"
timeout 300 $NCU -k regex:gn_ -s 2 -c 3 -f -o gpurun_out/prof2_gn python tools/profile_step.py 1 1 1 > /dev/null 2>&1; echo "gn $?"
timeout 300 $NCU -k regex:fl_ -s 1 -c 2 -f -o gpurun_out/prof2_fl python tools/profile_fatllama.py 3 > /dev/null 2>&1; echo "fl $?"
