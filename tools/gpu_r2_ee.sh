#!/bin/bash
# round 2, session 2: packed-f32x2 snake, GroupNorm stats/apply with 8 loads in flight, fast SiLU
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_mega_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2ee_ops.log 2>&1; rc=$?; echo "ops+mega exit $rc"; tail -n 3 gpurun_out/r2ee_ops.log
[ $rc -ne 0 ] && exit 1
timeout 300 python tools/op_times.py 1 > gpurun_out/r2ee_ops_b1.tsv 2>/dev/null
timeout 300 python tools/op_times.py 8 > gpurun_out/r2ee_ops_b8.tsv 2>/dev/null
for f in gpurun_out/r2ee_ops_b1.tsv profiles/r2_g_ops_b1.tsv gpurun_out/r2dd_ops_b1.tsv gpurun_out/r2ee_ops_b8.tsv profiles/r2_g_ops_b8.tsv; do
  echo "$f: GN us $(grep -i 'norm' $f | awk -F'\t' '{s+=$8} END {print s}')  snake us $(grep -i '\.act\|snake\|activation' $f | awk -F'\t' '{s+=$8} END {print s}') total us $(awk -F'\t' '{s+=$8} END {print s}' $f)"
done
grep "vae.decoder.up.0.block.0.norm1\|vae.decoder.up.0.block.1.norm1" gpurun_out/r2ee_ops_b1.tsv
timeout 300 python tools/section_times.py 1 1 2>/dev/null | tail -6
timeout 300 python tools/section_times.py 8 1 2>/dev/null | tail -6
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
timeout 300 $NCU -k regex:snake_aa_kernel -s 80 -c 1 -f -o gpurun_out/r2ee_prof_snake python tools/profile_step.py 1 1 1 > gpurun_out/r2ee_ncu_snake.log 2>&1; echo "snake exit $?"
timeout 300 $NCU -k regex:gn_apply -s 50 -c 1 -f -o gpurun_out/r2ee_prof_gn_apply python tools/profile_step.py 1 1 1 > gpurun_out/r2ee_ncu_gna.log 2>&1; echo "gn_apply exit $?"
timeout 300 $NCU -k regex:gn_stats -s 50 -c 1 -f -o gpurun_out/r2ee_prof_gn_stats python tools/profile_step.py 1 1 1 > gpurun_out/r2ee_ncu_gns.log 2>&1; echo "gn_stats exit $?"
timeout 900 python -m pytest tests/test_flashsr_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2ee_e2e.log 2>&1; echo "e2e exit $?"; tail -n 3 gpurun_out/r2ee_e2e.log
