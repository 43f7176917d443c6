#!/bin/bash
# Round 2 validation + artefacts: whole GPU suite, smoke, bench, launch lists, traffic pass, --set full captures, traces.
mkdir -p gpurun_out
rm -f gpurun_out/parity_full.json gpurun_out/parity_c4.json
timeout 1800 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 12 gpurun_out/r2f_pytest_gpu.log
cat gpurun_out/parity_full.json gpurun_out/parity_c4.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/r2f_smoke.log
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench exit $?"; cat gpurun_out/r2f_bench.json; tail -n 5 gpurun_out/r2f_bench.err
timeout 300 python tools/mega_trace.py 1 1 > gpurun_out/r2f_mega_trace_b1.txt 2>&1; grep -v Warning gpurun_out/r2f_mega_trace_b1.txt | head -n 16
timeout 300 python tools/mega_trace.py 8 1 > gpurun_out/r2f_mega_trace_b8.txt 2>&1; grep -v Warning gpurun_out/r2f_mega_trace_b8.txt | head -n 16
timeout 300 python tools/section_times.py 1 1 2>/dev/null | tail -7 | tee gpurun_out/r2f_sections_b1.txt
timeout 300 python tools/section_times.py 8 1 2>/dev/null | tail -7 | tee gpurun_out/r2f_sections_b8.txt
timeout 300 python tools/section_times.py 8 4 2>/dev/null | tail -7 | tee gpurun_out/r2f_sections_b8_s4.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_c2_b1.csv python tools/profile_step.py 1 1 1 > gpurun_out/r2f_ncu.log 2>&1; echo "ncu launches b1 exit $?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_c2_b8.csv python tools/profile_step.py 8 1 0 > gpurun_out/r2f_ncu8.log 2>&1; echo "ncu launches b8 exit $?"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r2f_traffic.csv python tools/profile_step.py 1 1 1 > gpurun_out/r2f_traffic.log 2>&1; echo "traffic exit $?"
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
for idx in 0 19 100 150; do
  timeout 300 $NCU -k regex:gemm_tc_kernel -s $idx -c 1 -f -o gpurun_out/r2f_prof_gemm_$idx python tools/profile_step.py 1 1 1 > gpurun_out/r2f_ncufull_$idx.log 2>&1; echo "gemm $idx exit $?"
done
timeout 300 $NCU -k regex:unet_mega_kernel -c 1 -f -o gpurun_out/r2f_prof_mega python tools/profile_step.py 1 1 1 > gpurun_out/r2f_ncufull_mega.log 2>&1; echo "mega full exit $?"
timeout 300 $NCU -k regex:snake_aa_kernel -s 80 -c 1 -f -o gpurun_out/r2f_prof_snake python tools/profile_step.py 1 1 1 > /dev/null 2>&1; echo "snake $?"
timeout 300 $NCU -k regex:fl_ -s 1 -c 2 -f -o gpurun_out/r2f_prof_fl python tools/profile_fatllama.py 3 > /dev/null 2>&1; echo "fl $?"
ls -la gpurun_out/*.ncu-rep | head
