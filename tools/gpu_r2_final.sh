#!/bin/bash
# Round 2 validation + artefacts (final): whole GPU suite, smoke, bench, launch lists, traffic pass, --set full captures, traces.
P=${1:-r2g}
mkdir -p gpurun_out
rm -f gpurun_out/parity_full.json gpurun_out/parity_c4.json
timeout 1800 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/${P}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 12 gpurun_out/${P}_pytest_gpu.log
cat gpurun_out/parity_full.json gpurun_out/parity_c4.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/${P}_smoke.log
timeout 900 python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; echo "bench exit $?"; cat gpurun_out/${P}_bench.json; tail -n 5 gpurun_out/${P}_bench.err
timeout 300 python tools/mega_trace.py 1 1 > gpurun_out/${P}_mega_trace_b1.txt 2>&1
timeout 300 python tools/mega_trace.py 8 1 > gpurun_out/${P}_mega_trace_b8.txt 2>&1
(timeout 300 python tools/section_times.py 1 1 2>/dev/null | tail -7; timeout 300 python tools/section_times.py 8 1 2>/dev/null | tail -7; timeout 300 python tools/section_times.py 16 4 2>/dev/null | tail -7) | tee gpurun_out/${P}_sections.txt
timeout 300 python tools/op_times.py 1 > gpurun_out/${P}_ops_b1.tsv 2>/dev/null
timeout 300 python tools/op_times.py 8 > gpurun_out/${P}_ops_b8.tsv 2>/dev/null
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${P}_launches_c2_b1.csv python tools/profile_step.py 1 1 1 > gpurun_out/${P}_ncu.log 2>&1; echo "ncu launches b1 exit $?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${P}_launches_c2_b8.csv python tools/profile_step.py 8 1 0 > gpurun_out/${P}_ncu8.log 2>&1; echo "ncu launches b8 exit $?"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/${P}_traffic.csv python tools/profile_step.py 1 1 1 > gpurun_out/${P}_traffic.log 2>&1; echo "traffic exit $?"
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
for idx in 0 60 120; do
  timeout 300 $NCU -k regex:gemm_tc_kernel -s $idx -c 1 -f -o gpurun_out/${P}_prof_gemm_$idx python tools/profile_step.py 1 1 1 > gpurun_out/${P}_ncufull_$idx.log 2>&1; echo "gemm $idx exit $?"
done
for idx in 0 10; do
  timeout 300 $NCU -k regex:gemm_tc_pair_kernel -s $idx -c 1 -f -o gpurun_out/${P}_prof_pair_$idx python tools/profile_step.py 1 1 1 > gpurun_out/${P}_ncufull_pair_$idx.log 2>&1; echo "pair $idx exit $?"
done
timeout 300 $NCU -k regex:gemm_tc_pair_kernel -s 20 -c 1 -f -o gpurun_out/${P}_prof_pair_b8 python tools/profile_step.py 8 1 0 > /dev/null 2>&1; echo "pair b8 $?"
timeout 300 $NCU -k regex:unet_mega_kernel -c 1 -f -o gpurun_out/${P}_prof_mega python tools/profile_step.py 1 1 1 > gpurun_out/${P}_ncufull_mega.log 2>&1; echo "mega full exit $?"
timeout 300 $NCU -k regex:snake_aa_kernel -s 80 -c 1 -f -o gpurun_out/${P}_prof_snake python tools/profile_step.py 1 1 1 > /dev/null 2>&1; echo "snake $?"
ls -la gpurun_out/${P}_*.ncu-rep | head -20
