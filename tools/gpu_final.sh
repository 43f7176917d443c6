#!/bin/bash
# Driver-style validation + the round's final artefacts in one gpurun call.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "cpus: $(nproc)" >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 1 1 1 > gpurun_out/profile_step.log 2>&1; echo "ncu exit $?"
python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/launches.summary.txt 2>&1; head -n 12 gpurun_out/launches.summary.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/traffic.csv python tools/profile_step.py 1 1 1 > gpurun_out/traffic.log 2>&1; echo "traffic exit $?"
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
for idx in 0 19 391; do
  timeout 200 $NCU -k regex:gemm_tc_kernel -s $idx -c 1 -f -o gpurun_out/prof4_gemm_$idx python tools/profile_step.py 1 1 1 > /dev/null 2>&1; echo "gemm $idx exit $?"
done
timeout 200 $NCU -k regex:lowpass_kernel -c 1 -f -o gpurun_out/prof4_lowpass python tools/profile_step.py 1 1 1 > /dev/null 2>&1; echo "lowpass $?"
timeout 200 $NCU -k regex:fl_ -s 1 -c 2 -f -o gpurun_out/prof4_fl python tools/profile_fatllama.py 3 > /dev/null 2>&1; echo "fl $?"
