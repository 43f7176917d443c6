#!/bin/bash
# Round 2, first GPU call: the new parity-at-size tests first (their numbers land in gpurun_out/parity_full.json), then the
# whole suite, smoke, the default bench line, the A/B of the split-K policy and of the fused QKV / embedding plans, and
# the launch list of one c2 pass.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests/test_flashsr_gpu.py tests/test_noise.py tests/test_checkpoint_gpu.py tests/test_fatllama_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2a_new_tests.log 2>&1; echo "new tests exit $?"; tail -n 25 gpurun_out/r2a_new_tests.log
cat gpurun_out/parity_full.json
timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider --deselect tests/test_flashsr_gpu.py --deselect tests/test_checkpoint_gpu.py --deselect tests/test_fatllama_gpu.py --deselect tests/test_noise.py > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rest exit $?"; tail -n 8 gpurun_out/r2a_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/r2a_smoke.log
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench exit $?"; cat gpurun_out/r2a_bench.json; tail -n 5 gpurun_out/r2a_bench.err
LEAN="EGR_BENCH_CPU=0 EGR_BENCH_PATHB=0 EGR_BENCH_C3=0 EGR_BENCH_C5=0 EGR_BENCH_EAGER=0"
env $LEAN EGR_TC_SPLIT_POLICY=launch timeout 600 python bench.py > gpurun_out/r2a_bench_split_launch.json 2> gpurun_out/r2a_bench_split_launch.err; echo "bench (split by launch) exit $?"; cat gpurun_out/r2a_bench_split_launch.json
env $LEAN EGR_FUSE_QKV=1 EGR_FUSE_EMB=1 timeout 600 python bench.py > gpurun_out/r2a_bench_fused.json 2> gpurun_out/r2a_bench_fused.err; echo "bench (fused qkv+emb) exit $?"; cat gpurun_out/r2a_bench_fused.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_launches_c2_b1.csv python tools/profile_step.py 1 1 1 > gpurun_out/r2a_ncu.log 2>&1; echo "ncu b1 exit $?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_launches_c2_b8.csv python tools/profile_step.py 8 1 0 > gpurun_out/r2a_ncu8.log 2>&1; echo "ncu b8 exit $?"
timeout 300 python tools/section_times.py 1 1 > gpurun_out/r2a_sections_b1.txt 2>&1; cat gpurun_out/r2a_sections_b1.txt | tail -8
timeout 300 python tools/section_times.py 8 1 > gpurun_out/r2a_sections_b8.txt 2>&1; cat gpurun_out/r2a_sections_b8.txt | tail -8
