#!/bin/bash
# First gpurun call of the next round: run what was written after round 1's GPU budget ran out (egr_eval_lsd, egr_eval_lufs, the example-graph chain), then the
# driver-style validation.  `--runxfail` turns the xfail marks off so a failure shows its traceback.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_eval_lsd_gpu.py tests/test_zz_eval_lufs_gpu.py tests/test_zz_eval_hf_gpu.py tests/test_zz_null_full_gpu.py tests/test_zz_fused_qkv_gpu.py tests/test_zz_chain_gpu.py -q -m gpu --runxfail -p no:cacheprovider > gpurun_out/lsd_gpu.log 2>&1; echo "lsd exit $?"; tail -n 15 gpurun_out/lsd_gpu.log
timeout 300 python tools/eval_probe.py > gpurun_out/eval_probe.log 2>&1; echo "eval probe exit $?"; tail -n 5 gpurun_out/eval_probe.log
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
# opt-in plan variants, to be compared with the default bench line above
EGR_BENCH_CPU=0 EGR_BENCH_PATHB=0 EGR_FUSE_QKV=1 EGR_FUSE_EMB=1 timeout 600 python bench.py > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; echo "bench (EGR_FUSE_QKV=1 EGR_FUSE_EMB=1) exit $?"; cat gpurun_out/bench_fused.json
