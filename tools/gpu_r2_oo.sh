#!/bin/bash
# final check of the default sub-batch of 48: full-size FlashSR tests (incl. the 18-rows-in-one-launch identity) + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_flashsr_gpu.py tests/test_checkpoint_gpu.py tests/test_zz_chain_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2oo_e2e.log 2>&1; echo "e2e exit $?"; tail -n 3 gpurun_out/r2oo_e2e.log
timeout 900 python bench.py > gpurun_out/r2oo_bench.json 2> gpurun_out/r2oo_bench.err; echo "bench exit $?"
python3 - <<'PY'
import json
d=json.loads(open('gpurun_out/r2oo_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print('c3',d['c3']['value'],d['c3']['seconds'],d['c3']['phases_max_over_ranks'])
print('c5',d['chain_c5']['value'],'batched',d['batched'])
print('frac',d['roofline']['frac'],'b8',d['roofline']['batch8']['frac'])
PY
tail -n 3 gpurun_out/r2oo_bench.err
