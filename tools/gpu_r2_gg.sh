#!/bin/bash
# round 2, session 2, third call: whole GPU suite + smoke + bench on the new CUDA-core kernels
P=r2gg
mkdir -p gpurun_out
rm -f gpurun_out/parity_full.json gpurun_out/parity_c4.json
timeout 300 python tools/snake_sweep.py 1 2>/dev/null
timeout 300 python tools/snake_sweep.py 8 2>/dev/null
timeout 1800 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/${P}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 6 gpurun_out/${P}_pytest_gpu.log
cat gpurun_out/parity_full.json gpurun_out/parity_c4.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/${P}_smoke.log
timeout 900 python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; echo "bench exit $?"; cat gpurun_out/${P}_bench.json; tail -n 5 gpurun_out/${P}_bench.err
