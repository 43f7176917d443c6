#!/bin/bash
# N-GPU check of bench.py (pads the tail rank's span block when 130 spans do not divide by N)
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2n_bench_n$N.json 2> gpurun_out/r2n_bench_n$N.err
echo "bench N=$N rc=$?"
python tools/bench_brief.py < gpurun_out/r2n_bench_n$N.json
python - <<PY
import json
d=json.loads(open("gpurun_out/r2n_bench_n$N.json").read().strip().splitlines()[-1])
print("c3", d["c3"]); print("c5", d["chain_c5"])
PY
tail -3 gpurun_out/r2n_bench_n$N.err
