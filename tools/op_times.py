"""Warm device time of every op of the FlashSR plan, one op at a time (egr_plan_run(h, i, i+1), back to back).
    python tools/op_times.py [batch] [name filter] > ops.tsv        (EGREGORA_B200_LIB picks the library build)"""
import os, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ.setdefault("EGREGORA_FLASHSR_RANDOM_INIT", "1")
import bench
bench.load_pkg()
from egregora_b200 import _abi, egregora_audio_super_resolution as N
dev = torch.device("cuda", 0)
eng = N.get_engine(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
x = bench.synth_audio(N.CHUNK_SAMPLES, B).to(dev)
eng.infer(x, lowpass=True, steps=1)
be, h = eng.plan(B, 1, True)
es = eng.stream; st = es.cuda_stream; lib = eng.lib
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
lt = {l["name"]: l for l in be.layer_table}
mega = [i for i, o in enumerate(be.ops) if o.flags & 1]
lo, hi = (mega[0], mega[-1] + 1) if mega else (0, 0)
flt = sys.argv[2] if len(sys.argv) > 2 else ""   # optional substring filter on the op name
for i, o in enumerate(be.ops):
    if lo <= i < hi or (flt and flt not in o.name):
        continue
    for _ in range(2): _abi.check(lib.egr_plan_run(h, i, i + 1, st))
    torch.cuda.synchronize()
    reps = 10
    e0.record(es)
    for _ in range(reps): _abi.check(lib.egr_plan_run(h, i, i + 1, st))
    e1.record(es); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps
    l = lt.get(o.name, {})
    print(f"{i}\t{o.name}\t{l.get('kind', '-')}\t{l.get('M', 0)}\t{l.get('N', 0)}\t{l.get('K', 0)}\t{l.get('taps', 0)}\t{us:.2f}")
