"""Warm device time of the vocoder's anti-aliased SnakeBeta ops: the one-channel-per-thread kernel (EGR_SNAKE_SCALAR=1)
against the two-channel packed-f32x2 kernel.
    python tools/snake_sweep.py [batch]"""
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
idx = [i for i, o in enumerate(be.ops) if "activation" in o.name]
def total():
    t = 0.0
    for i in idx:
        for _ in range(2): _abi.check(lib.egr_plan_run(h, i, i + 1, st))
        torch.cuda.synchronize()
        e0.record(es)
        for _ in range(10): _abi.check(lib.egr_plan_run(h, i, i + 1, st))
        e1.record(es); torch.cuda.synchronize()
        t += 1e3 * e0.elapsed_time(e1) / 10
    return t
os.environ["EGR_SNAKE_SCALAR"] = "1"
print(f"batch {B}: {len(idx)} ops; scalar kernel {total():.1f} us")
del os.environ["EGR_SNAKE_SCALAR"]
print(f"batch {B}: packed kernel {total():.1f} us")
