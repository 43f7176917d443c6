"""One profiled pass of the FlashSR node path for ncu (run with --profile-from-start off):
warm-up passes, then cudaProfilerStart .. one pass .. cudaProfilerStop.  Usage:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python tools/profile_step.py [batch] [steps] [lowpass]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import os  # noqa: E402
os.environ.setdefault("EGREGORA_FLASHSR_RANDOM_INIT", "1")   # no checkpoint in this environment
import bench  # noqa: E402

bench.load_pkg()
from egregora_b200 import egregora_audio_super_resolution as N  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lowpass = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
dev = torch.device("cuda", 0)
engine = N.get_engine(dev)
win, hop = N._win_hop()
x = bench.synth_audio(win, batch).to(dev)
model = lambda c, row0=0: engine.infer(c, lowpass=lowpass, steps=steps, seed=4321, row0=row0)  # noqa: E731
for _ in range(2):
    N.upscale_48k(x, model)
torch.cuda.synchronize()
torch.cuda.profiler.start()
y = N.upscale_48k(x, model)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", tuple(y.shape), float(y.abs().max()))
