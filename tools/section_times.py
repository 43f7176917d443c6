"""Real (warm, back-to-back) device time of contiguous plan sections, by op-name prefix."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import os
os.environ.setdefault("EGREGORA_FLASHSR_RANDOM_INIT", "1")   # no checkpoint in this environment
import bench
bench.load_pkg()
from egregora_b200 import _abi, egregora_audio_super_resolution as N
dev = torch.device("cuda", 0)
eng = N.get_engine(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
x = bench.synth_audio(N.CHUNK_SAMPLES, B).to(dev)
eng.infer(x, lowpass=True, steps=steps)
be, h = eng.plan(B, steps, True)
names = [o.name for o in be.ops]
def sec(n):
    if n.startswith("vae.encoder") or n.startswith("vae.quant"): return "vae.encoder"
    if n.startswith("unet") or n in ("t_emb", "axpby", "geglu") or n.startswith("attn.small"): return "unet"
    if n.startswith("vae.decoder") or n.startswith("vae.post"): return "vae.decoder"
    if n.startswith("vocoder"): return "vocoder"
    if n.startswith("lp.") or n in ("lowpass", "stft_mel"): return "frontend"
    return None
# assign sections by forward fill
labels = []; cur = "frontend"
for n in names:
    s = sec(n)
    if s: cur = s
    labels.append(cur)
bounds = []
start = 0
for i in range(1, len(names) + 1):
    if i == len(names) or labels[i] != labels[start]:
        bounds.append((labels[start], start, i)); start = i
es = eng.stream
st = es.cuda_stream
lib = eng.lib
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot = {}
for lab, a, b in bounds:
    for _ in range(2): _abi.check(lib.egr_plan_run(h, a, b, st))
    torch.cuda.synchronize()
    reps = 10
    e0.record(es)
    for _ in range(reps): _abi.check(lib.egr_plan_run(h, a, b, st))
    e1.record(es); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tot[lab] = tot.get(lab, [0, 0.0]); tot[lab][0] += b - a; tot[lab][1] += ms
for k, v in tot.items(): print(f"{k:14s} ops={v[0]:4d} ms={v[1]:7.3f}")
e0.record(es)
for _ in range(10): _abi.check(lib.egr_plan_run(h, 0, -1, st))
e1.record(es); torch.cuda.synchronize()
print("whole plan ms", e0.elapsed_time(e1) / 10, "ops", len(names))
