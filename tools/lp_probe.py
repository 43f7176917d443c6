import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from harness import MiniPlan
from egregora_b200 import flashsr_model as M
spec = M.default_spec()
T = spec["chunk"]
g = torch.Generator().manual_seed(5)
wav = (0.1 * torch.randn(1, T, generator=g)).cumsum(1) * 0.05
wav = wav - wav.mean(1, keepdim=True); wav = wav / wav.abs().max() * 0.5
mp = MiniPlan({}, spec=spec)
wi = mp.input(wav[:, None, None, :])
lp = mp.be.lowpass(wi)
mp.run_gpu()
print("ok", float(mp.read(lp).abs().max()))
