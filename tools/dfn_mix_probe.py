"""Device time of egr_dfn_mix at clip scale (c5: 5 min stereo at 48 kHz) with CUDA events on the launching stream,
against the HBM roofline (algorithmic bytes = 24*T per channel: rms reads 4T, mix reads 8T + writes 4T, limiter
reads + writes 8T).   python tools/dfn_mix_probe.py"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_pkg  # noqa: E402

load_pkg()
from egregora_b200 import egregora_dfn_mix as M  # noqa: E402

dev = torch.device("cuda", 0)
for secs, C in [(300, 2), (600, 2), (5.12, 1)]:
    T = int(48000 * secs)
    dry = torch.randn((C, T), device=dev) * 0.3
    wet = dry * 0.8
    for _ in range(3):
        y = M.adaptive_mix(dry, wet, 48000)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        y = M.adaptive_mix(dry, wet, 48000)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gb = 24.0 * C * T / 1e9
    print(f"{secs}s x{C}: {ms:8.3f} ms  {gb / ms * 1e3:8.1f} GB/s algorithmic ({gb / ms * 1e3 / 6532.9:.2%} of HBM peak)")
