"""Device time of egr_eval_null_test and egr_eval_lsd at clip scale (5 / 10 min stereo at 48 kHz), CUDA events on the launching stream,
against the HBM roofline (null test: algorithmic bytes = 20*C*N — pass 1 reads 8CN, pass 2 reads 8CN and writes 4CN;
LSD: 8*C*N, both clips read once, the 4x frame overlap is served by L2).
    python tools/eval_probe.py"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_pkg  # noqa: E402

load_pkg()
from egregora_b200 import egregora_eval_metrics as M  # noqa: E402

dev = torch.device("cuda", 0)
for secs, C in [(300, 2), (600, 2)]:
    N = int(48000 * secs)
    a = torch.randn((C, N), device=dev) * 0.2
    b = a * 0.9 + 0.01 * torch.randn((C, N), device=dev)
    for _ in range(3):
        M.null_test(a, b, least_squares_scale=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        M.null_test(a, b, least_squares_scale=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gb = 20.0 * C * N / 1e9
    print(f"{secs}s x{C}: {ms:8.3f} ms  {gb / ms * 1e3:8.1f} GB/s algorithmic ({gb / ms * 1e3 / 6532.9:.2%} of HBM peak)")
    for _ in range(3):
        M.lsd(a, b)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        M.lsd(a, b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gb = 8.0 * C * N / 1e9
    print(f"{secs}s x{C} lsd: {ms:8.3f} ms  {gb / ms * 1e3:8.1f} GB/s algorithmic ({gb / ms * 1e3 / 6532.9:.2%} of HBM peak)")
