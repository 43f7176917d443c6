"""Device time of egr_resample_poly at clip scale (CUDA events on the launching stream), against the HBM
roofline (algorithmic bytes = 4*(n_in + n_out) per channel).   python tools/resample_probe.py"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_pkg  # noqa: E402

load_pkg()
from egregora_b200 import egregora_audio_super_resolution as N  # noqa: E402

dev = torch.device("cuda", 0)
for (src, dst, secs, C) in [(44100, 48000, 600, 2), (48000, 44100, 600, 2), (16000, 48000, 600, 2), (48000, 96000, 600, 2),
                            (44100, 48000, 5.12, 1)]:
    n = int(src * secs)
    x = torch.randn((C, n), device=dev) * 0.1
    for _ in range(3):
        y = N._resample_hq(x, src, dst)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        y = N._resample_hq(x, src, dst)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gb = 4.0 * C * (n + y.shape[1]) / 1e9
    print(f"{src}->{dst} {secs}s x{C}: {ms:8.3f} ms  {gb / ms * 1e3:8.1f} GB/s algorithmic ({gb / ms * 1e3 / 6532.9:.2%} of HBM peak)")
