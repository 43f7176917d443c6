#!/usr/bin/env python
"""Generates tests/golden/fatllama_c4_golden.npz: BASELINE config c4 (Fat-Llama 3 min 44.1 kHz stereo, 300 iterations,
threshold 0.6, normalize + autoscale on) evaluated ONCE at full size by the float64 oracle (oracle/fat_llama_oracle.py,
~10 minutes of host time, channels in parallel), reduced to a committed fixture the GPU test can check at the
north_star tolerance (1e-5 per sample, pre-quantisation):

  idx        sorted sample positions (both edges + 20 000 seeded random positions)
  pre[C,n]   float64 oracle output at idx, before the final PCM-16 step
  blk[C,64]  float64 sums of the oracle output over 64 equal blocks      (a checksum of the whole clip)
  blk2[C,64] float64 sums of squares over the same blocks
  in_sha     sha256 of the float32 input, so the test knows it regenerated the same clip

Run from the repo root:  python tests/golden/make_fatllama_c4_golden.py
"""
import hashlib
import sys
from concurrent.futures import ProcessPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from fatllama_cases import C4, audio, c4_sample_index  # noqa: E402
from oracle import fat_llama_oracle as O  # noqa: E402


def _channel(args):
    ch, iters, thr = args
    expanded = ch.astype(np.float64)
    return expanded + O.ist(expanded, iters, thr, np.float64)


def main():
    c = C4
    x = audio(c["C"], c["S"], c["seed"], c["sr"])
    samples = O.pcm16_write(x.T).astype(np.float64)                 # [S,C] integer-scaled (upstream read_audio)
    assert O.upscale_factor(c["sr"], c["C"], c["kbps"]) == 1
    with ProcessPoolExecutor(c["C"]) as ex:
        cols = list(ex.map(_channel, [(samples[:, k].copy(), c["iters"], c["thr"]) for k in range(c["C"])]))
    y = np.stack(cols, 1)
    # autoscale + normalize + patched write scale: exactly O.upscale / O.node_run after upscale_channels
    for k in range(c["C"]):
        y[:, k] = (y[:, k] / np.max(np.abs(y[:, k]))) * np.max(np.abs(samples[:, k]))
    y = y / np.max(np.abs(y))
    if float(np.max(np.abs(y))) > 1.0:
        y = y / 32768.0
    y = y.T                                                           # [C,S]
    idx = c4_sample_index(c["S"])
    nb = 64
    edges = np.linspace(0, c["S"], nb + 1).astype(np.int64)
    blk = np.stack([[y[k, edges[b]:edges[b + 1]].sum() for b in range(nb)] for k in range(c["C"])])
    blk2 = np.stack([[(y[k, edges[b]:edges[b + 1]] ** 2).sum() for b in range(nb)] for k in range(c["C"])])
    out = ROOT / "tests" / "golden" / "fatllama_c4_golden.npz"
    np.savez_compressed(out, idx=idx, pre=y[:, idx], blk=blk, blk2=blk2, edges=edges,
                        in_sha=np.frombuffer(hashlib.sha256(x.tobytes()).digest(), np.uint8))
    print("wrote", out, out.stat().st_size, "bytes; max|y| =", float(np.max(np.abs(y))))


if __name__ == "__main__":
    main()
