"""Seeded signals of tests/golden/make_eval_hf_golden.py (shared by the HF-band tests)."""
import numpy as np


def signal(name, c):
    rng = np.random.default_rng(sum(map(ord, name)))
    C, N, sr = c["C"], c["N"], c["sr"]
    t = np.arange(N) / sr
    lo = 0.3 * np.sin(2 * np.pi * 440.0 * t)[None] + 0.05 * rng.standard_normal((C, N)).cumsum(axis=1) / 50.0
    hi = c["hf_gain"] * rng.standard_normal((C, N))
    return (lo + hi).astype(np.float32)
