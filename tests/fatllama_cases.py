"""Shared input generators of the path-B tests and of tests/golden/make_fatllama_c4_golden.py."""
import numpy as np

C4 = dict(C=2, S=7938000, sr=44100, iters=300, thr=0.6, kbps=1411, seed=44)   # BASELINE.json configs[3]


def audio(C, S, seed, sr=16000):
    rng = np.random.default_rng(seed)
    t = np.arange(S) / sr
    x = 0.05 * rng.standard_normal((C, S))
    for h, a in ((220.0, 0.3), (440.0, 0.15), (1760.0, 0.05)):
        x += a * np.sin(2 * np.pi * h * t + rng.uniform(0, 6.28, (C, 1)))
    return np.clip(x, -1, 1).astype(np.float32)


def c4_sample_index(S, n=20000, edge=1000, seed=123):
    """Positions of the committed subsample: both edges plus a seeded random draw (sorted, unique)."""
    rng = np.random.default_rng(seed)
    idx = np.concatenate([np.arange(edge), np.arange(S - edge, S), rng.integers(0, S, n)])
    return np.unique(idx)
