"""Seeded signals of tests/golden/make_eval_lufs_golden.py (shared by the LUFS tests)."""
import numpy as np


def signal(name, c):
    rng = np.random.default_rng(sum(map(ord, name)))
    C, N, sr, kind = c["C"], c["N"], c["sr"], c["kind"]
    t = np.arange(N) / sr
    if kind == "noise":
        x = 0.1 * rng.standard_normal((C, N))
    elif kind == "tone":
        x = 0.3 * np.sin(2 * np.pi * 50.0 * t)[None] + 0.2 * np.sin(2 * np.pi * 3000.0 * t)[None] + 0.01 * rng.standard_normal((C, N))
    else:
        env = np.where((t * 2).astype(int) % 2 == 0, 0.5, 0.004)[None]
        x = env * rng.standard_normal((C, N)) + 0.05 * np.sin(2 * np.pi * 30.0 * t)[None]
    return x.astype(np.float32)
