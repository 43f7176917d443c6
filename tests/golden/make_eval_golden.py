"""Generates tests/golden/eval_golden.json by importing the REFERENCE modules (/root/reference/
egregora_null_test_suite.py and egregora_audio_eval_pack.py) and calling Audio_Null_Test.execute (LUFS / LSD off) and
_si_sdr on seeded inputs; stores the metrics and a sha256 + probes of the null signal.  Run here only.
    python tests/golden/make_eval_golden.py
"""
import hashlib
import importlib.util
import json
import sys
from pathlib import Path

import numpy as np
import torch

OUT = Path(__file__).resolve().parent

CASES = {  # name: (C, N_ref, N_proc, invert_b, least_squares_scale, gain, noise)
    "stereo_ls": (2, 48000, 48000, True, True, 0.9, 0.01),
    "mono_plain": (1, 30011, 30011, True, False, 1.0, 0.002),
    "noinvert": (2, 20000, 19000, False, False, -0.7, 0.05),
    "hot": (2, 9999, 12000, True, True, 3.0, 0.6),
    "long": (2, 48000 * 20, 48000 * 20, True, True, 0.8, 0.003),
}


def signals(name, C, Na, Nb, gain, noise):
    rng = np.random.default_rng(sum(map(ord, name)))
    N = max(Na, Nb)
    t = np.arange(N) / 48000.0
    base = (0.4 * np.sin(2 * np.pi * 220 * t) + 0.2 * rng.standard_normal((C, N))).astype(np.float32)
    proc = (gain * base + noise * rng.standard_normal((C, N)) + 0.01).astype(np.float32)
    return base[:, :Na].copy(), proc[:, :Nb].copy()


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def main():
    nt = load("ref_null", "/root/reference/egregora_null_test_suite.py")
    ev = load("ref_eval", "/root/reference/egregora_audio_eval_pack.py")
    node = nt.Audio_Null_Test()
    aud = lambda x: {"waveform": torch.from_numpy(x)[None], "sample_rate": 48000}  # noqa: E731
    G = {}
    for name, (C, Na, Nb, inv, ls, gain, noise) in CASES.items():
        A, B = signals(name, C, Na, Nb, gain, noise)
        out, m = node.execute(aud(A), aud(B), invert_b=inv, least_squares_scale=ls, compute_null_lufs=False,
                              compute_lsd=False)
        null = out["samples"]
        n = min(Na, Nb)
        idx = np.linspace(0, n - 1, 257).astype(np.int64)
        G[name] = {"C": C, "Na": Na, "Nb": Nb, "invert_b": inv, "least_squares_scale": ls, "gain": gain, "noise": noise,
                   "metrics": m, "si_sdr_db": float(ev._si_sdr(A, B)),
                   "null_sha256": hashlib.sha256(np.ascontiguousarray(null).tobytes()).hexdigest(),
                   "probe_idx": idx.tolist(), "probe": null[:, idx].astype(np.float64).tolist()}
    (OUT / "eval_golden.json").write_text(json.dumps(G, indent=1))
    print("wrote", OUT / "eval_golden.json")


if __name__ == "__main__":
    main()
