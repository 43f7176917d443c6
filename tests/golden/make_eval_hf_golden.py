"""Generates tests/golden/eval_hf_golden.json by importing the REFERENCE module /root/reference/
egregora_null_test_suite.py and calling _band_energy_hi_db on seeded inputs.  Run here only.
    python tests/golden/make_eval_hf_golden.py
"""
import importlib.util
import json
import sys
from pathlib import Path

import numpy as np

OUT = Path(__file__).resolve().parent
CASES = {  # name: (C, N, sr, lo_hz, hf_gain)  — hf_gain scales the content above 10 kHz
    "stereo_smooth_len": (2, 96000, 48000, 8000, 0.1),
    "mono_odd_len": (1, 30011, 44100, 5000, 0.01),        # odd N: not a packed-real length
    "prime_len_bluestein": (2, 10007, 48000, 12000, 1.0),  # prime N: Bluestein path
    "lo_above_content": (1, 48000, 48000, 20000, 0.001),
    "lo_1k": (3, 20000, 16000, 1000, 0.3),
}


def signal(name, C, N, sr, hf_gain):
    rng = np.random.default_rng(sum(map(ord, name)))
    t = np.arange(N) / sr
    lo = 0.3 * np.sin(2 * np.pi * 440.0 * t)[None] + 0.05 * rng.standard_normal((C, N)).cumsum(axis=1) / 50.0
    hi = hf_gain * rng.standard_normal((C, N))
    return (lo + hi).astype(np.float32)


def main():
    spec = importlib.util.spec_from_file_location("ref_null", "/root/reference/egregora_null_test_suite.py")
    nt = importlib.util.module_from_spec(spec)
    sys.modules["ref_null"] = nt
    spec.loader.exec_module(nt)
    G = {}
    for name, (C, N, sr, lo_hz, g) in CASES.items():
        x = signal(name, C, N, sr, g)
        v = nt._band_energy_hi_db(x, sr, lo_hz)
        G[name] = {"C": C, "N": N, "sr": sr, "lo_hz": lo_hz, "hf_gain": g, "hf_db": v,
                   "bins_hi": int(np.sum(np.fft.rfftfreq(N, d=1.0 / sr) >= lo_hz))}
        print(name, v)
    (OUT / "eval_hf_golden.json").write_text(json.dumps(G, indent=1))


if __name__ == "__main__":
    main()
