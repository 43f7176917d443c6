"""Generates tests/golden/null_full_golden.json: Audio_Null_Test.execute of the REFERENCE
(/root/reference/egregora_null_test_suite.py:421-467) with EVERY metric toggle on, on seeded full-band inputs.
Run here only:  python tests/golden/make_null_full_golden.py
"""
import importlib.util
import json
import sys
from pathlib import Path

import numpy as np
import torch

OUT = Path(__file__).resolve().parent
CASES = {  # name: (C, N, sr, invert_b, least_squares_scale, gain, noise)
    "stereo_ls_48k": (2, 60000, 48000, True, True, 0.8, 0.02),
    "mono_plain_44k1": (1, 50000, 44100, True, False, 1.0, 0.01),
    "stereo_noinvert_ls": (2, 40000, 48000, False, True, -0.6, 0.05),
}


def signals(name, C, N, sr, gain, noise):
    rng = np.random.default_rng(sum(map(ord, name)))
    t = np.arange(N) / sr
    base = (0.3 * np.sin(2 * np.pi * 440 * t) + 0.1 * rng.standard_normal((C, N))).astype(np.float32)
    proc = (gain * base + noise * rng.standard_normal((C, N))).astype(np.float32)
    return base, proc


def main():
    spec = importlib.util.spec_from_file_location("ref_null", "/root/reference/egregora_null_test_suite.py")
    nt = importlib.util.module_from_spec(spec)
    sys.modules["ref_null"] = nt
    spec.loader.exec_module(nt)
    node = nt.Audio_Null_Test()
    G = {}
    for name, (C, N, sr, inv, ls, gain, noise) in CASES.items():
        A, B = signals(name, C, N, sr, gain, noise)
        aud = lambda x: {"waveform": torch.from_numpy(x)[None], "sample_rate": sr}  # noqa: E731
        _, m = node.execute(aud(A), aud(B), invert_b=inv, least_squares_scale=ls, compute_corr=True, compute_null_rms=True,
                            compute_null_lufs=True, compute_lsd=True, compute_hf_residual=True, n_fft=2048, hop=512, hf_band_hz=8000)
        G[name] = {"C": C, "N": N, "sr": sr, "invert_b": inv, "least_squares_scale": ls, "gain": gain, "noise": noise, "metrics": m}
        print(name, m)
    (OUT / "null_full_golden.json").write_text(json.dumps(G, indent=1))


if __name__ == "__main__":
    main()
