"""Generates tests/golden/eval_lsd_golden.json by importing the REFERENCE module /root/reference/
egregora_audio_eval_pack.py and calling Metrics_LSD_SISDR.execute (which runs _stft_mag + _lsd, :389-411, :453-467) on
seeded inputs; also checks that egregora_null_test_suite.py's copies of the two functions give the same numbers.
Run here only:  python tests/golden/make_eval_lsd_golden.py
"""
import importlib.util
import json
import sys
from pathlib import Path

import numpy as np
import torch

OUT = Path(__file__).resolve().parent

CASES = {  # name: (C, N_ref, N_proc, n_fft, hop, gain, noise, band) — band < 1 zeroes the top of BOTH spectra exactly
    "fullband_stereo": (2, 96000, 96000, 2048, 512, 0.8, 0.02, 1.0),
    "short_one_frame": (1, 1000, 1000, 2048, 512, 0.5, 0.01, 1.0),
    "ragged_nfft512": (2, 30011, 29000, 512, 128, 1.0, 0.05, 1.0),
    "mono_nfft1024": (1, 48000, 50000, 1024, 256, 1.3, 0.1, 1.0),
    "nfft4096": (2, 60000, 60000, 4096, 1024, 0.9, 0.03, 1.0),
    "nfft8192_hop64": (1, 20000, 20000, 8192, 64, 0.7, 0.02, 1.0),
    "nfft64": (1, 5000, 5000, 64, 64, 1.1, 0.2, 1.0),
    "bandlimited_noise_floor": (2, 96000, 96000, 2048, 512, 0.9, 0.0, 0.25),
}


def signals(name, C, Na, Nb, gain, noise, band):
    rng = np.random.default_rng(sum(map(ord, name)))
    N = max(Na, Nb)
    t = np.arange(N) / 48000.0
    base = (0.3 * np.sin(2 * np.pi * 440 * t) + 0.1 * rng.standard_normal((C, N))).astype(np.float32)
    proc = (gain * base + noise * rng.standard_normal((C, N))).astype(np.float32)
    if band < 1.0:
        def cut(x):
            X = np.fft.rfft(x.astype(np.float64))
            X[..., int(X.shape[-1] * band):] = 0
            return np.fft.irfft(X, n=x.shape[-1]).astype(np.float32)
        base, proc = cut(base), cut(proc)
    return base[:, :Na].copy(), proc[:, :Nb].copy()


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def main():
    ev = load("ref_eval", "/root/reference/egregora_audio_eval_pack.py")
    nt = load("ref_null", "/root/reference/egregora_null_test_suite.py")
    node = ev.Metrics_LSD_SISDR()
    aud = lambda x: {"waveform": torch.from_numpy(x)[None], "sample_rate": 48000}  # noqa: E731
    G = {}
    for name, (C, Na, Nb, n_fft, hop, gain, noise, band) in CASES.items():
        A, B = signals(name, C, Na, Nb, gain, noise, band)
        (m,) = node.execute(aud(A), aud(B), n_fft=n_fft, hop=hop, compute_lsd=True, compute_si_sdr=False)
        a, b = A.mean(axis=0), B.mean(axis=0)
        n = min(a.size, b.size)
        SA, SB = ev._stft_mag(a[:n], n_fft, hop), ev._stft_mag(b[:n], n_fft, hop)
        assert ev._lsd(SA, SB) == (m["lsd_mean_db"], m["lsd_p95_db"])
        assert nt._lsd(nt._stft_mag(a[:n], n_fft, hop), nt._stft_mag(b[:n], n_fft, hop)) == ev._lsd(SA, SB)
        LA, LB = 20 * np.log10(SA + 1e-12), 20 * np.log10(SB + 1e-12)
        per = np.sqrt(np.mean((LA - LB) ** 2, axis=0) + 1e-12)
        idx = np.unique(np.linspace(0, per.size - 1, 33).astype(np.int64))
        G[name] = {"C": C, "Na": Na, "Nb": Nb, "n_fft": n_fft, "hop": hop, "gain": gain, "noise": noise, "band": band,
                   "lsd_mean_db": m["lsd_mean_db"], "lsd_p95_db": m["lsd_p95_db"], "frames": int(per.size),
                   "probe_idx": idx.tolist(), "per_probe": per[idx].astype(np.float64).tolist()}
    (OUT / "eval_lsd_golden.json").write_text(json.dumps(G, indent=1))
    print("wrote", OUT / "eval_lsd_golden.json")


if __name__ == "__main__":
    main()
