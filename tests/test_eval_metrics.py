"""Evaluation reductions: Audio_Null_Test arithmetic (egregora_null_test_suite.py:421-467)
, _si_sdr (egregora_audio_eval_pack.py:414-429) and the LSD pair _stft_mag / _lsd (:389-411).

CPU: the numpy restatement against the golden file produced by the reference functions (null signal bit-exact,
metrics to 1e-12 relative — the float64 dot products go through the same BLAS here).  GPU: egr_eval_null_test through
`egregora_eval_metrics`: null signal bit-exact, float64 metrics within 1e-9 relative of the reference's (different
summation order), corr_coef within 3e-6 (the reference forms it in float32).

LSD: the oracle and a numpy float32 emulation of the kernel's own FFT scheme (tests/lsd_cases.py) are checked here on
the CPU against goldens made by the reference node; the GPU run of egr_eval_lsd is tests/test_zz_eval_lsd_gpu.py.
Tolerance 1e-4 dB wherever every STFT bin holds signal.  Where bins hold nothing but FFT rounding noise (the
`bandlimited_noise_floor` case: exact zeros above a quarter of the band) the reference's own value is set by ITS FFT's
rounding noise — another correct float32 FFT lands several dB away — so that case is only bounded, not matched.
"""
import hashlib
import json

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from lsd_cases import emulate_kernel_lsd, signals as lsd_signals
from lufs_cases import signal as lufs_signal
from hf_cases import signal as hf_signal
from oracle import eval_oracle as O


@pytest.fixture(scope="module")
def egold():
    return json.loads((GOLDEN / "eval_golden.json").read_text())


def _signals(name, c):  # same generator as tests/golden/make_eval_golden.py
    rng = np.random.default_rng(sum(map(ord, name)))
    N = max(c["Na"], c["Nb"])
    t = np.arange(N) / 48000.0
    base = (0.4 * np.sin(2 * np.pi * 220 * t) + 0.2 * rng.standard_normal((c["C"], N))).astype(np.float32)
    proc = (c["gain"] * base + c["noise"] * rng.standard_normal((c["C"], N)) + 0.01).astype(np.float32)
    return base[:, :c["Na"]].copy(), proc[:, :c["Nb"]].copy()


def _close(a, b, rel, abs_=0.0):
    return abs(a - b) <= abs_ + rel * max(abs(a), abs(b))


def _check_null(c, null):
    assert hashlib.sha256(np.ascontiguousarray(null).tobytes()).hexdigest() == c["null_sha256"]
    assert np.array_equal(null[:, c["probe_idx"]].astype(np.float64), np.asarray(c["probe"]))


def test_oracle_matches_reference_golden(egold):
    for name, c in egold.items():
        A, B = _signals(name, c)
        null, m = O.null_test(A, B, c["invert_b"], c["least_squares_scale"])
        _check_null(c, null)
        for k, v in c["metrics"].items():
            assert _close(m[k], v, 1e-12, 1e-300), (name, k, m[k], v)
        assert _close(O.si_sdr(A, B), c["si_sdr_db"], 1e-12)


@pytest.mark.gpu
def test_kernel_matches_reference_golden(egold, cuda_dev, pkg):
    from egregora_b200 import egregora_eval_metrics as M
    for name, c in egold.items():
        A, B = _signals(name, c)
        null, m = M.null_test(torch.from_numpy(A), torch.from_numpy(B), invert_b=c["invert_b"],
                              least_squares_scale=c["least_squares_scale"])
        _check_null(c, null.cpu().numpy())
        ref = c["metrics"]
        assert m["overshoot_count"] == ref["overshoot_count"]
        assert _close(m["clipped_pct"], ref["clipped_pct"], 1e-12)
        assert _close(m["scale_k"], ref["scale_k"], 1e-9)
        assert _close(m["null_rms_dbfs"], ref["null_rms_dbfs"], 1e-9)
        assert _close(m["corr_coef"], ref["corr_coef"], 0.0, 3e-6)  # the reference rounds in float32 (up to 5e-7 off the float64 value)
        assert _close(m["si_sdr_db"], c["si_sdr_db"], 1e-9, 1e-9), (name, m["si_sdr_db"], c["si_sdr_db"])
        assert _close(M.si_sdr(torch.from_numpy(A), torch.from_numpy(B)), c["si_sdr_db"], 1e-9, 1e-9)


@pytest.mark.gpu
def test_kernel_properties_at_clip_scale(cuda_dev, pkg):
    """c5-sized clip (5 min stereo): identical inputs null to exact zero (rms floor -200 dBFS, corr 1), the metrics are
    deterministic launch to launch, and a scaled copy is recovered by the least-squares scale."""
    from egregora_b200 import egregora_eval_metrics as M
    g = torch.Generator().manual_seed(4)
    x = (torch.randn((2, 48000 * 300), generator=g) * 0.2).to(cuda_dev)
    null, m = M.null_test(x, x)
    assert float(null.abs().max()) == 0.0 and m["overshoot_count"] == 0
    assert abs(m["null_rms_dbfs"] + 200.0) < 1e-9 and abs(m["corr_coef"] - 1.0) < 1e-9
    y = x * 0.5
    _, m1 = M.null_test(x, y, least_squares_scale=True)
    _, m2 = M.null_test(x, y, least_squares_scale=True)
    assert m1 == m2
    assert abs(m1["scale_k"] - 2.0) < 1e-6 and m1["null_rms_dbfs"] < -120.0
    assert m1["si_sdr_db"] > 120.0
    with pytest.raises(RuntimeError):
        M.null_test(torch.zeros(1, 10), torch.zeros(2, 10))


# ------------------------------------------------------------------------------------------------ LSD
@pytest.fixture(scope="module")
def lgold():
    return json.loads((GOLDEN / "eval_lsd_golden.json").read_text())


def test_lsd_oracle_matches_reference_golden(lgold):
    for name, c in lgold.items():
        A, B = lsd_signals(name, c)
        mean, p95 = O.lsd(A, B, c["n_fft"], c["hop"])
        tol = 1e-5 if c["band"] >= 1.0 else 0.5  # same numpy FFT -> same rounding noise, but do not rely on it
        assert abs(mean - c["lsd_mean_db"]) <= tol and abs(p95 - c["lsd_p95_db"]) <= tol, (name, mean, p95)
        a, b = A.mean(axis=0), B.mean(axis=0)
        n = min(a.size, b.size)
        per = O.lsd_per_frame(O.stft_mag(a[:n], c["n_fft"], c["hop"]), O.stft_mag(b[:n], c["n_fft"], c["hop"]))
        assert per.size == c["frames"]
        assert np.allclose(per[c["probe_idx"]], c["per_probe"], rtol=0, atol=tol)


def test_lsd_kernel_arithmetic_emulated_matches_reference_golden(lgold):
    """The kernel's FFT scheme restated in numpy float32 (thread loops vectorised) against the reference numbers."""
    for name, c in lgold.items():
        A, B = lsd_signals(name, c)
        mean, p95, per = emulate_kernel_lsd(A, B, c["n_fft"], c["hop"])
        assert per.size == c["frames"]
        if c["band"] >= 1.0:
            assert abs(mean - c["lsd_mean_db"]) <= 1e-4 and abs(p95 - c["lsd_p95_db"]) <= 1e-4, (name, mean, p95)
            assert np.allclose(per[c["probe_idx"]], c["per_probe"], rtol=0, atol=1e-4)
        else:  # noise-floor bins: both implementations are "right"; they only have to be in the same region
            assert abs(mean - c["lsd_mean_db"]) <= 8.0, (name, mean, c["lsd_mean_db"])


# ------------------------------------------------------------------------------------------------ LUFS
def test_lufs_oracle_matches_reference_golden():
    """oracle.integrated_lufs / k_weight against the reference's functions: the filtered signal bit for bit (sha256),
    the loudness to 1e-12 (same float64 arithmetic)."""
    lg = json.loads((GOLDEN / "eval_lufs_golden.json").read_text())
    for name, c in lg.items():
        x = lufs_signal(name, c)
        y = O.k_weight(c["sr"], x)
        assert hashlib.sha256(np.ascontiguousarray(y).tobytes()).hexdigest() == c["kweight_sha256"], name
        assert abs(O.integrated_lufs(x, c["sr"]) - c["lufs"]) <= 1e-12, name


# ------------------------------------------------------------------------------------------------ HF band ratio
def test_hf_band_oracle_matches_reference_golden():
    hg = json.loads((GOLDEN / "eval_hf_golden.json").read_text())
    for name, c in hg.items():
        assert abs(O.hf_band_db(hf_signal(name, c), c["sr"], c["lo_hz"]) - c["hf_db"]) <= 1e-9, name
