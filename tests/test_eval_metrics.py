"""Evaluation reductions: Audio_Null_Test arithmetic (egregora_null_test_suite.py:421-467)
and _si_sdr (egregora_audio_eval_pack.py:414-429).

CPU: the numpy restatement against the golden file produced by the reference functions (null signal bit-exact,
metrics to 1e-12 relative — the float64 dot products go through the same BLAS here).  GPU: egr_eval_null_test through
`egregora_eval_metrics`: null signal bit-exact, float64 metrics within 1e-9 relative of the reference's (different
summation order), corr_coef within 3e-6 (the reference forms it in float32).
"""
import hashlib
import json

import numpy as np
import pytest
import torch

from conftest import GOLDEN
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
