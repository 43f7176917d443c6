"""Adaptive wet/dry mix of the reference's DeepFilterNet node:
egregora_audio_enhance_extras.py:548-605 (helpers) and :657-704 (mix, post gain, limiter).

CPU part: the numpy restatement (oracle/dfn_mix_oracle.py) is pinned bit for bit against the golden fixture made by
calling the reference's own methods (tests/golden/make_dfn_mix_golden.py) and — when /root/reference exists (this
container, not the GPU box) — against those methods directly on fresh inputs.  GPU part: egr_dfn_mix through the
package's `adaptive_mix`; bit-exact for the linear mix curve, <= 1e-6 absolute with the equal-power curve (the only
difference is the last bit of sinf / cosf).
"""
import hashlib
import importlib.util
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import dfn_mix_oracle as O

REF = Path("/root/reference/egregora_audio_enhance_extras.py")
DEFAULTS = dict(strength=0.65, mix_curve="equal_power", adaptive_mode="more_on_noise", adaptive_amount=0.45,
                vad_threshold=0.90, vad_smooth_ms=60, post_gain_db=0.5, limit_ceiling=True, ceiling=0.98)


@pytest.fixture(scope="module")
def mgold():
    return np.load(GOLDEN / "dfn_mix_golden.npz"), json.loads((GOLDEN / "dfn_mix_cases.json").read_text())


def _signals(name, C, T):  # same generator as tests/golden/make_dfn_mix_golden.py
    rng = np.random.default_rng(sum(map(ord, name)))
    t = np.arange(T) / 48000.0
    env = (0.15 + np.abs(np.sin(2 * np.pi * 0.7 * t)) ** 3)[None, :]
    speech = (rng.standard_normal((C, T)) * 0.35 * env).astype(np.float32)
    noise = (rng.standard_normal((C, T)) * 0.05).astype(np.float32)
    return np.clip(speech + noise, -1, 1).astype(np.float32), (speech * 0.9).astype(np.float32)


def _check(G, name, y, exact):
    if exact:
        assert hashlib.sha256(np.ascontiguousarray(y).tobytes()).digest() == G[f"{name}_sha256"].tobytes(), name
    if f"{name}_out" in G.files:
        want, got = G[f"{name}_out"], y
    else:
        idx = G[f"{name}_probe_idx"]
        want, got = G[f"{name}_probe"], y[:, idx]
    assert got.shape == want.shape
    if exact:
        assert np.array_equal(got, want), name
    else:
        assert float(np.abs(got - want).max()) <= 1e-6, (name, float(np.abs(got - want).max()))


# ------------------------------------------------------------------------------------------------ CPU
def test_oracle_matches_reference_golden(mgold):
    G, cases = mgold
    for name, c in cases.items():
        dry, wet = _signals(name, c["C"], c["T"])
        p = dict(DEFAULTS)
        p.update(c["kwargs"])
        y = O.adaptive_mix(dry, wet, 48000, **p)
        _check(G, name, y, exact=True)


@pytest.mark.skipif(not REF.exists(), reason="reference tree only exists in the build container")
def test_oracle_matches_reference_methods_directly():
    spec = importlib.util.spec_from_file_location("ref_extras", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cls = mod.Egregora_DeepFilterNet_Denoise
    node = cls.__new__(cls)
    rng = np.random.default_rng(11)
    for n in (1, 7, 479, 480, 481, 5000, 96137):
        x = (rng.standard_normal(n) * 0.1 * (np.abs(np.sin(np.arange(n) / 3000.0)) + 0.05)).astype(np.float32)
        a, b = node._vad_probs_rms_48k(x), O.vad_probs_rms(x)
        assert np.array_equal(a, b)
        for ms in (60, 0, 5, 300):
            sa, sb = node._smooth_probs(a, ms), O.smooth_probs(b, ms)
            assert np.array_equal(sa, sb)
            for mode in ("off", "more_on_noise", "more_on_speech", "gate_on_noise", "something_else"):
                ra = node._strength_per_frame(0.65, sa, mode, 0.45, 0.9)
                rb = O.strength_per_frame(0.65, sb, mode, 0.45, 0.9)
                assert np.array_equal(ra, rb)
                for curve in ("equal_power", "linear"):
                    ga, gb = node._gains_from_strength(ra, curve), O.gains_from_strength(rb, curve)
                    assert np.array_equal(ga[0], gb[0]) and np.array_equal(ga[1], gb[1])


def test_pairwise_sum_model_is_numpys():
    rng = np.random.default_rng(3)
    for n in (1, 5, 8, 9, 100, 128, 129, 240, 479, 480, 1000):
        a = (rng.standard_normal(n) * 3).astype(np.float32)
        assert O.pairwise_sum_f32(a) == np.add.reduce(a), n
    fr = (rng.standard_normal((17, 480))).astype(np.float32)
    want = np.asarray([np.add.reduce((f * f).astype(np.float32)) for f in fr], np.float32)
    assert np.array_equal(O._full_frames_sumsq(fr), want)


def test_percentile_model_is_numpys():
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 20, 21, 300, 301, 3001, 30000):
        v = np.abs(rng.standard_normal(n)).astype(np.float32)
        for q in (95.0, 50.0, 0.0, 100.0, 37.3):
            assert O.percentile_f32(v, q) == np.percentile(v, q), (n, q)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_kernel_matches_reference_golden(mgold, cuda_dev, pkg):
    from egregora_b200 import egregora_dfn_mix as M
    G, cases = mgold
    for name, c in cases.items():
        dry, wet = _signals(name, c["C"], c["T"])
        p = dict(DEFAULTS)
        p.update(c["kwargs"])
        y = M.adaptive_mix(torch.from_numpy(dry), torch.from_numpy(wet), 48000, **p)
        assert y.is_cuda and tuple(y.shape) == dry.shape
        _check(G, name, y.cpu().numpy(), exact=(p["mix_curve"] == "linear"))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["off", "more_on_noise", "more_on_speech", "gate_on_noise"])
@pytest.mark.parametrize("T", [1, 479, 480, 481, 4800 * 7 + 3, 48000 * 4])
def test_kernel_matches_oracle_linear_curve_bit_exact(mode, T, cuda_dev, pkg):
    """With the linear curve no transcendental is involved: frame RMS (numpy's pairwise sums), percentile (exact order
    statistics + float32 lerp), smoothing, strength, mix, gain, limiter must all agree to the bit."""
    from egregora_b200 import egregora_dfn_mix as M
    dry, wet = _signals(f"{mode}{T}", 2, T)
    kw = dict(DEFAULTS, adaptive_mode=mode, mix_curve="linear", vad_threshold=0.6, post_gain_db=1.5, ceiling=0.4)
    want = O.adaptive_mix(dry, wet, 48000, **kw)
    got = M.adaptive_mix(torch.from_numpy(dry).to(cuda_dev), torch.from_numpy(wet).to(cuda_dev), 48000, **kw).cpu().numpy()
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_kernel_clip_scale_and_errors(cuda_dev, pkg):
    """c5-sized clip (5 min stereo at 48 kHz): output bounded by the ceiling, silence stays silence, and the error
    paths raise RuntimeError as everywhere else in the package."""
    from egregora_b200 import egregora_dfn_mix as M
    T = 48000 * 300
    g = torch.Generator().manual_seed(9)
    dry = torch.randn((2, T), generator=g) * 0.3
    wet = dry * 0.8
    y = M.adaptive_mix(dry.to(cuda_dev), wet.to(cuda_dev), 48000)
    assert tuple(y.shape) == (2, T) and bool(torch.isfinite(y).all())
    assert float(y.abs().max()) <= 0.98 + 1e-6
    z = M.adaptive_mix(torch.zeros(1, 9600), torch.zeros(1, 9600), 48000)
    assert float(z.abs().max()) == 0.0
    with pytest.raises(RuntimeError):
        M.adaptive_mix(torch.zeros(1, 100), torch.zeros(1, 100), 44100)
    with pytest.raises(RuntimeError):
        M.adaptive_mix(torch.zeros(1, 100), torch.zeros(2, 100), 48000)
