"""Generates tests/golden/dfn_mix_golden.npz by importing the REFERENCE module
(/root/reference/egregora_audio_enhance_extras.py) and calling the DeepFilterNet node's own helper methods
(_vad_probs_rms_48k :548-559, _smooth_probs :561-573, _strength_per_frame :575-594, _gains_from_strength :596-605),
then carrying out steps 5-6 of its execute() (:657-704: expand the 10 ms frame gains, mix, clip, post gain, peak
limiter, clamp) on those outputs with the same numpy / torch calls.  The DeepFilterNet model is absent, so `wet` is a
seeded synthetic signal.  Run here only; the GPU box reads the committed fixture.
    python tests/golden/make_dfn_mix_golden.py
"""
import hashlib
import importlib.util
from pathlib import Path

import numpy as np
import torch

OUT = Path(__file__).resolve().parent
REF = Path("/root/reference/egregora_audio_enhance_extras.py")

CASES = {  # name: (C, T, kwargs)
    "default_2ch": (2, 48000 * 2 + 137, {}),
    "speech_linear": (1, 30011, {"adaptive_mode": "more_on_speech", "mix_curve": "linear", "vad_smooth_ms": 200}),
    "gate": (1, 48000, {"adaptive_mode": "gate_on_noise", "vad_threshold": 0.5, "post_gain_db": 0.0}),
    "off_hot": (2, 9600, {"adaptive_mode": "off", "strength": 1.0, "post_gain_db": 6.0, "ceiling": 0.5}),
    "nosmooth_nolimit": (1, 4800 + 1, {"vad_smooth_ms": 0, "limit_ceiling": False, "post_gain_db": 3.0}),
    "short": (1, 100, {}),
    "long": (2, 48000 * 30, {}),
}
DEFAULTS = dict(strength=0.65, mix_curve="equal_power", adaptive_mode="more_on_noise", adaptive_amount=0.45,
                vad_threshold=0.90, vad_smooth_ms=60, post_gain_db=0.5, limit_ceiling=True, ceiling=0.98)


def signals(name, C, T):
    rng = np.random.default_rng(sum(map(ord, name)))
    t = np.arange(T) / 48000.0
    env = (0.15 + np.abs(np.sin(2 * np.pi * 0.7 * t)) ** 3)[None, :]
    speech = (rng.standard_normal((C, T)) * 0.35 * env).astype(np.float32)
    noise = (rng.standard_normal((C, T)) * 0.05).astype(np.float32)
    dry = np.clip(speech + noise, -1, 1).astype(np.float32)
    wet = (speech * 0.9).astype(np.float32)
    return dry, wet


def reference_mix(node, dry, wet, p):
    sr = 48000
    hop = int(sr * 0.010)
    out = []
    for ch in range(dry.shape[0]):
        probs = node._vad_probs_rms_48k(dry[ch])
        vad_s = node._smooth_probs(probs, p["vad_smooth_ms"])
        s_eff = node._strength_per_frame(p["strength"], vad_s, p["adaptive_mode"], p["adaptive_amount"], p["vad_threshold"])
        s_per = np.repeat(s_eff, max(1, hop))[:dry.shape[1]].astype(np.float32)
        g_dry, g_wet = node._gains_from_strength(s_per, p["mix_curve"])
        y = np.clip(g_dry * dry[ch] + g_wet * wet[ch], -1.0, 1.0)
        out.append(torch.from_numpy(y))
    y = torch.stack(out, dim=0)
    if p["post_gain_db"] != 0.0:
        y = y * float(10.0 ** (p["post_gain_db"] / 20.0))
    if p["limit_ceiling"]:
        peak = torch.max(torch.abs(y)).item()
        if peak > p["ceiling"] and peak > 0:
            y = y * (p["ceiling"] / peak)
    return torch.clamp(y, -1.0, 1.0).numpy(), s_eff


def main():
    spec = importlib.util.spec_from_file_location("ref_extras", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cls = mod.Egregora_DeepFilterNet_Denoise
    node = cls.__new__(cls)
    G = {}
    for name, (C, T, kw) in CASES.items():
        p = dict(DEFAULTS)
        p.update(kw)
        dry, wet = signals(name, C, T)
        y, s_last = reference_mix(node, dry, wet, p)
        assert y.dtype == np.float32
        G[f"{name}_meta"] = np.asarray([C, T, sum(map(ord, name))], np.int64)
        G[f"{name}_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(y).tobytes()).digest(), np.uint8)
        G[f"{name}_strength_lastch"] = s_last
        if C * T <= 100000:
            G[f"{name}_out"] = y
        else:
            idx = np.linspace(0, T - 1, 1025).astype(np.int64)
            G[f"{name}_probe_idx"] = idx
            G[f"{name}_probe"] = y[:, idx]
    np.savez_compressed(OUT / "dfn_mix_golden.npz", **G)
    (OUT / "dfn_mix_cases.json").write_text(__import__("json").dumps(
        {k: {"C": v[0], "T": v[1], "kwargs": v[2]} for k, v in CASES.items()}, indent=1))
    print("wrote", OUT / "dfn_mix_golden.npz", sum(v.nbytes for v in G.values()), "bytes raw")


if __name__ == "__main__":
    main()
