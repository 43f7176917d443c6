"""Adaptive wet/dry mix of the reference's 🧼 DeepFilterNet node, on the device.

Mirrors steps 5-6 of `Egregora_DeepFilterNet_Denoise.execute` (/root/reference/egregora_audio_enhance_extras.py
:657-704) and the helpers it calls — `_vad_probs_rms_48k` :548-559, `_smooth_probs` :561-573, `_strength_per_frame`
:575-594, `_gains_from_strength` :596-605 — with the node's own parameter names and defaults (:609-625).  The
DeepFilterNet model (third-party `deepfilternet`, not part of the reference tree) is NOT here: `wet` is whatever
produced the denoised signal.  This is the prefix of BASELINE config c5 (DFN3 -> FlashSR -> Fat-Llama) that the
reference itself implements; one C-ABI call (`egr_dfn_mix`) replaces its per-channel numpy loops.  48 kHz only:
at other rates the reference's VAD branch goes through `df.io.resample`, which is third-party too.
No CPU fallback; host code is plumbing.
"""
from __future__ import annotations

import torch

from . import _abi

_MODES = {"off": "EGR_MIX_OFF", "more_on_noise": "EGR_MIX_MORE_ON_NOISE", "more_on_speech": "EGR_MIX_MORE_ON_SPEECH",
          "gate_on_noise": "EGR_MIX_GATE_ON_NOISE"}


def adaptive_mix(dry: torch.Tensor, wet: torch.Tensor, sr: int, *, strength: float = 0.65,
                 mix_curve: str = "equal_power", adaptive_vad_source: str = "rms",
                 adaptive_mode: str = "more_on_noise", adaptive_amount: float = 0.45, vad_threshold: float = 0.90,
                 vad_smooth_ms: int = 60, post_gain_db: float = 0.5, limit_ceiling: bool = True,
                 ceiling: float = 0.98) -> torch.Tensor:
    """dry, wet: [C,T] float32 (host or device) at `sr` -> mixed, gained and limited [C,T] DEVICE tensor."""
    if not torch.cuda.is_available():
        raise RuntimeError("CUDA GPU not detected. The B200-native mix has no CPU fallback (sm_100a kernels only).")
    if dry.dim() != 2 or wet.shape != dry.shape:
        raise RuntimeError(f"dry/wet must both be [C, T]; got {tuple(dry.shape)} and {tuple(wet.shape)}")
    if int(sr) != 48000:
        raise RuntimeError("adaptive_mix runs at 48 kHz only (the reference resamples its VAD branch through df.io otherwise)")
    if adaptive_vad_source == "rnnoise":
        raise RuntimeError("adaptive_vad_source='rnnoise' needs the third-party pyrnnoise model; use 'rms'")
    K = _abi.K
    device = torch.device("cuda", torch.cuda.current_device())
    lib = _abi.init(device.index or 0)
    d = dry.detach().to(device=device, dtype=torch.float32).contiguous()
    w = wet.detach().to(device=device, dtype=torch.float32).contiguous()
    C, T = d.shape
    out = torch.empty_like(d)
    if T == 0:
        return out
    wb = int(lib.egr_dfn_mix_workspace_bytes(C, T))
    work = torch.empty(wb, dtype=torch.uint8, device=device)
    mode = K[_MODES.get(adaptive_mode, "EGR_MIX_OFF")]   # unknown modes behave like "off" in the reference (:592-593)
    curve = K["EGR_CURVE_EQUAL_POWER"] if mix_curve == "equal_power" else K["EGR_CURVE_LINEAR"]
    vad = K["EGR_VAD_RMS"] if adaptive_vad_source == "rms" else K["EGR_VAD_NONE"]
    _abi.check(lib.egr_dfn_mix(d.data_ptr(), w.data_ptr(), out.data_ptr(), C, T, int(sr), float(strength), curve, vad, mode,
                               float(adaptive_amount), float(vad_threshold), int(vad_smooth_ms), float(post_gain_db),
                               1 if limit_ceiling else 0, float(ceiling), work.data_ptr(), wb,
                               torch.cuda.current_stream().cuda_stream), "egr_dfn_mix")
    return out
