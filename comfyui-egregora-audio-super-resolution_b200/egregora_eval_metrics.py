"""Clip-scale evaluation reductions on the device.

Mirrors the arithmetic of the reference's `Audio_Null_Test.execute` (/root/reference/egregora_null_test_suite.py
:421-467 — trim to the shorter clip, optional least-squares scale, inversion, null = A + B, corr_coef, null_rms_dbfs,
overshoot_count, clipped_pct, scale_k) and of `_si_sdr` (egregora_audio_eval_pack.py:414-429).  Not here: the LUFS,
LSD and HF-band options of those nodes (K-weighting / STFT paths).  One C-ABI call, two streaming
passes; no CPU fallback.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import _abi

_KEYS = {"si_sdr_db": "EGR_EVAL_SI_SDR_DB", "corr_coef": "EGR_EVAL_CORR", "null_rms_dbfs": "EGR_EVAL_NULL_RMS_DBFS",
         "overshoot_count": "EGR_EVAL_OVERSHOOT", "clipped_pct": "EGR_EVAL_CLIPPED_PCT", "scale_k": "EGR_EVAL_SCALE_K"}


def null_test(ref: torch.Tensor, proc: torch.Tensor, *, invert_b: bool = True, least_squares_scale: bool = False,
              want_null: bool = True) -> Tuple[torch.Tensor, Dict[str, float]]:
    """ref, proc: [C,N] float32 (host or device; lengths may differ, both are trimmed to the shorter, ref :426-428).
    Returns (null [C,N] device tensor or None, metrics dict with the reference's key names + si_sdr_db)."""
    if not torch.cuda.is_available():
        raise RuntimeError("CUDA GPU not detected. The B200-native metrics have no CPU fallback (sm_100a kernels only).")
    if ref.dim() != 2 or proc.dim() != 2 or ref.shape[0] != proc.shape[0]:
        raise RuntimeError(f"ref/proc must be [C, N] with equal channel counts; got {tuple(ref.shape)} and {tuple(proc.shape)}")
    device = torch.device("cuda", torch.cuda.current_device())
    lib = _abi.init(device.index or 0)
    a = ref.detach().to(device=device, dtype=torch.float32).contiguous()
    b = proc.detach().to(device=device, dtype=torch.float32).contiguous()
    C, n = a.shape[0], min(a.shape[1], b.shape[1])
    if n == 0:
        raise RuntimeError("empty audio")
    null = torch.empty((C, n), dtype=torch.float32, device=device) if want_null else None
    met = torch.zeros(_abi.K["EGR_EVAL_NUM"], dtype=torch.float64, device=device)
    wb = int(lib.egr_eval_workspace_bytes())
    work = torch.empty(wb, dtype=torch.uint8, device=device)
    _abi.check(lib.egr_eval_null_test(a.data_ptr(), a.shape[1], b.data_ptr(), b.shape[1], C, n, 1 if invert_b else 0,
                                      1 if least_squares_scale else 0, null.data_ptr() if want_null else None,
                                      met.data_ptr(), work.data_ptr(), wb, torch.cuda.current_stream().cuda_stream),
               "egr_eval_null_test")
    m = met.cpu().tolist()
    out = {k: float(m[_abi.K[v]]) for k, v in _KEYS.items()}
    out["overshoot_count"] = int(out["overshoot_count"])
    return null, out


def si_sdr(ref: torch.Tensor, est: torch.Tensor) -> float:
    """Scale-invariant SDR in dB of `est` against `ref` ([C,N] or [N]; channels are averaged, ref :414-429)."""
    r = ref if ref.dim() == 2 else ref[None, :]
    e = est if est.dim() == 2 else est[None, :]
    if r.shape[0] != e.shape[0]:  # the reference averages each side's channels separately
        r, e = r.double().mean(0, keepdim=True).float(), e.double().mean(0, keepdim=True).float()
    return null_test(r, e, want_null=False)[1]["si_sdr_db"]
