"""Clip-scale evaluation reductions on the device.

Mirrors the arithmetic of the reference's `Audio_Null_Test.execute` (/root/reference/egregora_null_test_suite.py
:421-467 — trim to the shorter clip, optional least-squares scale, inversion, null = A + B, corr_coef, null_rms_dbfs,
overshoot_count, clipped_pct, scale_k), of `_si_sdr` (egregora_audio_eval_pack.py:414-429) and of `_stft_mag` +
`_lsd` (egregora_audio_eval_pack.py:389-411; the same two functions again at egregora_null_test_suite.py:167-189).
of `integrated_lufs` + `_k_weight` (egregora_null_test_suite.py:125-164) and of `_band_energy_hi_db` (:190-197) —
every metric `Audio_Null_Test.execute` can return.  One C-ABI call per metric group; no CPU fallback.
"""
from __future__ import annotations

import ctypes
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


def lsd(ref: torch.Tensor, proc: torch.Tensor, n_fft: int = 2048, hop: int = 512, proc_gain: float = 1.0) -> Tuple[float, float]:
    """Log-spectral distance (lsd_mean_db, lsd_p95_db) of `proc` against `ref`, as Metrics_LSD_SISDR.execute computes
    it (egregora_audio_eval_pack.py:453-467): channel means, both trimmed to the shorter clip, `_stft_mag` frames
    (symmetric Hann, no centring), `_lsd`.  ref, proc: [C,N] or [N] float32, host or device.  proc_gain: the null test's
    least-squares scale (its `B = (B * k).astype(float32)` step, :436) when the LSD is taken after it."""
    if not torch.cuda.is_available():
        raise RuntimeError("CUDA GPU not detected. The B200-native metrics have no CPU fallback (sm_100a kernels only).")
    r = ref if ref.dim() == 2 else ref[None, :]
    e = proc if proc.dim() == 2 else proc[None, :]
    if r.dim() != 2 or e.dim() != 2:
        raise RuntimeError(f"ref/proc must be [C, N] or [N]; got {tuple(ref.shape)} and {tuple(proc.shape)}")
    device = torch.device("cuda", torch.cuda.current_device())
    lib = _abi.init(device.index or 0)
    a = r.detach().to(device=device, dtype=torch.float32).contiguous()
    b = e.detach().to(device=device, dtype=torch.float32).contiguous()
    if a.shape[0] != b.shape[0]:  # the reference takes each side's channel mean separately (float32, :456-457)
        a, b = a.mean(0, keepdim=True), b.mean(0, keepdim=True)
    C, n = a.shape[0], min(a.shape[1], b.shape[1])
    if n == 0:
        raise RuntimeError("empty audio")
    met = torch.zeros(_abi.K["EGR_LSD_NUM"], dtype=torch.float64, device=device)
    wb = int(lib.egr_eval_lsd_workspace_bytes(n, int(n_fft), int(hop)))
    work = torch.empty(max(wb, 256), dtype=torch.uint8, device=device)
    _abi.check(lib.egr_eval_lsd(a.data_ptr(), a.shape[1], b.data_ptr(), b.shape[1], C, n, int(n_fft), int(hop), float(proc_gain),
                                met.data_ptr(), work.data_ptr(), wb, torch.cuda.current_stream().cuda_stream), "egr_eval_lsd")
    m = met.cpu().tolist()
    return float(m[_abi.K["EGR_LSD_MEAN_DB"]]), float(m[_abi.K["EGR_LSD_P95_DB"]])


def integrated_lufs(x: torch.Tensor, sample_rate: int) -> float:
    """The reference's `integrated_lufs` (egregora_null_test_suite.py:143-164 = egregora_audio_eval_pack.py:153-167) of
    x [C,N] or [N] float32, host or device: its one-pole high-pass + first-difference tilt (`_k_weight`, float32, bit
    exact), channel mean, 400 ms / 100 ms blocks, -10 LU relative gate."""
    if not torch.cuda.is_available():
        raise RuntimeError("CUDA GPU not detected. The B200-native metrics have no CPU fallback (sm_100a kernels only).")
    a = x if x.dim() == 2 else x[None, :]
    if a.dim() != 2 or a.shape[1] == 0:
        raise RuntimeError(f"audio must be a non-empty [C, N] or [N] tensor; got {tuple(x.shape)}")
    device = torch.device("cuda", torch.cuda.current_device())
    lib = _abi.init(device.index or 0)
    a = a.detach().to(device=device, dtype=torch.float32).contiguous()
    C, n = a.shape
    met = torch.zeros(_abi.K["EGR_LUFS_NUM"], dtype=torch.float64, device=device)
    wb = int(lib.egr_eval_lufs_workspace_bytes(C, n, int(sample_rate)))
    work = torch.empty(wb, dtype=torch.uint8, device=device)
    _abi.check(lib.egr_eval_lufs(a.data_ptr(), n, C, n, int(sample_rate), met.data_ptr(), work.data_ptr(), wb,
                                 torch.cuda.current_stream().cuda_stream), "egr_eval_lufs")
    return float(met.cpu()[_abi.K["EGR_LUFS_INTEGRATED"]])


def hf_band_db(x: torch.Tensor, sample_rate: int, lo_hz: float) -> float:
    """`_band_energy_hi_db` (egregora_null_test_suite.py:190-197): energy of the channel mean at and above `lo_hz` over
    its total energy, in dB, from one whole-clip FFT (the hand-written path-B transform).  x [C,N] or [N] float32."""
    if not torch.cuda.is_available():
        raise RuntimeError("CUDA GPU not detected. The B200-native metrics have no CPU fallback (sm_100a kernels only).")
    a = x if x.dim() == 2 else x[None, :]
    if a.dim() != 2 or a.shape[1] == 0:
        raise RuntimeError(f"audio must be a non-empty [C, N] or [N] tensor; got {tuple(x.shape)}")
    device = torch.device("cuda", torch.cuda.current_device())
    lib = _abi.init(device.index or 0)
    a = a.detach().to(device=device, dtype=torch.float32).contiguous()
    C, n = a.shape
    plan = ctypes.c_void_p()
    _abi.check(lib.egr_fft_plan_create(n, 1, ctypes.byref(plan)), "egr_fft_plan_create")
    try:
        met = torch.zeros(_abi.K["EGR_HF_NUM"], dtype=torch.float64, device=device)
        wb = int(lib.egr_eval_hf_band_workspace_bytes(plan, n))
        work = torch.empty(wb, dtype=torch.uint8, device=device)
        _abi.check(lib.egr_eval_hf_band(plan, a.data_ptr(), n, C, n, int(sample_rate), float(lo_hz), met.data_ptr(),
                                        work.data_ptr(), wb, torch.cuda.current_stream().cuda_stream), "egr_eval_hf_band")
        out = float(met.cpu()[_abi.K["EGR_HF_RESIDUAL_DB"]])   # the read-back also orders the plan's destruction after the work
    finally:
        lib.egr_fft_plan_destroy(plan)
    return out


def audio_null_test(ref: torch.Tensor, proc: torch.Tensor, sample_rate: int, *, invert_b: bool = True,
                    least_squares_scale: bool = False, compute_corr: bool = True, compute_null_rms: bool = True,
                    compute_null_lufs: bool = True, compute_lsd: bool = True, compute_hf_residual: bool = False,
                    n_fft: int = 2048, hop: int = 512, hf_band_hz: float = 8000) -> Tuple[torch.Tensor, Dict[str, float]]:
    """Everything `Audio_Null_Test.execute` computes (egregora_null_test_suite.py:421-467), same keyword names, same
    metric keys, same toggles; ref / proc are [C,N] float32 at `sample_rate` (already aligned and matched, as the node's
    second input says).  Returns (null [C,N] device tensor, metrics dict)."""
    null, m = null_test(ref, proc, invert_b=invert_b, least_squares_scale=least_squares_scale)
    out: Dict[str, float] = {}
    if compute_corr:
        out["corr_coef"] = m["corr_coef"]
    if compute_null_rms:
        out["null_rms_dbfs"] = m["null_rms_dbfs"]
    if compute_null_lufs:
        out["null_lufs"] = integrated_lufs(null, sample_rate)
    if compute_lsd:   # of the channel means of A and of the scaled B (:439-440, :456-458); a sign does not change |X|
        gain = float(torch.tensor(m["scale_k"], dtype=torch.float32)) if least_squares_scale else 1.0
        out["lsd_mean_db"], out["lsd_p95_db"] = lsd(ref, proc, n_fft, hop, proc_gain=gain)
    if compute_hf_residual:
        out["hf_residual_db"] = hf_band_db(null, sample_rate, hf_band_hz)
    out["overshoot_count"] = m["overshoot_count"]
    out["clipped_pct"] = m["clipped_pct"]
    out["scale_k"] = m["scale_k"]
    return null, out
