"""🎛️ Spectral Enhance (Fat Llama — GPU) — B200-native drop-in for the reference node of the same ID.

Mirrors /root/reference/egregora_fat_llama_gpu.py:228-294 (`EgregoraFatLlamaGPU`): same INPUT_TYPES,
RETURN_TYPES, FUNCTION, CATEGORY, input coercion and error messages.  The reference shells the audio
through two 16-bit files around `fat_llama.audio_fattener.feed.upscale` (CuPy/cuFFT); here the same
wire-format arithmetic (PCM-16 quantise on the way in and on the way out, :34-37 / :291) runs as CUDA
kernels in memory, and the iterative FFT -> gate -> inverse-FFT loop is egr_fatllama_run (hand-written
mixed-radix FFT, no cuFFT/CuPy).  Host code is plumbing only; there is no CPU fallback.
"""
from __future__ import annotations

import tempfile
import time
import wave
from pathlib import Path
from typing import Tuple

import numpy as np
import torch

from . import _abi

RETURN_TYPES = ("AUDIO",)
FUNCTION = "run"
CATEGORY = "Egregora/Audio"


# ------------------------------------------------------------------------------------ coercion (host)
def _to_cs(x) -> np.ndarray:
    """channels-first f32 [C,S] from [S], [S,C] or [C,S]; rescales by the peak when it exceeds 1 (ref :18-32)."""
    a = np.asarray(x, dtype=np.float32)
    if a.ndim == 1:
        a = a[None, :]
    elif a.ndim == 2:
        rows, cols = a.shape
        if cols <= 8 and rows > cols:
            a = a.T
    else:
        a = a.reshape(-1)[None, :]
    peak = float(np.max(np.abs(a))) if a.size else 0.0
    if peak > 1.0:
        a = a / (peak + 1e-8)
    return a.astype(np.float32)


def _read_pcm_wav(path: str) -> Tuple[np.ndarray, int]:
    """WAV reader (PCM 8/16/24/32-bit) -> float32 frames-first, like sf.read(dtype='float32')."""
    try:
        with wave.open(path, "rb") as w:
            sr, ch, sw, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
            raw = w.readframes(n)
    except (wave.Error, EOFError, OSError) as e:
        raise RuntimeError(f"Failed to read audio file {path}: {e} (without the soundfile package only PCM WAV files "
                           "can be decoded; FLAC/OGG/float WAV need `pip install soundfile`)") from e
    if sw == 2:
        x = np.frombuffer(raw, "<i2").astype(np.float32) / 32768.0
    elif sw == 4:
        x = np.frombuffer(raw, "<i4").astype(np.float64) / 2147483648.0
    elif sw == 3:
        b = np.frombuffer(raw, np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v & 0x800000, v - 0x1000000, v)
        x = v.astype(np.float64) / 8388608.0
    elif sw == 1:
        x = (np.frombuffer(raw, np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise RuntimeError(f"Failed to read audio file {path}: unsupported WAV sample width {sw}")
    x = x.astype(np.float32)
    return (x.reshape(-1, ch) if ch > 1 else x), sr


def _read_audio_file(path: str) -> Tuple[np.ndarray, int]:
    """sf.read(path, dtype='float32', always_2d=False) of the reference (:62-66, :74-78) — through soundfile itself when
    it is importable (any libsndfile format: WAV incl. float / extensible, FLAC, OGG), else the stdlib PCM-WAV reader.
    Failures are RuntimeErrors, the pack's only error type."""
    try:
        import soundfile as sf  # type: ignore
    except Exception:
        return _read_pcm_wav(path)
    try:
        y, sr = sf.read(path, dtype="float32", always_2d=False)
    except Exception as e:
        raise RuntimeError(f"Failed to read audio file {path}: {e}") from e
    return np.asarray(y, np.float32), int(sr)


def _normalize_audio_input(AUDIO=None, audio_path: str = "", audio_url: str = "") -> Tuple[torch.Tensor, int]:
    """AUDIO dict / (array, sr) / path / url -> ([C,S] f32 tensor, sr).  Branch order, the dict branch's lack
    of peak rescale, and the error strings follow ref :40-80."""
    if isinstance(AUDIO, dict) and "waveform" in AUDIO and "sample_rate" in AUDIO:
        wf = AUDIO["waveform"]
        if not isinstance(wf, torch.Tensor):
            wf = torch.as_tensor(np.asarray(wf))
        if wf.dim() == 3:
            wf = wf[0]
        if wf.dim() != 2:
            raise RuntimeError(f"Unexpected AUDIO tensor shape: {tuple(wf.shape)} (want [C,T])")
        return wf.detach().float(), int(AUDIO["sample_rate"])
    if isinstance(AUDIO, (list, tuple)) and len(AUDIO) == 2:
        arr, sr = AUDIO
        return torch.from_numpy(_to_cs(np.asarray(arr))), int(sr)
    if audio_path:
        p = Path(audio_path)
        if not p.exists():
            raise RuntimeError(f"audio_path not found: {audio_path}")
        y, sr = _read_audio_file(str(p))
        return torch.from_numpy(_to_cs(y)), int(sr)
    if audio_url:
        import requests
        r = requests.get(audio_url, timeout=60)
        r.raise_for_status()
        p = Path(tempfile.gettempdir()) / f"eg_url_{int(time.time() * 1000)}.wav"
        p.write_bytes(r.content)
        y, sr = _read_audio_file(str(p))
        return torch.from_numpy(_to_cs(y)), int(sr)
    raise RuntimeError("No AUDIO provided.")


def _ensure_gpu_stack() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "CUDA GPU not detected. Fat Llama (GPU) requires an NVIDIA GPU. "
            "If you need CPU, use the separate Fat Llama — CPU/FFTW node."
        )
    return torch.device("cuda", torch.cuda.current_device())


# ------------------------------------------------------------------------------------ device pipeline
SAMPLE_WIDTH = 2  # the temp WAV the reference writes is PCM_16 (sf.write default subtype)


def upscale_factor(sample_rate: int, channels: int, target_bitrate_kbps: int) -> int:
    """upstream: factor = round(target bitrate / source bitrate), at least 1 ([RECALL], see oracle header)."""
    src = sample_rate * channels * 8 * SAMPLE_WIDTH
    return max(1, int(round(target_bitrate_kbps * 1000.0 / src)))


def fat_llama_device(x_dev: torch.Tensor, sr: int, max_iterations: int, threshold_value: float,
                     target_bitrate_kbps: int, toggle_normalize: bool, toggle_autoscale: bool,
                     return_prequant: bool = False):
    """[C,S] device f32 in [-1,1] -> ([C,S*U] device f32, sr*U): the whole reference node body on device."""
    dev = x_dev.device
    lib = _abi.init(dev.index or 0)
    st = torch.cuda.current_stream().cuda_stream
    C, S = x_dev.shape
    x_dev = x_dev.contiguous()
    U = upscale_factor(sr, C, target_bitrate_kbps)
    # temp-WAV write (:34-37) + upstream read: f32 -> PCM-16 -> integer-scaled f32
    q = torch.empty((C, S), dtype=torch.int16, device=dev)
    _abi.check(lib.egr_pcm16_quantize(x_dev.data_ptr(), q.data_ptr(), C * S, st), "egr_pcm16_quantize")
    samples = torch.empty((C, S), dtype=torch.float32, device=dev)
    _abi.check(lib.egr_pcm16_to_float(q.data_ptr(), samples.data_ptr(), C * S, 1.0, st), "egr_pcm16_to_float")
    # feed.upscale arithmetic
    y = torch.empty((C, S * U), dtype=torch.float32, device=dev)
    wbytes = lib.egr_fatllama_workspace_bytes(C, S, U)
    work = torch.empty((max(int(wbytes), 16),), dtype=torch.uint8, device=dev)
    flags = (_abi.K["EGR_FL_NORMALIZE"] if toggle_normalize else 0) | (_abi.K["EGR_FL_AUTOSCALE"] if toggle_autoscale else 0)
    _abi.check(lib.egr_fatllama_run(samples.data_ptr(), y.data_ptr(), C, S, U, int(max_iterations),
                                    float(threshold_value), flags, work.data_ptr(), int(wbytes), st), "egr_fatllama_run")
    # patched write_audio (:188-208): integer-scaled data is brought back to [-1,1] by 2**(8*sw-1)
    # — decided and applied on the device (the branch needs max|y|; reading it back would stall the stream)
    peak = torch.empty((1,), dtype=torch.float32, device=dev)
    _abi.check(lib.egr_absmax(y.data_ptr(), C * S * U, peak.data_ptr(), st), "egr_absmax")
    _abi.check(lib.egr_scale_if_above(y.data_ptr(), C * S * U, peak.data_ptr(), 1.0, 1.0 / float(2 ** (8 * SAMPLE_WIDTH - 1)), st),
               "egr_scale_if_above")
    # output file (PCM-16) + sf.read(float32) (:291)
    q2 = torch.empty((C, S * U), dtype=torch.int16, device=dev)
    _abi.check(lib.egr_pcm16_quantize(y.data_ptr(), q2.data_ptr(), C * S * U, st), "egr_pcm16_quantize")
    out = torch.empty((C, S * U), dtype=torch.float32, device=dev)
    _abi.check(lib.egr_pcm16_to_float(q2.data_ptr(), out.data_ptr(), C * S * U, 1.0 / 32768.0, st), "egr_pcm16_to_float")
    if return_prequant:
        return out, sr * U, y
    return out, sr * U


def _to_host(t: torch.Tensor) -> torch.Tensor:
    """device [C,T] -> fresh pinned CPU tensor (PCIe-speed copy, one sync — the only one of the node body)."""
    host = torch.empty(t.shape, dtype=torch.float32, pin_memory=True)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return host


class EgregoraFatLlamaGPU:
    """Spectral Enhance (Fat Llama — GPU only); adaptive filter stays disabled as in the reference (:223)."""

    @classmethod
    def INPUT_TYPES(cls):
        return {
            "required": {
                "target_format": (["wav", "flac"],),
                "max_iterations": ("INT", {"default": 300, "min": 1, "max": 5000}),
                "threshold_value": ("FLOAT", {"default": 0.6, "min": 0.0, "max": 1.0, "step": 0.01}),
                "target_bitrate_kbps": ("INT", {"default": 1411, "min": 64, "max": 5000}),
                "toggle_normalize": ("BOOLEAN", {"default": True}),
                "toggle_autoscale": ("BOOLEAN", {"default": True}),
            },
            "optional": {
                "AUDIO": ("AUDIO",),
                "audio_path": ("STRING", {"default": ""}),
                "audio_url": ("STRING", {"default": ""}),
            },
        }

    RETURN_TYPES = RETURN_TYPES
    FUNCTION = FUNCTION
    CATEGORY = CATEGORY
    OUTPUT_NODE = False

    def run(self, target_format, max_iterations, threshold_value, target_bitrate_kbps, toggle_normalize,
            toggle_autoscale, AUDIO=None, audio_path="", audio_url=""):
        device = _ensure_gpu_stack()
        cs, in_sr = _normalize_audio_input(AUDIO, audio_path, audio_url)
        # target_format only picked the temp container (wav / 16-bit flac): both are lossless PCM-16 wires
        x_dev = cs.to(device=device, dtype=torch.float32)
        out, sr = fat_llama_device(x_dev, in_sr, int(max_iterations), float(threshold_value),
                                   int(target_bitrate_kbps), bool(toggle_normalize), bool(toggle_autoscale))
        return ({"waveform": _to_host(out).unsqueeze(0).contiguous(), "sample_rate": int(sr)},)  # [1,C,T]


NODE_CLASS_MAPPINGS = {"EgregoraFatLlamaGPU": EgregoraFatLlamaGPU}
NODE_DISPLAY_NAME_MAPPINGS = {"EgregoraFatLlamaGPU": "🎛️ Spectral Enhance (Fat Llama — GPU)"}
