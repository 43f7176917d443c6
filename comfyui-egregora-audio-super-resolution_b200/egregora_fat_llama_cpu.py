"""🎛️ Spectral Enhance (Fat Llama — CPU/FFTW) — node ID kept for existing graphs.

Mirrors /root/reference/egregora_fat_llama_cpu.py:136-191: same INPUT_TYPES (no toggle inputs,
max_iterations default 800 / max 10000) and the same call binding — `upscale` is invoked WITHOUT the
toggle kwargs (:126-134), so upstream's defaults (normalize on, autoscale on) apply.  In this pack the
node shares the B200 kernels with the GPU node: the product path has no CPU implementation (the CPU
restatement lives in oracle/ and is only the parity checker and the timed baseline).
"""
from __future__ import annotations

import torch

from .egregora_fat_llama_gpu import (CATEGORY, FUNCTION, RETURN_TYPES, _normalize_audio_input, _to_host,
                                     fat_llama_device)


def _ensure_device() -> torch.device:
    """The reference's CPU node needs no GPU (egregora_fat_llama_cpu.py:77-134: fat_llama_fftw + pyFFTW).  This build
    ships no CPU arithmetic at all (north_star: "no CPU fallback"), so on a machine without CUDA this node ID cannot
    run — say exactly that instead of the GPU node's "use the CPU node" hint, which would point back here."""
    if not torch.cuda.is_available():
        raise RuntimeError(
            "Fat Llama — CPU/FFTW: this B200-native build of the pack keeps the node ID for existing graphs but runs it "
            "on the CUDA kernels of the GPU node; it ships no CPU/FFTW implementation and no CUDA GPU was detected. "
            "Install the original pack (fat-llama-fftw) for CPU-only machines."
        )
    return torch.device("cuda", torch.cuda.current_device())


class EgregoraFatLlamaCPU:
    @classmethod
    def INPUT_TYPES(cls):
        return {
            "required": {
                "target_format": (["wav", "flac"],),
                "max_iterations": ("INT", {"default": 800, "min": 1, "max": 10000}),
                "threshold_value": ("FLOAT", {"default": 0.6, "min": 0.0, "max": 1.0, "step": 0.01}),
                "target_bitrate_kbps": ("INT", {"default": 1411, "min": 64, "max": 5000}),
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

    def run(self, target_format, max_iterations, threshold_value, target_bitrate_kbps, AUDIO=None,
            audio_path="", audio_url=""):
        device = _ensure_device()
        cs, in_sr = _normalize_audio_input(AUDIO, audio_path, audio_url)
        out, sr = fat_llama_device(cs.to(device=device, dtype=torch.float32), in_sr, int(max_iterations),
                                   float(threshold_value), int(target_bitrate_kbps), True, True)
        return ({"waveform": _to_host(out).unsqueeze(0).contiguous(), "sample_rate": int(sr)},)


NODE_CLASS_MAPPINGS = {"EgregoraFatLlamaCPU": EgregoraFatLlamaCPU}
NODE_DISPLAY_NAME_MAPPINGS = {"EgregoraFatLlamaCPU": "🎛️ Spectral Enhance (Fat Llama — CPU/FFTW)"}
