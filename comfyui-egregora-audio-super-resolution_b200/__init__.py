"""B200-native drop-in for the two hot paths of ComfyUI-Egregora-Audio-Super-Resolution.

Registers the same node IDs as the reference's __init__.py:33-52 for the three hot nodes.  The other 16
node IDs of the reference pack (enhance extras / eval pack / null-test suite) are outside this build's
scope; when the original pack is installed beside this one (or EGREGORA_REFERENCE_PACK
points at it) their mappings are merged in unchanged so existing graphs keep resolving every ID.
"""
import importlib.util
import os
import sys
from pathlib import Path

from .egregora_audio_super_resolution import EgregoraAudioSuperResolution
from .egregora_fat_llama_gpu import EgregoraFatLlamaGPU
from .egregora_fat_llama_cpu import EgregoraFatLlamaCPU

NODE_CLASS_MAPPINGS = {
    "EgregoraAudioUpscaler": EgregoraAudioSuperResolution,
    "EgregoraFatLlamaGPU": EgregoraFatLlamaGPU,
    "EgregoraFatLlamaCPU": EgregoraFatLlamaCPU,
}

NODE_DISPLAY_NAME_MAPPINGS = {
    "EgregoraAudioUpscaler": "🎧 Audio Super Resolution (FlashSR)",
    "EgregoraFatLlamaGPU": "🎛️ Spectral Enhance (Fat Llama — GPU)",
    "EgregoraFatLlamaCPU": "🎛️ Spectral Enhance (Fat Llama — CPU/FFTW)",
}

_AUX_MODULES = ("egregora_audio_enhance_extras", "egregora_audio_eval_pack", "egregora_null_test_suite")


def _merge_reference_aux_nodes():
    """Soft re-export of the non-hot nodes from an installed copy of the original pack (never required)."""
    here = Path(__file__).resolve().parent
    cands = [os.environ.get("EGREGORA_REFERENCE_PACK", ""),
             str(here.parent / "ComfyUI-Egregora-Audio-Super-Resolution")]
    for root in cands:
        if not root or not Path(root).is_dir():
            continue
        for name in _AUX_MODULES:
            f = Path(root) / f"{name}.py"
            if not f.exists():
                continue
            try:
                spec = importlib.util.spec_from_file_location(f"_egregora_ref_{name}", f)
                mod = importlib.util.module_from_spec(spec)
                sys.modules[spec.name] = mod
                spec.loader.exec_module(mod)
                for k, v in getattr(mod, "NODE_CLASS_MAPPINGS", {}).items():
                    NODE_CLASS_MAPPINGS.setdefault(k, v)
                for k, v in getattr(mod, "NODE_DISPLAY_NAME_MAPPINGS", {}).items():
                    NODE_DISPLAY_NAME_MAPPINGS.setdefault(k, v)
            except Exception:
                continue  # same soft-import policy as the reference's __init__.py:8-30
        return


_merge_reference_aux_nodes()

__all__ = ["NODE_CLASS_MAPPINGS", "NODE_DISPLAY_NAME_MAPPINGS"]
