"""Generates tests/golden/resample_golden.npz by importing the REFERENCE module (read-only tree at
/root/reference) and calling its own `_resample_hq` (egregora_audio_super_resolution.py:159-207; soxr is
absent in this image, so its scipy.signal.resample_poly branch :181-191 runs).  Run here only; the GPU box
has no /root/reference and reads the committed fixture.
    python tests/golden/make_resample_golden.py
"""
import hashlib
import importlib.util
import sys
from pathlib import Path

import numpy as np

OUT = Path(__file__).resolve().parent
REF = Path("/root/reference/egregora_audio_super_resolution.py")

CASES = {  # name: (channels, n_in, src_sr, dst_sr, keep full output?)
    "16k_48k": (1, 4000, 16000, 48000, True),
    "441_48": (2, 4410, 44100, 48000, True),
    "48_441": (2, 4800, 48000, 44100, True),
    "48_96": (1, 2000, 48000, 96000, True),
    "22_48": (1, 2205, 22050, 48000, True),
    "8k_48k": (1, 801, 8000, 48000, True),
    "48_8": (1, 4801, 48000, 8000, True),
    "tiny7": (1, 7, 44100, 48000, True),
    "one": (1, 1, 16000, 48000, True),
    "long_441_48": (2, 441000, 44100, 48000, False),
    "long_48_441": (1, 480001, 48000, 44100, False),
}


def main():
    spec = importlib.util.spec_from_file_location("ref_sr", REF)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_sr"] = mod
    spec.loader.exec_module(mod)
    try:
        import soxr  # noqa: F401
        raise SystemExit("soxr present: the reference would not take its scipy branch")
    except ImportError:
        pass
    G = {}
    for name, (C, n, src, dst, full) in CASES.items():
        rng = np.random.default_rng(sum(map(ord, name)))
        x = (rng.standard_normal((C, n)) * 0.25).astype(np.float32)
        y = mod._resample_hq(x, src, dst)
        assert y.dtype == np.float32
        G[f"{name}_meta"] = np.asarray([C, n, src, dst, y.shape[1], sum(map(ord, name))], np.int64)
        G[f"{name}_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(y).tobytes()).digest(), np.uint8)
        if full:
            G[f"{name}_out"] = y
        else:
            idx = np.linspace(0, y.shape[1] - 1, 257).astype(np.int64)
            G[f"{name}_probe_idx"] = idx
            G[f"{name}_probe"] = y[:, idx]
    np.savez_compressed(OUT / "resample_golden.npz", **G)
    print("wrote", OUT / "resample_golden.npz", sum(v.nbytes for v in G.values()), "bytes raw")


if __name__ == "__main__":
    main()
