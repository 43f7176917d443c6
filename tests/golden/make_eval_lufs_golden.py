"""Generates tests/golden/eval_lufs_golden.json by importing the REFERENCE module /root/reference/
egregora_null_test_suite.py and calling its integrated_lufs / _k_weight on seeded inputs (the same functions again in
egregora_audio_eval_pack.py are checked to agree).  The reference filters sample by sample in Python, so the cases are
short.  Run here only:  python tests/golden/make_eval_lufs_golden.py
"""
import hashlib
import importlib.util
import json
import sys
from pathlib import Path

import numpy as np

OUT = Path(__file__).resolve().parent

CASES = {  # name: (C, N, sr, kind)
    "stereo_bursts_48k": (2, 96000, 48000, "bursts"),       # loud / quiet sections: the -10 LU gate removes blocks
    "mono_44k1": (1, 50000, 44100, "noise"),
    "short_under_one_block": (2, 9000, 48000, "noise"),     # N < 400 ms: one block over what there is
    "three_channels_16k": (3, 40000, 16000, "tone"),
    "ragged_tail": (2, 19200 + 4800 * 3 + 17, 48000, "bursts"),
}


def signal(name, C, N, sr, kind):
    rng = np.random.default_rng(sum(map(ord, name)))
    t = np.arange(N) / sr
    if kind == "noise":
        x = 0.1 * rng.standard_normal((C, N))
    elif kind == "tone":
        x = 0.3 * np.sin(2 * np.pi * 50.0 * t)[None] + 0.2 * np.sin(2 * np.pi * 3000.0 * t)[None] + 0.01 * rng.standard_normal((C, N))
    else:
        env = np.where((t * 2).astype(int) % 2 == 0, 0.5, 0.004)[None]
        x = env * rng.standard_normal((C, N)) + 0.05 * np.sin(2 * np.pi * 30.0 * t)[None]
    return x.astype(np.float32)


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def main():
    nt = load("ref_null", "/root/reference/egregora_null_test_suite.py")
    ev = load("ref_eval", "/root/reference/egregora_audio_eval_pack.py")
    G = {}
    for name, (C, N, sr, kind) in CASES.items():
        x = signal(name, C, N, sr, kind)
        audio = {"sample_rate": sr, "samples": x, "meta": {}}
        v = nt.integrated_lufs(audio)
        assert v == ev.integrated_lufs(audio)
        y = nt._k_weight(sr, x)
        G[name] = {"C": C, "N": N, "sr": sr, "kind": kind, "lufs": v,
                   "kweight_sha256": hashlib.sha256(np.ascontiguousarray(y).tobytes()).hexdigest()}
        print(name, v)
    (OUT / "eval_lufs_golden.json").write_text(json.dumps(G, indent=1))


if __name__ == "__main__":
    main()
