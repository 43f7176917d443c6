"""Generate golden vectors by importing the REFERENCE itself (run in the build container only:
`python tests/golden/make_golden.py`).  /root/reference does not exist on the GPU box, so the outputs
(driver_golden.npz, surface.json) are committed and nothing else reads /root/reference at test time.

What is pinned: span tables, Hann endpoint values, _wola_stitch on seeded inputs,
_from_audio_dict/_to_cs coercions, the full run() driver with an injected chunk model (hashes + probes),
the scipy resample branch shape/probes, and the complete 19-node surface.
"""
import hashlib
import importlib.util
import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def load_ref_package():
    sys.modules.setdefault("soundfile", types.ModuleType("soundfile"))
    spec = importlib.util.spec_from_file_location(
        "egregora_ref", REF / "__init__.py", submodule_search_locations=[str(REF)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["egregora_ref"] = mod
    spec.loader.exec_module(mod)
    return mod


def jsonable(x):
    if isinstance(x, dict):
        return {k: jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [jsonable(v) for v in x]
    return x


def main():
    pkg = load_ref_package()
    sr_mod = sys.modules["egregora_ref.egregora_audio_super_resolution"]
    fl_mod = sys.modules["egregora_ref.egregora_fat_llama_gpu"]
    G = {}

    # ---- surface
    surface = {}
    for nid, cls in pkg.NODE_CLASS_MAPPINGS.items():
        surface[nid] = {
            "class": cls.__name__,
            "INPUT_TYPES": jsonable(cls.INPUT_TYPES()),
            "RETURN_TYPES": list(cls.RETURN_TYPES),
            "RETURN_NAMES": list(getattr(cls, "RETURN_NAMES", ())),
            "FUNCTION": cls.FUNCTION,
            "CATEGORY": cls.CATEGORY,
            "OUTPUT_NODE": bool(getattr(cls, "OUTPUT_NODE", False)),
            "display": pkg.NODE_DISPLAY_NAME_MAPPINGS.get(nid),
        }
    (OUT / "surface.json").write_text(json.dumps(surface, indent=1, ensure_ascii=False, sort_keys=True))

    # ---- spans
    win, hop = 245760, int((5.12 - 0.5) * 48000)
    totals = [0, 1, 1000, 221760, 245759, 245760, 245761, 443520, 467520, 480000, 8640000, 14400000, 28800000]
    for t in totals:
        G[f"spans_{t}"] = np.asarray(sr_mod._iter_chunks(t, win, hop), np.int64).reshape(-1, 2)
    G["spans_small_1000_64_48"] = np.asarray(sr_mod._iter_chunks(1000, 64, 48), np.int64)
    G["spans_small_halfhop_1000_64"] = np.asarray(sr_mod._iter_chunks(1000, 64, 32), np.int64)

    # ---- hann
    h = sr_mod._hann(win)
    G["hann_probe_idx"] = np.asarray([0, 1, 2, 1000, 122879, 122880, 245758, 245759], np.int64)
    G["hann_probe_val"] = h[G["hann_probe_idx"]]
    G["hann_sha256"] = np.frombuffer(hashlib.sha256(h.tobytes()).digest(), np.uint8)
    G["hann_64"] = sr_mod._hann(64)

    # ---- wola on seeded small cases (win 64, hop 48; ragged tail; L_pred != win; C in {1,2,3})
    rng = np.random.default_rng(1234)
    for name, (total, w, hp, C, lpred) in {
        "a": (1000, 64, 48, 2, 64),
        "b": (777, 64, 48, 1, 64),
        "c": (500, 64, 32, 3, 64),
        "d": (300, 64, 48, 2, 50),     # model returns fewer samples than win
        "e": (64, 64, 48, 2, 64),      # exactly one chunk
        "f": (10, 64, 48, 1, 64),      # shorter than one window
    }.items():
        spans = sr_mod._iter_chunks(total, w, hp)
        preds = [(rng.standard_normal((C, lpred)).astype(np.float32), s, L) for s, L in spans]
        out = sr_mod._wola_stitch(preds, total, w)
        G[f"wola_{name}_meta"] = np.asarray([total, w, hp, C, lpred], np.int64)
        G[f"wola_{name}_spans"] = np.asarray(spans, np.int64)
        G[f"wola_{name}_preds"] = np.stack([p[0] for p in preds])
        G[f"wola_{name}_out"] = out
    G["wola_empty"] = sr_mod._wola_stitch([], 5, 64)

    # ---- full driver with injected chunk models (identity, gain+delay), big window: keep hashes + probes
    class FakeRunner:
        REQ_SR, CHUNK_S, OVERLAP_S, CHUNK_SAMPLES = 48000, 5.12, 0.5, 245760
        calls = []

        def __init__(self, lowpass=False):
            self.lowpass = lowpass

        def infer(self, x):
            FakeRunner.calls.append(x.shape)
            return FakeRunner.fn(x)

    sr_mod._FlashSRRunner = FakeRunner
    node = sr_mod.EgregoraAudioSuperResolution()
    x = (rng.standard_normal((2, 480000)) * 0.1).astype(np.float32)
    G["driver_in"] = x[:, ::997].copy()  # probe of the input (the test regenerates x from the same rng)
    probes = np.asarray([0, 1, 2, 1000, 221759, 221760, 230000, 245759, 245760, 443520, 450000, 479999], np.int64)
    G["driver_probe_idx"] = probes
    for mname, fn in {
        "identity": lambda c: c,
        "gain_roll": lambda c: (0.5 * np.roll(c, 3, axis=1)).astype(np.float32),
        "short": lambda c: c[:, :200000],  # L_pred < win
    }.items():
        FakeRunner.fn = staticmethod(fn)
        FakeRunner.calls = []
        (res,) = node.run(audio={"waveform": torch.from_numpy(x)[None], "sample_rate": 48000})
        y = res["waveform"].numpy()[0]
        G[f"driver_{mname}_probe"] = y[:, probes]
        G[f"driver_{mname}_sha256"] = np.frombuffer(hashlib.sha256(y.tobytes()).digest(), np.uint8)
        G[f"driver_{mname}_ncalls"] = np.asarray([len(FakeRunner.calls)], np.int64)

    # 16 kHz mono in, 44.1 kHz out through the scipy resample branch
    FakeRunner.fn = staticmethod(lambda c: c)
    x16 = (rng.standard_normal((1, 160000)) * 0.1).astype(np.float32)
    (res,) = node.run(audio={"waveform": torch.from_numpy(x16)[None], "sample_rate": 16000}, output_sr="44100")
    y = res["waveform"].numpy()[0]
    G["driver_16k_shape"] = np.asarray(y.shape, np.int64)
    G["driver_16k_sr"] = np.asarray([res["sample_rate"]], np.int64)
    G["driver_16k_probe_idx"] = np.asarray([0, 1, 100, 44100, 220500, 440999], np.int64)
    G["driver_16k_probe"] = y[:, G["driver_16k_probe_idx"]]

    # ---- coercions
    a = rng.standard_normal((5, 2)).astype(np.float32) * 3
    G["to_cs_in_frames_first"] = a
    G["to_cs_out_frames_first"] = fl_mod._to_cs(a)
    b = rng.standard_normal((2, 9)).astype(np.float32) * 0.3
    G["to_cs_in_cs"] = b
    G["to_cs_out_cs"] = fl_mod._to_cs(b)
    G["to_cs_out_1d"] = fl_mod._to_cs(b[0])
    cs, sr = sr_mod._from_audio_dict((a, 8000))
    G["from_tuple_out"] = cs
    cs, sr = sr_mod._from_audio_dict({"waveform": torch.from_numpy(rng.standard_normal((3, 2, 2)).astype(np.float32)), "sample_rate": 7})
    G["from_dict_b3_shape"] = np.asarray(cs.shape, np.int64)
    try:
        sr_mod._from_audio_dict({"waveform": torch.zeros(5), "sample_rate": 7})
    except RuntimeError as e:
        (OUT / "errors.json").write_text(json.dumps({"from_audio_dict_1d": str(e)}, indent=1))

    np.savez_compressed(OUT / "driver_golden.npz", **G)
    print("wrote", OUT / "driver_golden.npz", sum(v.nbytes for v in G.values()), "bytes raw;", len(surface), "nodes")


if __name__ == "__main__":
    main()
