"""The reference's example graph shape (Example/Audio Super Resolution.json: LoadAudio -> EgregoraAudioUpscaler ->
EgregoraFatLlamaGPU -> PreviewAudio) run through a minimal executor (tests/mini_executor.py) against this package's
NODE_CLASS_MAPPINGS: the prompt validates against the surface the reference generated (tests/golden/surface.json), the
AUDIO dict one node returns is what the next accepts, and the second stage matches its oracle on the first stage's output.

CPU part: prompt validation against both surfaces, executor semantics with stub nodes.  GPU part: the chain itself
(16 kHz stereo in -> device resampler -> FlashSR full spec, 1 chunk -> Fat-Llama 20 iterations).  The GPU part was
first run on hardware by the round-1 driver (passed); a plain hardware test since round 2.
"""
import json

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from mini_executor import PromptError, _validate, execute


def _prompt(sr_choice="48000", iters=20):
    return {
        "31": {"class_type": "LoadAudio", "inputs": {}},
        "27": {"class_type": "EgregoraAudioUpscaler", "inputs": {"audio": ["31", 0], "lowpass_input": False, "output_sr": sr_choice}},
        "26": {"class_type": "EgregoraFatLlamaGPU", "inputs": {
            "AUDIO": ["27", 0], "target_format": "wav", "max_iterations": iters, "threshold_value": 0.6,
            "target_bitrate_kbps": 1411, "toggle_normalize": True, "toggle_autoscale": True}},
    }


class _LoadAudioStub:
    RETURN_TYPES = ("AUDIO",)
    FUNCTION = "load"
    SR, C, S = 16000, 2, 24000

    @classmethod
    def INPUT_TYPES(cls):
        return {"required": {}}

    def load(self):
        rng = np.random.default_rng(5)
        t = np.arange(self.S) / self.SR
        x = 0.02 * rng.standard_normal((self.C, self.S))
        for h, a in ((220.0, 0.3), (660.0, 0.1), (3000.0, 0.05)):
            x += a * np.sin(2 * np.pi * h * t + rng.uniform(0, 6.28, (self.C, 1)))
        return ({"waveform": torch.from_numpy(x.astype(np.float32))[None], "sample_rate": self.SR},)


def test_prompt_is_valid_against_the_reference_surface(pkg):
    """Widget values of the prompt satisfy the INPUT_TYPES the REFERENCE declares (surface.json, produced by importing
    it) and the ones this package declares — same sockets, same combos, same ranges."""
    surface = json.loads((GOLDEN / "surface.json").read_text())
    p = _prompt()
    for nid in ("27", "26"):
        ct = p[nid]["class_type"]
        widgets = {k: v for k, v in p[nid]["inputs"].items() if not isinstance(v, list)}
        for spec in (surface[ct]["INPUT_TYPES"], pkg.NODE_CLASS_MAPPINGS[ct].INPUT_TYPES()):
            unlinked = {k: {n: d for n, d in (spec.get(k) or {}).items() if n in widgets or k == "optional"} for k in ("required", "optional")}
            _validate(ct, unlinked, widgets)
            for name, v in p[nid]["inputs"].items():
                if isinstance(v, list):
                    decl = {**(spec.get("required") or {}), **(spec.get("optional") or {})}[name]
                    assert decl[0] == "AUDIO"
    bad = _prompt(sr_choice="22050")
    with pytest.raises(PromptError):
        _validate("EgregoraAudioUpscaler", {"required": {k: v for k, v in surface["EgregoraAudioUpscaler"]["INPUT_TYPES"]["required"].items() if k != "audio"}},
                  {k: v for k, v in bad["27"]["inputs"].items() if not isinstance(v, list)})


def test_executor_semantics_with_stub_nodes():
    class Gain:
        RETURN_TYPES = ("AUDIO",)
        FUNCTION = "go"

        @classmethod
        def INPUT_TYPES(cls):
            return {"required": {"audio": ("AUDIO",), "g": ("FLOAT", {"default": 2.0, "min": 0.0, "max": 4.0})}}

        def go(self, audio, g):
            return ({"waveform": audio["waveform"] * g, "sample_rate": audio["sample_rate"]},)

    maps = {"LoadAudio": _LoadAudioStub, "Gain": Gain}
    out = execute({"a": {"class_type": "LoadAudio", "inputs": {}}, "b": {"class_type": "Gain", "inputs": {"audio": ["a", 0]}},
                   "c": {"class_type": "Gain", "inputs": {"audio": ["b", 0], "g": 0.5}}}, maps)
    assert torch.equal(out["c"][0]["waveform"], out["a"][0]["waveform"])   # default 2.0 then 0.5
    with pytest.raises(PromptError):
        execute({"a": {"class_type": "LoadAudio", "inputs": {}}, "b": {"class_type": "Gain", "inputs": {"audio": ["a", 0], "g": 9.0}}}, maps)
    with pytest.raises(PromptError):
        execute({"b": {"class_type": "Gain", "inputs": {"g": 1.0}}}, maps)
    with pytest.raises(PromptError):
        execute({"b": {"class_type": "Nope", "inputs": {}}}, maps)


@pytest.mark.gpu
def test_example_graph_chain_on_device(cuda_dev, pkg):
    from oracle import fat_llama_oracle as O
    maps = dict(pkg.NODE_CLASS_MAPPINGS)
    maps["LoadAudio"] = _LoadAudioStub
    out1 = execute(_prompt(), maps)
    up, fl = out1["27"][0], out1["26"][0]
    n48 = _LoadAudioStub.S * 3
    for a in (up, fl):
        assert set(a.keys()) == {"waveform", "sample_rate"} and a["sample_rate"] == 48000
        w = a["waveform"]
        assert w.device.type == "cpu" and w.dtype == torch.float32 and w.is_contiguous() and tuple(w.shape) == (1, 2, n48)
        assert bool(torch.isfinite(w).all())
    assert float(fl["waveform"].abs().max()) <= 1.0
    # stage 2 against its oracle on stage 1's actual output (PCM-16 wire: at most 1 LSB, nearly all samples equal)
    want, sr = O.node_run(up["waveform"][0].numpy(), 48000, 20, 0.6, 1411, True, True, dtype=np.float64)
    diff = np.abs(fl["waveform"][0].numpy() - want)
    assert sr == 48000 and np.max(diff) <= (1.0 / 32768.0) * 1.0001 and np.mean(diff > 0) < 0.05
    # the graph is deterministic run to run (fixed diffusion seed, deterministic kernels)
    out2 = execute(_prompt(), maps)
    assert torch.equal(out2["26"][0]["waveform"], fl["waveform"])
    # inputs are not mutated and nothing is retained (SURVEY §8b ownership)
    src = out1["31"][0]["waveform"]
    assert torch.equal(src, _LoadAudioStub().load()[0]["waveform"])
