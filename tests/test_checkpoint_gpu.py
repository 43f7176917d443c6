"""The node with real checkpoint FILES: three synthetic .pth files in upstream naming are resolved through
EGREGORA_FLASHSR_WEIGHTS / flashsr_min --ckpt-dir, loaded by flashsr_weights.py, and the node's output matches the fp32
oracle loaded from the SAME files (reference: egregora_audio_super_resolution.py:260-265, :346-359, :388-431)."""
import wave

import numpy as np
import pytest
import torch

from conftest import ROOT, load_pkg

load_pkg()
pytestmark = pytest.mark.gpu


def _clip(total, seed=1234):
    import bench
    return bench.synth_audio(total, 1, seed=seed)


def test_node_loads_checkpoint_and_matches_oracle(cuda_dev, synthetic_ckpt_dir, monkeypatch):
    from egregora_b200 import egregora_audio_super_resolution as N, flashsr_weights as FW, flashsr_model as M
    from oracle import driver_oracle, flashsr_oracle as O
    d, W_written = synthetic_ckpt_dir
    monkeypatch.delenv("EGREGORA_FLASHSR_RANDOM_INIT", raising=False)   # the checkpoint is the ONLY source of weights here
    monkeypatch.setenv("EGREGORA_FLASHSR_WEIGHTS", str(d))
    node = N.EgregoraAudioSuperResolution()
    node.NUM_STEPS, node.SEED = 1, 99
    x = _clip(N.CHUNK_SAMPLES)                                           # one span: c2's shape through the whole node
    (res,) = node.run(audio={"waveform": x[None], "sample_rate": 48000}, lowpass_input=True, output_sr="48000")
    eng = N.get_engine(cuda_dev)
    assert eng.weights_tag.startswith("ckpt:") and str(d.resolve()) in eng.weights_tag
    # oracle side: weights from the same files, the same noise, the reference driver around the model
    W = FW.load_checkpoint(d, M.default_spec())
    noise = eng.make_noise(1, 99, 0).cpu()
    model = lambda c: O.run_flashsr(eng.spec, W, torch.from_numpy(np.ascontiguousarray(c)), noise, steps=1, lowpass=True)[0].numpy()  # noqa: E731
    want, sr = driver_oracle.run_driver(x.numpy(), 48000, model)
    got = res["waveform"][0].numpy()
    assert res["sample_rate"] == sr == 48000 and got.shape == want.shape
    rms = float(np.sqrt(np.mean((got - want) ** 2)))
    assert rms < 1e-3, rms
    # what was loaded is what was written (weight-norm folded back), not a random re-initialisation
    for k in ("unet.input_blocks.1.0.in_layers.2.weight", "vocoder.ups.0.0.weight", "vae.decoder.mid.attn_1.q.weight"):
        assert torch.allclose(eng.weights[k], W_written[k], rtol=1e-6, atol=1e-7), k


def test_node_without_checkpoint_raises(cuda_dev, tmp_path, monkeypatch):
    from egregora_b200 import egregora_audio_super_resolution as N
    monkeypatch.delenv("EGREGORA_FLASHSR_RANDOM_INIT", raising=False)
    monkeypatch.setenv("EGREGORA_FLASHSR_WEIGHTS", str(tmp_path))
    with pytest.raises(RuntimeError, match="FlashSR weights missing"):
        N.EgregoraAudioSuperResolution().run(audio={"waveform": torch.zeros(1, 1, 4800), "sample_rate": 48000})


def test_flashsr_min_cli(cuda_dev, synthetic_ckpt_dir, tmp_path, monkeypatch, capsys):
    """flashsr_min --ckpt-dir DIR --in a.wav --out b.wav: same flags as the reference CLI (flashsr_min.py:6-11), prints OK,
    --ckpt-dir decides the weights, output = the node's result for the mono mix, written at --target-sr."""
    from egregora_b200 import egregora_audio_super_resolution as N, flashsr_min
    d, _ = synthetic_ckpt_dir
    monkeypatch.delenv("EGREGORA_FLASHSR_RANDOM_INIT", raising=False)
    monkeypatch.delenv("EGREGORA_FLASHSR_WEIGHTS", raising=False)
    x = (_clip(60000, seed=5)[0].numpy() * 0.8).astype(np.float32)
    stereo = np.stack([x, 0.5 * x[::-1]], 1)
    src, dst = tmp_path / "in.wav", tmp_path / "out.wav"
    flashsr_min._write_audio(str(src), stereo, 48000)
    flashsr_min.main(["--ckpt-dir", str(d), "--in", str(src), "--out", str(dst), "--target-sr", "48000"])
    assert capsys.readouterr().out.strip().endswith("OK")
    got, sr = flashsr_min._read_audio(str(dst))
    assert sr == 48000 and got.shape == (60000, 1)
    back, _ = flashsr_min._read_audio(str(src))
    node = N.EgregoraAudioSuperResolution()
    node.CKPT_DIR = str(d)
    (res,) = node.run(audio={"waveform": torch.from_numpy(back.mean(1))[None, None], "sample_rate": 48000})
    want = res["waveform"][0, 0].numpy()
    assert np.max(np.abs(got[:, 0] - want)) <= 1.0 / 32768 + 1e-7       # PCM-16 file
    with pytest.raises(RuntimeError, match="FlashSR weights missing"):
        flashsr_min.main(["--ckpt-dir", str(tmp_path), "--in", str(src), "--out", str(dst)])
