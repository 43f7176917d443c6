"""Checkpoint loading (reference egregora_audio_super_resolution.py:260-265, :282-320, :346-359):
three files in upstream naming -> the model's parameter set, loud failures, never silent random weights."""
import os
import sys

import pytest
import torch

from conftest import ROOT, load_pkg

load_pkg()
sys.path.insert(0, str(ROOT / "tools"))
import make_synthetic_ckpt as S  # noqa: E402
from egregora_b200 import flashsr_model as M, flashsr_weights as FW  # noqa: E402


@pytest.fixture(scope="module")
def tiny_ckpt(tmp_path_factory):
    spec = M.tiny_spec()
    W = M.init_weights(spec, 3)
    d = tmp_path_factory.mktemp("ckpt_tiny")
    S.write(d, spec, W)
    return spec, W, d


def test_round_trip_upstream_naming(tiny_ckpt):
    spec, W, d = tiny_ckpt
    got = FW.load_checkpoint(d, spec)
    assert list(got) == list(M.param_shapes(spec))
    for k, v in W.items():
        assert got[k].shape == v.shape and got[k].dtype == torch.float32
        assert torch.allclose(got[k], v, rtol=1e-6, atol=1e-7), k      # weight-norm folding: g * v / ||v||
    # files really are in upstream naming, not ours
    raw = torch.load(str(d / "student_ldm.pth"), weights_only=True)["state_dict"]
    assert all(k.startswith(("model.diffusion_model.", "first_stage_model.", "cond_stage_model.")) for k in raw)
    voc = torch.load(str(d / "sr_vocoder.pth"), weights_only=True)["generator"]
    assert any(k.endswith(".weight_g") for k in voc) and "conv_pre.weight" not in voc


def test_missing_file_raises_reference_message(tiny_ckpt, tmp_path):
    spec, W, d = tiny_ckpt
    for f in ("student_ldm.pth", "vae.pth"):
        (tmp_path / f).write_bytes((d / f).read_bytes())
    with pytest.raises(RuntimeError) as e:
        FW.load_checkpoint(tmp_path, spec)
    assert "FlashSR weights missing" in str(e.value) and "sr_vocoder.pth" in str(e.value)   # reference :316-319
    assert "models/audio/flashsr" in str(e.value)


def test_missing_and_misshaped_tensors_are_named(tiny_ckpt, tmp_path):
    spec, W, d = tiny_ckpt
    for f in FW.HF_FILES:
        (tmp_path / f).write_bytes((d / f).read_bytes())
    sd = torch.load(str(d / "vae.pth"), weights_only=True)
    sd["state_dict"].pop("encoder.conv_in.weight")
    sd["state_dict"]["decoder.conv_out.bias"] = torch.zeros(7)
    torch.save(sd, str(tmp_path / "vae.pth"))
    with pytest.raises(RuntimeError) as e:
        FW.load_checkpoint(tmp_path, spec)
    msg = str(e.value)
    assert "vae.pth" in msg and "encoder.conv_in.weight" in msg and "decoder.conv_out.bias" in msg and "(7,)" in msg


def test_unreadable_file(tiny_ckpt, tmp_path):
    spec, W, d = tiny_ckpt
    for f in FW.HF_FILES:
        (tmp_path / f).write_bytes(b"not a checkpoint")
    with pytest.raises(RuntimeError, match="could not be read"):
        FW.load_checkpoint(tmp_path, spec)


def test_node_never_runs_random_weights_silently(tmp_path, monkeypatch):
    """No checkpoint + no explicit switch -> the reference's RuntimeError; the switch warns loudly."""
    monkeypatch.delenv("EGREGORA_FLASHSR_RANDOM_INIT", raising=False)
    monkeypatch.setenv("EGREGORA_FLASHSR_WEIGHTS", str(tmp_path))        # an empty directory
    with pytest.raises(RuntimeError, match="FlashSR weights missing"):
        FW.weights_for_node(spec=M.tiny_spec())
    monkeypatch.delenv("EGREGORA_FLASHSR_WEIGHTS")
    monkeypatch.setattr(FW, "models_dir", lambda: tmp_path)
    with pytest.raises(RuntimeError, match="FlashSR weights missing"):
        FW.weights_for_node(spec=M.tiny_spec())
    monkeypatch.setenv("EGREGORA_FLASHSR_RANDOM_INIT", "1")
    with pytest.warns(RuntimeWarning, match="RANDOM"):
        W, tag = FW.weights_for_node(spec=M.tiny_spec())
    assert tag == "random:0" and list(W) == list(M.param_shapes(M.tiny_spec()))


def test_resolution_order(tmp_path, monkeypatch):
    monkeypatch.setenv("EGREGORA_FLASHSR_WEIGHTS", "/somewhere/else")
    assert str(FW.resolve_ckpt_dir(str(tmp_path))) == str(tmp_path)              # flashsr_min --ckpt-dir wins
    assert str(FW.resolve_ckpt_dir()) == "/somewhere/else"
    monkeypatch.delenv("EGREGORA_FLASHSR_WEIGHTS")
    assert FW.resolve_ckpt_dir().parts[-3:] == ("models", "audio", "flashsr")     # reference :27-35, :265


def test_explicit_dir_with_partial_files_raises_even_with_random_switch(tiny_ckpt, tmp_path, monkeypatch):
    spec, W, d = tiny_ckpt
    (tmp_path / "vae.pth").write_bytes((d / "vae.pth").read_bytes())
    monkeypatch.setenv("EGREGORA_FLASHSR_RANDOM_INIT", "1")
    with pytest.raises(RuntimeError, match="FlashSR weights missing"):
        FW.weights_for_node(str(tmp_path), spec)


def test_flashsr_min_cli_surface_and_script_import(tmp_path):
    """The CLI keeps the reference's flags (flashsr_min.py:6-11), requires --ckpt-dir, and imports both as a module and as
    a script (how the reference file is run)."""
    import subprocess
    p = ROOT / "comfyui-egregora-audio-super-resolution_b200" / "flashsr_min.py"
    r = subprocess.run([sys.executable, str(p), "--in", "a.wav", "--out", "b.wav"], capture_output=True, text=True)
    assert r.returncode == 2 and "--ckpt-dir" in r.stderr
    r = subprocess.run([sys.executable, str(p), "--help"], capture_output=True, text=True)
    for flag in ("--ckpt-dir", "--in", "--out", "--target-sr", "--device"):
        assert flag in r.stdout
    # script mode reaches the audio reader (unreadable input -> the pack's RuntimeError, not an ImportError)
    bad = tmp_path / "x.wav"
    bad.write_bytes(b"nope")
    r = subprocess.run([sys.executable, str(p), "--ckpt-dir", str(tmp_path), "--in", str(bad), "--out", str(tmp_path / "y.wav")],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "Failed to read audio file" in r.stderr and "ImportError" not in r.stderr
