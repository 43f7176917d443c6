"""CLI twin of the reference's flashsr_min.py:5-24 (same flags), routed through the real B200 path.

The reference file is an identity stub (mono-mix, pad 64, truncate, write at --target-sr); north_star names
it as the FlashSR entry, so this CLI keeps its argument surface and prints "OK", but runs the node's chunked
upsampler with the checkpoint files found in --ckpt-dir (student_ldm.pth, sr_vocoder.pth, vae.pth — the reference
runner's files, egregora_audio_super_resolution.py:260-261).  Audio I/O uses soundfile when it is importable (what
the reference uses: any libsndfile format) and the stdlib `wave` module (PCM WAV) otherwise; decode failures are
RuntimeErrors, the pack's only error type.  Runs as a module (`python -m <pkg>.flashsr_min`) or as a script
(`python flashsr_min.py`, how the reference file is run).
"""
import argparse
import importlib.util
import sys
import wave
from pathlib import Path

import numpy as np
import torch


def _read_audio(path):
    """-> ([S, C] float32, sr)."""
    try:
        import soundfile as sf  # type: ignore
    except Exception:
        sf = None
    if sf is not None:
        try:
            x, sr = sf.read(path, dtype="float32", always_2d=True)
            return np.asarray(x, np.float32), int(sr)
        except Exception as e:
            raise RuntimeError(f"Failed to read audio file {path}: {e}") from e
    try:
        with wave.open(str(path), "rb") as w:
            sr, ch, sw, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
            raw = w.readframes(n)
    except (wave.Error, EOFError, OSError) as e:
        raise RuntimeError(f"Failed to read audio file {path}: {e} (without the soundfile package only PCM WAV "
                           "files can be decoded)") from e
    if sw == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif sw == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif sw == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        x = np.where(v >= 1 << 23, v - (1 << 24), v).astype(np.float32) / 8388608.0
    elif sw == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise RuntimeError(f"Failed to read audio file {path}: unsupported PCM sample width {sw}")
    return x.reshape(-1, ch), sr


def _write_audio(path, x, sr):
    try:
        import soundfile as sf  # type: ignore
        sf.write(path, np.asarray(x, np.float32), int(sr))   # default subtype for .wav: PCM_16, as in the reference
        return
    except ImportError:
        pass
    y = np.clip(np.rint(np.asarray(x, np.float64) * 32768.0), -32768, 32767).astype("<i2")
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1 if y.ndim == 1 else y.shape[1])
        w.setsampwidth(2)
        w.setframerate(int(sr))
        w.writeframes(y.tobytes())


def _node_module():
    if __package__:
        from . import egregora_audio_super_resolution as N
        return N
    # run as a script: load the package this file lives in under a private name
    pkg_dir = Path(__file__).resolve().parent
    name = "egregora_b200"
    if name not in sys.modules:
        spec = importlib.util.spec_from_file_location(name, pkg_dir / "__init__.py", submodule_search_locations=[str(pkg_dir)])
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    return importlib.import_module(name + ".egregora_audio_super_resolution")


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--ckpt-dir", required=True)
    ap.add_argument("--in", dest="inp", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--target-sr", type=int, default=48000)
    ap.add_argument("--device", default="auto")
    args = ap.parse_args(argv)

    N = _node_module()
    wav, sr = _read_audio(args.inp)
    mono = wav.mean(axis=1) if wav.ndim == 2 else wav  # reference mono-mixes, flashsr_min.py:15-18
    audio = {"waveform": torch.from_numpy(np.ascontiguousarray(mono, np.float32))[None, None, :], "sample_rate": int(sr)}
    node = N.EgregoraAudioSuperResolution()
    node.CKPT_DIR = args.ckpt_dir
    if str(args.target_sr) in ("48000", "44100", "96000"):
        (res,) = node.run(audio=audio, lowpass_input=False, output_sr=str(args.target_sr))
        y = res["waveform"][0, 0]
    else:   # a rate the node's combo box does not offer: upsample to 48 kHz, then the node's own resampler
        (res,) = node.run(audio=audio, lowpass_input=False, output_sr="48000")
        y = N._resample_hq(res["waveform"][0], 48000, int(args.target_sr))[0].cpu()
    _write_audio(args.out, y.numpy(), args.target_sr)
    print("OK")


if __name__ == "__main__":
    main()
