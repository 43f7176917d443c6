"""CLI twin of the reference's flashsr_min.py:5-24 (same flags), routed through the real B200 path.

The reference file is an identity stub (mono-mix, pad 64, truncate, write at --target-sr); north_star names
it as the FlashSR entry, so this CLI keeps its argument surface and prints "OK", but runs the node's
chunked upsampler.  WAV I/O uses the stdlib `wave` module (PCM-16), soundfile is not required.
"""
import argparse
import wave

import numpy as np
import torch


def _read_wav(path):
    with wave.open(path, "rb") as w:
        sr, ch, sw, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
        raw = w.readframes(n)
    if sw != 2:
        raise RuntimeError(f"only PCM-16 WAV is supported by this CLI (sample width {sw})")
    x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    return x.reshape(-1, ch), sr


def _write_wav(path, x, sr):
    y = np.clip(np.rint(np.asarray(x, np.float64) * 32768.0), -32768, 32767).astype("<i2")
    with wave.open(path, "wb") as w:
        w.setnchannels(1 if y.ndim == 1 else y.shape[1])
        w.setsampwidth(2)
        w.setframerate(int(sr))
        w.writeframes(y.tobytes())


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--ckpt-dir", required=True)
    ap.add_argument("--in", dest="inp", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--target-sr", type=int, default=48000)
    ap.add_argument("--device", default="auto")
    args = ap.parse_args(argv)

    from .egregora_audio_super_resolution import EgregoraAudioSuperResolution
    wav, sr = _read_wav(args.inp)
    mono = wav.mean(axis=1) if wav.ndim == 2 else wav  # reference mono-mixes, flashsr_min.py:15-18
    audio = {"waveform": torch.from_numpy(mono.astype(np.float32))[None, None, :], "sample_rate": int(sr)}
    sr_choice = str(args.target_sr) if str(args.target_sr) in ("48000", "44100", "96000") else "48000"
    (res,) = EgregoraAudioSuperResolution().run(audio=audio, lowpass_input=False, output_sr=sr_choice)
    _write_wav(args.out, res["waveform"][0, 0].numpy(), args.target_sr)
    print("OK")


if __name__ == "__main__":
    main()
