"""ORACLE (test infrastructure, never shipped, never on the product path) — **parity unpinned**.

CPU restatement of the Fat-Llama spectral-enhance path that the reference reaches through ONE call,
`feed.upscale(...)` (egregora_fat_llama_gpu.py:213-224, egregora_fat_llama_cpu.py:126-134).  The arithmetic
lives in third-party PyPI packages that are absent from /root/reference and cannot be installed here:

    fat-llama      >= 1.1.0    (requirements.txt:10)  fat_llama.audio_fattener.feed       (CuPy / cuFFT)
    fat-llama-fftw >= 1.0.4.4  (requirements.txt:11)  fat_llama_fftw.audio_fattener.feed  (pyFFTW)

Both are version FLOORS, not pins, and the reference holds no test or golden vector for this path, so this
file restates the packages' published algorithm (README "iterative soft thresholding" description and
function names) and anchors everything it can on the reference's own call sites:

  * read side   — reference writes a PCM-16 WAV (sf.write default subtype, :34-37) and upstream reads it back as
                  integer-valued samples (the write-side patch divides by 2**(8*sample_width-1), :195-200,
                  which only makes sense for integer-scaled data);
  * upscale()   — keyword arguments exactly as bound at :213-224 (GPU) / :126-134 (CPU, toggles default True);
  * write side  — patched write_audio (:188-208), then PCM-16 file, then sf.read(float32) (:291) and _to_cs (:292).

Restated upstream algorithm (per channel, float32 on integer-scaled samples):
    expanded = repeat each sample `U` times            (new_interpolation_algorithm)
    x0       = where(|expanded| > thr, expanded, 0)     (initialize_ist)
    repeat max_iterations times:                        (iterative_soft_thresholding)
        X = fft(x); X = where(|X| > thr, X, 0); x = ifft(X).real
    y = expanded + x
    autoscale : y = (y / max|y|) * max|channel_in|      (scale_amplitude, per channel)
    normalize : y /= max|y| over all channels           (normalize_signal)
    U = max(1, round(target_bitrate / source_bitrate)),  output sample rate = sr * U.
Anything here that later proves to differ from the real packages is a one-line change in both this file and
the CUDA host code; the kernels are agnostic (gate threshold, U, flags are arguments).
"""
from __future__ import annotations

import numpy as np

try:  # scipy.fft keeps float32 transforms in float32 (like cuFFT / FFTW single), numpy would promote
    from scipy import fft as _fft
except Exception:  # pragma: no cover
    _fft = None


def pcm16_write(x: np.ndarray) -> np.ndarray:
    """float -> int16 exactly as libsndfile does for float input to a PCM_16 file with clipping enabled
    (soundfile's behaviour for sf.write(path, float_array, sr)): scale by 0x8000, clip, lrint."""
    s = np.asarray(x, np.float32) * np.float32(32768.0)
    s = np.clip(s, -32768.0, 32767.0)
    return np.rint(s).astype(np.int16)


def pcm16_read(q: np.ndarray) -> np.ndarray:
    """sf.read(dtype='float32') of a PCM_16 file."""
    return (q.astype(np.float32) / np.float32(32768.0)).astype(np.float32)


def upscale_factor(sample_rate: int, channels: int, target_bitrate_kbps: int, sample_width: int = 2) -> int:
    src = sample_rate * channels * 8 * sample_width
    return max(1, int(round(target_bitrate_kbps * 1000.0 / src)))


def ist(expanded: np.ndarray, max_iter: int, thr: float, dtype=np.float32) -> np.ndarray:
    cdt = np.complex64 if dtype == np.float32 else np.complex128
    x = np.where(np.abs(expanded) > thr, expanded, 0).astype(dtype)
    for _ in range(max_iter):
        X = _fft.fft(x.astype(cdt))
        X = np.where(np.abs(X) > thr, X, 0).astype(cdt)
        x = _fft.ifft(X).real.astype(dtype)
    return x


def upscale_channels(samples_sc: np.ndarray, U: int, max_iter: int, thr: float, dtype=np.float32) -> np.ndarray:
    """samples_sc [S,C] integer-scaled -> [S*U, C]."""
    outs = []
    for ch in samples_sc.T:
        expanded = np.repeat(ch.astype(dtype), U)
        outs.append(expanded + ist(expanded, max_iter, thr, dtype))
    return np.stack(outs, axis=1)


def upscale(samples_sc: np.ndarray, U: int, max_iterations: int, threshold_value: float,
            toggle_normalize: bool = True, toggle_autoscale: bool = True, dtype=np.float32) -> np.ndarray:
    """The arithmetic of feed.upscale between read_audio and write_audio.  [S,C] -> [S*U,C]."""
    x = np.asarray(samples_sc, dtype)
    if x.ndim == 1:
        x = x[:, None]
    y = upscale_channels(x, U, max_iterations, threshold_value, dtype)
    if toggle_autoscale:                                   # scale_amplitude, per channel
        cols = []
        for c in range(x.shape[1]):
            normalized = (y[:, c] / np.max(np.abs(y[:, c]))).astype(dtype)
            cols.append((normalized * np.max(np.abs(x[:, c]))).astype(dtype))
        y = np.stack(cols, axis=1).astype(dtype)
    if toggle_normalize:
        y = (y / np.max(np.abs(y))).astype(dtype)
    return y


def node_run(cs: np.ndarray, sr: int, max_iterations: int, threshold_value: float, target_bitrate_kbps: int,
             toggle_normalize: bool = True, toggle_autoscale: bool = True, dtype=np.float32,
             return_prequant: bool = False):
    """EgregoraFatLlamaGPU.run for the AUDIO-dict branch (egregora_fat_llama_gpu.py:257-294):
    cs [C,S] float32 in [-1,1] -> ([C,S*U] float32, sr*U)."""
    q_in = pcm16_write(cs.T)                                   # _save_temp_wav (:34-37): [S,C] PCM-16
    samples = q_in.astype(np.float32)                          # upstream read_audio: integer-scaled
    sample_width = 2
    U = upscale_factor(sr, cs.shape[0], target_bitrate_kbps, sample_width)
    y = upscale(samples, U, max_iterations, threshold_value, toggle_normalize, toggle_autoscale, dtype)
    m = float(np.max(np.abs(y))) if y.size else 0.0            # patched write_audio (:188-208)
    if m > 1.0:
        y = y / float(2 ** (8 * sample_width - 1))
    y = y.astype(np.float32)
    out = pcm16_read(pcm16_write(y))                           # file write + sf.read(float32) (:291)
    out_cs = np.ascontiguousarray(out.T)                       # _to_cs (:292): [S,C] -> [C,S]; peak <= 1
    if return_prequant:
        return out_cs, sr * U, np.ascontiguousarray(y.T)
    return out_cs, sr * U
