"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement in numpy of the reference's path-A *driver* arithmetic — everything around the model
call in /root/reference/egregora_audio_super_resolution.py — and of the path-B input coercion.
Pinned: tests/test_oracle_golden.py checks every function here against vectors produced by importing
the reference itself in the build container (tests/golden/make_golden.py -> tests/golden/driver_golden.npz).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

REQ_SR = 48000            # egregora_audio_super_resolution.py:255
CHUNK_S = 5.12            # :256
OVERLAP_S = 0.50          # :257
CHUNK_SAMPLES = int(REQ_SR * CHUNK_S)  # :258  -> 245760


def win_hop(req_sr: int = REQ_SR) -> Tuple[int, int]:
    """egregora_audio_super_resolution.py:400-404."""
    win = CHUNK_SAMPLES
    hop = int((CHUNK_S - OVERLAP_S) * req_sr)
    if hop <= 0 or hop >= win:
        hop = win // 2
    return win, hop


def iter_chunks(total_samples: int, win: int, hop: int) -> List[Tuple[int, int]]:
    """egregora_audio_super_resolution.py:213-225."""
    spans = []
    i = 0
    while i < total_samples:
        L = min(win, total_samples - i)
        spans.append((i, L))
        if i + L >= total_samples:
            break
        i += hop
    return spans


def hann(L: int) -> np.ndarray:
    """egregora_audio_super_resolution.py:210-211 — symmetric Hann, computed in f64, cast to f32."""
    return np.hanning(L).astype(np.float32)


def gather_chunks(in_cs: np.ndarray, spans, win: int) -> np.ndarray:
    """egregora_audio_super_resolution.py:411-416: slice [C,L], right zero-pad to win. -> [n,C,win]."""
    C = in_cs.shape[0]
    out = np.zeros((len(spans), C, win), np.float32)
    for k, (s, L) in enumerate(spans):
        out[k, :, :L] = in_cs[:, s:s + L]
    return out


def wola_stitch(chunks_pred, total_len: int, win: int) -> np.ndarray:
    """egregora_audio_super_resolution.py:227-251.  chunks_pred: list of (pred [C,L_pred], start, L_in)."""
    if not chunks_pred:
        return np.zeros((1, max(1, total_len)), np.float32)
    C = chunks_pred[0][0].shape[0]
    acc = np.zeros((C, total_len), np.float32)
    wsum = np.zeros(total_len, np.float32)
    w_full = hann(win)
    for y_cs, start, L_in in chunks_pred:
        L = min(L_in, y_cs.shape[1])
        w = w_full[:L] if L <= win else np.ones(L, np.float32)
        acc[:, start:start + L] += y_cs[:, :L] * w[None, :]
        wsum[start:start + L] += w
    wsum[wsum == 0] = 1.0
    return (acc / wsum[None, :]).astype(np.float32)


def from_audio_array(arr) -> np.ndarray:
    """(array, sr) branch of _from_audio_dict, egregora_audio_super_resolution.py:140-155."""
    arr = np.asarray(arr, dtype=np.float32)
    if arr.ndim == 1:
        cs = arr[None, :]
    elif arr.ndim == 2:
        if arr.shape[0] >= arr.shape[1] and arr.shape[1] <= 8:
            cs = arr.T
        else:
            cs = arr
    else:
        cs = arr.reshape(1, -1)
    return cs.astype(np.float32)


def to_cs(x) -> np.ndarray:
    """_to_cs, egregora_fat_llama_gpu.py:18-32 (identical in egregora_fat_llama_cpu.py:12-26)."""
    a = np.asarray(x, dtype=np.float32)
    if a.ndim == 1:
        a = a[None, :]
    elif a.ndim == 2:
        h, w = a.shape
        if w <= 8 and h > w:
            a = a.T
    else:
        a = a.reshape(-1)[None, :]
    m = float(np.max(np.abs(a))) if a.size else 0.0
    if m > 1.0:
        a = a / (m + 1e-8)
    return a.astype(np.float32)


def resample_poly_ref(x_cs: np.ndarray, src_sr: int, dst_sr: int) -> np.ndarray:
    """scipy branch of _resample_hq, egregora_audio_super_resolution.py:181-191 (soxr is absent here)."""
    if src_sr == dst_sr:
        return x_cs.astype(np.float32)
    from math import gcd
    from scipy.signal import resample_poly
    g = gcd(src_sr, dst_sr)
    up, down = dst_sr // g, src_sr // g
    out = [resample_poly(x_cs[c], up=up, down=down).astype(np.float32) for c in range(x_cs.shape[0])]
    L = min(map(len, out))
    return np.stack([ch[:L] for ch in out], axis=0)


def run_driver(in_cs: np.ndarray, in_sr: int, chunk_model, output_sr: int = 48000) -> Tuple[np.ndarray, int]:
    """EgregoraAudioSuperResolution.run, egregora_audio_super_resolution.py:388-431, with the model call
    (`runner.infer`, :417) replaced by `chunk_model([C,win]) -> [C,L_pred]`."""
    if in_sr != REQ_SR:
        in_cs = resample_poly_ref(in_cs, in_sr, REQ_SR)
        in_sr = REQ_SR
    win, hop = win_hop()
    total = in_cs.shape[1]
    spans = iter_chunks(total, win, hop)
    preds = []
    for start, L in spans:
        chunk = in_cs[:, start:start + L]
        if L < win:
            chunk = np.concatenate([chunk, np.zeros((in_cs.shape[0], win - L), np.float32)], axis=1)
        preds.append((chunk_model(chunk), start, L))
    out = wola_stitch(preds, total, win)
    if output_sr != in_sr:
        return resample_poly_ref(out, in_sr, output_sr), output_sr
    return out, in_sr
