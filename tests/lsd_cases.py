"""Shared by the LSD tests: the seeded signal generator of tests/golden/make_eval_lsd_golden.py and a numpy float32
emulation of the arithmetic of eval_lsd_frames_kernel / eval_lsd_final_kernel (csrc/eval_metrics.cu), thread loops
vectorised.  The emulation runs on the CPU and pins the kernel's index math (packed-real Stockham radix-2, Hermitian
split, log, per-frame RMS) against the reference goldens without a GPU; it is test infrastructure only."""
import numpy as np

f32 = np.float32


def signals(name, c):
    rng = np.random.default_rng(sum(map(ord, name)))
    N = max(c["Na"], c["Nb"])
    t = np.arange(N) / 48000.0
    base = (0.3 * np.sin(2 * np.pi * 440 * t) + 0.1 * rng.standard_normal((c["C"], N))).astype(f32)
    proc = (c["gain"] * base + c["noise"] * rng.standard_normal((c["C"], N))).astype(f32)
    if c["band"] < 1.0:
        def cut(x):
            X = np.fft.rfft(x.astype(np.float64))
            X[..., int(X.shape[-1] * c["band"]):] = 0
            return np.fft.irfft(X, n=x.shape[-1]).astype(f32)
        base, proc = cut(base), cut(proc)
    return base[:, :c["Na"]].copy(), proc[:, :c["Nb"]].copy()


def _tables(n_fft):
    i = np.arange(n_fft)
    window = (0.5 + 0.5 * np.cos(np.pi * (1 - n_fft + 2 * i) / (n_fft - 1))).astype(f32)   # eval_lsd_tables_kernel
    tw = np.exp(-2j * np.pi * np.arange(n_fft // 2 + 1) / n_fft).astype(np.complex64)
    return window, tw


def _frame_logmag(mono, start, n_fft, window, tw):
    M = n_fft // 2
    fr = np.zeros(n_fft, f32)
    seg = mono[start:start + n_fft]
    fr[:seg.size] = seg
    v = fr * window
    src = (v[0::2] + 1j * v[1::2]).astype(np.complex64)
    Ns = 1
    while Ns < M:
        tstep = M // (2 * Ns)
        j = np.arange(M // 2)
        k = j & (Ns - 1)
        a, bb = src[j], (src[j + M // 2] * tw[2 * k * tstep]).astype(np.complex64)
        j0 = ((j - k) << 1) + k
        dst = np.empty(M, np.complex64)
        dst[j0], dst[j0 + Ns] = a + bb, a - bb
        src, Ns = dst, Ns << 1
    k = np.arange(M + 1)
    zk, zm = src[np.where(k == M, 0, k)], src[np.where(k == 0, 0, M - k)]
    e = (f32(0.5) * (zk.real + zm.real) + 1j * (f32(0.5) * (zk.imag - zm.imag))).astype(np.complex64)
    o = (f32(0.5) * (zk.real - zm.real) + 1j * (f32(0.5) * (zk.imag + zm.imag))).astype(np.complex64)
    wo = (tw[k] * o).astype(np.complex64)
    re, im = e.real + wo.imag, e.imag - wo.real
    return (f32(20) * np.log10(np.sqrt(re * re + im * im).astype(f32) + f32(1e-12))).astype(f32)


def emulate_kernel_lsd(A, B, n_fft, hop):
    a = A.mean(axis=0) if A.ndim > 1 else A
    b = B.mean(axis=0) if B.ndim > 1 else B
    n = min(a.size, b.size)
    a, b = a[:n], b[:n]
    window, tw = _tables(n_fft)
    frames = 1 + max(0, (n - n_fft) // hop)
    per = np.empty(frames, f32)
    for f in range(frames):
        d = _frame_logmag(a, f * hop, n_fft, window, tw) - _frame_logmag(b, f * hop, n_fft, window, tw)
        per[f] = np.sqrt(f32((d * d).astype(np.float64).sum() / (n_fft // 2 + 1)) + f32(1e-12))
    return float(f32(per.astype(np.float64).sum() / frames)), float(np.percentile(per, 95)), per
