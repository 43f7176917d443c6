"""ORACLE (test infrastructure, never shipped, never on the product path) — **parity unpinned**.

Torch fp32 interpretation of the FlashSR graph (`flashsr_model.FlashSRGraph`): every op the walker emits is
evaluated here with plain torch.nn.functional calls on CPU (or eager CUDA for the "baseline only" timing).
The model body is upstream jakeoneijk/FlashSR_Inference@main (un-pinned, absent; reference call site
egregora_audio_super_resolution.py:361-369), restated from the published AudioSR / BigVGAN components it is
built from — see the header of flashsr_model.py for what is restated and why parity is unpinned.

The CUDA plan (flashsr_plan.py) interprets the SAME graph with hand-written kernels; tests compare the two on
identical weights, inputs and diffusion noise.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this module.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F


class TorchBackend:
    """Tensors are NCHW (2-D stages) or [B,C,T] (1-D stages) float32."""

    def __init__(self, spec: dict, weights: Dict[str, torch.Tensor], device="cpu", dtype=torch.float32):
        from importlib import import_module
        self.s = spec
        self.dev = torch.device(device)
        self.dt = dtype
        self.W = {k: v.to(self.dev, dtype) for k, v in weights.items()}
        import sys
        M = sys.modules.get("egregora_b200.flashsr_model")
        if M is None:  # bench.py / smoke import path
            raise RuntimeError("load the package as `egregora_b200` before building the oracle backend")
        self.M = M
        m = spec["mel"]
        self.window = torch.from_numpy(M.hann_periodic(m["win"])).to(self.dev, dtype)
        self.mel_basis = torch.from_numpy(M.mel_filterbank(spec["sr"], m["n_fft"], m["n_mels"], m["fmin"], m["fmax"])).to(self.dev, dtype)
        k = spec["vocoder"]["aa_kernel"]
        self.aa_filter = torch.from_numpy(M.kaiser_sinc_filter1d(0.25, 0.3, k)).to(self.dev, dtype)
        self.trace = None  # optional dict name -> tensor for layer-wise comparisons

    def _rec(self, name, y):
        if self.trace is not None:
            self.trace[name] = y.detach()
        return y

    # ------------------------------------------------------------------ 2-D
    def conv2d(self, x, name, cin, cout, k, stride=1, pad="same", add=None, rowbias=None, act=None, out=None,
               transposed=False):
        w, b = self.W[name + ".weight"], self.W.get(name + ".bias")
        if stride == 2 and pad == "ldm_down":
            x = F.pad(x, (0, 1, 0, 1))
            y = F.conv2d(x, w, b, stride=2)
        else:
            y = F.conv2d(x, w, b, stride=stride, padding=k // 2)
        if rowbias is not None:
            y = y + rowbias.reshape(rowbias.shape[0], -1, 1, 1)
        if add is not None:
            y = y + add
        return self._rec(name, y)

    def groupnorm(self, x, name, c, groups, eps, silu=False):
        y = F.group_norm(x, groups, self.W[name + ".weight"], self.W[name + ".bias"], eps)
        return self._rec(name, F.silu(y) if silu else y)

    def upsample2x(self, x):
        return F.interpolate(x, scale_factor=2.0, mode="nearest")

    def concat(self, a, b):
        return torch.cat([a, b], dim=1)

    def slice_channels(self, x, lo, n):
        return x[:, lo:lo + n]

    def axpby(self, x, y, a, b):
        return a * x + b * y

    # tokens: the walker hands 2-D feature maps to linear/layernorm inside transformers
    def _tokens(self, x):
        if x.dim() == 4:
            B, C, H, W = x.shape
            return x.permute(0, 2, 3, 1).reshape(B, H * W, C), (H, W)
        return x, None

    def _untokens(self, t, hw):
        if hw is None:
            return t
        B, S, C = t.shape
        return t.reshape(B, hw[0], hw[1], C).permute(0, 3, 1, 2)

    def linear(self, x, name, cin, cout, bias=True, small=False, act=None, out=None, add=None):
        w, b = self.W[name + ".weight"], self.W.get(name + ".bias") if bias else None
        if small:
            y = F.linear(x, w, b)
        else:
            t, hw = self._tokens(x)
            y = self._untokens(F.linear(t, w, b), hw)
        if act == "silu":
            y = F.silu(y)
        if add is not None:
            y = y + add
        return self._rec(name, y)

    def layernorm(self, x, name, c, eps):
        t, hw = self._tokens(x)
        y = F.layer_norm(t, (c,), self.W[name + ".weight"], self.W[name + ".bias"], eps)
        return self._rec(name, self._untokens(y, hw))

    def geglu(self, x, inner):
        t, hw = self._tokens(x)
        a, g = t[..., :inner], t[..., inner:]
        return self._untokens(a * F.gelu(g), hw)

    def attention(self, q, k, v, heads, head_dim, v_transposed=False):
        tq, hw = self._tokens(q)
        tk, _ = self._tokens(k)
        tv, _ = self._tokens(v)
        B, S, C = tq.shape
        def split(t):
            return t.reshape(B, S, heads, head_dim).permute(0, 2, 1, 3)
        s = torch.matmul(split(tq), split(tk).transpose(-1, -2)) * (head_dim ** -0.5)
        o = torch.matmul(torch.softmax(s, dim=-1), split(tv))
        return self._untokens(o.permute(0, 2, 1, 3).reshape(B, S, C), hw)

    def time_embedding(self, t_value, dim):
        half = dim // 2
        freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half).to(self.dev)
        args = float(t_value) * freqs
        return torch.cat([torch.cos(args), torch.sin(args)])[None].to(self.dt)

    # ------------------------------------------------------------------ 1-D
    def upsample_conv2d(self, x, name, cin, cout):
        """ldm Upsample: nearest 2x then conv3x3 (the CUDA plan evaluates the same thing as four 2x2-tap phases)."""
        return self.conv2d(self.upsample2x(x), name, cin, cout, 3)

    def conv1d(self, x, name, cin, cout, k, dilation=1, add=None, act=None, add2=None, post=1.0):
        y = F.conv1d(x, self.W[name + ".weight"], self.W[name + ".bias"], padding=dilation * (k // 2), dilation=dilation)
        if add is not None:
            y = y + add
        if add2 is not None:   # running sum of the parallel residual blocks (BigVGAN: xs += resblock(x); x = xs / nk)
            y = y + add2
        if post != 1.0:
            y = y * post
        if act == "tanh":
            y = torch.tanh(y)
        return self._rec(name, y)

    def conv1d_strided(self, x, name, cin, cout, k, stride):
        return self._rec(name, F.conv1d(x, self.W[name + ".weight"], self.W[name + ".bias"], stride=stride,
                                        padding=(stride + 1) // 2))

    def convT1d(self, x, name, cin, cout, k, stride, add=None):
        y = F.conv_transpose1d(x, self.W[name + ".weight"], self.W[name + ".bias"], stride=stride,
                               padding=(k - stride) // 2)
        y = y[..., : x.shape[-1] * stride]
        if add is not None:
            y = y + add
        return self._rec(name, y)

    def snake_aa(self, x, name, c):
        """BigVGAN Activation1d(SnakeBeta, logscale): up 2x (kaiser-sinc, replicate pad) -> snake -> down 2x."""
        K = self.aa_filter.numel()
        ratio = 2
        B, C, T = x.shape
        f = self.aa_filter.view(1, 1, K).expand(C, 1, K)
        pad = K // ratio - 1
        pad_l = pad * ratio + (K - ratio) // 2
        pad_r = pad * ratio + (K - ratio + 1) // 2
        u = F.pad(x, (pad, pad), mode="replicate")
        u = ratio * F.conv_transpose1d(u, f, stride=ratio, groups=C)
        u = u[..., pad_l:-pad_r]
        alpha = torch.exp(self.W[name + ".act.alpha"]).view(1, C, 1)
        beta = torch.exp(self.W[name + ".act.beta"]).view(1, C, 1)
        u = u + (1.0 / (beta + 1e-9)) * torch.sin(u * alpha) ** 2
        even = K % 2 == 0
        u = F.pad(u, (K // 2 - int(even), K // 2), mode="replicate")
        return self._rec(name, F.conv1d(u, f, stride=ratio, groups=C))

    def add(self, a, b):
        return a + b

    def scale(self, a, s):
        return a * s

    def mean_blocks(self, ys):
        xs = ys[0]
        for y in ys[1:]:
            xs = xs + y
        return xs * (1.0 / len(ys))

    # ------------------------------------------------------------------ front end
    def stft_mag(self, wav):
        m = self.s["mel"]
        p = (m["n_fft"] - m["hop"]) // 2
        y = F.pad(wav[:, None, :], (p, p), mode="reflect")[:, 0]
        st = torch.stft(y, m["n_fft"], hop_length=m["hop"], win_length=m["win"], window=self.window, center=False,
                        return_complex=True)
        return torch.sqrt(st.real ** 2 + st.imag ** 2 + m["mag_eps"])  # [B, n_freq, frames]

    def stft_mel(self, wav):
        m = self.s["mel"]
        mel = torch.matmul(self.mel_basis, self.stft_mag(wav))
        logmel = torch.log(torch.clamp(mel, min=m["log_clamp"]))
        return self._rec("mel_lr", logmel.permute(0, 2, 1)[:, None])  # [B,1,T,F]

    def lowpass(self, wav):
        from scipy.signal import sosfiltfilt
        lp = self.s["lowpass"]
        tab = self.M.lowpass_sos_table(self.s)
        mag = self.stft_mag(wav).double()
        energy = torch.cumsum(mag.sum(dim=2), dim=1)  # [B, n_freq]
        thr = energy[:, -1:] * lp["energy_percentile"]
        idx = torch.clamp((energy < thr).sum(dim=1) - 1, min=0).cpu().numpy()
        x = wav.detach().cpu().double().numpy()
        y = np.stack([sosfiltfilt(tab[int(idx[b])], x[b]) for b in range(x.shape[0])])
        self.cutoff_bins = idx
        return self._rec("lowpass", torch.from_numpy(y).to(self.dev, self.dt))

    def mel_as_sequence(self, mel):      # [B,1,T,F] -> [B,F,T]
        return mel[:, 0].permute(0, 2, 1)

    def wav_as_sequence(self, wav):      # [B,T] -> [B,1,T]
        return wav[:, None, :]

    def sequence_as_wav(self, y):        # [B,1,T] -> [B,T]
        return y[:, 0]


def run_flashsr(spec, weights, wav: torch.Tensor, noise: torch.Tensor, steps=1, lowpass=False, device="cpu",
                trace=None, dtype=torch.float32, backend=None):
    """wav [B,chunk] f32, noise [B,z,T/8,F/8] (NCHW) -> [B,chunk].  `backend`: a TorchBackend to reuse (its weights are
    already on its device — bench.py's GPU-eager baseline builds it once, as an eager user would)."""
    import sys
    M = sys.modules["egregora_b200.flashsr_model"]
    be = backend if backend is not None else TorchBackend(spec, weights, device, dtype)
    be.trace = trace
    with torch.inference_mode():
        y = M.FlashSRGraph(spec).forward(be, wav.to(be.dev, dtype), noise.to(be.dev, dtype), steps=steps, lowpass=lowpass)
    return y.float().cpu(), be
