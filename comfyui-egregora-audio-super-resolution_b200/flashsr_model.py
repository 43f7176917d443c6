"""FlashSR model definition as data: hyper-parameters (`default_spec`) + a backend-agnostic graph walker.

The reference pack does not contain the model: it imports `FlashSR.FlashSR.FlashSR` from an un-pinned
download of github.com/jakeoneijk/FlashSR_Inference (egregora_audio_super_resolution.py:65-68, :323) and calls
`model(x[B,245760], lowpass_input=bool) -> [B,245760]` (:366-369).  That source and its three weight files
are absent here, so this file RESTATES the published architecture FlashSR is assembled from — **parity
unpinned** (SURVEY.md §0.3, App. A-1):

  * mel front-end + AutoencoderKL + latent UNet: AudioSR (Liu et al. 2023) `basic` config — 48 kHz, n_fft 2048,
    hop 480, 256 mels 20-24000 Hz; VAE ch 128, ch_mult (1,2,4,8), 2 res-blocks, z 16, mid attention;
    UNet in 32 (noise ⊕ LR latent), out 16, model_channels 128, channel_mult (1,2,3,5), self-attention
    transformers at ds 2/4/8 with 32-dim heads;
  * one-/few-step sampling: v-prediction on a cosine schedule, deterministic (DDIM, eta 0) updates,
    x_T supplied by the caller so runs are reproducible;
  * SR vocoder: BigVGAN-style generator (anti-aliased SnakeBeta AMP blocks, transposed-conv upsampling x480)
    conditioned on the decoded mel AND on the low-resolution waveform through a strided-conv encoder whose
    features are added at each resolution;
  * optional input low-pass: roll-off detection on the STFT energy + order-8 Chebyshev-I zero-phase filter.

Every width/depth/rate is a spec entry, nothing is hard-coded in kernels.  The walker below describes the
computation against an abstract backend `be`; two backends exist:
    oracle/flashsr_oracle.py : torch fp32 (the parity checker — test infrastructure only)
    flashsr_plan.py          : emits the egr_op list executed by libegregora_b200.so (the product)
Weight names follow the upstream ldm / BigVGAN state-dict conventions so real checkpoints can be mapped.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Tuple

import numpy as np
import torch


def default_spec() -> dict:
    return {
        "sr": 48000,
        "chunk": 245760,
        "mel": {"n_fft": 2048, "hop": 480, "win": 2048, "n_mels": 256, "fmin": 20.0, "fmax": 24000.0,
                "mag_eps": 1e-9, "log_clamp": 1e-5},
        "vae": {"ch": 128, "ch_mult": [1, 2, 4, 8], "num_res_blocks": 2, "z_channels": 16, "embed_dim": 16,
                "in_channels": 1, "groups": 32, "eps": 1e-6},
        "unet": {"in_channels": 32, "out_channels": 16, "model_channels": 128, "channel_mult": [1, 2, 3, 5],
                 "num_res_blocks": 2, "attention_ds": [2, 4, 8], "head_dim": 32, "groups": 32, "eps": 1e-5,
                 "attn_norm_eps": 1e-6, "ff_mult": 4},
        "diffusion": {"T": 1000, "cosine_s": 0.008},
        "vocoder": {"num_mels": 256, "upsample_rates": [6, 5, 4, 2, 2], "upsample_kernel_sizes": [12, 10, 8, 4, 4],
                    "upsample_initial_channel": 1536, "resblock_kernel_sizes": [3, 7, 11],
                    "resblock_dilations": [1, 3, 5], "aa_kernel": 12, "wave_cond": True},
        "lowpass": {"order": 8, "ripple_db": 0.1, "energy_percentile": 0.985, "min_cutoff_hz": 1000.0},
    }


def tiny_spec() -> dict:
    """Same topology, small widths/lengths: the sizes the CPU oracle finishes in seconds (parity tests)."""
    s = default_spec()
    s["chunk"] = 3840 * 2
    s["mel"].update({"n_fft": 256, "hop": 60, "win": 256, "n_mels": 64, "fmax": 24000.0})
    s["vae"].update({"ch": 32, "ch_mult": [1, 2, 2, 4], "groups": 8, "z_channels": 8, "embed_dim": 8})
    s["unet"].update({"in_channels": 16, "out_channels": 8, "model_channels": 32, "channel_mult": [1, 2, 2, 3],
                      "head_dim": 16, "groups": 8})
    s["vocoder"].update({"num_mels": 64, "upsample_rates": [5, 3, 2, 2], "upsample_kernel_sizes": [10, 6, 4, 4],
                         "upsample_initial_channel": 128})
    return s


# ------------------------------------------------------------------------------------------------ constants
def hann_periodic(n: int) -> np.ndarray:
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)).astype(np.float64)


def mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """Slaney-scale, Slaney-normalised triangular filterbank [n_mels, n_fft//2+1] (the librosa.filters.mel
    default AudioSR uses), restated in numpy."""
    def hz_to_mel(f):
        f = np.asarray(f, np.float64)
        f_sp = 200.0 / 3
        mels = f / f_sp
        min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)

    def mel_to_hz(m):
        m = np.asarray(m, np.float64)
        f_sp = 200.0 / 3
        min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    n_freq = n_fft // 2 + 1
    fftfreqs = np.linspace(0, sr / 2.0, n_freq)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return (w * enorm[:, None]).astype(np.float32)


def kaiser_sinc_filter1d(cutoff: float, half_width: float, kernel_size: int) -> np.ndarray:
    """BigVGAN alias-free-torch filter (float32 taps)."""
    even = kernel_size % 2 == 0
    half_size = kernel_size // 2
    delta_f = 4 * half_width
    A = 2.285 * (half_size - 1) * math.pi * delta_f + 7.95
    if A > 50.0:
        beta = 0.1102 * (A - 8.7)
    elif A >= 21.0:
        beta = 0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0)
    else:
        beta = 0.0
    window = torch.kaiser_window(kernel_size, beta=beta, periodic=False, dtype=torch.float32)
    if even:
        time = torch.arange(-half_size, half_size, dtype=torch.float32) + 0.5
    else:
        time = torch.arange(kernel_size, dtype=torch.float32) - half_size
    if cutoff == 0:
        return np.zeros(kernel_size, np.float32)
    filt = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
    filt = filt / filt.sum()
    return filt.numpy().astype(np.float32)


def cosine_alphas_cumprod(T: int, s: float) -> np.ndarray:
    steps = T + 1
    x = np.linspace(0, T, steps, dtype=np.float64)
    ac = np.cos(((x / T) + s) / (1 + s) * np.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = np.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
    return np.cumprod(1.0 - betas)


def ddim_schedule(T: int, s: float, num_steps: int) -> List[Tuple[int, float, float]]:
    """[(t, alpha_bar_t, alpha_bar_prev)] for `num_steps` evenly spaced steps from t = T-1 down; the last
    step lands on x_0 (alpha_bar_prev = 1)."""
    ac = cosine_alphas_cumprod(T, s)
    ts = np.round(np.linspace(T - 1, 0, num_steps + 1)[:-1]).astype(int)
    out = []
    for i, t in enumerate(ts):
        prev = float(ac[ts[i + 1]]) if i + 1 < len(ts) else 1.0
        out.append((int(t), float(ac[t]), prev))
    return out


def lowpass_sos_table(spec: dict) -> np.ndarray:
    """One zero-phase Chebyshev-I design per possible cutoff bin, so the kernel can pick a row on device
    without a host round trip: [n_freq, n_sections, 6] float64 (+ per-row sosfilt_zi appended by the caller)."""
    from scipy.signal import cheby1
    lp, mel = spec["lowpass"], spec["mel"]
    n_freq = mel["n_fft"] // 2 + 1
    nyq = spec["sr"] / 2.0
    nsec = (lp["order"] + 1) // 2
    tab = np.zeros((n_freq, nsec, 6), np.float64)
    for b in range(n_freq):
        fc = min(max(b / (n_freq - 1) * nyq, lp["min_cutoff_hz"]), nyq * 0.999)
        tab[b] = cheby1(lp["order"], lp["ripple_db"], fc / nyq, btype="low", output="sos")
    return tab


# ------------------------------------------------------------------------------------------------ parameters
class ParamSet:
    """Ordered name -> (shape, kind) registry; `kind` drives init and blob packing."""

    def __init__(self):
        self.items: "OrderedDict[str, Tuple[tuple, str]]" = OrderedDict()

    def add(self, name, shape, kind):
        assert name not in self.items, name
        self.items[name] = (tuple(int(v) for v in shape), kind)

    def conv2d(self, name, cin, cout, k, bias=True):
        self.add(name + ".weight", (cout, cin, k, k), "conv2d")
        if bias:
            self.add(name + ".bias", (cout,), "bias")

    def conv1d(self, name, cin, cout, k, bias=True):
        self.add(name + ".weight", (cout, cin, k), "conv1d")
        if bias:
            self.add(name + ".bias", (cout,), "bias")

    def convT1d(self, name, cin, cout, k):
        self.add(name + ".weight", (cin, cout, k), "convT1d")
        self.add(name + ".bias", (cout,), "bias")

    def linear(self, name, cin, cout, bias=True):
        self.add(name + ".weight", (cout, cin), "linear")
        if bias:
            self.add(name + ".bias", (cout,), "bias")

    def norm(self, name, c):
        self.add(name + ".weight", (c,), "norm_w")
        self.add(name + ".bias", (c,), "norm_b")

    def snake(self, name, c):
        self.add(name + ".act.alpha", (c,), "snake")
        self.add(name + ".act.beta", (c,), "snake")


class _ParamBackend:
    """Backend that only records which parameters the walker touches (shapes come from the call sites)."""

    def __init__(self):
        self.P = ParamSet()

    # every op returns a dummy "tensor" (channel count is all the walker needs)
    def conv2d(self, x, name, cin, cout, k, **kw):
        self.P.conv2d(name, cin, cout, k)
        return None

    def conv1d(self, x, name, cin, cout, k, **kw):
        self.P.conv1d(name, cin, cout, k)
        return None

    def upsample_conv2d(self, x, name, cin, cout, **kw):
        self.P.conv2d(name, cin, cout, 3)
        return None

    def conv1d_strided(self, x, name, cin, cout, k, stride, **kw):
        self.P.conv1d(name, cin, cout, k)
        return None

    def convT1d(self, x, name, cin, cout, k, stride, **kw):
        self.P.convT1d(name, cin, cout, k)
        return None

    def linear(self, x, name, cin, cout, bias=True, **kw):
        self.P.linear(name, cin, cout, bias)
        return None

    def groupnorm(self, x, name, c, *a, **kw):
        self.P.norm(name, c)
        return None

    def layernorm(self, x, name, c, *a, **kw):
        self.P.norm(name, c)
        return None

    def snake_aa(self, x, name, c, **kw):
        self.P.snake(name, c)
        return None

    def __getattr__(self, item):  # concat, add, attention, geglu, upsample2x, ... carry no parameters
        return lambda *a, **kw: None


def param_shapes(spec: dict) -> "OrderedDict[str, Tuple[tuple, str]]":
    be = _ParamBackend()
    FlashSRGraph(spec).forward(be, None, None, steps=1, lowpass=False)
    return be.P.items


def init_weights(spec: dict, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded random weights of the spec'd architecture (no checkpoint exists in this environment; `data` in
    bench.py says so).  Variance-preserving fan-in init so activations stay O(1) through ~100 layers.

    Round 2: the vocoder's residual branches are scaled (convs2 x 0.35).  With unit-gain branches every AMP block
    doubled the stream's variance and the last stage ran at rms 9-13 (tools/parity_diag.py, round-2 run): the snake
    term sin^2(alpha*x)/beta then sees phases of +-10 rad, where ANY 1e-3 relative perturbation of x — f16 operand
    rounding, TF32, a different summation order — moves the phase by 1e-2 rad and the relative error of the waveform
    grew from 1.6e-3 (vocoder input) to 7.8e-3 (output): 1.14e-3 RMS against the fp32 oracle at c2, a property of that
    synthetic network, not of a trained one, whose weight-normed branches are small corrections of the stream.  With
    the scaled branches the stream stays at rms 0.6-1.0 through all five stages."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = OrderedDict()
    for name, (shape, kind) in param_shapes(spec).items():
        if kind in ("conv2d", "conv1d", "linear"):
            fan_in = int(np.prod(shape[1:]))
            w = (torch.rand(shape, generator=g) * 2 - 1) * math.sqrt(3.0 / fan_in)
        elif kind == "convT1d":
            cin, cout, k = shape
            w = (torch.rand(shape, generator=g) * 2 - 1) * math.sqrt(3.0 / (cin * 2))  # 2 taps hit each output
        elif kind == "bias":
            w = torch.randn(shape, generator=g) * 0.02
        elif kind == "norm_w":
            w = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "norm_b":
            w = 0.1 * torch.randn(shape, generator=g)
        elif kind == "snake":
            w = 0.2 * torch.randn(shape, generator=g)
        else:
            raise ValueError(kind)
        if name.startswith("vocoder.resblocks.") and ".convs2." in name and kind == "conv1d":
            w = w * 0.35  # residual-branch gain: keeps the vocoder stream O(1) (see the docstring)
        if name == "vocoder.conv_post.weight":
            w = w * 0.2   # keep the synthetic waveform out of tanh saturation (rms ~0.1, like programme audio)
        out[name] = w.float().contiguous()
    return out


def name_prefix(i: int, nk: int, j: int) -> str:
    return f"vocoder.resblocks.{i * nk + j}"


# ------------------------------------------------------------------------------------------------ the graph
class FlashSRGraph:
    def __init__(self, spec: dict):
        self.s = spec

    # ---------------------------------------------------------------- VAE (ldm AutoencoderKL)
    def _vae_res(self, be, x, name, cin, cout):
        v = self.s["vae"]
        h = be.groupnorm(x, f"{name}.norm1", cin, v["groups"], v["eps"], silu=True)
        h = be.conv2d(h, f"{name}.conv1", cin, cout, 3)
        h = be.groupnorm(h, f"{name}.norm2", cout, v["groups"], v["eps"], silu=True)
        sc = x if cin == cout else be.conv2d(x, f"{name}.nin_shortcut", cin, cout, 1)
        return be.conv2d(h, f"{name}.conv2", cout, cout, 3, add=sc)

    def _vae_attn(self, be, x, name, c):
        v = self.s["vae"]
        hn = be.groupnorm(x, f"{name}.norm", c, v["groups"], v["eps"], silu=False)
        q = be.conv2d(hn, f"{name}.q", c, c, 1, out="f16")
        k = be.conv2d(hn, f"{name}.k", c, c, 1, out="f16")
        vv = be.conv2d(hn, f"{name}.v", c, c, 1, out="f16", transposed=True)
        a = be.attention(q, k, vv, heads=1, head_dim=c, v_transposed=True)
        return be.conv2d(a, f"{name}.proj_out", c, c, 1, add=x)

    def vae_encode(self, be, mel):
        """mel [B,T,F,1] -> posterior mean [B,T/8,F/8,embed] (mode of the diagonal Gaussian)."""
        v = self.s["vae"]
        ch, mult, nrb = v["ch"], v["ch_mult"], v["num_res_blocks"]
        h = be.conv2d(mel, "vae.encoder.conv_in", v["in_channels"], ch, 3)
        cin = ch
        for i, m in enumerate(mult):
            cout = ch * m
            for j in range(nrb):
                h = self._vae_res(be, h, f"vae.encoder.down.{i}.block.{j}", cin, cout)
                cin = cout
            if i != len(mult) - 1:
                h = be.conv2d(h, f"vae.encoder.down.{i}.downsample.conv", cin, cin, 3, stride=2, pad="ldm_down")
        h = self._vae_res(be, h, "vae.encoder.mid.block_1", cin, cin)
        h = self._vae_attn(be, h, "vae.encoder.mid.attn_1", cin)
        h = self._vae_res(be, h, "vae.encoder.mid.block_2", cin, cin)
        h = be.groupnorm(h, "vae.encoder.norm_out", cin, v["groups"], v["eps"], silu=True)
        h = be.conv2d(h, "vae.encoder.conv_out", cin, 2 * v["z_channels"], 3)
        mom = be.conv2d(h, "vae.quant_conv", 2 * v["z_channels"], 2 * v["embed_dim"], 1)
        return be.slice_channels(mom, 0, v["embed_dim"])

    def vae_decode(self, be, z):
        """z [B,T/8,F/8,embed] -> mel [B,T,F,1]."""
        v = self.s["vae"]
        ch, mult, nrb = v["ch"], v["ch_mult"], v["num_res_blocks"]
        cin = ch * mult[-1]
        h = be.conv2d(z, "vae.post_quant_conv", v["embed_dim"], v["z_channels"], 1)
        h = be.conv2d(h, "vae.decoder.conv_in", v["z_channels"], cin, 3)
        h = self._vae_res(be, h, "vae.decoder.mid.block_1", cin, cin)
        h = self._vae_attn(be, h, "vae.decoder.mid.attn_1", cin)
        h = self._vae_res(be, h, "vae.decoder.mid.block_2", cin, cin)
        for i in reversed(range(len(mult))):
            cout = ch * mult[i]
            for j in range(nrb + 1):
                h = self._vae_res(be, h, f"vae.decoder.up.{i}.block.{j}", cin, cout)
                cin = cout
            if i != 0:
                h = be.upsample_conv2d(h, f"vae.decoder.up.{i}.upsample.conv", cin, cin)
        h = be.groupnorm(h, "vae.decoder.norm_out", cin, v["groups"], v["eps"], silu=True)
        return be.conv2d(h, "vae.decoder.conv_out", cin, v["in_channels"], 3)

    # ---------------------------------------------------------------- UNet (ldm openaimodel.UNetModel)
    def _unet_res(self, be, x, emb_act, name, cin, cout):
        u = self.s["unet"]
        h = be.groupnorm(x, f"{name}.in_layers.0", cin, u["groups"], u["eps"], silu=True)
        if isinstance(emb_act, dict):   # plan backend, EGR_FUSE_EMB=1: all blocks' projections came out of one GEMV
            e = emb_act[name]
        else:
            e = be.linear(emb_act, f"{name}.emb_layers.1", 4 * u["model_channels"], cout, small=True)
        h = be.conv2d(h, f"{name}.in_layers.2", cin, cout, 3, rowbias=e)
        h = be.groupnorm(h, f"{name}.out_layers.0", cout, u["groups"], u["eps"], silu=True)
        sc = x if cin == cout else be.conv2d(x, f"{name}.skip_connection", cin, cout, 1)
        return be.conv2d(h, f"{name}.out_layers.3", cout, cout, 3, add=sc)

    def _unet_attn(self, be, x, name, c):
        """SpatialTransformer, depth 1, self-attention only (context=None on both attention layers)."""
        u = self.s["unet"]
        heads, hd = c // u["head_dim"], u["head_dim"]
        hn = be.groupnorm(x, f"{name}.norm", c, u["groups"], u["attn_norm_eps"], silu=False)
        t = be.conv2d(hn, f"{name}.proj_in", c, c, 1)
        blk = f"{name}.transformer_blocks.0"
        for a in ("attn1", "attn2"):
            n = be.layernorm(t, f"{blk}.norm{1 if a == 'attn1' else 2}", c, 1e-5)
            if getattr(be, "fuse_qkv", False) is True and hd in (16, 32):   # plan backend only, opt-in (EGR_FUSE_QKV=1)
                q, k, v = be.linear_qkv(n, f"{blk}.{a}", c)
            else:
                q = be.linear(n, f"{blk}.{a}.to_q", c, c, bias=False, out="f16")
                k = be.linear(n, f"{blk}.{a}.to_k", c, c, bias=False, out="f16")
                v = be.linear(n, f"{blk}.{a}.to_v", c, c, bias=False, out="f16")
            o = be.attention(q, k, v, heads=heads, head_dim=hd, v_transposed=False)
            t = be.linear(o, f"{blk}.{a}.to_out.0", c, c, add=t)
        n = be.layernorm(t, f"{blk}.norm3", c, 1e-5)
        inner = c * u["ff_mult"]
        g = be.linear(n, f"{blk}.ff.net.0.proj", c, 2 * inner)
        g = be.geglu(g, inner)
        t = be.linear(g, f"{blk}.ff.net.2", inner, c, add=t)
        return be.conv2d(t, f"{name}.proj_out", c, c, 1, add=x)

    def unet(self, be, x, t_value):
        """x [B,H,W,in_channels], scalar timestep -> v-prediction [B,H,W,out_channels]."""
        u = self.s["unet"]
        mc, mult, nrb = u["model_channels"], u["channel_mult"], u["num_res_blocks"]
        te = be.time_embedding(t_value, mc)
        te = be.linear(te, "unet.time_embed.0", mc, 4 * mc, small=True, act="silu")
        te = be.linear(te, "unet.time_embed.2", 4 * mc, 4 * mc, small=True, act="silu")  # SiLU of emb_layers.0 folded in
        if getattr(be, "fuse_emb", False) is True:
            te = be.linear_emb_all(te, 4 * mc)
        hs = []
        h = be.conv2d(x, "unet.input_blocks.0.0", u["in_channels"], mc, 3)
        hs.append((h, mc))
        ch, ds, idx = mc, 1, 1
        for level, m in enumerate(mult):
            for _ in range(nrb):
                h = self._unet_res(be, h, te, f"unet.input_blocks.{idx}.0", ch, m * mc)
                ch = m * mc
                if ds in u["attention_ds"]:
                    h = self._unet_attn(be, h, f"unet.input_blocks.{idx}.1", ch)
                hs.append((h, ch))
                idx += 1
            if level != len(mult) - 1:
                h = be.conv2d(h, f"unet.input_blocks.{idx}.0.op", ch, ch, 3, stride=2, pad="same")
                hs.append((h, ch))
                idx += 1
                ds *= 2
        h = self._unet_res(be, h, te, "unet.middle_block.0", ch, ch)
        h = self._unet_attn(be, h, "unet.middle_block.1", ch)
        h = self._unet_res(be, h, te, "unet.middle_block.2", ch, ch)
        idx = 0
        for level, m in list(enumerate(mult))[::-1]:
            for i in range(nrb + 1):
                skip, sch = hs.pop()
                h = be.concat(h, skip)
                h = self._unet_res(be, h, te, f"unet.output_blocks.{idx}.0", ch + sch, m * mc)
                ch = m * mc
                sub = 1
                if ds in u["attention_ds"]:
                    h = self._unet_attn(be, h, f"unet.output_blocks.{idx}.{sub}", ch)
                    sub += 1
                if level and i == nrb:
                    h = be.upsample_conv2d(h, f"unet.output_blocks.{idx}.{sub}.conv", ch, ch)
                    ds //= 2
                idx += 1
        h = be.groupnorm(h, "unet.out.0", ch, u["groups"], u["eps"], silu=True)
        return be.conv2d(h, "unet.out.2", ch, u["out_channels"], 3)

    # ---------------------------------------------------------------- SR vocoder (BigVGAN-style)
    def vocoder(self, be, mel, wav):
        """mel [B,Tm,n_mels] (channels-last), wav [B,Tm*prod(rates),1] -> waveform [B,T,1]."""
        vc = self.s["vocoder"]
        rates, ksz = vc["upsample_rates"], vc["upsample_kernel_sizes"]
        c0 = vc["upsample_initial_channel"]
        chans = [c0 // (2 ** (i + 1)) for i in range(len(rates))]
        feats = {}
        if vc["wave_cond"]:
            w = be.conv1d(wav, "vocoder.wave_pre", 1, chans[-1], 7)
            feats[len(rates) - 1] = w
            for i in reversed(range(1, len(rates))):
                w = be.conv1d_strided(w, f"vocoder.wave_downs.{i}", chans[i], chans[i - 1], 2 * rates[i], rates[i])
                feats[i - 1] = w
        x = be.conv1d(mel, "vocoder.conv_pre", vc["num_mels"], c0, 7)
        cin = c0
        nk = len(vc["resblock_kernel_sizes"])
        for i, (r, k) in enumerate(zip(rates, ksz)):
            x = be.convT1d(x, f"vocoder.ups.{i}.0", cin, chans[i], k, r, add=feats.get(i))
            cin = chans[i]
            ys = []
            for j, rk in enumerate(vc["resblock_kernel_sizes"]):
                name = f"{name_prefix(i, nk, j)}"
                y = x
                for di, d in enumerate(vc["resblock_dilations"]):
                    yt = be.snake_aa(y, f"{name}.activations.{2 * di}", cin)
                    yt = be.conv1d(yt, f"{name}.convs1.{di}", cin, cin, rk, dilation=d)
                    yt = be.snake_aa(yt, f"{name}.activations.{2 * di + 1}", cin)
                    y = be.conv1d(yt, f"{name}.convs2.{di}", cin, cin, rk, add=y)
                ys.append(y)
            x = be.mean_blocks(ys)   # BigVGAN: xs = sum of the parallel blocks, x = xs / num_kernels
        x = be.snake_aa(x, "vocoder.activation_post", cin)
        return be.conv1d(x, "vocoder.conv_post", cin, 1, 7, act="tanh")

    # ---------------------------------------------------------------- FlashSR.forward
    def forward(self, be, wav, noise, steps: int = 1, lowpass: bool = False):
        """wav [B,chunk] -> [B,chunk].  `noise` is x_T [B,T/8,F/8,z] (caller-supplied)."""
        d = self.s["diffusion"]
        if lowpass:
            wav = be.lowpass(wav)
        mel_lr = be.stft_mel(wav)
        z_lr = self.vae_encode(be, mel_lr)
        x = noise
        region = getattr(be, "region", None)   # plan backend: the denoising loop is one persistent-kernel region
        if region is not None:
            region("unet", True)
        for (t, a_t, a_prev) in ddim_schedule(d["T"], d["cosine_s"], steps):
            v = self.unet(be, be.concat(x, z_lr), t)
            # v-prediction: x0 = sqrt(a)x - sqrt(1-a)v ; eps = sqrt(a)v + sqrt(1-a)x ; x' = sqrt(a')x0 + sqrt(1-a')eps
            sa, s1a = math.sqrt(a_t), math.sqrt(1.0 - a_t)
            sp, s1p = math.sqrt(a_prev), math.sqrt(1.0 - a_prev)
            x = be.axpby(x, v, sp * sa + s1p * s1a, -sp * s1a + s1p * sa)
        if region is not None:
            region("unet", False)
        mel_hat = self.vae_decode(be, x)
        y = self.vocoder(be, be.mel_as_sequence(mel_hat), be.wav_as_sequence(wav))
        return be.sequence_as_wav(y)
