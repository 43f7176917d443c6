"""Test infrastructure: a slow torch/CPU interpreter of the egr_op list (same semantics as the CUDA kernels,
including f16 rounding of tensor-core operands).  It lets the CPU suite validate everything the plan builder
decides — views, taps, packed weights, crops, buffer offsets/aliasing — against the fp32 oracle without a GPU,
and predicts the numerical gap of the f16-operand / f32-accumulate scheme.  Never used by the product."""
import math

import numpy as np
import torch
import torch.nn.functional as F


class Interp:
    def __init__(self, K, ops, ws_bytes, blob_bytes):
        self.K = K
        self.ops = ops
        self.ws = torch.zeros(ws_bytes + 4096, dtype=torch.uint8)
        self.wt = torch.frombuffer(bytearray(blob_bytes), dtype=torch.uint8)
        # poison the workspace so reads of never-written memory show up as NaN
        self.ws.view(torch.float32)[:] = float("nan")

    # ---------------------------------------------------------------- memory
    def _space(self, addr):
        space, off = addr >> 60, addr & 0x0FFFFFFFFFFFFFFF
        if space == self.K["EGR_SPACE_WS"]:
            return self.ws, off
        if space == self.K["EGR_SPACE_WT"]:
            return self.wt, off
        return None, 0

    def flat(self, addr, dtype, n=None):
        mem, off = self._space(addr)
        if mem is None:
            return None
        es = torch.empty((), dtype=dtype).element_size()
        assert off % es == 0
        end = mem.numel() - (mem.numel() - off) % es
        v = mem[off:end].view(dtype)
        return v if n is None else v[:n]

    def p(self, op, slot, dtype, n=None):
        return self.flat(op.ptr[self.K["EGR_P_" + slot]], dtype, n)

    def i(self, op, key):
        return int(op.i[self.K["EGR_I_" + key]])

    def f(self, op, key):
        return float(op.f[self.K["EGR_F_" + key]])

    # ---------------------------------------------------------------- ops
    def gemm(self, op, tc):
        i = lambda k: self.i(op, k)
        t = op.x0
        dt = torch.float16 if t.elem else torch.float32
        A = self.flat(t.addr, dt)
        dims = [t.dim[d] if d < t.rank else 1 for d in range(5)]
        strides = [t.stride[d] if d < t.rank else 0 for d in range(5)]
        Wo, Ho, Bo, K_, N, ntaps = i("WO"), i("HO"), i("BO"), i("K"), i("N"), i("NTAPS")
        dimW, dimH, dimB = i("DIMW"), i("DIMH"), i("DIMB")
        b, h, w, k = torch.meshgrid(torch.arange(Bo), torch.arange(Ho), torch.arange(Wo), torch.arange(K_), indexing="ij")
        wdt = torch.float16 if tc else torch.float32
        Wm = self.p(op, "W", wdt)
        acc = torch.zeros(Bo, Ho, Wo, N, dtype=torch.float32)
        for tp in range(ntaps):
            c = [torch.full_like(b, int(op.tap[tp][d])) for d in range(5)]
            c[0] = c[0] + k
            c[dimW] = c[dimW] + w
            c[dimH] = c[dimH] + h
            c[dimB] = c[dimB] + b
            valid = torch.ones_like(b, dtype=torch.bool)
            off = torch.zeros_like(b)
            for d in range(5):
                valid &= (c[d] >= 0) & (c[d] < dims[d])
                off = off + c[d] * strides[d]
            a = torch.where(valid, A[torch.where(valid, off, torch.zeros_like(off))].float(), torch.zeros(()))
            if tc:
                a = a.half().float()  # K-tail beyond K is zero-filled by TMA: same as not reading it
            if i("WZ_BATCH"):
                wt = torch.stack([Wm[bb * i("WSTRIDE_Z"):][: N * i("WSTRIDE_N")].view(N, i("WSTRIDE_N"))[:, :K_] for bb in range(Bo)]).float()
                acc += torch.einsum("bhwk,bnk->bhwn", a, wt)
            else:
                base = tp * i("WSTRIDE_Z")
                wt = torch.as_strided(Wm, (N, K_), (i("WSTRIDE_N"), 1), Wm.storage_offset() + base).float()
                acc += torch.einsum("bhwk,nk->bhwn", a, wt)
        v = acc * self.f(op, "ALPHA")
        bias = self.p(op, "BIAS", torch.float32, N)
        if bias is not None:
            v = v + bias
        rb = self.p(op, "ROWBIAS", torch.float32)
        if rb is not None:
            rs = i("ROWBIAS_STRIDE")
            v = v + torch.stack([rb[bb * rs: bb * rs + N] for bb in range(Bo)])[:, None, None, :]
        act = i("ACT")
        if act == self.K["EGR_ACT_SILU"]:
            v = F.silu(v)
        elif act == self.K["EGR_ACT_TANH"]:
            v = torch.tanh(v)
        bI, hI, wI, nI = torch.meshgrid(torch.arange(Bo), torch.arange(Ho), torch.arange(Wo), torch.arange(N), indexing="ij")
        pix = hI * Wo + wI
        if i("TRANSPOSED"):
            idx = bI * i("OUT_BATCH_STRIDE") + nI * i("OUT_N_STRIDE") + pix + i("OUT_OFFSET")
            keep = torch.ones_like(idx, dtype=torch.bool)
        else:
            hs = i("OUT_H_STRIDE") if op.code == self.K["EGR_OP_GEMM_TC"] else 0
            flat = (hI * hs + wI * i("OUT_PIX_STRIDE") if hs else pix * i("OUT_PIX_STRIDE")) + i("OUT_OFFSET") + nI
            keep = (flat >= i("OUT_LO")) & (flat < i("OUT_HI"))
            idx = bI * i("OUT_BATCH_STRIDE") + flat
        idx, v = idx[keep], v[keep]
        res = self.p(op, "RESID", torch.float32)
        if res is not None:
            v = v + res[idx]
        res2 = self.p(op, "RESID2", torch.float32) if op.code in (self.K["EGR_OP_GEMM_TC"], self.K["EGR_OP_GEMM_SIMT"]) else None
        if res2 is not None:
            v = v + res2[idx]
        post = self.f(op, "POST")
        if post != 0.0:
            v = v * post
        o32, o16 = self.p(op, "OUT32", torch.float32), self.p(op, "OUT16", torch.float16)
        if o32 is not None:
            o32[idx] = v
        if o16 is not None:
            o16[idx] = v.half()

    def _cat(self, op):
        C0, C1, B, P = self.i(op, "C0"), self.i(op, "C1"), self.i(op, "BATCH"), self.i(op, "ROWS")
        x0 = self.flat(op.x0.addr, torch.float32, B * P * C0).view(B, P, C0)
        if C1:
            x1 = self.flat(op.x1.addr, torch.float32, B * P * C1).view(B, P, C1)
            return torch.cat([x0, x1], -1)
        return x0

    def gn_stats(self, op):
        x = self._cat(op).double()
        B, P, Cc = x.shape
        G = self.i(op, "GROUPS")
        xg = x.view(B, P, G, Cc // G)
        st = self.p(op, "STATS", torch.float64, B * G * 2).view(B, G, 2)
        st[:, :, 0] = xg.sum((1, 3))
        st[:, :, 1] = (xg * xg).sum((1, 3))

    def gn_apply(self, op):
        x = self._cat(op)
        B, P, Cc = x.shape
        G = self.i(op, "GROUPS")
        st = self.p(op, "STATS", torch.float64, B * G * 2).view(B, G, 2)
        cnt = P * (Cc // G)
        mean = st[:, :, 0] / cnt
        var = (st[:, :, 1] / cnt - mean * mean).clamp(min=0)
        rstd = (1.0 / torch.sqrt(var + self.f(op, "EPS"))).float()
        gam, bet = self.p(op, "GAMMA", torch.float32, Cc), self.p(op, "BETA", torch.float32, Cc)
        sc = rstd.repeat_interleave(Cc // G, 1) * gam
        sh = bet - mean.float().repeat_interleave(Cc // G, 1) * sc
        y = x * sc[:, None, :] + sh[:, None, :]
        if self.i(op, "MODE"):
            y = F.silu(y)
        o32, o16 = self.p(op, "OUT32", torch.float32), self.p(op, "OUT16", torch.float16)
        if o32 is not None:
            o32[: y.numel()] = y.reshape(-1)
        if o16 is not None:
            o16[: y.numel()] = y.reshape(-1).half()

    def layernorm(self, op):
        R, Cc = self.i(op, "ROWS"), self.i(op, "COLS")
        x = self.flat(op.x0.addr, torch.float32, R * Cc).view(R, Cc)
        y = F.layer_norm(x, (Cc,), self.p(op, "GAMMA", torch.float32, Cc), self.p(op, "BETA", torch.float32, Cc), self.f(op, "EPS"))
        self.p(op, "OUT16", torch.float16)[: R * Cc] = y.reshape(-1).half()

    def softmax(self, op):
        R, Cc = self.i(op, "ROWS"), self.i(op, "COLS")
        x = self.flat(op.x0.addr, torch.float32, R * Cc).view(R, Cc)
        self.p(op, "OUT16", torch.float16)[: R * Cc] = torch.softmax(x * self.f(op, "ALPHA"), -1).reshape(-1).half()

    def attn_small(self, op):
        S, Hh, hd, B = self.i(op, "SEQ"), self.i(op, "HEADS"), self.i(op, "HEADDIM"), self.i(op, "BATCH")
        n = B * S * Hh * hd
        ld = self.i(op, "AUX0") or Hh * hd   # row stride of q / k / v (column blocks of a fused projection when > C)

        def rows(t):
            return torch.as_strided(t, (B * S, Hh * hd), (ld, 1)).float().reshape(B, S, Hh, hd).permute(0, 2, 1, 3)
        q = rows(self.flat(op.x0.addr, torch.float16))
        k = rows(self.flat(op.x1.addr, torch.float16))
        v = rows(self.p(op, "AUX", torch.float16))
        o = torch.softmax(q @ k.transpose(-1, -2) * self.f(op, "ALPHA"), -1) @ v
        self.p(op, "OUT16", torch.float16)[:n] = o.permute(0, 2, 1, 3).reshape(-1).half()

    def geglu(self, op):
        R, D = self.i(op, "ROWS"), self.i(op, "COLS")
        x = self.flat(op.x0.addr, torch.float32, R * 2 * D).view(R, 2 * D)
        self.p(op, "OUT16", torch.float16)[: R * D] = (x[:, :D] * F.gelu(x[:, D:])).reshape(-1).half()

    def eltwise(self, op):
        K = self.K
        mode = self.i(op, "MODE")
        o32, o16 = self.p(op, "OUT32", torch.float32), self.p(op, "OUT16", torch.float16)
        if mode in (K["EGR_ELT_CAST16"], K["EGR_ELT_COPY32"]):
            C0, C1, R = self.i(op, "C0"), self.i(op, "C1"), self.i(op, "ROWS")
            ld0, ld1 = self.i(op, "AUX0") or C0, self.i(op, "AUX1") or C1
            x0 = torch.as_strided(self.flat(op.x0.addr, torch.float32), (R, C0), (ld0, 1))
            y = x0
            if C1:
                y = torch.cat([x0, torch.as_strided(self.flat(op.x1.addr, torch.float32), (R, C1), (ld1, 1))], 1)
        elif mode == K["EGR_ELT_AXPBY"]:
            n = self.i(op, "ROWS")
            y = self.f(op, "A") * self.flat(op.x0.addr, torch.float32, n) + self.f(op, "B") * self.flat(op.x1.addr, torch.float32, n)
        elif mode == K["EGR_ELT_SCALE_SHIFT"]:
            n = self.i(op, "ROWS")
            y = self.f(op, "A") * self.flat(op.x0.addr, torch.float32, n) + self.f(op, "B")
        elif mode == K["EGR_ELT_SUM3"]:
            n = self.i(op, "ROWS")
            y = self.f(op, "A") * ((self.flat(op.x0.addr, torch.float32, n) + self.flat(op.x1.addr, torch.float32, n))
                                   + self.p(op, "AUX", torch.float32, n))
        elif mode == K["EGR_ELT_UPSAMPLE2X"]:
            B, H, W, Cc = self.i(op, "BATCH"), self.i(op, "AUX0"), self.i(op, "AUX1"), self.i(op, "C0")
            x = self.flat(op.x0.addr, torch.float32, B * H * W * Cc).view(B, H, W, Cc)
            y = x.repeat_interleave(2, 1).repeat_interleave(2, 2)
        else:
            raise ValueError(mode)
        y = y.reshape(-1)
        if o32 is not None:
            o32[: y.numel()] = y
        if o16 is not None:
            o16[: y.numel()] = y.half()

    def snake(self, op):
        B, T, Cc = self.i(op, "BATCH"), self.i(op, "ROWS"), self.i(op, "COLS")
        x = self.flat(op.x0.addr, torch.float32, B * T * Cc).view(B, T, Cc).permute(0, 2, 1)
        f = self.p(op, "AUX", torch.float32, 12)
        Kk, ratio = 12, 2
        ff = f.view(1, 1, Kk).expand(Cc, 1, Kk)
        pad = Kk // ratio - 1
        pl, pr = pad * ratio + (Kk - ratio) // 2, pad * ratio + (Kk - ratio + 1) // 2
        u = ratio * F.conv_transpose1d(F.pad(x, (pad, pad), mode="replicate"), ff, stride=ratio, groups=Cc)[..., pl:-pr]
        al = torch.exp(self.p(op, "GAMMA", torch.float32, Cc)).view(1, Cc, 1)
        be = torch.exp(self.p(op, "BETA", torch.float32, Cc)).view(1, Cc, 1)
        u = u + (1.0 / (be + 1e-9)) * torch.sin(u * al) ** 2
        y = F.conv1d(F.pad(u, (Kk // 2 - 1, Kk // 2), mode="replicate"), ff, stride=ratio, groups=Cc)
        y = y.permute(0, 2, 1).reshape(-1)
        o32, o16 = self.p(op, "OUT32", torch.float32), self.p(op, "OUT16", torch.float16)
        if o32 is not None:
            o32[: y.numel()] = y
        if o16 is not None:
            o16[: y.numel()] = y.half()

    def stft(self, op):
        B, T, n_fft, hop, n_mels = self.i(op, "BATCH"), self.i(op, "ROWS"), self.i(op, "AUX0"), self.i(op, "AUX1"), self.i(op, "AUX2")
        frames, mode = self.i(op, "SEQ"), self.i(op, "MODE")
        wav = self.flat(op.x0.addr, torch.float32, B * T).view(B, T)
        cst = self.p(op, "AUX", torch.uint8)
        window = cst[: n_fft * 4].view(torch.float32)
        o = n_fft * 4 + (n_fft // 4) * 8 + (n_fft // 2 + 1) * 8
        lo = cst[o: o + n_mels * 4].view(torch.int32)
        hi = cst[o + n_mels * 4: o + n_mels * 8].view(torch.int32)
        n_freq = n_fft // 2 + 1
        basis = cst[o + n_mels * 8: o + n_mels * 8 + n_mels * n_freq * 4].view(torch.float32).view(n_mels, n_freq)
        pad = (n_fft - hop) // 2
        y = F.pad(wav[:, None], (pad, pad), mode="reflect")[:, 0]
        st = torch.stft(y, n_fft, hop_length=hop, win_length=n_fft, window=window, center=False, return_complex=True)
        mag = torch.sqrt(st.real ** 2 + st.imag ** 2 + self.f(op, "A"))
        assert mag.shape[2] == frames
        if mode == 0:
            mask = torch.zeros_like(basis)
            for m in range(n_mels):
                mask[m, lo[m]: hi[m]] = 1
            mel = torch.log(torch.clamp((basis * mask) @ mag, min=self.f(op, "B")))
            self.p(op, "OUT32", torch.float32)[: B * frames * n_mels] = mel.permute(0, 2, 1).reshape(-1)
        else:
            e = self.p(op, "STATS", torch.float64, B * n_freq).view(B, n_freq)
            e += mag.double().sum(2)

    def lowpass(self, op):
        from scipy.signal import sosfiltfilt
        B, T, n_freq, nsec = self.i(op, "BATCH"), self.i(op, "ROWS"), self.i(op, "COLS"), self.i(op, "AUX0")
        wav = self.flat(op.x0.addr, torch.float32, B * T).view(B, T)
        e = self.p(op, "STATS", torch.float64, B * n_freq).view(B, n_freq)
        cum = torch.cumsum(e, 1)
        idx = torch.clamp((cum < cum[:, -1:] * self.f(op, "A")).sum(1) - 1, min=0)
        tab = self.p(op, "W", torch.float64, n_freq * nsec * 6).view(n_freq, nsec, 6).numpy()
        y = np.stack([sosfiltfilt(tab[int(idx[b])], wav[b].double().numpy()) for b in range(B)])
        self.p(op, "OUT32", torch.float32)[: B * T] = torch.from_numpy(y).float().reshape(-1)
        self.p(op, "OUT16", torch.int32)[:B] = idx.int()

    def time_embed(self, op):
        dim = self.i(op, "COLS")
        half = dim // 2
        fr = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
        a = self.f(op, "A") * fr
        self.p(op, "OUT32", torch.float32)[:dim] = torch.cat([torch.cos(a), torch.sin(a)])

    def run(self, first=0, last=None):
        K = self.K
        table = {K["EGR_OP_GEMM_TC"]: lambda o: self.gemm(o, True), K["EGR_OP_GEMM_SIMT"]: lambda o: self.gemm(o, False),
                 K["EGR_OP_GN_STATS"]: self.gn_stats, K["EGR_OP_GN_APPLY"]: self.gn_apply, K["EGR_OP_LAYERNORM"]: self.layernorm,
                 K["EGR_OP_SOFTMAX"]: self.softmax, K["EGR_OP_ATTN_SMALL"]: self.attn_small, K["EGR_OP_GEGLU"]: self.geglu,
                 K["EGR_OP_ELTWISE"]: self.eltwise, K["EGR_OP_SNAKE_AA"]: self.snake, K["EGR_OP_STFT_MEL"]: self.stft,
                 K["EGR_OP_LOWPASS"]: self.lowpass, K["EGR_OP_TIME_EMBED"]: self.time_embed,
                 K["EGR_OP_ZERO"]: lambda o: self.p(o, "OUT32", torch.uint8)[: self.i(o, "ROWS")].zero_()}
        ops = self.ops[first: last if last is not None else len(self.ops)]
        for op in ops:
            table[op.code](op)
