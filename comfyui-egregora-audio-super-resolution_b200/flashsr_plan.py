"""Plan backend: interprets `flashsr_model.FlashSRGraph` into the egr_op list executed by libegregora_b200.so.

Host plumbing only — shapes, buffer liveness / offsets inside one workspace, weight repacking into the layouts
the kernels consume (f16 [taps][Cout][Cin] for the tcgen05 tap-GEMM, f32 for the CUDA-core layers, constant
tables for the front end).  No arithmetic on activations happens here.

Layout rule: every activation is channels-last — [B,H,W,C] for the 2-D stages, [B,1,T,C] for the 1-D stages —
so a conv tap is a shifted TMA box over a rank-4/5 view and no im2col buffer exists.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _abi
from . import flashsr_model as M

K = _abi.K
ALIGN = 256
BIG = 1 << 62


class Buf:
    __slots__ = ("nbytes", "first", "last", "offset", "persistent", "tag")

    def __init__(self, nbytes: int, tag: str = "", persistent: bool = False):
        self.nbytes = int((nbytes + ALIGN - 1) // ALIGN * ALIGN)
        self.first = None
        self.last = None
        self.offset = None
        self.persistent = persistent
        self.tag = tag


class PT:
    """Plan tensor: logical [B,H,W,C] with optional f32 / f16 storage, or a virtual channel concat."""

    def __init__(self, B, H, W, Cc, f32: Optional[Buf] = None, f16: Optional[Buf] = None, ld: Optional[int] = None,
                 coff: int = 0, parts=None, f16_transposed: bool = False):
        self.B, self.H, self.W, self.C = int(B), int(H), int(W), int(Cc)
        self.f32, self.f16 = f32, f16
        self.ld = int(ld) if ld else int(Cc)   # channel stride of a pixel row (views of wider tensors)
        self.coff = int(coff)                   # channel offset inside the row
        self.parts = parts                      # (PT, PT) for a virtual concat
        self.f16_transposed = f16_transposed    # f16 stored as [B][C][H*W]
        self.producer = None                    # RawOp that writes .f32 and can also emit an f16 copy at the same indices

    @property
    def P(self):
        return self.H * self.W


class RawOp:
    def __init__(self, code, name):
        self.code, self.name = code, name
        self.x0 = None   # (buf, byte_off, rank, elem, dims, strides)
        self.x1 = None
        self.ptr = {}    # slot -> ("ws", buf, byte_off) | ("wt", off)
        self.i = {}
        self.f = {}
        self.taps = []
        self.reads: List[Buf] = []
        self.writes: List[Buf] = []
        self.index = -1
        self.flags = 0


class WeightBlob:
    """Packs parameters/constants into one byte blob; entries are cached by key so every plan (any batch size)
    refers to the same offsets."""

    def __init__(self):
        self.chunks: List[bytes] = []
        self.size = 0
        self.index: Dict[str, Tuple[int, int]] = {}
        self.frozen = False

    def put(self, key: str, arr: np.ndarray) -> int:
        if key in self.index:
            return self.index[key][0]
        if self.frozen:
            raise RuntimeError(f"weight blob is frozen; unexpected new entry {key}")
        raw = np.ascontiguousarray(arr).tobytes()
        off = self.size
        pad = (-len(raw)) % ALIGN
        self.chunks.append(raw + b"\0" * pad)
        self.size += len(raw) + pad
        self.index[key] = (off, len(raw))
        return off

    def tobytes(self) -> bytes:
        return b"".join(self.chunks)


def _pow2_le(n: int) -> int:
    p = 1
    while p * 2 <= n:
        p *= 2
    return p


def tile_geometry(Wo: int, Ho: int, Bo: int) -> Tuple[int, int, int]:
    bw = min(128, _pow2_le(max(1, Wo)))
    bh = min(128 // bw, _pow2_le(max(1, Ho)))
    bb = 128 // (bw * bh)
    return bw, bh, bb


def gn_slabs(P: int) -> Tuple[int, int]:
    """(slab, n_slabs) of the GroupNorm reduction — must match slab_for() in csrc/ops.cu; a function of the pixel
    count only, so results are bit-identical for any batch size."""
    slab = min(max((P + 591) // 592, 16), P)
    return slab, (P + slab - 1) // slab


def pick_block_n(N: int) -> int:
    best = 0
    for bn in range(16, 257, 16):
        if N % bn == 0:
            best = bn
    return best


class PlanBackend:
    def __init__(self, spec: dict, weights: Dict[str, torch.Tensor], blob: WeightBlob, batch: int):
        self.s = spec
        self.Wt = weights
        self.blob = blob
        self.B = int(batch)
        self.ops: List[RawOp] = []
        self.bufs: List[Buf] = []
        self.inputs: Dict[str, PT] = {}
        self.output: Optional[PT] = None
        self.taps_used = 0
        self.flops = 0.0
        self.tc_flops = 0.0
        self.layer_table: List[dict] = []
        # The three attention projections of a transformer block as ONE GEMM (N = 3C) whose output the attention kernel
        # reads as strided column blocks — 2 launches fewer per attention.  Measured (round 2, c2 B=1, together with the
        # fused embedding GEMV below): 17.12 -> 16.48 ms per chunk; on by default, EGR_FUSE_QKV=0 restores three GEMMs.
        self.fuse_qkv = os.environ.get("EGR_FUSE_QKV", "1") == "1"
        # EGR_FUSE_EMB (default on): the time-embedding projections of ALL UNet ResBlocks (emb_layers.1, M = 1 GEMVs that depend on
        # nothing but the step's embedding) as one GEMV per diffusion step; each block's conv takes its slice as row bias.
        self.fuse_emb = os.environ.get("EGR_FUSE_EMB", "1") == "1"
        self._extra_w: Dict[str, torch.Tensor] = {}   # weights derived at plan time (fused projections)
        self.cutoff_buf: Optional[Buf] = None
        self.debug = False
        self.named: Dict[str, PT] = {}
        # ops emitted while a region is open carry EGR_FLAG_MEGA: the library may run the whole stretch (the UNet of every
        # diffusion step) as ONE persistent kernel (csrc/mega.cu).  EGR_NO_MEGA=1 (read by the library) keeps one launch per op.
        self.mega_region = False

    # ------------------------------------------------------------------ buffers / ops
    def buf(self, nbytes, tag="", persistent=False) -> Buf:
        b = Buf(nbytes, tag, persistent)
        self.bufs.append(b)
        return b

    def new(self, B, H, W, Cc, f32=False, f16=False, tag="", persistent=False, f16_transposed=False) -> PT:
        n = B * H * W * Cc
        persistent = persistent or self.debug
        t = PT(B, H, W, Cc, self.buf(n * 4, tag + ".f32", persistent) if f32 else None,
               self.buf(n * 2, tag + ".f16", persistent) if f16 else None, f16_transposed=f16_transposed)
        if self.debug and tag:
            self.named[tag] = t
        return t

    def region(self, name: str, on: bool):
        """FlashSRGraph brackets the denoising loop with region("unet", True/False)."""
        if name == "unet":
            self.mega_region = bool(on)

    def emit(self, op: RawOp) -> RawOp:
        idx = len(self.ops)
        op.index = idx
        if self.mega_region:
            op.flags |= K["EGR_FLAG_MEGA"]
        for b in op.reads + op.writes:
            if b.first is None:
                b.first = idx
            b.last = idx
        self.ops.append(op)
        return op

    def _ws(self, op: RawOp, slot: str, buf: Buf, off: int = 0, write=False):
        op.ptr[slot] = ("ws", buf, off)
        (op.writes if write else op.reads).append(buf)

    def _wt(self, op: RawOp, slot: str, off: int):
        op.ptr[slot] = ("wt", off)

    # ------------------------------------------------------------------ weights
    def w_bias(self, name) -> Optional[int]:
        key = name + ".bias"
        if key in self._extra_w:
            return self.blob.put("f32:" + key, self._extra_w[key].float().numpy())
        if key not in self.Wt:
            return None
        return self.blob.put("f32:" + key, self.Wt[key].float().numpy())

    def w_vec(self, key) -> int:
        return self.blob.put("f32:" + key, self.Wt[key].float().numpy())

    def w_taps(self, name: str, kind: str, f16: bool, **kw) -> Tuple[int, int, int]:
        """Returns (offset, ntaps, K) of the packed [taps][N][K] weight."""
        w = (self._extra_w[name + ".weight"] if name + ".weight" in self._extra_w else self.Wt[name + ".weight"]).float()
        if kind == "conv2d":          # [cout, cin, kh, kw] -> [kh*kw][cout][cin]
            co, ci, kh, kw_ = w.shape
            p = w.permute(2, 3, 0, 1).reshape(kh * kw_, co, ci)
        elif kind == "conv1d":        # [cout, cin, k] -> [k][cout][cin]
            p = w.permute(2, 0, 1)
        elif kind == "linear":        # [cout, cin] -> [1][cout][cin]
            p = w[None]
        elif kind == "convT1d":       # [cin, cout, 2u] -> [2][u*cout][cin]: tap 0 hits x[q], tap 1 hits x[q-1]
            ci, co, k = w.shape
            u = kw["stride"]
            assert k == 2 * u, "transposed conv is built for kernel == 2*stride"
            p = torch.stack([w[:, :, 0:u], w[:, :, u:2 * u]], 0)      # [2][ci][co][u]
            p = p.permute(0, 3, 2, 1).reshape(2, u * co, ci)           # n = r*cout + co
        elif kind == "conv1d_strided":  # [cout, cin, 2r] stride r pad P -> 3 taps over the [T/r, r*cin] view
            co, ci, k = w.shape
            r = kw["stride"]
            P = (r + 1) // 2
            assert k == 2 * r
            p = torch.zeros(3, co, r * ci)
            for ti, tap in enumerate((-1, 0, 1)):
                for ph in range(r):
                    j = tap * r + ph + P
                    if 0 <= j < k:
                        p[ti, :, ph * ci:(ph + 1) * ci] = w[:, :, j]
        else:
            raise ValueError(kind)
        p = p.contiguous()
        key = ("f16:" if f16 else "f32:") + name + ".weight:" + kind
        arr = p.numpy().astype(np.float16) if f16 else p.numpy().astype(np.float32)
        return self.blob.put(key, arr), p.shape[0], p.shape[2]

    # ------------------------------------------------------------------ helpers
    def _materialize16(self, x: PT, tag="cast") -> PT:
        """Return a PT whose .f16 holds x (concats and f32-only tensors go through one CAST16 pass)."""
        if x.f16 is not None and x.parts is None and not x.f16_transposed and x.ld == x.C:
            return x
        prods = x.producer if isinstance(x.producer, list) else ([x.producer] if x.producer is not None else [])
        if (x.f16 is None and x.parts is None and x.ld == x.C and x.coff == 0 and prods
                and all(pr.index >= 0 and "OUT16" not in pr.ptr for pr in prods)):
            # the op(s) that produced the f32 tensor write the f16 copy in their own epilogue (same element indices as the
            # f32 store, so the four interleaved phase GEMMs of an up-sampling conv fill one buffer): no separate cast pass
            buf = self.buf(x.B * x.P * x.C * 2, tag + ".f16", persistent=self.debug)
            buf.first = min(pr.index for pr in prods)
            buf.last = max(pr.index for pr in prods)
            for pr in prods:
                pr.ptr["OUT16"] = ("ws", buf, 0)
                pr.writes.append(buf)
            x.f16 = buf
            return x
        parts = x.parts if x.parts else (x, None)
        a, b = parts
        out = self.new(x.B, x.H, x.W, x.C, f16=True, tag=tag)
        op = RawOp(K["EGR_OP_ELTWISE"], tag)
        op.i = {"MODE": K["EGR_ELT_CAST16"], "C0": a.C, "C1": b.C if b else 0, "ROWS": x.B * x.P,
                "AUX0": a.ld, "AUX1": b.ld if b else 0}
        assert a.f32 is not None and (b is None or b.f32 is not None)
        op.x0 = (a.f32, a.coff * 4, 1, 0, [a.C], [1])
        op.reads.append(a.f32)
        if b:
            op.x1 = (b.f32, b.coff * 4, 1, 0, [b.C], [1])
            op.reads.append(b.f32)
        self._ws(op, "OUT16", out.f16, write=True)
        self.emit(op)
        return out

    def _gemm(self, name, a: PT, a_view, taps, Kdim, N, w_off, *, simt: bool, dimW, dimH, dimB, Wo, Ho, Bo,
              out_pt: PT, out16: bool, out32: bool, bias_off=None, rowbias: Optional[PT] = None, resid: Optional[PT] = None,
              act=None, alpha=1.0, out_pix_stride=None, out_batch_stride=None, out_offset=0, out_lo=0, out_hi=BIG,
              transposed=False, out_n_stride=0, wstride_n=None, wstride_z=None, wz_batch=False, w_buf: Optional[Buf] = None,
              a_elem=1, resid2: Optional[PT] = None, post=1.0, out_h_stride=0, flops_scale=1.0):
        op = RawOp(K["EGR_OP_GEMM_SIMT"] if simt else K["EGR_OP_GEMM_TC"], name)
        buf, off, dims, strides = a_view
        op.x0 = (buf, off, len(dims), a_elem, dims, strides)
        op.reads.append(buf)
        bw, bh, bb = tile_geometry(Wo, Ho, Bo)
        ntaps = len(taps)
        assert ntaps <= _abi.MAX_TAPS, name
        bn = pick_block_n(N) if not simt else 32
        if not simt:
            assert bn >= 16, (name, N)
        op.i = {"DIMW": dimW, "DIMH": dimH, "DIMB": dimB, "BW": bw, "BH": bh, "BB": bb, "WO": Wo, "HO": Ho, "BO": Bo,
                "NTAPS": ntaps, "K": Kdim, "N": N, "BLOCKN": bn,
                "WSTRIDE_N": wstride_n if wstride_n is not None else Kdim,
                "WSTRIDE_Z": wstride_z if wstride_z is not None else N * Kdim,
                "WZ_BATCH": 1 if wz_batch else 0,
                "ROWBIAS_STRIDE": 0,
                "OUT_PIX_STRIDE": out_pix_stride if out_pix_stride is not None else N,
                "OUT_BATCH_STRIDE": out_batch_stride if out_batch_stride is not None else Ho * Wo * N,
                "OUT_OFFSET": out_offset, "OUT_LO": out_lo, "OUT_HI": out_hi,
                "TRANSPOSED": 1 if transposed else 0, "OUT_N_STRIDE": out_n_stride,
                "ACT": {None: K["EGR_ACT_NONE"], "silu": K["EGR_ACT_SILU"], "tanh": K["EGR_ACT_TANH"]}[act],
                "KBLOCK": 64}
        if out_h_stride:
            assert not simt
            op.i["OUT_H_STRIDE"] = out_h_stride
        op.f = {"ALPHA": alpha, "POST": post}
        op.taps = taps
        if w_buf is not None:
            self._ws(op, "W", w_buf, w_off)
        else:
            self._wt(op, "W", w_off)
        if bias_off is not None:
            self._wt(op, "BIAS", bias_off)
        if rowbias is not None:
            self._ws(op, "ROWBIAS", rowbias.f32, 4 * rowbias.coff)
        if resid is not None:
            assert resid.f32 is not None and resid.parts is None and resid.ld == resid.C
            self._ws(op, "RESID", resid.f32)
        if resid2 is not None:
            assert resid is not None and resid2.f32 is not None and resid2.parts is None and resid2.ld == resid2.C
            self._ws(op, "RESID2", resid2.f32)
        if out32:
            self._ws(op, "OUT32", out_pt.f32, write=True)
        if out16:
            self._ws(op, "OUT16", out_pt.f16, write=True)
        fl = 2.0 * Bo * Ho * Wo * N * Kdim * ntaps * flops_scale   # algorithmic FLOPs (zero-padded taps excluded)
        self.flops += fl
        if not simt:
            self.tc_flops += fl
        self.layer_table.append({"name": name, "kind": "simt" if simt else "tc", "M": Bo * Ho * Wo, "N": N, "K": Kdim * ntaps,
                                 "taps": ntaps, "flops": fl, "block_n": bn, "mega": bool(self.mega_region)})
        self.emit(op)
        if out32 and not out16 and not transposed:
            out_pt.producer = op   # f16 copy on demand (same indices as the f32 store)
        return op

    @staticmethod
    def _use_tc(cin, cout):
        return cin >= 16 and cin % 8 == 0 and cout >= 16 and cout % 16 == 0

    # ------------------------------------------------------------------ 2-D ops
    def conv2d(self, x: PT, name, cin, cout, k, stride=1, pad="same", add=None, rowbias=None, act=None, out=None,
               transposed=False):
        assert x.C == cin, (name, x.C, cin)
        tc = self._use_tc(cin, cout)
        B, H, W = x.B, x.H, x.W
        if tc:
            xa = self._materialize16(x, name + ".in16")
            abuf, elem = xa.f16, 1
        else:
            if x.parts is not None or x.ld != x.C:
                xa = self._materialize16(x, name + ".in16")
                abuf, elem = xa.f16, 1
            elif x.f32 is not None:
                abuf, elem = x.f32, 0
            else:
                abuf, elem = x.f16, 1
        if stride == 1:
            Ho, Wo = H, W
            dims, strides = [cin, W, H, B], [1, cin, W * cin, H * W * cin]
            dimW, dimH, dimB = 1, 2, 3
            taps = [[0, dx - k // 2, dy - k // 2, 0, 0] for dy in range(k) for dx in range(k)]
        else:
            assert stride == 2 and k == 3 and H % 2 == 0 and W % 2 == 0
            Ho, Wo = H // 2, W // 2
            pt = 0 if pad == "ldm_down" else 1
            dims = [2 * cin, W // 2, 2, H // 2, B]
            strides = [1, 2 * cin, W * cin, 2 * W * cin, H * W * cin]
            dimW, dimH, dimB = 1, 3, 4
            taps = []
            for dy in range(3):
                for dx in range(3):
                    a, b = dy - pt, dx - pt
                    taps.append([(b % 2) * cin, b // 2, a % 2, a // 2, 0])
        w_off, ntaps, Kd = self.w_taps(name, "conv2d", f16=tc)
        want16 = out == "f16"
        o = self.new(B, Ho, Wo, cout, f32=not want16, f16=want16, tag=name, f16_transposed=transposed)
        kw = {}
        if transposed:
            assert want16
            kw = dict(transposed=True, out_n_stride=Ho * Wo, out_batch_stride=cout * Ho * Wo, out_pix_stride=1)
        self._gemm(name, x, (abuf, 0, dims, strides), taps, cin, cout, w_off, simt=not tc, dimW=dimW, dimH=dimH, dimB=dimB,
                   Wo=Wo, Ho=Ho, Bo=B, out_pt=o, out16=want16, out32=not want16, bias_off=self.w_bias(name),
                   rowbias=rowbias, resid=add, act=act, a_elem=elem, **kw)
        return o

    def upsample_conv2d(self, x: PT, name, cin, cout):
        """conv3x3(nearest_upsample2x(x)) without the 4x larger intermediate: output pixel (2y+a, 2x+b) only sees a
        2x2 window of the low-resolution map, so each of the four output phases is a 4-tap GEMM over x whose weights
        are sums of the 3x3 taps that land on the same source pixel (2.25x fewer FLOPs, no up-sampled copy in HBM)."""
        assert x.C == cin and self._use_tc(cin, cout)
        xa = self._materialize16(x, name + ".in16")
        B, H, W = x.B, x.H, x.W
        o = self.new(B, 2 * H, 2 * W, cout, f32=True, tag=name)
        w = self.Wt[name + ".weight"].float()           # [cout, cin, 3, 3]
        groups = {0: [(-1, [0]), (0, [1, 2])], 1: [(0, [0, 1]), (1, [2])]}   # phase -> [(source offset, kernel rows/cols)]
        dims, strides = [cin, W, H, B], [1, cin, W * cin, H * W * cin]
        phase_ops = []
        for a in (0, 1):
            for b in (0, 1):
                taps, mats = [], []
                for dy, kys in groups[a]:
                    for dx, kxs in groups[b]:
                        taps.append([0, dx, dy, 0, 0])
                        mats.append(sum(w[:, :, ky, kx] for ky in kys for kx in kxs))   # [cout, cin]
                pk = torch.stack(mats, 0).contiguous()
                w_off = self.blob.put(f"f16:{name}.weight:up{a}{b}", pk.numpy().astype(np.float16))
                phase_ops.append(self._gemm(f"{name}.p{a}{b}", x, (xa.f16, 0, dims, strides), taps, cin, cout, w_off, simt=False, dimW=1, dimH=2,
                           dimB=3, Wo=W, Ho=H, Bo=B, out_pt=o, out16=False, out32=True, bias_off=self.w_bias(name),
                           out_pix_stride=2 * cout, out_h_stride=4 * W * cout, out_batch_stride=4 * H * W * cout,
                           out_offset=(a * 2 * W + b) * cout, out_lo=0, out_hi=4 * H * W * cout))
        o.producer = phase_ops   # four writers: an f16 copy on demand is written by all of them (_materialize16)
        return o

    def linear(self, x: PT, name, cin, cout, bias=True, small=False, act=None, out=None, add=None):
        assert x.C == cin, (name, x.C, cin)
        tc = self._use_tc(cin, cout) and not small
        if tc:
            xa = self._materialize16(x, name + ".in16")
            abuf, elem = xa.f16, 1
        else:
            abuf, elem = (x.f32, 0) if x.f32 is not None else (x.f16, 1)
        B, H, W = x.B, x.H, x.W
        dims, strides = [cin, W, H, B], [1, cin, W * cin, H * W * cin]
        w_off, _, _ = self.w_taps(name, "linear", f16=tc)
        want16 = out == "f16"
        o = self.new(B, H, W, cout, f32=not want16, f16=want16, tag=name)
        self._gemm(name, x, (abuf, 0, dims, strides), [[0, 0, 0, 0, 0]], cin, cout, w_off, simt=not tc, dimW=1, dimH=2, dimB=3,
                   Wo=W, Ho=H, Bo=B, out_pt=o, out16=want16, out32=not want16, bias_off=self.w_bias(name) if bias else None,
                   resid=add, act=act, a_elem=elem)
        return o

    def linear_emb_all(self, emb: PT, cin) -> Dict[str, PT]:
        """emb_layers.1 of every UNet ResBlock in one GEMV: {block name: [1,1,1,cout] f32 slice of the joint output}."""
        names = sorted(k[:-len(".emb_layers.1.weight")] for k in self.Wt if k.startswith("unet.") and k.endswith(".emb_layers.1.weight"))
        key = "unet.emb_layers_all"
        if key + ".weight" not in self._extra_w:
            self._extra_w[key + ".weight"] = torch.cat([self.Wt[n + ".emb_layers.1.weight"] for n in names], 0)
            self._extra_w[key + ".bias"] = torch.cat([self.Wt[n + ".emb_layers.1.bias"] for n in names], 0)
        total = int(self._extra_w[key + ".weight"].shape[0])
        o = self.linear(emb, key, cin, total, small=True)
        out, off = {}, 0
        for n in names:
            c = int(self.Wt[n + ".emb_layers.1.weight"].shape[0])
            out[n] = PT(o.B, 1, 1, c, f32=o.f32, ld=total, coff=off)
            off += c
        return out

    def linear_qkv(self, x: PT, prefix, c):
        """to_q / to_k / to_v of one attention layer as a single [3c, c] projection -> three f16 column-block views."""
        name = prefix + ".to_qkv"
        if name + ".weight" not in self._extra_w:
            self._extra_w[name + ".weight"] = torch.cat([self.Wt[f"{prefix}.to_{s}.weight"] for s in "qkv"], 0)
        o = self.linear(x, name, c, 3 * c, bias=False, out="f16")
        return tuple(PT(o.B, o.H, o.W, c, f16=o.f16, ld=3 * c, coff=i * c) for i in range(3))

    def groupnorm(self, x: PT, name, c, groups, eps, silu=False):
        assert x.C == c
        a, b = x.parts if x.parts else (x, None)
        assert a.f32 is not None and a.ld == a.C and (b is None or (b.f32 is not None and b.ld == b.C))
        slab, ns = gn_slabs(x.P)
        # [B][G][2] f64 totals followed by [B][ns][G][2] per-slab partials (deterministic two-stage reduction)
        stats = self.buf(x.B * groups * 2 * 8 * (1 + ns), name + ".stats")
        common = {"C0": a.C, "C1": b.C if b else 0, "GROUPS": groups, "BATCH": x.B, "ROWS": x.P, "AUX0": slab, "AUX1": ns}
        st = RawOp(K["EGR_OP_GN_STATS"], name + ".stats")
        st.i = dict(common)
        st.x0 = (a.f32, 0, 1, 0, [a.C], [1]); st.reads.append(a.f32)
        if b:
            st.x1 = (b.f32, 0, 1, 0, [b.C], [1]); st.reads.append(b.f32)
        self._ws(st, "STATS", stats, write=True)
        self.emit(st)
        o = self.new(x.B, x.H, x.W, c, f16=True, tag=name)
        ap = RawOp(K["EGR_OP_GN_APPLY"], name)
        ap.i = dict(common, MODE=1 if silu else 0)
        ap.f = {"EPS": eps}
        ap.x0 = st.x0; ap.reads.append(a.f32)
        if b:
            ap.x1 = st.x1; ap.reads.append(b.f32)
        self._ws(ap, "STATS", stats)
        self._wt(ap, "GAMMA", self.w_vec(name + ".weight"))
        self._wt(ap, "BETA", self.w_vec(name + ".bias"))
        self._ws(ap, "OUT16", o.f16, write=True)
        self.emit(ap)
        return o

    def layernorm(self, x: PT, name, c, eps):
        assert x.C == c and x.f32 is not None
        o = self.new(x.B, x.H, x.W, c, f16=True, tag=name)
        op = RawOp(K["EGR_OP_LAYERNORM"], name)
        op.i = {"ROWS": x.B * x.P, "COLS": c}
        op.f = {"EPS": eps}
        op.x0 = (x.f32, 0, 1, 0, [c], [1]); op.reads.append(x.f32)
        self._wt(op, "GAMMA", self.w_vec(name + ".weight"))
        self._wt(op, "BETA", self.w_vec(name + ".bias"))
        self._ws(op, "OUT16", o.f16, write=True)
        self.emit(op)
        return o

    def geglu(self, x: PT, inner):
        assert x.C == 2 * inner and x.f32 is not None
        o = self.new(x.B, x.H, x.W, inner, f16=True, tag="geglu")
        op = RawOp(K["EGR_OP_GEGLU"], "geglu")
        op.i = {"ROWS": x.B * x.P, "COLS": inner}
        op.x0 = (x.f32, 0, 1, 0, [x.C], [1]); op.reads.append(x.f32)
        self._ws(op, "OUT16", o.f16, write=True)
        self.emit(op)
        return o

    def attention(self, q: PT, k: PT, v: PT, heads, head_dim, v_transposed=False):
        B, S, Cc = q.B, q.P, q.C
        scale = head_dim ** -0.5
        o = self.new(B, q.H, q.W, Cc, f16=True, tag="attn.out")
        if heads == 1 and head_dim > 64:
            assert v_transposed and v.f16_transposed and S >= 16 and S % 16 == 0
            scores = self.new(B, 1, S, S, f32=True, tag="attn.scores")
            self._gemm("attn.qk", q, (q.f16, 0, [Cc, S, 1, B], [1, Cc, S * Cc, S * Cc]), [[0, 0, 0, 0, 0]], Cc, S, 0,
                       simt=False, dimW=1, dimH=2, dimB=3, Wo=S, Ho=1, Bo=B, out_pt=scores, out16=False, out32=True,
                       wstride_n=Cc, wstride_z=S * Cc, wz_batch=True, w_buf=k.f16)
            probs = self.new(B, 1, S, S, f16=True, tag="attn.probs")
            sm = RawOp(K["EGR_OP_SOFTMAX"], "attn.softmax")
            sm.i = {"ROWS": B * S, "COLS": S}
            sm.f = {"ALPHA": scale}
            sm.x0 = (scores.f32, 0, 1, 0, [S], [1]); sm.reads.append(scores.f32)
            self._ws(sm, "OUT16", probs.f16, write=True)
            self.emit(sm)
            self._gemm("attn.pv", probs, (probs.f16, 0, [S, S, 1, B], [1, S, S * S, S * S]), [[0, 0, 0, 0, 0]], S, Cc, 0,
                       simt=False, dimW=1, dimH=2, dimB=3, Wo=S, Ho=1, Bo=B, out_pt=o, out16=True, out32=False,
                       wstride_n=S, wstride_z=Cc * S, wz_batch=True, w_buf=v.f16)
            return o
        assert not v_transposed and head_dim <= 64
        assert q.ld == k.ld == v.ld and (q.ld == Cc or head_dim in (16, 32)), "strided q/k/v needs the 16/32-dim head kernel"
        op = RawOp(K["EGR_OP_ATTN_SMALL"], "attn.small")
        op.i = {"SEQ": S, "HEADS": heads, "HEADDIM": head_dim, "BATCH": B, "AUX0": q.ld if q.ld != Cc else 0}
        op.f = {"ALPHA": scale}
        op.x0 = (q.f16, 2 * q.coff, 1, 1, [Cc], [1]); op.reads.append(q.f16)
        op.x1 = (k.f16, 2 * k.coff, 1, 1, [Cc], [1]); op.reads.append(k.f16)
        self._ws(op, "AUX", v.f16, 2 * v.coff)
        self._ws(op, "OUT16", o.f16, write=True)
        self.emit(op)
        return o

    def upsample2x(self, x: PT):
        assert x.f32 is not None and x.parts is None
        o = self.new(x.B, 2 * x.H, 2 * x.W, x.C, f16=True, tag="up2x")
        op = RawOp(K["EGR_OP_ELTWISE"], "up2x")
        op.i = {"MODE": K["EGR_ELT_UPSAMPLE2X"], "BATCH": x.B, "AUX0": x.H, "AUX1": x.W, "C0": x.C}
        op.x0 = (x.f32, 0, 1, 0, [x.C], [1]); op.reads.append(x.f32)
        self._ws(op, "OUT16", o.f16, write=True)
        self.emit(op)
        return o

    def concat(self, a: PT, b: PT):
        assert (a.B, a.H, a.W) == (b.B, b.H, b.W) and a.parts is None and b.parts is None
        return PT(a.B, a.H, a.W, a.C + b.C, parts=(a, b))

    def slice_channels(self, x: PT, lo, n):
        assert x.f32 is not None and x.parts is None
        return PT(x.B, x.H, x.W, n, f32=x.f32, ld=x.ld, coff=x.coff + lo)

    def _elt(self, name, mode, x: PT, y: Optional[PT], a, b, n):
        o = self.new(x.B, x.H, x.W, x.C, f32=True, tag=name)
        op = RawOp(K["EGR_OP_ELTWISE"], name)
        op.i = {"MODE": mode, "ROWS": n}
        op.f = {"A": a, "B": b}
        op.x0 = (x.f32, 0, 1, 0, [x.C], [1]); op.reads.append(x.f32)
        if y is not None:
            op.x1 = (y.f32, 0, 1, 0, [y.C], [1]); op.reads.append(y.f32)
        self._ws(op, "OUT32", o.f32, write=True)
        self.emit(op)
        if mode in (K["EGR_ELT_AXPBY"], K["EGR_ELT_SCALE_SHIFT"]):
            o.producer = op
        return o

    def axpby(self, x: PT, y: PT, a, b):
        assert x.f32 is not None and y.f32 is not None and x.ld == x.C and y.ld == y.C
        return self._elt("axpby", K["EGR_ELT_AXPBY"], x, y, a, b, x.B * x.P * x.C)

    def add(self, a: PT, b: PT):
        return self._elt("add", K["EGR_ELT_AXPBY"], a, b, 1.0, 1.0, a.B * a.P * a.C)

    def mean_blocks(self, ys):
        """(y0 + y1 + ...) / n of the vocoder's parallel residual blocks; three blocks are one fused pass."""
        n = len(ys)
        if n == 3 and all(y.f32 is not None and y.parts is None and y.ld == y.C for y in ys) and (ys[0].B * ys[0].P * ys[0].C) % 4 == 0:
            x = ys[0]
            o = self.new(x.B, x.H, x.W, x.C, f32=True, tag="mean3")
            op = RawOp(K["EGR_OP_ELTWISE"], "mean3")
            op.i = {"MODE": K["EGR_ELT_SUM3"], "ROWS": x.B * x.P * x.C}
            op.f = {"A": 1.0 / n}
            op.x0 = (ys[0].f32, 0, 1, 0, [x.C], [1]); op.reads.append(ys[0].f32)
            op.x1 = (ys[1].f32, 0, 1, 0, [x.C], [1]); op.reads.append(ys[1].f32)
            self._ws(op, "AUX", ys[2].f32)
            self._ws(op, "OUT32", o.f32, write=True)
            self.emit(op)
            o.producer = op
            return o
        xs = ys[0]
        for y in ys[1:]:
            xs = self.add(xs, y)
        return self.scale(xs, 1.0 / n)

    def scale(self, a: PT, s):
        return self._elt("scale", K["EGR_ELT_SCALE_SHIFT"], a, None, s, 0.0, a.B * a.P * a.C)

    def time_embedding(self, t_value, dim):
        o = self.new(1, 1, 1, dim, f32=True, tag="t_emb")
        op = RawOp(K["EGR_OP_TIME_EMBED"], "t_emb")
        op.i = {"COLS": dim}
        op.f = {"A": float(t_value)}
        self._ws(op, "OUT32", o.f32, write=True)
        self.emit(op)
        return o

    # ------------------------------------------------------------------ 1-D ops ([B,1,T,C])
    def conv1d(self, x: PT, name, cin, cout, k, dilation=1, add=None, act=None, add2=None, post=1.0):
        assert x.C == cin and x.H == 1, name
        tc = self._use_tc(cin, cout)
        if tc:
            xa = self._materialize16(x, name + ".in16")
            abuf, elem = xa.f16, 1
        else:
            abuf, elem = (x.f32, 0) if x.f32 is not None else (x.f16, 1)
        B, T = x.B, x.W
        dims, strides = [cin, T, 1, B], [1, cin, T * cin, T * cin]
        taps = [[0, (j - k // 2) * dilation, 0, 0, 0] for j in range(k)]
        w_off, _, _ = self.w_taps(name, "conv1d", f16=tc)
        o = self.new(B, 1, T, cout, f32=True, tag=name)
        self._gemm(name, x, (abuf, 0, dims, strides), taps, cin, cout, w_off, simt=not tc, dimW=1, dimH=2, dimB=3, Wo=T, Ho=1,
                   Bo=B, out_pt=o, out16=False, out32=True, bias_off=self.w_bias(name), resid=add, act=act, a_elem=elem,
                   resid2=add2, post=post)
        return o

    def conv1d_strided(self, x: PT, name, cin, cout, k, stride):
        assert x.C == cin and x.H == 1 and x.W % stride == 0 and k == 2 * stride
        B, T = x.B, x.W
        To, Kd = T // stride, stride * cin
        tc = self._use_tc(Kd, cout)
        if tc:
            xa = self._materialize16(x, name + ".in16")
            abuf, elem = xa.f16, 1
        else:
            abuf, elem = (x.f32, 0) if x.f32 is not None else (x.f16, 1)
        dims, strides = [Kd, To, 1, B], [1, Kd, To * Kd, To * Kd]
        taps = [[0, t, 0, 0, 0] for t in (-1, 0, 1)]
        w_off, _, _ = self.w_taps(name, "conv1d_strided", f16=tc, stride=stride)
        o = self.new(B, 1, To, cout, f32=True, tag=name)
        self._gemm(name, x, (abuf, 0, dims, strides), taps, Kd, cout, w_off, simt=not tc, dimW=1, dimH=2, dimB=3, Wo=To, Ho=1,
                   Bo=B, out_pt=o, out16=False, out32=True, bias_off=self.w_bias(name), a_elem=elem,
                   flops_scale=2.0 / 3.0)   # k = 2*stride real taps inside the 3*stride-wide padded window
        return o

    def convT1d(self, x: PT, name, cin, cout, k, stride, add=None):
        assert x.C == cin and x.H == 1 and k == 2 * stride
        B, T = x.B, x.W
        u, p = stride, (k - stride) // 2
        N = u * cout
        tc = self._use_tc(cin, N)
        if tc:
            xa = self._materialize16(x, name + ".in16")
            abuf, elem = xa.f16, 1
        else:
            abuf, elem = (x.f32, 0) if x.f32 is not None else (x.f16, 1)
        dims, strides = [cin, T, 1, B], [1, cin, T * cin, T * cin]
        taps = [[0, 0, 0, 0, 0], [0, -1, 0, 0, 0]]
        w_off, _, _ = self.w_taps(name, "convT1d", f16=tc, stride=stride)
        # bias is per output channel, N = u*cout columns -> tile it u times
        bkey = "f32:" + name + ".bias:tiled"
        b_off = self.blob.put(bkey, np.tile(self.Wt[name + ".bias"].float().numpy(), u))
        o = self.new(B, 1, T * u, cout, f32=True, tag=name)
        self._gemm(name, x, (abuf, 0, dims, strides), taps, cin, N, w_off, simt=not tc, dimW=1, dimH=2, dimB=3, Wo=T + 1, Ho=1,
                   Bo=B, out_pt=o, out16=False, out32=True, bias_off=b_off, resid=add, a_elem=elem,
                   out_pix_stride=N, out_batch_stride=T * u * cout, out_offset=-p * cout, out_lo=0, out_hi=T * u * cout)
        return o

    def snake_aa(self, x: PT, name, c):
        assert x.C == c and x.H == 1 and x.f32 is not None
        kk = self.s["vocoder"]["aa_kernel"]
        f_off = self.blob.put("f32:aa_filter", M.kaiser_sinc_filter1d(0.25, 0.3, kk))
        o = self.new(x.B, 1, x.W, c, f16=True, tag=name)
        op = RawOp(K["EGR_OP_SNAKE_AA"], name)
        op.i = {"BATCH": x.B, "ROWS": x.W, "COLS": c, "AUX0": kk}
        op.x0 = (x.f32, 0, 1, 0, [c], [1]); op.reads.append(x.f32)
        self._wt(op, "GAMMA", self.w_vec(name + ".act.alpha"))
        self._wt(op, "BETA", self.w_vec(name + ".act.beta"))
        self._wt(op, "AUX", f_off)
        self._ws(op, "OUT16", o.f16, write=True)
        self.emit(op)
        return o

    # ------------------------------------------------------------------ front end
    def _stft_consts(self) -> int:
        m = self.s["mel"]
        n_fft, n_mels = m["n_fft"], m["n_mels"]
        window = M.hann_periodic(m["win"]).astype(np.float32)
        assert m["win"] == n_fft
        Mh = n_fft // 2
        kk = np.arange(Mh // 2, dtype=np.float64)
        tw_half = np.stack([np.cos(2 * np.pi * kk / Mh), -np.sin(2 * np.pi * kk / Mh)], -1).astype(np.float32)
        kf = np.arange(Mh + 1, dtype=np.float64)
        tw_full = np.stack([np.cos(2 * np.pi * kf / n_fft), -np.sin(2 * np.pi * kf / n_fft)], -1).astype(np.float32)
        basis = M.mel_filterbank(self.s["sr"], n_fft, n_mels, m["fmin"], m["fmax"])
        nz = basis != 0
        lo = np.where(nz.any(1), nz.argmax(1), 0).astype(np.int32)
        hi = np.where(nz.any(1), basis.shape[1] - nz[:, ::-1].argmax(1), 0).astype(np.int32)
        raw = window.tobytes() + tw_half.tobytes() + tw_full.tobytes() + lo.tobytes() + hi.tobytes() + basis.tobytes()
        return self.blob.put("stft_consts", np.frombuffer(raw, np.uint8))

    def _stft(self, wav: PT, mode: int, out: Optional[PT], energy: Optional[Buf], name):
        m = self.s["mel"]
        T = wav.W
        frames = T // m["hop"]
        op = RawOp(K["EGR_OP_STFT_MEL"], name)
        op.i = {"BATCH": wav.B, "ROWS": T, "AUX0": m["n_fft"], "AUX1": m["hop"], "AUX2": m["n_mels"], "MODE": mode, "SEQ": frames}
        op.f = {"A": m["mag_eps"], "B": m["log_clamp"]}
        op.x0 = (wav.f32, 0, 1, 0, [1], [1]); op.reads.append(wav.f32)
        self._wt(op, "AUX", self._stft_consts())
        if out is not None:
            self._ws(op, "OUT32", out.f32, write=True)
        if energy is not None:
            self._ws(op, "STATS", energy, write=True)
        self.emit(op)

    def stft_mel(self, wav: PT):
        m = self.s["mel"]
        frames = wav.W // m["hop"]
        o = self.new(wav.B, frames, m["n_mels"], 1, f32=True, tag="mel_lr")
        self._stft(wav, 0, o, None, "stft_mel")
        return o

    def lowpass(self, wav: PT):
        from scipy.signal import sosfilt_zi
        m, lp = self.s["mel"], self.s["lowpass"]
        n_freq = m["n_fft"] // 2 + 1
        tab = M.lowpass_sos_table(self.s)
        nsec = tab.shape[1]
        zi = np.stack([sosfilt_zi(tab[b]) for b in range(n_freq)])
        sos_off = self.blob.put("lp_sos", tab.astype(np.float64))
        zi_off = self.blob.put("lp_zi", zi.astype(np.float64))
        B, T = wav.B, wav.W
        energy = self.buf(B * n_freq * 8, "lp.energy")
        z = RawOp(K["EGR_OP_ZERO"], "lp.zero")
        z.i = {"ROWS": B * n_freq * 8}
        self._ws(z, "OUT32", energy, write=True)
        self.emit(z)
        self._stft(wav, 1, None, energy, "lp.stft_energy")
        edge = 3 * (2 * nsec + 1)
        scratch = self.buf(B * (T + 2 * edge + 8192) * 8, "lp.scratch")  # chunk-transposed f64 signal, padded to 8192 chunks
        cut = self.buf(B * 4, "lp.cutoff", persistent=True)
        self.cutoff_buf = cut
        o = self.new(B, 1, T, 1, f32=True, tag="lp.out")
        op = RawOp(K["EGR_OP_LOWPASS"], "lowpass")
        op.i = {"BATCH": B, "ROWS": T, "COLS": n_freq, "AUX0": nsec}
        op.f = {"A": lp["energy_percentile"]}
        op.x0 = (wav.f32, 0, 1, 0, [1], [1]); op.reads.append(wav.f32)
        self._ws(op, "STATS", energy)
        self._wt(op, "W", sos_off)
        self._wt(op, "BIAS", zi_off)
        self._ws(op, "AUX", scratch, write=True)
        self._ws(op, "OUT32", o.f32, write=True)
        self._ws(op, "OUT16", cut, write=True)
        self.emit(op)
        return o

    def mel_as_sequence(self, mel: PT):   # [B,T,F,1] -> [B,1,T,F]: same bytes
        return PT(mel.B, 1, mel.H, mel.W, f32=mel.f32)

    def wav_as_sequence(self, wav: PT):
        return wav

    def sequence_as_wav(self, y: PT):
        return y

    # ------------------------------------------------------------------ finalize
    def allocate(self) -> int:
        """Greedy first-fit interval allocation of every buffer inside one workspace."""
        n_ops = len(self.ops)
        for b in self.bufs:
            if b.first is None:
                b.first, b.last = 0, 0
            if b.persistent:
                b.first, b.last = 0, n_ops
        order = sorted(self.bufs, key=lambda b: (b.first, -b.nbytes))
        live: List[Buf] = []
        top = 0
        for b in order:
            live = [x for x in live if x.last >= b.first]
            live.sort(key=lambda x: x.offset)
            pos = 0
            for x in live:
                if x.offset - pos >= b.nbytes:
                    break
                pos = max(pos, x.offset + x.nbytes)
            b.offset = pos
            live.append(b)
            top = max(top, pos + b.nbytes)
        return top

    def build_ops(self):
        arr = (_abi.Op * len(self.ops))()
        for n, r in enumerate(self.ops):
            o = arr[n]
            o.code = r.code
            o.flags = r.flags
            o.name = r.name.encode()[:47]

            def fill(t, v):
                if v is None:
                    return
                buf, off, rank, elem, dims, strides = v
                t.addr = _abi.ws(buf.offset + off)
                t.rank, t.elem = rank, elem
                for d in range(rank):
                    t.dim[d], t.stride[d] = dims[d], strides[d]

            fill(o.x0, r.x0)
            fill(o.x1, r.x1)
            for slot, v in r.ptr.items():
                idx = K["EGR_P_" + slot]
                o.ptr[idx] = _abi.ws(v[1].offset + v[2]) if v[0] == "ws" else _abi.wt(v[1])
            for key, v in r.i.items():
                o.i[K["EGR_I_" + key]] = int(v)
            for key, v in r.f.items():
                o.f[K["EGR_F_" + key]] = float(v)
            for ti, tp in enumerate(r.taps):
                for d in range(5):
                    o.tap[ti][d] = int(tp[d])
        return arr


def build_plan(spec, weights, blob: WeightBlob, batch: int, steps: int, lowpass: bool, debug: bool = False):
    """Walk the graph once for a given (batch, steps, lowpass) and return the PlanBackend with ops + layout.
    debug=True keeps every intermediate alive (no buffer reuse) and indexes them by op name for layer-wise checks."""
    be = PlanBackend(spec, weights, blob, batch)
    be.debug = debug
    T = spec["chunk"]
    frames = T // spec["mel"]["hop"]
    z = spec["vae"]["embed_dim"]
    wav = be.new(batch, 1, T, 1, f32=True, tag="in.wav", persistent=True)
    noise = be.new(batch, frames // 8, spec["mel"]["n_mels"] // 8, z, f32=True, tag="in.noise", persistent=True)
    be.inputs = {"wav": wav, "noise": noise}
    y = M.FlashSRGraph(spec).forward(be, wav, noise, steps=steps, lowpass=lowpass)
    y.f32.persistent = True
    be.output = y
    be.ws_bytes = be.allocate()
    return be
