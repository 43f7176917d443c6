"""GPU parity of the plan's op kernels against plain torch fp32 references of the same op (called through
egr_plan_create / egr_plan_run, i.e. the C ABI).  Tolerances: CUDA-core f32 paths 1e-5 relative; tensor-core
paths round operands to f16 (11-bit mantissa) and accumulate in f32 -> 2e-3 relative RMS."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from harness import MiniPlan, rel_err

pytestmark = pytest.mark.gpu
TC_TOL, F32_TOL = 2e-3, 2e-5


def _w(shape, seed):
    g = torch.Generator().manual_seed(seed)
    fan = int(np.prod(shape[1:]))
    return (torch.rand(shape, generator=g) * 2 - 1) * math.sqrt(3.0 / fan)


def _x(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


@pytest.mark.parametrize("cin,cout,H,W,B,k", [
    (64, 64, 8, 16, 1, 1),      # one tile, one k-iteration
    (128, 128, 16, 16, 2, 3),   # taps + halo (TMA zero fill)
    (16, 32, 8, 16, 1, 3),      # K tail (box wider than the tensor)
    (48, 96, 4, 32, 3, 3),      # K tail 48, N 96, batch not filling a tile
    (96, 160, 8, 8, 2, 1),      # K = 64 + 32 tail, BLOCK_N 160
    (64, 512, 8, 16, 1, 1),     # two N tiles of 256
    (256, 240, 2, 64, 1, 3),    # BLOCK_N 240
    (32, 16, 4, 4, 5, 3),       # tiny spatial extent: tile spans 8 batch items (B=5 -> OOB rows)
])
def test_conv2d_tc(cin, cout, H, W, B, k, cuda_dev):
    Wt = {"c.weight": _w((cout, cin, k, k), 1), "c.bias": _x((cout,), 2) * 0.1}
    x = _x((B, cin, H, W), 3)
    res = _x((B, cout, H, W), 4)
    mp = MiniPlan(Wt)
    xi, ri = mp.input(x), mp.input(res)
    y = mp.be.conv2d(xi, "c", cin, cout, k, add=ri)
    assert mp.be.ops[-1].code == 1
    mp.run_gpu()
    ref = F.conv2d(x, Wt["c.weight"], Wt["c.bias"], padding=k // 2) + res
    assert rel_err(mp.read(y), ref) < TC_TOL


@pytest.mark.parametrize("cin,cout,H,W,B,k", [
    (640, 640, 8, 4, 1, 3),      # UNet bottom: 32 pixels, K = 5760 -> split-K, last-arriver reduction
    (640, 640, 8, 4, 3, 1),      # same tile shared by 3 batch items (bb = 4)
    (384, 384, 16, 8, 2, 3),     # one 128-row tile per item, 2 items
    (1280, 640, 8, 4, 1, 3),     # concat-width K
    (256, 2048, 32, 16, 1, 1),   # wide N (GEGLU projection), 4 m-tiles
    (128, 128, 64, 32, 1, 3),    # 16 m-tiles, 18 k-steps
])
def test_conv2d_tc_splitk(cin, cout, H, W, B, k, cuda_dev):
    Wt = {"c.weight": _w((cout, cin, k, k), 1), "c.bias": _x((cout,), 2) * 0.1}
    x = _x((B, cin, H, W), 3)
    res = _x((B, cout, H, W), 4)
    mp = MiniPlan(Wt)
    xi, ri = mp.input(x), mp.input(res)
    y = mp.be.conv2d(xi, "c", cin, cout, k, add=ri)
    y16 = mp.be.conv2d(xi, "c", cin, cout, k, out="f16", act="silu")
    mp.run_gpu()
    ref = F.conv2d(x, Wt["c.weight"], Wt["c.bias"], padding=k // 2)
    assert rel_err(mp.read(y), ref + res) < TC_TOL
    assert rel_err(mp.read(y16), F.silu(ref)) < TC_TOL
    # second launch of the same plan ops reuses the arrival counters: must give the same answer
    mp2 = MiniPlan(Wt)
    y2 = mp2.be.conv2d(mp2.input(x), "c", cin, cout, k)
    y3 = mp2.be.conv2d(mp2.input(x), "c", cin, cout, k)
    mp2.run_gpu()
    assert torch.equal(mp2.read(y2), mp2.read(y3))  # deterministic reduction order


@pytest.mark.parametrize("cin,cout,H,W,B,k", [
    (32, 64, 256, 128, 2, 3),    # 512 tiles of 128 pixels > 148 SMs: persistent loop, TMEM double buffering
    (64, 48, 200, 130, 1, 3),    # ragged W (tile tail), N = 48
    (16, 256, 128, 256, 1, 1),   # BLOCK_N 256 (single sub-tile), 256 tiles
])
def test_conv2d_tc_persistent(cin, cout, H, W, B, k, cuda_dev):
    Wt = {"c.weight": _w((cout, cin, k, k), 1), "c.bias": _x((cout,), 2) * 0.1}
    x = _x((B, cin, H, W), 3)
    res = _x((B, cout, H, W), 4)
    mp = MiniPlan(Wt)
    y = mp.be.conv2d(mp.input(x), "c", cin, cout, k, add=mp.input(res))
    mp.run_gpu()
    ref = F.conv2d(x, Wt["c.weight"], Wt["c.bias"], padding=k // 2) + res
    assert rel_err(mp.read(y), ref) < TC_TOL


@pytest.mark.parametrize("cin,cout,k,d,T,B", [(64, 64, 3, 1, 40000, 2), (96, 96, 11, 5, 50000, 1), (48, 48, 7, 3, 70001, 1)])
def test_conv1d_tc_persistent(cin, cout, k, d, T, B, cuda_dev):
    Wt = {"c.weight": _w((cout, cin, k), 1), "c.bias": _x((cout,), 2) * 0.1}
    x = _x((B, cin, 1, T), 3)
    res = _x((B, cout, 1, T), 4)
    mp = MiniPlan(Wt)
    y = mp.be.conv1d(mp.input(x), "c", cin, cout, k, dilation=d, add=mp.input(res))
    mp.run_gpu()
    ref = F.conv1d(x[:, :, 0], Wt["c.weight"], Wt["c.bias"], padding=d * (k // 2), dilation=d) + res[:, :, 0]
    assert rel_err(mp.read(y)[:, :, 0], ref) < TC_TOL


@pytest.mark.parametrize("cin,cout,H,W,B", [(64, 64, 8, 16, 2), (128, 128, 32, 64, 1), (32, 48, 5, 12, 3)])
def test_upsample_conv2d_tc(cin, cout, H, W, B, cuda_dev):
    """nearest-2x + conv3x3 evaluated as four 2x2-tap phase GEMMs with pre-summed weights and strided output."""
    Wt = {"c.weight": _w((cout, cin, 3, 3), 1), "c.bias": _x((cout,), 2) * 0.1}
    x = _x((B, cin, H, W), 3)
    mp = MiniPlan(Wt)
    y = mp.be.upsample_conv2d(mp.input(x), "c", cin, cout)
    mp.run_gpu()
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), Wt["c.weight"], Wt["c.bias"], padding=1)
    assert rel_err(mp.read(y), ref) < TC_TOL


@pytest.mark.parametrize("cin,cout,pad", [(64, 64, "ldm_down"), (128, 128, "same"), (32, 32, "same")])
def test_conv2d_stride2_tc(cin, cout, pad, cuda_dev):
    Wt = {"c.weight": _w((cout, cin, 3, 3), 1), "c.bias": _x((cout,), 2) * 0.1}
    x = _x((2, cin, 16, 32), 3)
    mp = MiniPlan(Wt)
    y = mp.be.conv2d(mp.input(x), "c", cin, cout, 3, stride=2, pad=pad)
    mp.run_gpu()
    if pad == "ldm_down":
        ref = F.conv2d(F.pad(x, (0, 1, 0, 1)), Wt["c.weight"], Wt["c.bias"], stride=2)
    else:
        ref = F.conv2d(x, Wt["c.weight"], Wt["c.bias"], stride=2, padding=1)
    assert rel_err(mp.read(y), ref) < TC_TOL


@pytest.mark.parametrize("cin,cout", [(1, 32), (128, 1), (32, 2), (3, 5)])
def test_conv2d_simt(cin, cout, cuda_dev):
    Wt = {"c.weight": _w((cout, cin, 3, 3), 1), "c.bias": _x((cout,), 2) * 0.1}
    x = _x((2, cin, 16, 24), 3)
    mp = MiniPlan(Wt)
    y = mp.be.conv2d(mp.input(x), "c", cin, cout, 3)
    assert mp.be.ops[-1].code == 2
    mp.run_gpu()
    assert rel_err(mp.read(y), F.conv2d(x, Wt["c.weight"], Wt["c.bias"], padding=1)) < F32_TOL


@pytest.mark.parametrize("cin,H,W,B", [(128, 20, 37, 2), (32, 9, 300, 1), (64, 3, 129, 3)])
def test_conv2d_one_output_channel_f16_input(cin, H, W, B, cuda_dev):
    """vae.decoder.conv_out as the plan runs it: GroupNorm + SiLU (f16 out) -> 3x3 conv to ONE channel (the input-stationary
    kernel: row tiles with ragged ends, zero padding on all four borders, batch items)."""
    Wt = {"n.weight": 1 + 0.1 * _x((cin,), 1), "n.bias": 0.1 * _x((cin,), 2),
          "c.weight": _w((1, cin, 3, 3), 3), "c.bias": _x((1,), 4) * 0.1}
    x = _x((B, cin, H, W), 5)
    mp = MiniPlan(Wt)
    hmid = mp.be.groupnorm(mp.input(x), "n", cin, 32, 1e-6, silu=True)
    y = mp.be.conv2d(hmid, "c", cin, 1, 3)
    assert mp.be.ops[-1].code == 2
    mp.run_gpu()
    mid = mp.read(hmid)     # the f16 values the conv actually consumed
    assert rel_err(mp.read(y), F.conv2d(mid, Wt["c.weight"], Wt["c.bias"], padding=1)) < F32_TOL


@pytest.mark.parametrize("cin,k,T,B", [(48, 7, 1000, 2), (16, 7, 250, 1), (24, 3, 777, 1), (96, 11, 260, 2)])
def test_conv1d_one_output_channel_f16_input(cin, k, T, B, cuda_dev):
    """vocoder.conv_post as the plan runs it: anti-aliased SnakeBeta (f16 out) -> k-tap conv to ONE channel + tanh."""
    Wt = {"a.act.alpha": 0.2 * _x((cin,), 1), "a.act.beta": 0.2 * _x((cin,), 2),
          "c.weight": _w((1, cin, k), 3), "c.bias": _x((1,), 4) * 0.1}
    x = _x((B, cin, 1, T), 5)
    mp = MiniPlan(Wt)
    hmid = mp.be.snake_aa(mp.input(x), "a", cin)
    y = mp.be.conv1d(hmid, "c", cin, 1, k, act="tanh")
    assert mp.be.ops[-1].code == 2
    mp.run_gpu()
    mid = mp.read(hmid)[:, :, 0]
    ref = torch.tanh(F.conv1d(mid, Wt["c.weight"], Wt["c.bias"], padding=k // 2))
    assert rel_err(mp.read(y)[:, :, 0], ref) < F32_TOL


def test_conv2d_f16_transposed_and_rowbias(cuda_dev):
    cin, cout = 64, 128
    Wt = {"c.weight": _w((cout, cin, 1, 1), 1), "c.bias": _x((cout,), 2) * 0.1,
          "e.weight": _w((cout, 32), 5), "e.bias": _x((cout,), 6) * 0.1}
    x = _x((2, cin, 8, 16), 3)
    emb = _x((1, 32, 1, 1), 7)
    mp = MiniPlan(Wt)
    xi = mp.input(x)
    e = mp.be.linear(mp.input(emb), "e", 32, cout, small=True, act="silu")
    y1 = mp.be.conv2d(xi, "c", cin, cout, 1, rowbias=e)
    y2 = mp.be.conv2d(xi, "c", cin, cout, 1, out="f16", transposed=True)
    mp.run_gpu()
    eref = F.silu(F.linear(emb.view(1, 32), Wt["e.weight"], Wt["e.bias"]))
    assert rel_err(mp.read(e).view(1, cout), eref) < F32_TOL
    ref = F.conv2d(x, Wt["c.weight"], Wt["c.bias"])
    assert rel_err(mp.read(y1), ref + eref.view(1, cout, 1, 1)) < TC_TOL
    assert rel_err(mp.read(y2), ref) < TC_TOL


@pytest.mark.parametrize("cin,cout,k,d,T", [(64, 64, 3, 1, 256), (48, 48, 7, 3, 300), (96, 96, 11, 5, 128), (192, 192, 3, 5, 512)])
def test_conv1d_dilated_tc(cin, cout, k, d, T, cuda_dev):
    Wt = {"c.weight": _w((cout, cin, k), 1), "c.bias": _x((cout,), 2) * 0.1}
    x = _x((2, cin, 1, T), 3)
    res = _x((2, cout, 1, T), 4)
    mp = MiniPlan(Wt)
    y = mp.be.conv1d(mp.input(x), "c", cin, cout, k, dilation=d, add=mp.input(res))
    mp.run_gpu()
    ref = F.conv1d(x[:, :, 0], Wt["c.weight"], Wt["c.bias"], padding=d * (k // 2), dilation=d) + res[:, :, 0]
    assert rel_err(mp.read(y)[:, :, 0], ref) < TC_TOL


@pytest.mark.parametrize("cin,cout,u,T", [(128, 64, 5, 32), (64, 32, 2, 200), (96, 48, 6, 64), (64, 32, 3, 40)])
def test_conv_transpose1d_tc(cin, cout, u, T, cuda_dev):
    k = 2 * u
    Wt = {"c.weight": _w((cin, cout, k), 1) * 3, "c.bias": _x((cout,), 2) * 0.1}
    x = _x((2, cin, 1, T), 3)
    add = _x((2, cout, 1, T * u), 4)
    mp = MiniPlan(Wt)
    y = mp.be.convT1d(mp.input(x), "c", cin, cout, k, u, add=mp.input(add))
    mp.run_gpu()
    ref = F.conv_transpose1d(x[:, :, 0], Wt["c.weight"], Wt["c.bias"], stride=u, padding=(k - u) // 2)[..., : T * u] + add[:, :, 0]
    assert rel_err(mp.read(y)[:, :, 0], ref) < TC_TOL


@pytest.mark.parametrize("cin,cout,r,T", [(48, 96, 2, 256), (32, 64, 5, 320), (1, 16, 2, 64)])
def test_conv1d_strided(cin, cout, r, T, cuda_dev):
    Wt = {"c.weight": _w((cout, cin, 2 * r), 1), "c.bias": _x((cout,), 2) * 0.1}
    x = _x((2, cin, 1, T), 3)
    mp = MiniPlan(Wt)
    y = mp.be.conv1d_strided(mp.input(x), "c", cin, cout, 2 * r, r)
    mp.run_gpu()
    ref = F.conv1d(x[:, :, 0], Wt["c.weight"], Wt["c.bias"], stride=r, padding=(r + 1) // 2)
    tol = TC_TOL if mp.be.ops[-1].code == 1 else F32_TOL
    assert rel_err(mp.read(y)[:, :, 0], ref) < tol


@pytest.mark.parametrize("C0,C1,G,silu", [(32, 0, 8, True), (128, 0, 32, False), (64, 32, 32, True), (384, 256, 32, True), (16, 16, 32, False)])
def test_groupnorm(C0, C1, G, silu, cuda_dev):
    Cc = C0 + C1
    Wt = {"n.weight": 1 + 0.1 * _x((Cc,), 1), "n.bias": 0.1 * _x((Cc,), 2)}
    a = _x((2, C0, 8, 12), 3) * 3 + 1.5
    mp = MiniPlan(Wt)
    t = mp.input(a)
    x = a
    if C1:
        b = _x((2, C1, 8, 12), 4) - 0.5
        t = mp.be.concat(t, mp.input(b))
        x = torch.cat([a, b], 1)
    y = mp.be.groupnorm(t, "n", Cc, G, 1e-6, silu=silu)
    mp.run_gpu()
    ref = F.group_norm(x, G, Wt["n.weight"], Wt["n.bias"], 1e-6)
    ref = F.silu(ref) if silu else ref
    assert rel_err(mp.read(y), ref) < 1e-3  # f16 output rounding


@pytest.mark.parametrize("C,H,W,B", [(1024, 64, 32, 1), (512, 64, 32, 2), (256, 64, 32, 3), (1024, 24, 20, 2)])
def test_groupnorm_cluster_sizes(C, H, W, B, cuda_dev):
    """Small-map GroupNorm whose groups are split over a thread-block cluster (8 / 4 / 2 CTAs per group, and a pixel
    count that does not divide by the cluster size); batch items must not influence each other (bit-equal alone)."""
    Wt = {"n.weight": 1 + 0.1 * _x((C,), 1), "n.bias": 0.1 * _x((C,), 2)}
    x = _x((B, C, H, W), 3) * 2 + 0.7
    mp = MiniPlan(Wt)
    y = mp.be.groupnorm(mp.input(x), "n", C, 32, 1e-6, silu=True)
    mp.run_gpu()
    got = mp.read(y)
    ref = F.silu(F.group_norm(x, 32, Wt["n.weight"], Wt["n.bias"], 1e-6))
    assert rel_err(got, ref) < 1e-3
    if B > 1:
        mp1 = MiniPlan(Wt)
        y1 = mp1.be.groupnorm(mp1.input(x[:1]), "n", C, 32, 1e-6, silu=True)
        mp1.run_gpu()
        assert torch.equal(mp1.read(y1), got[:1])


@pytest.mark.parametrize("C0,C1,H,W,B,silu", [(128, 0, 96, 64, 2, True), (64, 64, 83, 53, 1, False), (256, 128, 70, 61, 2, True), (32, 0, 211, 97, 3, True)])
def test_groupnorm_large_maps(C0, C1, H, W, B, silu, cuda_dev):
    """Maps above 4096 pixels take the slab path (gn_stats + gn_finalize + gn_apply): slab tails, pixel counts that are
    not multiples of the unrolled load groups, a virtual concat, and batch items that must not influence each other."""
    Cc = C0 + C1
    Wt = {"n.weight": 1 + 0.1 * _x((Cc,), 1), "n.bias": 0.1 * _x((Cc,), 2)}
    a = _x((B, C0, H, W), 3) * 2 + 0.7

    def build(a_, b_):
        mp = MiniPlan(Wt)
        t = mp.input(a_)
        if b_ is not None:
            t = mp.be.concat(t, mp.input(b_))
        y = mp.be.groupnorm(t, "n", Cc, 32, 1e-6, silu=silu)
        mp.run_gpu()
        return mp.read(y)

    b = (_x((B, C1, H, W), 4) - 0.5) if C1 else None
    x = torch.cat([a, b], 1) if C1 else a
    got = build(a, b)
    ref = F.group_norm(x, 32, Wt["n.weight"], Wt["n.bias"], 1e-6)
    ref = F.silu(ref) if silu else ref
    assert rel_err(got, ref) < 1e-3  # f16 output rounding
    if B > 1:
        assert torch.equal(build(a[:1], b[:1] if C1 else None), got[:1])


def test_layernorm_geglu_softmax_cast_upsample(cuda_dev):
    Cc = 96
    Wt = {"n.weight": 1 + 0.1 * _x((Cc,), 1), "n.bias": 0.1 * _x((Cc,), 2)}
    x = _x((2, Cc, 4, 8), 3) * 2 + 0.3
    mp = MiniPlan(Wt)
    xi = mp.input(x)
    ln = mp.be.layernorm(xi, "n", Cc, 1e-5)
    gg = mp.be.geglu(xi, Cc // 2)
    up = mp.be.upsample2x(xi)
    c16 = mp.be._materialize16(mp.be.concat(xi, mp.be.slice_channels(xi, 8, 16)), "cat")
    ax = mp.be.axpby(xi, xi, 0.25, -1.5)
    sc = mp.be.scale(xi, 1.0 / 3)
    mp.run_gpu()
    xt = x.permute(0, 2, 3, 1)
    assert rel_err(mp.read(ln), F.layer_norm(xt, (Cc,), Wt["n.weight"], Wt["n.bias"], 1e-5).permute(0, 3, 1, 2)) < 1e-3
    assert rel_err(mp.read(gg), (xt[..., :48] * F.gelu(xt[..., 48:])).permute(0, 3, 1, 2)) < 1e-3
    assert rel_err(mp.read(up), F.interpolate(x, scale_factor=2.0, mode="nearest")) < 1e-3
    assert rel_err(mp.read(c16), torch.cat([x, x[:, 8:24]], 1)) < 1e-3
    assert rel_err(mp.read(ax), -1.25 * x) < F32_TOL
    assert rel_err(mp.read(sc), x / 3) < F32_TOL


@pytest.mark.parametrize("S,heads,hd", [(32, 4, 16), (128, 2, 32), (512, 8, 32), (64, 1, 64)])
def test_attention_small(S, heads, hd, cuda_dev):
    Cc = heads * hd
    q, k, v = (_x((2, Cc, 1, S), s) for s in (1, 2, 3))
    mp = MiniPlan()
    o = mp.be.attention(mp.input(q, f16=True), mp.input(k, f16=True), mp.input(v, f16=True), heads, hd)
    mp.run_gpu()

    def sp(t):
        return t.half().float()[:, :, 0].permute(0, 2, 1).reshape(2, S, heads, hd).permute(0, 2, 1, 3)
    ref = torch.softmax(sp(q) @ sp(k).transpose(-1, -2) * hd ** -0.5, -1) @ sp(v)
    ref = ref.permute(0, 2, 1, 3).reshape(2, S, Cc).permute(0, 2, 1)
    assert rel_err(mp.read(o)[:, :, 0], ref) < 2e-3


@pytest.mark.parametrize("S,Cc", [(128, 128), (256, 192), (2048, 128)])
def test_attention_gemm_path(S, Cc, cuda_dev):
    """VAE mid-block attention: QK^T and PV on the tensor-core GEMM with a batch-indexed B operand."""
    Wt = {n + ".weight": _w((Cc, Cc, 1, 1), i) for i, n in enumerate("qkv")}
    Wt.update({n + ".bias": _x((Cc,), 10 + i) * 0.1 for i, n in enumerate("qkv")})
    x = _x((2, Cc, S // 16, 16), 3)
    mp = MiniPlan(Wt)
    xi = mp.input(x)
    q = mp.be.conv2d(xi, "q", Cc, Cc, 1, out="f16")
    k = mp.be.conv2d(xi, "k", Cc, Cc, 1, out="f16")
    v = mp.be.conv2d(xi, "v", Cc, Cc, 1, out="f16", transposed=True)
    o = mp.be.attention(q, k, v, heads=1, head_dim=Cc, v_transposed=True)
    mp.run_gpu()

    def tok(name):
        return F.conv2d(x, Wt[name + ".weight"], Wt[name + ".bias"]).flatten(2).permute(0, 2, 1)
    ref = torch.softmax(tok("q") @ tok("k").transpose(1, 2) * Cc ** -0.5, -1) @ tok("v")
    assert rel_err(mp.read(o).flatten(2).permute(0, 2, 1), ref) < 4e-3


@pytest.mark.parametrize("Cc,T", [(48, 1000), (24, 77), (96, 4096), (8, 13), (7, 50)])
def test_snake_aa(Cc, T, cuda_dev):
    from oracle.flashsr_oracle import TorchBackend
    from egregora_b200 import flashsr_model as M
    Wt = {"a.act.alpha": 0.2 * _x((Cc,), 1), "a.act.beta": 0.2 * _x((Cc,), 2)}
    x = _x((2, Cc, 1, T), 3) * 2
    mp = MiniPlan(Wt)
    y = mp.be.snake_aa(mp.input(x), "a", Cc)
    mp.run_gpu()
    ref = TorchBackend(M.tiny_spec(), Wt).snake_aa(x[:, :, 0], "a", Cc)
    got = mp.read(y)
    assert rel_err(got[:, :, 0], ref) < 1e-3
    if Cc % 2 == 0:
        # even channel counts take the two-channels-per-thread kernel (packed f32x2 arithmetic): every lane must round
        # exactly like the one-channel kernel
        os.environ["EGR_SNAKE_SCALAR"] = "1"
        try:
            mp2 = MiniPlan(Wt)
            y2 = mp2.be.snake_aa(mp2.input(x), "a", Cc)
            mp2.run_gpu()
            assert torch.equal(mp2.read(y2), got)
        finally:
            del os.environ["EGR_SNAKE_SCALAR"]


@pytest.mark.parametrize("spec_name", ["tiny", "full"])
def test_stft_mel_and_lowpass(spec_name, cuda_dev):
    from oracle.flashsr_oracle import TorchBackend
    from egregora_b200 import flashsr_model as M
    spec = M.tiny_spec() if spec_name == "tiny" else M.default_spec()
    T = spec["chunk"]
    g = torch.Generator().manual_seed(5)
    wav = (0.1 * torch.randn(2, T, generator=g)).cumsum(1) * 0.05
    wav = (wav - wav.mean(1, keepdim=True))
    wav = wav / wav.abs().max() * 0.5
    mp = MiniPlan({}, spec=spec)
    wi = mp.input(wav[:, None, None, :])
    mel = mp.be.stft_mel(wi)
    lp = mp.be.lowpass(wi)
    mp.run_gpu()
    ob = TorchBackend(spec, {})
    ref_mel = ob.stft_mel(wav)  # [B,1,T,F]
    got = mp.read(mel)[:, 0]    # plan tensor is [B, frames, n_mels, 1] -> NCHW read gives [B,1,frames,n_mels]
    assert float((got - ref_mel[:, 0]).abs().max()) < 2e-3
    ref_lp = ob.lowpass(wav)
    cut = mp.view(mp.be.cutoff_buf, torch.int32, (2,)).cpu().numpy()
    assert list(cut) == list(ob.cutoff_bins)
    assert float((mp.read(lp)[:, 0, 0] - ref_lp).abs().max()) < 1e-5
