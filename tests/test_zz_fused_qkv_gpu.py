"""The opt-in EGR_FUSE_QKV=1 plan variant on the device: to_q / to_k / to_v of every UNet attention layer as ONE GEMM
(N = 3C) whose output attn_rows_kernel<HD, STRIDED=true> reads as column blocks — 64 launches fewer per UNet pass (c2 plan:
924 -> 860 ops).  Off by default: written after the round's GPU budget was spent, so it is unmeasured; its numerics are
pinned on the CPU (plan interpreter: tests/test_plan_cpu.py, real kernel under the emulator: tests/test_cusim.py).  These
are the first hardware runs — collected last, verified on hardware at the end of round 1 (plain tests since round 2); XPASS = verified, then measure with
`EGR_FUSE_QKV=1 EGR_FUSE_EMB=1 python bench.py` against the default (EGR_FUSE_EMB: the 22 ResBlock time-embedding GEMVs as
one per diffusion step; both together 924 -> 839 ops)."""
import pytest
import torch

from harness import MiniPlan, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("S,heads,hd", [(32, 4, 16), (128, 2, 32), (512, 8, 32)])
def test_attention_reads_fused_qkv_column_blocks(S, heads, hd, cuda_dev, pkg):
    from egregora_b200.flashsr_plan import PT
    Cc, B = heads * hd, 2
    qkv = torch.randn((B, 3 * Cc, 1, S), generator=torch.Generator().manual_seed(S))
    mp = MiniPlan()
    t = mp.input(qkv, f16=True)
    q, k, v = (PT(B, 1, S, Cc, f16=t.f16, ld=3 * Cc, coff=i * Cc) for i in range(3))
    o = mp.be.attention(q, k, v, heads, hd)
    mp.run_gpu()

    def sp(i):
        x = qkv[:, i * Cc:(i + 1) * Cc].half().float()[:, :, 0].permute(0, 2, 1)
        return x.reshape(B, S, heads, hd).permute(0, 2, 1, 3)
    ref = torch.softmax(sp(0) @ sp(1).transpose(-1, -2) * hd ** -0.5, -1) @ sp(2)
    ref = ref.permute(0, 2, 1, 3).reshape(B, S, Cc).permute(0, 2, 1)
    assert rel_err(mp.read(o)[:, :, 0], ref) < 2e-3


@pytest.mark.parametrize("emb", [False, True])
def test_tiny_e2e_with_fused_qkv(cuda_dev, pkg, monkeypatch, emb):
    """emb: additionally EGR_FUSE_EMB=1 (all ResBlock time-embedding projections as one GEMV per step)."""
    from egregora_b200 import flashsr_model as M
    from egregora_b200.flashsr_engine import FlashSREngine
    from oracle import flashsr_oracle as O
    monkeypatch.setenv("EGR_FUSE_QKV", "1")
    monkeypatch.setenv("EGR_FUSE_EMB", "1" if emb else "0")
    spec = M.tiny_spec()
    W = M.init_weights(spec, 0)
    eng = FlashSREngine(cuda_dev, spec, W, max_batch=2)
    be, _ = eng.plan(2, 1, True)
    assert any(o.name.endswith(".to_qkv") for o in be.ops)
    assert any(o.name == "unet.emb_layers_all" for o in be.ops) == emb
    g = torch.Generator().manual_seed(9)
    wav = (0.1 * torch.randn(2, spec["chunk"], generator=g)).cumsum(1) * 0.05
    wav = wav - wav.mean(1, keepdim=True)
    wav = wav / wav.abs().max() * 0.5
    noise = eng.make_noise(2, 4321)
    y = eng.infer(wav.to(cuda_dev), lowpass=True, steps=1, noise=noise).cpu()
    yo, _ = O.run_flashsr(spec, W, wav, noise, steps=1, lowpass=True)
    assert float((y - yo).pow(2).mean().sqrt()) < 1e-3
    eng.close()
