"""CPU-side check of everything the plan builder decides (views, taps, packed weights, crops, buffer reuse):
the op list is executed by the torch interpreter in tests/plan_interp.py and compared with the fp32 oracle."""
import numpy as np
import pytest
import torch

from conftest import load_pkg

load_pkg()
from egregora_b200 import _abi, flashsr_model as M, flashsr_plan as P  # noqa: E402
from oracle import flashsr_oracle as O  # noqa: E402
from plan_interp import Interp  # noqa: E402


def _inputs(spec, B, seed=5):
    g = torch.Generator().manual_seed(seed)
    wav = (0.1 * torch.randn(B, spec["chunk"], generator=g)).cumsum(1) * 0.05
    wav = wav - wav.mean(1, keepdim=True)
    wav = wav / wav.abs().max() * 0.5
    fr = spec["chunk"] // spec["mel"]["hop"]
    noise = torch.randn(B, spec["vae"]["embed_dim"], fr // 8, spec["mel"]["n_mels"] // 8, generator=g)
    return wav, noise


@pytest.mark.parametrize("fuse_qkv,fuse_emb", [(False, False), (True, True)])
def test_tiny_plan_matches_oracle_through_interpreter(fuse_qkv, fuse_emb, monkeypatch):
    """The two plan variants (fused is the default since round 2): EGR_FUSE_QKV=1 runs the three attention projections as one GEMM and feeds the attention
    op strided column blocks (64 ops fewer); EGR_FUSE_EMB=1 runs the time-embedding projections of all ResBlocks as one
    GEMV per step (21 fewer at full size) — same numbers."""
    for var, on in (("EGR_FUSE_QKV", fuse_qkv), ("EGR_FUSE_EMB", fuse_emb)):
        monkeypatch.setenv(var, "1" if on else "0")
    spec = M.tiny_spec()
    W = M.init_weights(spec, 0)
    B, steps, lp = 1, 1, True
    blob = P.WeightBlob()
    be = P.build_plan(spec, W, blob, B, steps, lp)
    assert any(o.name.endswith(".to_qkv") for o in be.ops) == fuse_qkv
    assert any(o.name == "unet.emb_layers_all" for o in be.ops) == fuse_emb
    it = Interp(_abi.K, be.build_ops(), be.ws_bytes, blob.tobytes())
    wav, noise = _inputs(spec, B)

    def view(buf, dt, shape):
        n = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
        return it.ws[buf.offset: buf.offset + n].view(dt).view(*shape)

    view(be.inputs["wav"].f32, torch.float32, wav.shape).copy_(wav)
    view(be.inputs["noise"].f32, torch.float32, noise.permute(0, 2, 3, 1).shape).copy_(noise.permute(0, 2, 3, 1))
    it.run()
    y = view(be.output.f32, torch.float32, wav.shape).clone()
    yo, obe = O.run_flashsr(spec, W, wav, noise, steps=steps, lowpass=lp)
    assert not torch.isnan(y).any()
    rms = float((y - yo).pow(2).mean().sqrt())
    assert rms < 1e-3, rms  # north_star tolerance: FlashSR waveform within 1e-3 RMS
    assert 0.02 < float(yo.pow(2).mean().sqrt()) < 0.5  # the synthetic model is not degenerate / saturated
    assert list(view(be.cutoff_buf, torch.int32, (B,))) == list(obe.cutoff_bins)


def test_param_shapes_cover_all_weights():
    spec = M.tiny_spec()
    shapes = M.param_shapes(spec)
    W = M.init_weights(spec, 0)
    assert list(shapes) == list(W)
    full = M.param_shapes(M.default_spec())
    n = sum(int(np.prod(s)) for s, _ in full.values())
    assert 4e8 < n < 7e8  # ~0.5 B parameters: AudioSR-class VAE + UNet + BigVGAN-class vocoder


def test_ddim_schedule():
    s1 = M.ddim_schedule(1000, 0.008, 1)
    assert len(s1) == 1 and s1[0][0] == 999 and s1[0][2] == 1.0
    s4 = M.ddim_schedule(1000, 0.008, 4)
    assert [t for t, _, _ in s4] == [999, 749, 500, 250] and s4[-1][2] == 1.0
    assert all(s4[i][2] == s4[i + 1][1] for i in range(3))


def test_tile_geometry_and_block_n():
    for (w, h, b) in [(256, 512, 1), (32, 64, 8), (4, 8, 3), (3072, 1, 2), (513, 1, 1), (1, 1, 1)]:
        bw, bh, bb = P.tile_geometry(w, h, b)
        assert bw * bh * bb == 128
    assert [P.pick_block_n(n) for n in (16, 48, 96, 128, 384, 640, 1024, 1920, 4608, 5120)] == [16, 48, 96, 128, 192, 160, 256, 240, 256, 256]


def test_committed_layer_table_matches_the_plan():
    """roofline/flashsr_layers.json (written by tools/make_layer_table.py) is the table bench.py's
    roofline.achieved rests on: it must list exactly the GEMM ops of the c2 plan built now, and its FLOP count must
    agree with the oracle's hook count once the two documented differences are restated (phase GEMMs of the
    up-sampling convs, attention products)."""
    import json
    from conftest import ROOT
    tab = json.loads((ROOT / "roofline" / "flashsr_layers.json").read_text())
    spec = M.default_spec()
    be = P.build_plan(spec, M.init_weights(spec, 0), P.WeightBlob(), 1, 1, True)
    assert len(tab["layers"]) == len(be.layer_table) == tab["totals"]["gemm_ops"]
    for row, l in zip(tab["layers"], be.layer_table):
        assert (row["name"], row["kind"], row["M"], row["N"], row["K"], row["taps"]) == \
               (l["name"], l["kind"], l["M"], l["N"], l["K"], l["taps"])
        assert row["flops"] == l["flops"]
    assert tab["totals"]["tc_flops"] == be.tc_flops
    assert tab["reconciliation"]["relative_difference"] < 1e-4


def test_upsampling_conv_writes_the_f16_copy_for_the_following_shortcut_conv():
    """The four phase GEMMs of an up-sampling conv write the f16 copy that the next block's 1x1 shortcut conv consumes
    (same element indices as their interleaved f32 stores), so the c2 plan holds no separate f32 -> f16 pass over the VAE
    decoder's largest maps (three cast passes, 170 us of a c2 pass, before)."""
    spec = M.default_spec()
    be = P.build_plan(spec, M.init_weights(spec, 0), P.WeightBlob(), 1, 1, True)
    names = [o.name for o in be.ops]
    assert not any(n.startswith("vae.decoder.up.") and n.endswith("nin_shortcut.in16") for n in names)
    n_checked = 0
    for lvl in (1, 2, 3):
        phases = [o for o in be.ops if o.name.startswith(f"vae.decoder.up.{lvl}.upsample.conv.p")]
        assert len(phases) == 4
        shortcut = [o for o in be.ops if o.name == f"vae.decoder.up.{lvl - 1}.block.0.nin_shortcut"]
        if not shortcut:
            continue
        bufs = {id(o.ptr["OUT16"][1]) for o in phases}
        assert len(bufs) == 1 and all("OUT32" in o.ptr for o in phases)      # one shared f16 buffer next to the f32 output
        f16 = phases[0].ptr["OUT16"][1]
        assert shortcut[0].x0[0] is f16                                      # ... which is the shortcut conv's A operand
        assert f16.first == min(o.index for o in phases) and f16.last >= shortcut[0].index
        n_checked += 1
    assert n_checked == 3


def test_groupnorm_slab_count_is_a_function_of_the_pixel_count_only():
    """gn_slabs() must match slab_for() in csrc/ops.cu (the plan sizes the partial-sum buffer with it and the kernel
    refuses a mismatch); the batch size does not enter, so a chunk-channel's statistics are summed in the same order in any
    batch."""
    for P_, want in ((131072, (222, 591)), (32768, (56, 586)), (8192, (16, 512)), (4097, (16, 257)), (20, (16, 2)), (7, (7, 1))):
        assert P.gn_slabs(P_) == want, (P_, P.gn_slabs(P_))
