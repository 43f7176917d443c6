"""The UNet as one persistent kernel (csrc/mega.cu): every denoising step of a plan is a single cooperative launch that
walks the UNet's ops with grid barriers in between.  Checked here: the plan really contains such runs, no barrier watchdog
fired, the whole plan still replays as a CUDA graph, the result matches the fp32 oracle (north_star: 1e-3 RMS) and the
one-launch-per-op path (EGR_NO_MEGA=1) to f16 noise, and it is deterministic.  Reference call site of the model:
egregora_audio_super_resolution.py:366-369."""
import ctypes as C

import pytest
import torch

from conftest import load_pkg

load_pkg()
pytestmark = pytest.mark.gpu


def _inputs(spec, B, seed=5):
    g = torch.Generator().manual_seed(seed)
    wav = (0.1 * torch.randn(B, spec["chunk"], generator=g)).cumsum(1) * 0.05
    wav = wav - wav.mean(1, keepdim=True)
    return wav / wav.abs().max() * 0.5


def _mega_info(eng, handle):
    out = (C.c_int * 64)()
    eng.lib.egr_debug_mega_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int]
    eng.lib.egr_debug_mega_aborted.argtypes = [C.c_void_p]
    eng.lib.egr_debug_plan_graphed.argtypes = [C.c_void_p]
    assert eng.lib.egr_debug_mega_info(handle, out, 8) == 0
    runs = [tuple(out[1 + 7 * r: 8 + 7 * r]) for r in range(min(out[0], 8))]
    return out[0], runs, eng.lib.egr_debug_mega_aborted(handle), eng.lib.egr_debug_plan_graphed(handle)


@pytest.mark.parametrize("B,steps", [(1, 1), (2, 2), (3, 4)])
def test_tiny_unet_runs_as_one_kernel_per_region(cuda_dev, monkeypatch, B, steps):
    from egregora_b200 import flashsr_model as M
    from egregora_b200.flashsr_engine import FlashSREngine
    from oracle import flashsr_oracle as O
    spec = M.tiny_spec()
    W = M.init_weights(spec, 0)
    wav = _inputs(spec, B, seed=20 + B)
    monkeypatch.delenv("EGR_NO_MEGA", raising=False)
    eng = FlashSREngine(cuda_dev, spec, W, max_batch=4)
    noise = eng.make_noise(B, 4321)
    ys = [eng.infer(wav.to(cuda_dev), lowpass=True, steps=steps, noise=noise).cpu() for _ in range(3)]
    be, handle = eng.plan(B, steps, True)
    n_runs, runs, aborted, graphed = _mega_info(eng, handle)
    assert n_runs >= 1 and aborted == 0, (n_runs, runs, aborted)
    assert graphed == 1                                   # eager, capture, replay: the plan is a CUDA graph again
    in_runs = sum(r[1] - r[0] for r in runs)   # plan ops covered (a run's own op count differs: split-K reduce passes are extra ops, GroupNorm pairs fuse)
    flagged = sum(1 for o in be.ops if o.flags & 1)
    assert in_runs >= 0.9 * flagged > 0, (in_runs, flagged, runs)   # nearly every flagged op is inside a persistent run
    assert torch.equal(ys[0], ys[1]) and torch.equal(ys[1], ys[2])    # eager pass, captured pass, replayed pass
    yo, _ = O.run_flashsr(spec, W, wav, noise.cpu(), steps=steps, lowpass=True)
    rms = float((ys[0] - yo).pow(2).mean().sqrt())
    assert rms < 1e-3, rms
    eng.close()
    # the one-launch-per-op path computes the same network (GroupNorm reduction orders differ: f16-noise level agreement)
    monkeypatch.setenv("EGR_NO_MEGA", "1")
    eng2 = FlashSREngine(cuda_dev, spec, W, max_batch=4)
    y2 = eng2.infer(wav.to(cuda_dev), lowpass=True, steps=steps, noise=noise).cpu()
    _, handle2 = eng2.plan(B, steps, True)
    assert _mega_info(eng2, handle2)[0] == 0
    assert float((y2 - yo).pow(2).mean().sqrt()) < 1e-3
    assert float((y2 - ys[0]).pow(2).mean().sqrt()) < 1e-3
    eng2.close()
