"""GPU parity of the path-A driver kernels (egr_chunk_gather / egr_wola_stitch) through the C ABI:
bit-exact against the oracle (numpy restatement pinned to the reference's own outputs)."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import driver_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", list("abcdef"))
def test_wola_golden_cases(case, golden, cuda_dev, pkg):
    from egregora_b200 import egregora_audio_super_resolution as N
    total, w, hp, C, lpred = (int(v) for v in golden[f"wola_{case}_meta"])
    spans = [(int(s), int(L)) for s, L in golden[f"wola_{case}_spans"]]
    preds = torch.from_numpy(golden[f"wola_{case}_preds"]).to(cuda_dev)
    out = N.wola_stitch(preds, spans, total, w).cpu().numpy()
    assert np.array_equal(out, golden[f"wola_{case}_out"])  # bit-exact vs the reference's _wola_stitch


def test_wola_empty(cuda_dev, pkg, golden):
    from egregora_b200 import egregora_audio_super_resolution as N
    out = N.wola_stitch(torch.zeros(0, 1, 64, device=cuda_dev), [], 5, 64)
    assert np.array_equal(out.cpu().numpy(), golden["wola_empty"])


@pytest.mark.parametrize("total,C", [(480000, 2), (245760, 1), (245761, 1), (1000003, 3), (100, 1)])
def test_gather_and_wola_vs_oracle_real_window(total, C, cuda_dev, pkg):
    from egregora_b200 import egregora_audio_super_resolution as N
    rng = np.random.default_rng(total)
    x = (rng.standard_normal((C, total)) * 0.1).astype(np.float32)
    win, hop = O.win_hop()
    spans = O.iter_chunks(total, win, hop)
    assert spans == N._iter_chunks(total, win, hop)
    xd = torch.from_numpy(x).to(cuda_dev)
    chunks = N.gather_chunks(xd, spans, win)
    assert np.array_equal(chunks.cpu().numpy(), O.gather_chunks(x, spans, win))
    y = (chunks * 0.5).contiguous()  # some "model"
    got = N.wola_stitch(y, spans, total, win).cpu().numpy()
    want = O.wola_stitch([(y[k].cpu().numpy(), s, L) for k, (s, L) in enumerate(spans)], total, win)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("model", ["identity", "gain_roll", "short"])
def test_driver_matches_reference_hashes(model, golden, cuda_dev, pkg):
    """The whole run() hot loop with an injected chunk model reproduces the REFERENCE's output bit for bit."""
    from egregora_b200 import egregora_audio_super_resolution as N
    rng = np.random.default_rng(1234)
    for total, w, hp, C, lpred in [(1000, 64, 48, 2, 64), (777, 64, 48, 1, 64), (500, 64, 32, 3, 64),
                                   (300, 64, 48, 2, 50), (64, 64, 48, 2, 64), (10, 64, 48, 1, 64)]:
        for _ in O.iter_chunks(total, w, hp):
            rng.standard_normal((C, lpred))
    x = (rng.standard_normal((2, 480000)) * 0.1).astype(np.float32)
    fn = {"identity": lambda c: c, "gain_roll": lambda c: 0.5 * torch.roll(c, 3, dims=1),
          "short": lambda c: c[:, :200000].contiguous()}[model]
    y = N.upscale_48k(torch.from_numpy(x).to(cuda_dev), fn).cpu().numpy()
    assert np.array_equal(y[:, golden["driver_probe_idx"]], golden[f"driver_{model}_probe"])
    assert np.array_equal(np.frombuffer(hashlib.sha256(y.tobytes()).digest(), np.uint8), golden[f"driver_{model}_sha256"])


def test_wola_full_size_identity_property(cuda_dev, pkg):
    """c3 size (10 min stereo, 130 spans): identity model => output == input except sample 0 (hann[0] = 0)."""
    from egregora_b200 import egregora_audio_super_resolution as N
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((2, 28_800_000), generator=g, device=cuda_dev) * 0.1
    y = N.upscale_48k(x, lambda c: c)
    assert y.shape == x.shape and bool((y[:, 0] == 0).all())
    assert float((y[:, 1:] - x[:, 1:]).abs().max()) < 1e-7


def test_pcm16_and_absmax(cuda_dev, pkg):
    from egregora_b200 import _abi
    from oracle import fat_llama_oracle as FO
    lib = _abi.init(0)
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-1.2, 1.2, 100001), [1.0, -1.0, 0.5 / 32768, 1.5 / 32768, 2.5 / 32768, 0.99999]]).astype(np.float32)
    xd = torch.from_numpy(x).to(cuda_dev)
    q = torch.empty(x.shape, dtype=torch.int16, device=cuda_dev)
    _abi.check(lib.egr_pcm16_quantize(xd.data_ptr(), q.data_ptr(), x.size, 0))
    assert np.array_equal(q.cpu().numpy(), FO.pcm16_write(x))
    f = torch.empty(x.shape, dtype=torch.float32, device=cuda_dev)
    _abi.check(lib.egr_pcm16_to_float(q.data_ptr(), f.data_ptr(), x.size, 1.0 / 32768.0, 0))
    assert np.array_equal(f.cpu().numpy(), FO.pcm16_read(FO.pcm16_write(x)))
    m = torch.empty(1, dtype=torch.float32, device=cuda_dev)
    _abi.check(lib.egr_absmax(xd.data_ptr(), x.size, m.data_ptr(), 0))
    assert float(m.item()) == float(np.max(np.abs(x)))
