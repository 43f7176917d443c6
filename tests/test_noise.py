"""Diffusion start noise (egr_noise_fill): Philox4x32-10 known answers, the oracle's statistics, the kernel's real source
under the CPU emulator, and (GPU) the kernel itself + row-offset consistency — the property that makes an N-GPU run equal
the 1-GPU run (upstream draws unseeded torch.randn inside forward, so this boundary is parity-unpinned)."""
import ctypes as C
import shutil
import sys

import numpy as np
import pytest

from conftest import ROOT, load_pkg

load_pkg()
from oracle import noise_oracle as NO  # noqa: E402


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32 10 rounds."""
    f = lambda c, k: [int(v) for v in NO.philox4x32_10(np.array([c], np.uint32), k)[0]]  # noqa: E731
    assert f([0, 0, 0, 0], (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert f([0xffffffff] * 4, (0xffffffff, 0xffffffff)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert f([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], (0xa4093822, 0x299f31d0)) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_oracle_is_standard_normal_and_row_keyed():
    z = NO.noise_rows(4321, 0, 6, 32768)
    assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1.0) < 0.01 and np.isfinite(z).all()
    assert abs(float(np.mean(z ** 4)) - 3.0) < 0.1                        # kurtosis of a Gaussian
    assert np.array_equal(NO.noise_rows(4321, 4, 1, 32768)[0], z[4])      # a row depends on its global index only
    assert not np.array_equal(z[0], z[1]) and not np.array_equal(NO.noise_rows(4322, 0, 1, 64), z[:1, :64])
    assert abs(float(np.corrcoef(z[0], z[1])[0, 1])) < 0.02


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ is needed to build the emulator")
def test_kernel_source_under_emulator():
    sys.path.insert(0, str(ROOT / "tests" / "cusim"))
    import build as cusim_build
    from egregora_b200 import _abi
    lib = C.CDLL(str(cusim_build.build()))
    fn = lib.egr_noise_fill
    fn.restype, fn.argtypes = _abi.signatures()["egr_noise_fill"]
    assert lib.egr_init(0) == 0
    for seed, row0, rows, elems in ((4321, 0, 3, 1000), (2 ** 40 + 7, 259, 2, 32768), (1, 5, 1, 3)):
        out = np.full((rows, elems), np.nan, np.float32)
        assert fn(seed, row0, rows, elems, out.ctypes.data, None) == 0
        want = NO.noise_rows(seed, row0, rows, elems)
        assert np.max(np.abs(out - want)) < 2e-6, (seed, row0)


@pytest.mark.gpu
def test_kernel_matches_oracle_and_rows_are_position_independent(cuda_dev):
    import torch
    from egregora_b200 import flashsr_model as M
    from egregora_b200.flashsr_engine import FlashSREngine
    spec = M.tiny_spec()
    eng = FlashSREngine(cuda_dev, spec, M.init_weights(spec, 0), max_batch=2)
    n = eng.make_noise(5, 4321, row0=3)
    torch.cuda.synchronize()
    want = NO.noise_rows(4321, 3, 5, n[0].numel()).reshape(n.shape)
    assert float((n.cpu() - torch.from_numpy(want)).abs().max()) < 2e-6
    assert torch.equal(eng.make_noise(1, 4321, row0=6)[0], n[3])        # same global row -> same bits, any batch position
    assert not torch.equal(eng.make_noise(1, 4322, row0=6)[0], n[3])
