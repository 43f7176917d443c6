"""GPU parity of the hand-written FFT (egr_fft_exec, no cuFFT) against numpy's f64 FFT.  Tolerance: relative
RMS 3e-6 (float32 transform, f64-computed twiddle tables)."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import load_pkg

load_pkg()
from egregora_b200 import _abi  # noqa: E402

pytestmark = pytest.mark.gpu


def _fft(x: np.ndarray, inverse=False, scale=True):
    """x [batch, n] complex64 -> transformed copy, through the C ABI."""
    lib = _abi.init(0)
    batch, n = x.shape
    plan = C.c_void_p()
    _abi.check(lib.egr_fft_plan_create(n, batch, C.byref(plan)), "egr_fft_plan_create")
    try:
        d = torch.from_numpy(np.ascontiguousarray(x).view(np.float32).reshape(batch, 2 * n).copy()).cuda()
        wb = lib.egr_fft_plan_workspace_bytes(plan)
        w = torch.empty(max(wb, 16), dtype=torch.uint8, device="cuda")
        _abi.check(lib.egr_fft_exec(plan, d.data_ptr(), w.data_ptr(), int(inverse), int(scale), 0), "egr_fft_exec")
        torch.cuda.synchronize()
        return d.cpu().numpy().view(np.complex64).reshape(batch, n), lib.egr_fft_plan_passes(plan)
    finally:
        lib.egr_fft_plan_destroy(plan)


def _rel(a, b):
    return float(np.sqrt(np.mean(np.abs(a - b) ** 2)) / (np.sqrt(np.mean(np.abs(b) ** 2)) + 1e-30))


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 11, 13, 16, 60, 210, 1000, 2100, 4096, 8192, 8190, 10000, 30030,
                               80000, 122880, 1890 * 2100])
def test_forward_smooth_lengths(n, cuda_dev):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n))).astype(np.complex64)
    got, _ = _fft(x)
    ref = np.fft.fft(x.astype(np.complex128), axis=1)
    assert _rel(got, ref) < 3e-6


@pytest.mark.parametrize("n", [17, 10007, 2 * 10007, 123457])
def test_forward_bluestein_lengths(n, cuda_dev):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal((1, n)) + 1j * rng.standard_normal((1, n))).astype(np.complex64)
    got, _ = _fft(x)
    ref = np.fft.fft(x.astype(np.complex128), axis=1)
    assert _rel(got, ref) < 1e-5


@pytest.mark.parametrize("n", [6, 1000, 80000, 10007])
def test_inverse_and_round_trip(n, cuda_dev):
    rng = np.random.default_rng(n + 1)
    x = (rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n))).astype(np.complex64)
    inv, _ = _fft(x, inverse=True, scale=True)
    assert _rel(inv, np.fft.ifft(x.astype(np.complex128), axis=1)) < 1e-5
    fwd, _ = _fft(x)
    back, _ = _fft(fwd, inverse=True, scale=True)
    assert _rel(back, x) < 1e-5
    raw, _ = _fft(x, inverse=True, scale=False)
    assert _rel(raw / n, inv) < 1e-6


def test_linearity_and_impulse_at_full_size(cuda_dev):
    """c4-sized transform (M = 3 969 000 = 1890 x 2100): impulse -> pure phase ramp, and linearity."""
    n = 1890 * 2100
    x = np.zeros((1, n), np.complex64)
    x[0, 12345] = 1.0
    got, passes = _fft(x)
    k = np.arange(n)
    ref = np.exp(-2j * np.pi * ((k * 12345) % n) / n)
    assert np.max(np.abs(got[0] - ref)) < 2e-5
    assert passes == 3
