"""CPU-side: the C-ABI library loads and exports every symbol include/egregora_b200.h declares."""
import ctypes

from conftest import load_pkg

load_pkg()
from egregora_b200 import _abi  # noqa: E402


def test_library_loads_and_exports_header_symbols():
    lib = _abi.load()
    names = _abi.exported_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert lib.egr_abi_version() == _abi.K["EGR_ABI_VERSION"]


def test_struct_sizes_match():
    lib = _abi.load()
    assert lib.egr_sizeof(0) == ctypes.sizeof(_abi.Tensor)
    assert lib.egr_sizeof(1) == ctypes.sizeof(_abi.Op)


def test_no_gpu_fails_loudly():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _abi.load()
    assert lib.egr_init(0) != 0
    assert b"no CUDA device" in lib.egr_last_error() or b"CPU fallback" in lib.egr_last_error()
