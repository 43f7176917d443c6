"""egr_eval_lufs on the device (integrated_lufs + _k_weight of the reference's analysis nodes,
egregora_null_test_suite.py:125-164) through `egregora_eval_metrics.integrated_lufs`, against goldens made by the
reference functions (tests/golden/make_eval_lufs_golden.py).

STATUS: written after this round's GPU budget was spent.  The kernels' real source runs under the CPU emulator
(tests/test_cusim.py::test_eval_lufs_reference_golden: filtered signal bit-identical to the reference, loudness to 1e-9
dB); the tests below are the first hardware run, collected last and verified on hardware at the end of round 1 (plain tests since round 2) like the LSD ones.
"""
import json

import pytest
import torch

from conftest import GOLDEN
from lufs_cases import signal

pytestmark = pytest.mark.gpu


def test_lufs_kernel_matches_reference_golden(cuda_dev, pkg):
    from egregora_b200 import egregora_eval_metrics as M
    lg = json.loads((GOLDEN / "eval_lufs_golden.json").read_text())
    for name, c in lg.items():
        got = M.integrated_lufs(torch.from_numpy(signal(name, c)), c["sr"])
        assert abs(got - c["lufs"]) <= 1e-9, (name, got, c["lufs"])


def test_lufs_properties_at_clip_scale(cuda_dev, pkg):
    """c5-sized clip (5 min stereo at 48 kHz — minutes of Python looping in the reference): a gain of g moves the
    loudness by 20*log10(g) (the whole chain is homogeneous), results are deterministic, silence floors at the epsilon."""
    from egregora_b200 import egregora_eval_metrics as M
    g = torch.Generator().manual_seed(12)
    x = (torch.randn((2, 48000 * 300), generator=g) * 0.1).to(cuda_dev)
    a, b = M.integrated_lufs(x, 48000), M.integrated_lufs(x, 48000)
    assert a == b and -40.0 < a < 0.0
    assert abs(M.integrated_lufs(x * 0.5, 48000) - (a - 6.020599913279624)) < 1e-4
    assert abs(M.integrated_lufs(torch.zeros(1, 96000), 48000) - (-0.691 - 200.0)) < 1e-9
    with pytest.raises(RuntimeError):
        M.integrated_lufs(torch.zeros(2, 0), 48000)
