"""egr_eval_hf_band on the device (_band_energy_hi_db, egregora_null_test_suite.py:190-197) through
`egregora_eval_metrics.hf_band_db`, against goldens made by the reference function (tests/golden/make_eval_hf_golden.py).
Written after this round's GPU budget was spent; verified under the CPU emulator (tests/test_cusim.py); collected last and
verified on hardware at the end of round 1 (plain tests since round 2) like the LSD / LUFS hardware tests."""
import json

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from hf_cases import signal

pytestmark = pytest.mark.gpu


def test_hf_band_kernel_matches_reference_golden(cuda_dev, pkg):
    from egregora_b200 import egregora_eval_metrics as M
    hg = json.loads((GOLDEN / "eval_hf_golden.json").read_text())
    for name, c in hg.items():
        got = M.hf_band_db(torch.from_numpy(signal(name, c)), c["sr"], c["lo_hz"])
        assert abs(got - c["hf_db"]) <= 1e-3, (name, got, c["hf_db"])


def test_hf_band_properties_at_clip_scale(cuda_dev, pkg):
    """c5-sized clip (5 min stereo, N = 14.4 M = 2^9 3^2 5^5): white noise has its energy spread evenly, so the ratio
    above f is (1 - 2f/sr); the ratio does not depend on gain; lo_hz = 0 gives 0 dB."""
    from egregora_b200 import egregora_eval_metrics as M
    g = torch.Generator().manual_seed(21)
    x = (torch.randn((2, 48000 * 300), generator=g) * 0.1).to(cuda_dev)
    v = M.hf_band_db(x, 48000, 8000)
    assert abs(v - 10 * np.log10(1 - 8000 / 24000)) < 0.02
    assert abs(M.hf_band_db(x * 0.25, 48000, 8000) - v) < 1e-4
    assert abs(M.hf_band_db(x, 48000, 0.0)) < 1e-9
