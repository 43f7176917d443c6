import importlib.util
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
PKG_DIR = ROOT / "comfyui-egregora-audio-super-resolution_b200"
GOLDEN = ROOT / "tests" / "golden"
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
# No FlashSR checkpoint exists in this environment: tests that go through the node run on seeded random weights and say so
# explicitly (the node itself raises without this switch; tests/test_weights.py and test_checkpoint_gpu.py cover loading).
import os  # noqa: E402
os.environ.setdefault("EGREGORA_FLASHSR_RANDOM_INIT", "1")


def load_pkg():
    """Import the (hyphenated) package directory under the alias `egregora_b200`, as ComfyUI does by path."""
    if "egregora_b200" in sys.modules:
        return sys.modules["egregora_b200"]
    spec = importlib.util.spec_from_file_location("egregora_b200", PKG_DIR / "__init__.py",
                                                  submodule_search_locations=[str(PKG_DIR)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["egregora_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def golden():
    return np.load(GOLDEN / "driver_golden.npz")


@pytest.fixture(scope="session")
def cuda_dev():
    import os
    import torch
    if os.environ.get("EGR_TEST_INTERP") or os.environ.get("EGR_TEST_CUSIM"):
        return torch.device("cpu")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    load_pkg()
    from egregora_b200 import _abi
    _abi.init(0)
    return torch.device("cuda", 0)


@pytest.fixture(scope="session")
def synthetic_ckpt_dir(tmp_path_factory):
    """Full-size synthetic checkpoint files in upstream naming (tools/make_synthetic_ckpt.py) -> (dir, weights)."""
    load_pkg()
    sys.path.insert(0, str(ROOT / "tools"))
    import make_synthetic_ckpt as S
    from egregora_b200 import flashsr_model as M
    spec = M.default_spec()
    W = M.init_weights(spec, 7)
    d = tmp_path_factory.mktemp("flashsr_ckpt")
    S.write(d, spec, W)
    return d, W
