"""N-GPU == 1-GPU with the REAL engine (reference loop egregora_audio_super_resolution.py:407-420): the same
clip through node.run() under torchrun with 2 NCCL ranks and in a single process gives torch.equal outputs on every rank —
split-K is a per-layer constant, the diffusion noise is keyed by the global chunk-channel row, the stitch is bit-exact.
Needs 2 GPUs (gpurun --gpus 2); on a 1-GPU box the two-rank part is skipped and the row-offset property is covered by
tests/test_flashsr_gpu.py::test_full_spec_rows_are_bit_identical_alone_and_batched."""
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_node_output_equals_single_gpu(cuda_dev, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    total, channels, steps = 245760 + 4 * 221760 + 5000, 2, 2      # 6 spans (ragged tail) x 2 channels
    worker = str(ROOT / "tests" / "sharded_worker.py")
    args = [str(tmp_path), str(total), str(channels), str(steps)]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), worker] + args, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    r1 = subprocess.run([sys.executable, worker] + args, capture_output=True, text=True, timeout=900)
    assert r1.returncode == 0, r1.stderr[-3000:]
    want = np.load(tmp_path / "rank0_of1.npy")
    assert want.shape == (channels, total) and np.isfinite(want).all()
    for rank in range(2):
        got = np.load(tmp_path / f"rank{rank}_of2.npy")
        assert np.array_equal(got, want), f"rank {rank}: max diff {np.abs(got - want).max()}"
