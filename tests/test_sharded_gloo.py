"""N>1 host logic on CPU: two `gloo` ranks run the sharded driver `upscale_48k` (contiguous span blocks per
rank, zero-padded tail block, ONE all_gather, stitch on every rank — SURVEY.md §8e).  The two device kernels
(egr_chunk_gather / egr_wola_stitch) cannot run without a GPU, so this test swaps them for the numpy oracle of
the same functions; what is under test is the partition / padding / gather ordering, which must reproduce the
single-process reference driver bit for bit on every rank.
"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


SEED = 4321


def _chunk_model(c: torch.Tensor, row0: int = 0) -> torch.Tensor:
    # position-dependent, non-linear, deterministic: any mis-ordered or mis-padded chunk shows up in the stitch.
    # It also CONSUMES NOISE keyed by the global chunk-channel row, as the FlashSR engine does (egr_noise_fill's numpy
    # restatement): a sharded driver that restarted its rows at 0 on every rank would change the stitched output.
    from oracle import noise_oracle as NO
    ramp = torch.linspace(0.5, 1.5, c.shape[1], dtype=torch.float32)
    z = torch.from_numpy(NO.noise_rows(SEED, row0, c.shape[0], 4096))
    return torch.tanh(c * 3.0) * ramp + 0.01 * c.flip(1) + 0.05 * z.repeat(1, c.shape[1] // 4096 + 1)[:, :c.shape[1]]


def _worker(rank: int, world: int, port: int, total: int, channels: int, out_dir: str):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    from conftest import load_pkg
    load_pkg()
    from egregora_b200 import egregora_audio_super_resolution as N
    from oracle import driver_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = {"rows": 0}

    def gather_chunks(x, spans, win):
        return torch.from_numpy(O.gather_chunks(x.numpy(), spans, win))

    def wola_stitch(preds, spans, total_, win):
        lst = [(preds[k].numpy(), s, L) for k, (s, L) in enumerate(spans)]
        return torch.from_numpy(O.wola_stitch(lst, total_, win))

    def model(c, row0=0):
        calls["rows"] += c.shape[0]
        calls.setdefault("row0s", []).append(int(row0))
        return _chunk_model(c, row0)

    N.gather_chunks, N.wola_stitch = gather_chunks, wola_stitch
    x = torch.from_numpy((np.random.default_rng(7).standard_normal((channels, total)) * 0.2).astype(np.float32))
    y = N.upscale_48k(x, model, device=torch.device("cpu"))
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), y.numpy())
    np.save(os.path.join(out_dir, f"rows{rank}.npy"), np.array([calls["rows"]]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total,channels", [(1_000_000, 2), (245_760 + 221_760 * 2, 1), (300_000, 2)])
def test_two_rank_sharded_driver_matches_reference_driver(tmp_path, total, channels):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, total, channels, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, str(ROOT))
    from oracle import driver_oracle as O
    x = (np.random.default_rng(7).standard_normal((channels, total)) * 0.2).astype(np.float32)
    # single-process reference driver: one model call per span, rows numbered span-major (span k, channel c -> k*C + c)
    state = {"row": 0}

    def ref_model(c):
        y = _chunk_model(torch.from_numpy(np.ascontiguousarray(c)), state["row"]).numpy()
        state["row"] += c.shape[0]
        return y

    want, _ = O.run_driver(x, 48000, ref_model)
    win, hop = O.win_hop()
    n = len(O.iter_chunks(total, win, hop))
    rows = 0
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npy")
        assert got.shape == want.shape
        assert np.array_equal(got, want), f"rank {r}: max diff {np.abs(got - want).max()}"
        rows += int(np.load(tmp_path / f"rows{r}.npy")[0])
    assert rows == n * channels  # every chunk-channel evaluated exactly once across the ranks (no replicated work)
