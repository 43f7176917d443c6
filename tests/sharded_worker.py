"""torchrun worker of tests/test_sharded_nccl_gpu.py: every rank runs EgregoraAudioSuperResolution.run() on the same host
clip with the REAL engine (NCCL, one rank per GPU) and saves what it returned."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def clip(total, channels):
    import bench
    return torch.cat([bench.synth_audio(total, 1, seed=100 + c) for c in range(channels)], 0)


def main():
    out_dir, total, channels, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    os.environ.setdefault("EGREGORA_FLASHSR_RANDOM_INIT", "1")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from conftest import load_pkg
    load_pkg()
    from egregora_b200 import egregora_audio_super_resolution as N
    node = N.EgregoraAudioSuperResolution()
    node.NUM_STEPS = steps
    x = clip(total, channels)
    (res,) = node.run(audio={"waveform": x[None], "sample_rate": 48000}, lowpass_input=True, output_sr="48000")
    np.save(os.path.join(out_dir, f"rank{int(os.environ.get('RANK', '0'))}_of{world}.npy"), res["waveform"][0].numpy())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
