"""Where does an end-to-end node.run() step go when the spans are sharded over ranks?  torchrun --nproc-per-node N.
Host wall clock around each phase (with a device sync after it), per rank, plus raw all_gather latency at two sizes."""
import os, sys, time
from pathlib import Path
import torch, torch.distributed as dist
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ.setdefault("EGREGORA_FLASHSR_RANDOM_INIT", "1")
import bench
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
bench.load_pkg()
from egregora_b200 import egregora_audio_super_resolution as N
engine = N.get_engine(dev)
node = N.EgregoraAudioSuperResolution(); node.NUM_STEPS = 1
win, hop = N._win_hop()
total = win + (world - 1) * hop
x_host = bench.synth_audio(total, 1).pin_memory()
audio = {"waveform": x_host[None], "sample_rate": N.REQ_SR}
for _ in range(4):
    node.run(audio=audio, lowpass_input=True, output_sr="48000")
torch.cuda.synchronize(); dist.barrier()

def timed(fn, reps=10):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    ts.sort(); return ts[len(ts) // 2], ts[0], ts[-1]

print(f"[r{rank}] node.run            med/min/max ms", timed(lambda: node.run(audio=audio, lowpass_input=True, output_sr="48000")), flush=True)
x_dev = x_host.to(dev)
cm = lambda c, row0=0: engine.infer(c, lowpass=True, steps=1, seed=4321, row0=row0)
print(f"[r{rank}] upscale_48k(device)  ", timed(lambda: N.upscale_48k(x_dev, cm)), flush=True)
print(f"[r{rank}] upscale_48k(host x)  ", timed(lambda: N.upscale_48k(x_host, cm, device=dev)), flush=True)
out = N.upscale_48k(x_dev, cm)
print(f"[r{rank}] _make_audio          ", timed(lambda: N._make_audio(48000, out)), flush=True)
print(f"[r{rank}] infer B=1            ", timed(lambda: cm(x_dev[:, :win])), flush=True)
for nbytes in (1 << 20, 128 << 20):
    loc = torch.zeros(nbytes // 4, device=dev); g = torch.empty(world * nbytes // 4, device=dev)
    for _ in range(3): dist.all_gather_into_tensor(g, loc)
    print(f"[r{rank}] all_gather {nbytes >> 20} MiB/rank", timed(lambda: dist.all_gather_into_tensor(g, loc)), flush=True)
# phases of one sharded pass by CUDA events
node._marks = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier(); torch.cuda.synchronize()
e0.record(); node.run(audio=audio, lowpass_input=True, output_sr="48000"); e1.record(); torch.cuda.synchronize()
prev = e0; ph = {}
for name, ev in node._marks:
    ph[name] = round(prev.elapsed_time(ev), 3); prev = ev
ph["d2h"] = round(prev.elapsed_time(e1), 3)
print(f"[r{rank}] phases", ph, flush=True)
dist.destroy_process_group()
