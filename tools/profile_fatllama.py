"""ncu target for path B: c4-sized loop with few iterations (run with --profile-from-start off)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
bench.load_pkg()
from egregora_b200 import _abi  # noqa: E402
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
S = int(sys.argv[2]) if len(sys.argv) > 2 else 7938000
C = 2
dev = torch.device("cuda", 0)
lib = _abi.init(0)
x = (bench.synth_audio(S, C, sr=44100, seed=77) * 32767).round().to(dev)
y = torch.empty_like(x)
wb = lib.egr_fatllama_workspace_bytes(C, S, 1)
w = torch.empty(wb, dtype=torch.uint8, device=dev)
for i in range(2):
    if i == 1:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    _abi.check(lib.egr_fatllama_run(x.data_ptr(), y.data_ptr(), C, S, 1, iters, 0.6, 3, w.data_ptr(), wb, 0))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(y.abs().max()))
