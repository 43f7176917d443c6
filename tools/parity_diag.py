"""Where does the f16 error of the full-size FlashSR plan come from?  (GPU, ~1 min.)
  1. layer-wise relative error of every named intermediate of the default-spec plan against the fp32 oracle trace
  2. attribution: the oracle re-run with every GEMM weight rounded to f16 (same rounding the plan applies), so
       |cuda - oracle(f16 weights)|  = what the activation roundings (f16 A operands) + kernel arithmetic contribute
       |oracle(f16 weights) - oracle| = what the weight rounding alone contributes
Writes gpurun_out/parity_diag.json."""
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
os.environ.setdefault("EGREGORA_FLASHSR_RANDOM_INIT", "1")
import bench  # noqa: E402

bench.load_pkg()
from egregora_b200 import flashsr_model as M  # noqa: E402
from egregora_b200.flashsr_engine import FlashSREngine  # noqa: E402
from oracle import flashsr_oracle as O  # noqa: E402

dev = torch.device("cuda", 0)
spec = M.default_spec()
W = M.init_weights(spec, 0)
steps, lowpass = int(sys.argv[1]) if len(sys.argv) > 1 else 1, True
eng = FlashSREngine(dev, spec, W, debug=True, max_batch=1)
wav = bench.synth_audio(spec["chunk"], 1)
noise = eng.make_noise(1, 4321, 0)
y = eng.infer(wav.to(dev), lowpass=lowpass, steps=steps, noise=noise).cpu()
tr = {}
yo, _ = O.run_flashsr(spec, W, wav, noise.cpu(), steps=steps, lowpass=lowpass, trace=tr)
be, _ = eng.plan(1, steps, lowpass)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-30))


rows = []
for name, ref in tr.items():
    if name not in be.named or ref.dim() < 3:
        continue
    got = eng.read(be, name)
    r = ref if ref.dim() == 4 else ref[:, :, None, :]
    if got.shape != r.shape:
        continue
    rows.append({"name": name, "rel": rel(got, r), "rms": float(r.pow(2).mean().sqrt())})
W16 = {k: (v.half().float() if (v.dim() >= 2) else v) for k, v in W.items()}
yw, _ = O.run_flashsr(spec, W16, wav, noise.cpu(), steps=steps, lowpass=lowpass)
rms = lambda a, b: float((a - b).pow(2).mean().sqrt())  # noqa: E731
out = {"rms_cuda_vs_oracle": rms(y, yo), "rms_cuda_vs_oracle_f16w": rms(y, yw), "rms_oracle_f16w_vs_oracle": rms(yw, yo),
       "rms_signal": float(yo.pow(2).mean().sqrt()), "layers": rows}
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "parity_diag.json").write_text(json.dumps(out, indent=0))
print({k: v for k, v in out.items() if k != "layers"})
step = max(1, len(rows) // 60)
for r in rows[::step] + rows[-8:]:
    print(f"{r['name']:60s} rel={r['rel']:.3e} rms={r['rms']:.3e}")
