"""Per-op timeline of the persistent UNet kernel (EGR_MEGA_TRACE=1: clock64 stamps of CTA 0, thread 0 — for a GEMM op that
is the A-producer warp, so `body` ends when the producer has issued its last load and `barrier` holds the rest).
    python tools/mega_trace.py [batch] [steps]"""
import collections
import ctypes as C
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ.setdefault("EGREGORA_FLASHSR_RANDOM_INIT", "1")
os.environ["EGR_MEGA_TRACE"] = "1"
import bench  # noqa: E402

bench.load_pkg()
from egregora_b200 import egregora_audio_super_resolution as N  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
eng = N.get_engine(dev)
x = bench.synth_audio(N.CHUNK_SAMPLES, B).to(dev)
for _ in range(3):
    eng.infer(x, lowpass=True, steps=steps)
torch.cuda.synchronize()
be, h = eng.plan(B, steps, True)
lib = eng.lib
info = (C.c_int * 64)()
lib.egr_debug_mega_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int]
lib.egr_debug_mega_info(h, info, 8)
n_ops = info[3]
stamps = (C.c_ulonglong * (4 * n_ops))()
codes = (C.c_int * n_ops)()
lib.egr_debug_mega_trace.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_ulonglong), C.POINTER(C.c_int), C.c_int]
n = lib.egr_debug_mega_trace(h, 0, stamps, codes, n_ops)
NAMES = {1: "gemm_tc", 2: "gemv", 3: "gn_stats", 4: "gn_apply", 5: "layernorm", 6: "attn", 7: "geglu", 8: "cat", 9: "axpby", 10: "time_embed", 11: "splitk_red", 12: "gn_fused"}
mhz = 1965.0
agg = collections.OrderedDict()
flagged = [o.name for o in be.ops if o.flags & 1]
names, fi = [], 0
for k in range(n):   # a deferred split-K layer is two table entries (GEMM, reduction) for one plan op
    if codes[k] == 11:
        names.append(names[-1] + " [reduce]")
    else:
        if codes[k] == 12:   # one-op GroupNorm: the plan's stats op was folded into it
            fi += 1
        names.append(flagged[fi] if fi < len(flagged) else "?")
        fi += 1
rows = []
for k in range(n):
    t0, t1, t2, t3 = (stamps[4 * k + j] for j in range(4))
    a = agg.setdefault(NAMES.get(codes[k], str(codes[k])), [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += (t1 - t0) / mhz; a[2] += (t2 - t1) / mhz; a[3] += (t3 - t2) / mhz
    rows.append((names[k] if k < len(names) else "?", NAMES.get(codes[k]), (t1 - t0) / mhz, (t2 - t1) / mhz, (t3 - t2) / mhz))
tot = (stamps[4 * (n - 1) + 3] - stamps[0]) / mhz
print(f"run 0: {n} ops, {info[5]} grid barriers, {tot:.1f} us on CTA 0 (batch {B}, {steps} step(s))")
print(f"{'op':12s} {'n':>4s} {'setup us':>9s} {'body us':>9s} {'barrier us':>10s} {'total us':>9s}")
for name, (c, a, b, d) in agg.items():
    print(f"{name:12s} {c:4d} {a / c:9.2f} {b / c:9.2f} {d / c:10.2f} {(a + b + d):9.1f}")
print("slowest ops:")
for r in sorted(rows, key=lambda r: -(r[2] + r[3] + r[4]))[:25]:
    print(f"  {r[0]:58s} {r[1]:10s} setup {r[2]:6.2f} body {r[3]:7.2f} barrier {r[4]:7.2f}")
