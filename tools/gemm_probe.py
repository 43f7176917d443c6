"""Micro-benchmark of single tap-GEMM ops through the C ABI (egr_plan_create / egr_plan_run): the shapes that
dominate one FlashSR pass.  Prints device time per launch (CUDA events on the launching stream), cold (L2 flushed
by writing a 512 MB buffer before every launch) and warm (back-to-back), with the achieved TFLOP/s.
    python tools/gemm_probe.py [filter]
"""
import ctypes as C
import math
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from harness import MiniPlan  # noqa: E402
from egregora_b200 import _abi  # noqa: E402

SHAPES = [
    # kind, cin, cout, k, (H, W) or T, dilation
    ("conv2d", 128, 128, 3, (512, 256), 1),
    ("conv2d", 256, 256, 3, (256, 128), 1),
    ("conv2d", 512, 512, 3, (128, 64), 1),
    ("conv2d", 1024, 1024, 3, (64, 32), 1),
    ("conv2d", 1024, 1024, 1, (64, 32), 1),
    ("conv2d", 128, 128, 3, (64, 32), 1),
    ("conv2d", 256, 256, 3, (32, 16), 1),
    ("conv2d", 256, 256, 1, (32, 16), 1),
    ("conv2d", 384, 384, 3, (16, 8), 1),
    ("conv2d", 384, 384, 1, (16, 8), 1),
    ("conv2d", 640, 640, 3, (8, 4), 1),
    ("conv2d", 640, 640, 1, (8, 4), 1),
    ("conv2d", 1280, 640, 3, (8, 4), 1),
    ("conv2d", 640, 5120, 1, (8, 4), 1),
    ("conv1d", 768, 768, 11, 3072, 1),
    ("conv1d", 384, 384, 11, 15360, 1),
    ("conv1d", 192, 192, 11, 61440, 1),
    ("conv1d", 192, 192, 3, 61440, 5),
    ("conv1d", 96, 96, 11, 122880, 1),
    ("conv1d", 96, 96, 3, 122880, 1),
    ("conv1d", 48, 48, 11, 245760, 1),
    ("conv1d", 48, 48, 3, 245760, 1),
]


def w(shape, seed):
    g = torch.Generator().manual_seed(seed)
    fan = 1
    for s in shape[1:]:
        fan *= s
    return (torch.rand(shape, generator=g) * 2 - 1) * math.sqrt(3.0 / fan)


def probe(kind, cin, cout, k, sp, dil, batch, dev, flush):
    if kind == "conv2d":
        H, Wd = sp
        Wt = {"c.weight": w((cout, cin, k, k), 1), "c.bias": torch.zeros(cout)}
        x = torch.randn(batch, cin, H, Wd)
        mp = MiniPlan(Wt)
        xi = mp.input(x)
        res = mp.be.new(batch, H, Wd, cout, f32=True, tag="res", persistent=True)
        mp.be.conv2d(xi, "c", cin, cout, k, add=res)
        M = batch * H * Wd
        flops = 2.0 * M * cout * cin * k * k
    else:
        T = sp
        Wt = {"c.weight": w((cout, cin, k), 1), "c.bias": torch.zeros(cout)}
        x = torch.randn(batch, cin, 1, T)
        mp = MiniPlan(Wt)
        xi = mp.input(x)
        res = mp.be.new(batch, 1, T, cout, f32=True, tag="res", persistent=True)
        mp.be.conv1d(xi, "c", cin, cout, k, dilation=dil, add=res)
        M = batch * T
        flops = 2.0 * M * cout * cin * k
    mp.be.debug = False
    mp._finish()
    lib = _abi.init(dev.index or 0)
    ws = torch.zeros(mp.ws_bytes + 4096, dtype=torch.uint8, device=dev)
    wt = torch.frombuffer(bytearray(mp.blob.tobytes()), dtype=torch.uint8).to(dev)
    mp.ws = ws
    for t, xx in mp.inputs:
        mp.view(t.f32, torch.float32, xx.shape).copy_(xx)
    h = C.c_void_p()
    _abi.check(lib.egr_plan_create(mp.ops, len(mp.be.ops), ws.data_ptr(), ws.numel(), wt.data_ptr(), wt.numel(), C.byref(h)))
    st = torch.cuda.current_stream(dev).cuda_stream
    cfg = (C.c_int * 8)()
    lib.egr_debug_tc_config.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.egr_debug_tc_config(h, len(mp.be.ops) - 1, cfg)
    probe.cfg = "bn%d mt%d sp%d h%d work%d grid%d SA%d SB%d" % tuple(cfg)
    n_ops = len(mp.be.ops)
    gi = n_ops - 1  # the GEMM is the last op (a CAST16 precedes it)
    _abi.check(lib.egr_plan_run(h, 0, -1, st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cold = []
    for _ in range(5):
        flush.fill_(1.0)
        e0.record()
        _abi.check(lib.egr_plan_run(h, gi, gi + 1, st))
        e1.record()
        torch.cuda.synchronize()
        cold.append(e0.elapsed_time(e1) * 1e3)
    reps = 20
    e0.record()
    for _ in range(reps):
        _abi.check(lib.egr_plan_run(h, gi, gi + 1, st))
    e1.record()
    torch.cuda.synchronize()
    warm = e0.elapsed_time(e1) * 1e3 / reps
    lib.egr_plan_destroy(h)
    cold.sort()
    c = cold[len(cold) // 2]
    return M, flops, c, warm


def main():
    dev = torch.device("cuda", 0)
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    batches = [int(b) for b in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["1"])]
    flush = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device=dev)
    print(f"{'shape':44s} {'B':>2s} {'M':>8s} {'GF':>8s} {'cold us':>9s} {'TF/s':>7s} {'warm us':>9s} {'TF/s':>7s}")
    for (kind, cin, cout, k, sp, dil) in SHAPES:
        name = f"{kind} {cin}->{cout} k{k} d{dil} {sp}"
        if flt and flt not in name:
            continue
        for b in batches:
            M, fl, c, wm = probe(kind, cin, cout, k, sp, dil, b, dev, flush)
            print(f"{name:44s} {b:2d} {M:8d} {fl / 1e9:8.1f} {c:9.1f} {fl / c / 1e6:7.1f} {wm:9.1f} {fl / wm / 1e6:7.1f}  {probe.cfg}", flush=True)


if __name__ == "__main__":
    main()
