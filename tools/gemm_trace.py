"""Clock trace of CTA (0,0) of one tap-GEMM launch (debug hook egr_debug_tc_trace)."""
import ctypes as C
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "tools"))
import gemm_probe as G
from egregora_b200 import _abi

def main():
    dev = torch.device("cuda", 0)
    lib = _abi.init(0)
    lib.egr_debug_tc_trace.restype = C.c_int
    lib.egr_debug_tc_trace.argtypes = [C.c_void_p, C.c_int]
    flush = torch.empty(1024, dtype=torch.float32, device=dev)
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    brief = len(sys.argv) > 3 and sys.argv[3] == "brief"
    lib.egr_debug_tc_trace(None, 0)  # enable
    buf = (C.c_ulonglong * 1400)()
    for (kind, cin, cout, k, sp, dil) in G.SHAPES:
        name = f"{kind} {cin}->{cout} k{k} d{dil} {sp}"
        if flt and flt not in name:
            continue
        M, fl, c, wm = G.probe(kind, cin, cout, k, sp, dil, batch, dev, flush)
        lib.egr_debug_tc_trace(buf, 1400)
        t = list(buf)
        t0 = t[0]
        rel = lambda v: (v - t0) if v else -1
        epi = [(rel(t[2 + 2 * i]), rel(t[3 + 2 * i])) for i in range(6) if t[2 + 2 * i]]
        print(f"== {name}: warm {wm:.1f} us {G.probe.cfg}; setup {rel(t[1])}; epilogue (acc ready, done) per item: {epi} (cycles)")
        st = [t[1100 + 2 * i] for i in range(148) if t[1100 + 2 * i]]
        en = [t[1101 + 2 * i] for i in range(148) if t[1101 + 2 * i]]
        if st and en:
            g0 = min(st)
            print(f"   CTAs {len(st)}: start spread {max(st) - g0} ns, first end {min(en) - g0} ns, last end {max(en) - g0} ns; per-CTA duration min/max {min(e - s_ for s_, e in zip(st, en))}/{max(e - s_ for s_, e in zip(st, en))} ns")
        mma = [(rel(t[528 + 2 * i]), rel(t[529 + 2 * i])) for i in range(250) if t[528 + 2 * i]]
        prod = [(rel(t[16 + 2 * i]), rel(t[17 + 2 * i])) for i in range(250) if t[16 + 2 * i]]
        if len(mma) > 40:   # steady-state k-step period of CTA 0 (cycles) and the SM clock it ran at
            per = (mma[120][0] - mma[20][0]) / 100.0 if len(mma) > 120 else (mma[-1][0] - mma[20][0]) / (len(mma) - 21)
            lag = sum(m[0] - p_[1] for m, p_ in zip(mma[20:120], prod[20:120])) / max(1, len(mma[20:120]))
            cyc_total = max((e for _, e in epi), default=0)
            ns_cta0 = (t[1101] - t[1100]) if t[1100] else 0
            print(f"   k-step period {per:.0f} clk; producer-issue -> mma-issue lag {lag:.0f} clk; CTA0 {cyc_total} clk in {ns_cta0} ns = {1e3 * cyc_total / max(ns_cta0, 1):.0f} MHz")
        if brief:
            continue
        ep = [(rel(t[800 + 5 * i]),) + tuple(t[800 + 5 * i + j] - t[800 + 5 * i] for j in range(1, 5)) for i in range(16) if t[800 + 5 * i]]
        print("   epilogue blocks of item 1 (start; +tmem ld done, +smem staged, +residual arrived, +stored):", ep[:10])
        print("   producer warp: role entry", rel(t[5]), "decoded", rel(t[6]), "first wait passed", rel(t[7]))
        prod = [(rel(t[16 + 2 * i]), rel(t[17 + 2 * i])) for i in range(250) if t[16 + 2 * i]]
        mma = [(rel(t[528 + 2 * i]), rel(t[529 + 2 * i])) for i in range(250) if t[528 + 2 * i]]
        print("   producer (slot free, issued):", prod[:12], "...", prod[-3:])
        print("   mma (data ready, committed):", mma[:12], "...", mma[-3:])

main()
