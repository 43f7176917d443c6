"""Join the plan's GEMM layer table (CPU-side, no GPU needed) with an ncu launch list: per-shape time / TFLOP/s.
    python tools/layer_times.py gpurun_out/launches.csv [batch steps lowpass]"""
import collections, csv, io, json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_pkg
load_pkg()
from egregora_b200 import flashsr_model as M, flashsr_plan as P

def load(path):
    txt = open(path).read(); start = txt.find('"ID"'); return list(csv.DictReader(io.StringIO(txt[start:])))

def main():
    path = sys.argv[1]
    B, steps, lp = (int(sys.argv[2]), int(sys.argv[3]), bool(int(sys.argv[4]))) if len(sys.argv) > 4 else (1, 1, True)
    spec = M.default_spec()
    W = M.init_weights(spec, 0)
    be = P.build_plan(spec, W, P.WeightBlob(), B, steps, lp)
    tc = [l for l in be.layer_table if l["kind"] == "tc"]
    rows = [r for r in load(path) if "gemm_tc" in r["Kernel Name"]]
    assert len(tc) == len(rows), (len(tc), len(rows))
    g = collections.OrderedDict(); stage = collections.OrderedDict()
    for l, r in zip(tc, rows):
        us = float(r["Metric Value"]) / 1e3
        k = (l["M"], l["N"], l["K"], l["taps"], r["Grid Size"])
        v = g.setdefault(k, [0, 0.0, 0.0, l["name"]]); v[0] += 1; v[1] += l["flops"]; v[2] += us
        nm = l["name"]; sk = ".".join(nm.split(".")[:2]) if nm.startswith(("vae", "unet")) else "vocoder"
        s = stage.setdefault(sk, [0, 0.0, 0.0]); s[0] += 1; s[1] += l["flops"]; s[2] += us
    print("      M     N     K taps           grid |   n      GF       us    TF/s  first layer")
    for k, v in sorted(g.items(), key=lambda kv: -kv[1][2])[:40]:
        print(f"{k[0]:7d} {k[1]:5d} {k[2]:5d} {k[3]:3d} {k[4]:>14s} | {v[0]:3d} {v[1] / 1e9:7.1f} {v[2]:8.1f} {v[1] / v[2] / 1e6:7.1f}  {v[3]}")
    for k, v in stage.items():
        print(f"{k:24s} n={v[0]:3d} GF={v[1] / 1e9:8.1f} us={v[2]:8.1f} TF/s={v[1] / v[2] / 1e6:7.1f}")
    print("total GF", sum(v[1] for v in stage.values()) / 1e9, "us", sum(v[2] for v in stage.values()))
main()
