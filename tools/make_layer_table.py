"""Per-layer FLOP / byte table of one FlashSR pass ("generate the per-layer FLOP/byte table from the
oracle module tree with forward hooks and commit it (roofline/flashsr_layers.json) before claiming fractions").

Two independent counts, no GPU needed:
  plan   : every GEMM op of the CUDA plan (flashsr_plan.PlanBackend.layer_table) for c2 (batch 1, 1 step, lowpass on):
           M (output pixels), N (output channels), K (taps x input channels), FLOPs = 2*M*N*K, algorithmic bytes =
           f16 activations read once + f16 weights once + f32 output (+ f32 residual), and which roofline bounds it
           at the measured peaks.  bench.py's roofline.achieved divides exactly these FLOPs by the measured time.
  oracle : the fp32 torch oracle run once on the same configuration with torch.nn.functional.conv1d / conv2d /
           conv_transpose1d / linear wrapped by counting hooks (the oracle is functional, so the hooks sit on F.*).
The two are reconciled in `reconciliation`: the plan computes nearest-2x-upsample + conv3x3 as four 2x2-tap phase GEMMs
(2.25x fewer FLOPs than the oracle's conv on the up-sampled map) and runs the VAE attention products as GEMM ops (the
oracle uses matmul, not hooked).
    python tools/make_layer_table.py        # writes roofline/flashsr_layers.json
"""
import json
import re
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_pkg  # noqa: E402

load_pkg()
from egregora_b200 import flashsr_model as M, flashsr_plan as P  # noqa: E402

HBM, TENSOR = 6532.9e9, 1402.7e12   # MEASURED_PEAKS.json of this pool (hbm_gbs, bf16_tflops_sustained)


def plan_table(spec, W, batch=1, steps=1, lowpass=True):
    be = P.build_plan(spec, W, P.WeightBlob(), batch, steps, lowpass)
    rows = []
    for l in be.layer_table:
        k1 = l["K"] // max(l["taps"], 1)
        byt = 2.0 * l["M"] * k1 + 2.0 * l["N"] * l["K"] + 8.0 * l["M"] * l["N"]
        t_t, t_h = l["flops"] / TENSOR, byt / HBM
        rows.append({"name": l["name"], "kind": l["kind"], "M": l["M"], "N": l["N"], "K": l["K"], "taps": l["taps"],
                     "flops": l["flops"], "algorithmic_bytes": byt, "bound": "tensor" if t_t >= t_h else "hbm",
                     "ideal_us": 1e6 * max(t_t, t_h)})
    return be, rows


def oracle_hooks(spec, W, steps=1, lowpass=True):
    from oracle import flashsr_oracle as O
    rec = {"conv2d": [0, 0.0], "conv1d": [0, 0.0], "conv_transpose1d": [0, 0.0], "linear": [0, 0.0], "depthwise_fir": [0, 0.0]}
    orig = {k: getattr(F, k) for k in ("conv2d", "conv1d", "conv_transpose1d", "linear")}

    def wrap(kind):
        fn = orig[kind]

        def hooked(x, w, *a, **kw):
            y = fn(x, w, *a, **kw)
            groups = kw.get("groups", 1)
            if kind == "linear":
                fl = 2.0 * y.numel() * w.shape[1]
            elif kind == "conv_transpose1d":   # weight [Cin, Cout/groups, k]: every input sample meets every tap
                fl = 2.0 * x.shape[0] * x.shape[2] * w.shape[0] * w.shape[1] * w.shape[2]
            else:                               # weight [Cout, Cin/groups, *k]
                fl = 2.0 * y.numel() * w[0].numel()
            key = "depthwise_fir" if groups > 1 else kind
            rec[key][0] += 1
            rec[key][1] += fl
            return y
        return hooked

    for k in orig:
        setattr(F, k, wrap(k))
    try:
        sys.path.insert(0, str(ROOT))
        g = torch.Generator().manual_seed(1)
        wav = 0.1 * torch.randn(1, spec["chunk"], generator=g)
        fr = spec["chunk"] // spec["mel"]["hop"]
        noise = torch.randn(1, spec["vae"]["embed_dim"], fr // 8, spec["mel"]["n_mels"] // 8, generator=g)
        O.run_flashsr(spec, W, wav, noise, steps=steps, lowpass=lowpass)
    finally:
        for k, fn in orig.items():
            setattr(F, k, fn)
    return {k: {"calls": v[0], "flops": v[1]} for k, v in rec.items()}


def build(spec_name="default"):
    spec = M.default_spec() if spec_name == "default" else M.tiny_spec()
    W = M.init_weights(spec, 0)
    be, rows = plan_table(spec, W)
    tc = [r for r in rows if r["kind"] == "tc"]
    simt = [r for r in rows if r["kind"] == "simt"]
    up = [r for r in rows if re.search(r"\.p[01][01]$", r["name"])]           # phase GEMMs of upsample + conv3x3 (VAE and UNet)
    attn = [r for r in rows if r["name"] in ("attn.qk", "attn.pv")]           # VAE mid-block attention products
    hooks = oracle_hooks(spec, W)
    oracle_gemm = sum(hooks[k]["flops"] for k in ("conv2d", "conv1d", "conv_transpose1d", "linear"))
    plan_total = sum(r["flops"] for r in rows)
    up_f = sum(r["flops"] for r in up)
    attn_f = sum(r["flops"] for r in attn)
    return {
        "config": {"spec": spec_name, "workload": "c2: one 5.12 s mono chunk, 1 diffusion step, lowpass on (batch 1)" if spec_name == "default" else "tiny spec, batch 1",
                   "peaks": {"hbm_Bps": HBM, "tensor_FLOPps": TENSOR, "source": "MEASURED_PEAKS.json (sustained bf16)"}},
        "totals": {"gemm_ops": len(rows), "tc_ops": len(tc), "simt_ops": len(simt), "flops": plan_total,
                   "tc_flops": sum(r["flops"] for r in tc), "algorithmic_bytes_tc": sum(r["algorithmic_bytes"] for r in tc),
                   "ideal_ms_tc": 1e-3 * sum(r["ideal_us"] for r in tc), "hbm_bound_tc_ops": sum(r["bound"] == "hbm" for r in tc)},
        "oracle_hooks": hooks,
        "reconciliation": {
            "oracle_conv_linear_flops": oracle_gemm,
            "plan_flops": plan_total,
            "plan_upsample_phase_gemm_flops": up_f,
            "oracle_equivalent_of_those": 2.25 * up_f,
            "plan_attention_product_flops_not_hooked_in_oracle": attn_f,
            "plan_flops_restated_on_oracle_terms": plan_total - attn_f + 1.25 * up_f,
        },
        "layers": rows,
    }


def main():
    out = build("default")
    r = out["reconciliation"]
    rel = abs(r["plan_flops_restated_on_oracle_terms"] - r["oracle_conv_linear_flops"]) / r["oracle_conv_linear_flops"]
    out["reconciliation"]["relative_difference"] = rel
    (ROOT / "roofline").mkdir(exist_ok=True)
    (ROOT / "roofline" / "flashsr_layers.json").write_text(json.dumps(out, indent=1))
    print(json.dumps({k: out[k] for k in ("totals", "oracle_hooks", "reconciliation")}, indent=1))


if __name__ == "__main__":
    main()
