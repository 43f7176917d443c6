#!/usr/bin/env python
"""Write three synthetic FlashSR checkpoint files in UPSTREAM naming — student_ldm.pth, sr_vocoder.pth, vae.pth, the
files the reference runner loads (egregora_audio_super_resolution.py:260-261, :346-353) — from seeded random weights of
the spec'd architecture.  No real checkpoint exists in this environment; these files exercise the whole
load path (container unwrapping, prefix stripping, weight-norm folding, 1x1-conv/linear shape variants) that real files
take:

  student_ldm.pth  {"state_dict": {"model.diffusion_model.<unet name>": tensor, ...,
                                    "first_stage_model.decoy": ..., "cond_stage_model.decoy": ...}}   (ldm LatentDiffusion)
  sr_vocoder.pth   {"generator": {"<name>.weight_g", "<name>.weight_v" for every conv (torch weight_norm), snake
                                    alphas/betas and biases as they are}}                              (BigVGAN)
  vae.pth          {"state_dict": {"<vae name>": tensor}}  with the attention 1x1 convs stored as [C,C] linears

    python tools/make_synthetic_ckpt.py OUT_DIR [--tiny] [--seed 0]
"""
import argparse
import importlib.util
import sys
from collections import OrderedDict
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
PKG_DIR = ROOT / "comfyui-egregora-audio-super-resolution_b200"


def load_pkg():
    if "egregora_b200" in sys.modules:
        return sys.modules["egregora_b200"]
    spec = importlib.util.spec_from_file_location("egregora_b200", PKG_DIR / "__init__.py", submodule_search_locations=[str(PKG_DIR)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["egregora_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def write(out_dir, spec, weights, g_seed=99):
    """weights: {param_shapes name: tensor}.  Returns the dict of paths."""
    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    gen = torch.Generator().manual_seed(g_seed)
    unet, voc, vae = OrderedDict(), OrderedDict(), OrderedDict()
    for k, v in weights.items():
        if k.startswith("unet."):
            unet["model.diffusion_model." + k[len("unet."):]] = v.clone()
        elif k.startswith("vocoder."):
            name = k[len("vocoder."):]
            if name.endswith(".weight") and v.dim() == 3:     # conv / transposed conv under weight_norm (dim 0)
                base = name[:-len(".weight")]
                norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, 1, 1)
                scale = 0.5 + torch.rand(v.shape[0], 1, 1, generator=gen)   # v is any vector along the weight's direction
                voc[base + ".weight_g"] = norm.clone()
                voc[base + ".weight_v"] = (v * scale).clone()
            else:
                voc[name] = v.clone()
        elif k.startswith("vae."):
            name = k[len("vae."):]
            if ".attn_1." in name and name.endswith(".weight") and v.dim() == 4:
                vae[name] = v.reshape(v.shape[0], v.shape[1]).clone()      # stored as a linear
            else:
                vae[name] = v.clone()
        else:
            raise ValueError(k)
    unet["first_stage_model.decoy.weight"] = torch.zeros(3)
    unet["cond_stage_model.decoy.weight"] = torch.zeros(3)
    paths = {"student_ldm.pth": {"state_dict": unet, "global_step": 0}, "sr_vocoder.pth": {"generator": voc},
             "vae.pth": {"state_dict": vae}}
    for fname, obj in paths.items():
        torch.save(obj, str(out / fname))
    return {f: out / f for f in paths}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out_dir")
    ap.add_argument("--tiny", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    load_pkg()
    from egregora_b200 import flashsr_model as M
    spec = M.tiny_spec() if a.tiny else M.default_spec()
    for f, p in write(a.out_dir, spec, M.init_weights(spec, a.seed)).items():
        print(f, p.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
