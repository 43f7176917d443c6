"""FlashSR checkpoint loading — replaces the weight bootstrap of the reference's `_FlashSRRunner`
(/root/reference/egregora_audio_super_resolution.py:260-265 file names, :282-320 `_ensure_weights`, :346-359
`FlashSR(student_ldm.pth, sr_vocoder.pth, vae.pth)`).

The reference resolves `ComfyUI/models/audio/flashsr/{student_ldm,sr_vocoder,vae}.pth`, downloads them when they are
missing and raises `RuntimeError("FlashSR weights missing …")` when that fails.  This build has no network path:
it resolves the same directory (or `EGREGORA_FLASHSR_WEIGHTS`, or `flashsr_min --ckpt-dir`), `torch.load`s the three
files, maps the upstream state-dict names onto `flashsr_model.param_shapes` and raises the reference's message when a
file is missing, plus a precise one when a tensor is missing or mis-shaped.  It never substitutes random weights:
random initialisation exists only behind the explicit `EGREGORA_FLASHSR_RANDOM_INIT=1` switch that tests and bench.py
set (there is no checkpoint in this environment).

Name mapping.  The three files hold the state dicts of an ldm `LatentDiffusion` student (UNet under
`model.diffusion_model.`), a BigVGAN-style generator (possibly with weight-norm parametrisation: `weight_g/weight_v`
or `parametrizations.weight.original0/1`) and an ldm `AutoencoderKL`.  The graph in flashsr_model.py uses exactly those
module names under the section prefixes `unet.` / `vocoder.` / `vae.`, so the mapping is: unwrap the container, strip
the wrapper prefixes, fold weight norm, accept 1x1-conv <-> linear shape variants.  The upstream repo is an un-pinned
download that is absent here (**parity unpinned**); every rule below is therefore table-driven and reported on error.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from pathlib import Path
from typing import Dict, Iterable, Optional, Tuple

import torch

from . import flashsr_model as M

HF_FILES = ("student_ldm.pth", "sr_vocoder.pth", "vae.pth")   # reference :261, constructor order :350-353
MISSING_MSG = ("FlashSR weights missing and auto-download failed. "
               "Place these in models/audio/flashsr: student_ldm.pth, sr_vocoder.pth, vae.pth")   # reference :316-319

# file -> (section prefix in param_shapes, wrapper prefixes stripped from the file's keys, longest first;
#          `select`: when any key starts with it only those keys belong to the section — a full LatentDiffusion
#          checkpoint also carries first_stage_model.* / cond_stage_model.* copies)
SECTIONS = {
    "student_ldm.pth": ("unet.", ("model.diffusion_model.", "diffusion_model.", "student.", "unet.", "module."),
                        "model.diffusion_model."),
    "sr_vocoder.pth": ("vocoder.", ("generator.", "vocoder.", "module.", "model."), None),
    "vae.pth": ("vae.", ("first_stage_model.", "autoencoder.", "vae.", "module."), "first_stage_model."),
}
CONTAINERS = ("state_dict", "generator", "model", "ema", "params", "weights")


def _custom_root() -> Path:
    return Path(__file__).resolve().parent


def models_dir() -> Path:
    """.../ComfyUI/models — ComfyUI's own registry when it is importable, else the reference's rule
    (custom_nodes/<pack>/ -> parents[2]/models, reference :25-30)."""
    try:
        import folder_paths  # type: ignore  (ComfyUI)
        return Path(folder_paths.models_dir)
    except Exception:
        return _custom_root().parents[1] / "models"


def resolve_ckpt_dir(explicit: Optional[str] = None) -> Path:
    """Directory expected to hold the three files: explicit argument (flashsr_min --ckpt-dir) >
    EGREGORA_FLASHSR_WEIGHTS > ComfyUI/models/audio/flashsr (reference :32-35, :265)."""
    if explicit:
        return Path(explicit)
    env = os.environ.get("EGREGORA_FLASHSR_WEIGHTS", "")
    if env:
        return Path(env)
    return models_dir() / "audio" / "flashsr"


def random_init_allowed() -> bool:
    return os.environ.get("EGREGORA_FLASHSR_RANDOM_INIT", "") == "1"


def _unwrap(obj):
    """torch.load result -> flat {name: tensor}."""
    seen = 0
    while isinstance(obj, dict) and seen < 4:
        if obj and all(isinstance(v, torch.Tensor) for v in obj.values()):
            return obj
        nxt = next((obj[k] for k in CONTAINERS if k in obj and isinstance(obj[k], dict)), None)
        if nxt is None:
            break
        obj, seen = nxt, seen + 1
    if isinstance(obj, dict):
        flat = {k: v for k, v in obj.items() if isinstance(v, torch.Tensor)}
        if flat:
            return flat
    raise RuntimeError("checkpoint does not contain a state dict of tensors")


def _fold_weight_norm(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """`x.weight_g` + `x.weight_v` (torch.nn.utils.weight_norm, dim 0) or `x.parametrizations.weight.original0/1`
    -> `x.weight` = g * v / ||v|| (norm over every dim but 0), as torch computes it."""
    out: Dict[str, torch.Tensor] = OrderedDict()
    pairs = {}
    for k, v in sd.items():
        if k.endswith(".weight_g"):
            pairs.setdefault(k[:-len(".weight_g")], {})["g"] = v
        elif k.endswith(".weight_v"):
            pairs.setdefault(k[:-len(".weight_v")], {})["v"] = v
        elif k.endswith(".parametrizations.weight.original0"):
            pairs.setdefault(k[:-len(".parametrizations.weight.original0")], {})["g"] = v
        elif k.endswith(".parametrizations.weight.original1"):
            pairs.setdefault(k[:-len(".parametrizations.weight.original1")], {})["v"] = v
        else:
            out[k] = v
    for base, gv in pairs.items():
        if "g" not in gv or "v" not in gv:
            raise RuntimeError(f"weight-norm pair of '{base}' is incomplete")
        v = gv["v"].float()
        norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape([-1] + [1] * (v.dim() - 1))
        out[base + ".weight"] = gv["g"].float().reshape(norm.shape) * v / norm
    return out


def _strip(key: str, prefixes: Iterable[str]) -> str:
    changed = True
    while changed:
        changed = False
        for p in prefixes:
            if key.startswith(p):
                key, changed = key[len(p):], True
    return key


def _fit(t: torch.Tensor, shape: Tuple[int, ...]) -> Optional[torch.Tensor]:
    """Accept the tensor when it has the wanted shape up to singleton dims (ldm stores 1x1 convs as [o,i,1,1] where
    newer code has linears, and the other way round)."""
    if tuple(t.shape) == tuple(shape):
        return t
    squeeze = lambda s: tuple(d for d in s if d != 1)  # noqa: E731
    if t.numel() == int(torch.tensor(shape).prod()) and squeeze(t.shape) == squeeze(shape):
        return t.reshape(shape)
    return None


def map_section(fname: str, raw, spec: dict) -> Dict[str, torch.Tensor]:
    """One checkpoint file -> {param_shapes name: f32 tensor} for its section; raises on missing / mis-shaped."""
    section, prefixes, select = SECTIONS[fname]
    sd = _unwrap(raw)
    if select and any(k.startswith(select) for k in sd):
        sd = {k: v for k, v in sd.items() if k.startswith(select)}
    sd = _fold_weight_norm({_strip(k, prefixes): v for k, v in sd.items()})
    want = {k: shp for k, (shp, _) in M.param_shapes(spec).items() if k.startswith(section)}
    out, missing, bad = OrderedDict(), [], []
    for name, shape in want.items():
        t = sd.get(name[len(section):])
        if t is None:
            missing.append(name[len(section):])
            continue
        f = _fit(t.detach().float(), shape)
        if f is None:
            bad.append(f"{name[len(section):]}: checkpoint {tuple(t.shape)} vs model {tuple(shape)}")
            continue
        out[name] = f.contiguous()
    if missing or bad:
        msg = [f"FlashSR checkpoint {fname} does not match the model:"]
        if missing:
            msg.append(f"  {len(missing)} missing tensors, e.g. {', '.join(missing[:4])}")
        if bad:
            msg.append(f"  {len(bad)} mis-shaped tensors, e.g. {'; '.join(bad[:3])}")
        msg.append(f"  ({len(sd)} tensors in the file after prefix stripping; first: {', '.join(list(sd)[:3])})")
        raise RuntimeError("\n".join(msg))
    return out


def load_checkpoint(ckpt_dir, spec: Optional[dict] = None) -> Dict[str, torch.Tensor]:
    """The three reference files -> one weight dict in param_shapes order.  RuntimeError (reference text) when a file
    is missing."""
    spec = spec or M.default_spec()
    d = Path(ckpt_dir)
    missing = [f for f in HF_FILES if not (d / f).is_file()]
    if missing:
        raise RuntimeError(f"{MISSING_MSG} (looked in {d}; missing: {', '.join(missing)}; this build never downloads)")
    parts: Dict[str, torch.Tensor] = {}
    for f in HF_FILES:
        try:
            raw = torch.load(str(d / f), map_location="cpu", weights_only=True)
        except Exception as e:
            raise RuntimeError(f"FlashSR checkpoint {d / f} could not be read: {e}") from e
        parts.update(map_section(f, raw, spec))
    return OrderedDict((k, parts[k]) for k in M.param_shapes(spec))


def weights_for_node(explicit: Optional[str] = None, spec: Optional[dict] = None) -> Tuple[Dict[str, torch.Tensor], str]:
    """(weights, provenance tag) for the node's engine cache.  Never random unless explicitly allowed."""
    spec = spec or M.default_spec()
    d = resolve_ckpt_dir(explicit)
    have = [f for f in HF_FILES if (d / f).is_file()]
    if len(have) == len(HF_FILES) or explicit or os.environ.get("EGREGORA_FLASHSR_WEIGHTS"):
        return load_checkpoint(d, spec), f"ckpt:{d.resolve()}"
    if random_init_allowed():
        import warnings
        seed = int(os.environ.get("EGREGORA_FLASHSR_RANDOM_SEED", "0"))
        warnings.warn(f"FlashSR: EGREGORA_FLASHSR_RANDOM_INIT=1 and no checkpoint in {d}: running on SEEDED RANDOM weights "
                      "(tests / benchmarks only; the output is not super-resolved audio)", RuntimeWarning, stacklevel=2)
        return M.init_weights(spec, seed), f"random:{seed}"
    return load_checkpoint(d, spec), ""   # raises the reference's "weights missing" RuntimeError
