"""ctypes binding of libegregora_b200.so (the C ABI in include/egregora_b200.h).

Host plumbing only: loads the library, mirrors the two POD structs, and turns negative return codes
into RuntimeError (the reference pack's only error type, e.g. egregora_audio_super_resolution.py:136).
There is deliberately no fallback: if the shared library is missing or no sm_100 GPU is present the
calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_LIB_NAME = "libegregora_b200.so"
_HEADER_CANDIDATES = (_PKG.parent / "include" / "egregora_b200.h", _PKG / "egregora_b200.h")

MAX_TAPS = 16


class Tensor(C.Structure):
    _fields_ = [
        ("addr", C.c_uint64),
        ("rank", C.c_int32),
        ("elem", C.c_int32),
        ("dim", C.c_int64 * 5),
        ("stride", C.c_int64 * 5),
    ]


class Op(C.Structure):
    _fields_ = [
        ("code", C.c_int32),
        ("flags", C.c_int32),
        ("x0", Tensor),
        ("x1", Tensor),
        ("ptr", C.c_uint64 * 10),
        ("i", C.c_int64 * 40),
        ("f", C.c_double * 8),
        ("tap", (C.c_int16 * 5) * MAX_TAPS),
        ("name", C.c_char * 48),
    ]


def _parse_defines() -> dict:
    for h in _HEADER_CANDIDATES:
        if h.exists():
            txt = h.read_text()
            out = {}
            for m in re.finditer(r"^#define\s+(EGR_[A-Z0-9_]+)\s+(-?\d+)(?:ull|u)?\s*(?:/\*.*)?$", txt, re.M):
                out[m.group(1)] = int(m.group(2))
            return out
    raise RuntimeError(f"egregora_b200.h not found (looked in {[str(h) for h in _HEADER_CANDIDATES]})")


K = _parse_defines()  # EGR_* integer constants, single source of truth = the header


def addr(space: int, off: int) -> int:
    return (int(space) << 60) | (int(off) & 0x0FFFFFFFFFFFFFFF)


def ws(off: int) -> int:
    return addr(K["EGR_SPACE_WS"], off)


def wt(off: int) -> int:
    return addr(K["EGR_SPACE_WT"], off)


def absptr(p: int) -> int:
    return addr(K["EGR_SPACE_ABS"], p)


_lib = None


def lib_path() -> Path:
    return Path(os.environ.get("EGREGORA_B200_LIB", str(_PKG / _LIB_NAME)))


def signatures() -> dict:
    """name -> (restype, argtypes) of every entry point declared in include/egregora_b200.h."""
    vp, i32, i64, f32p = C.c_void_p, C.c_int, C.c_int64, C.c_void_p
    return {
        "egr_abi_version": (C.c_int, []),
        "egr_last_error": (C.c_char_p, []),
        "egr_init": (C.c_int, [i32]),
        "egr_sm_count": (C.c_int, []),
        "egr_launch_count": (C.c_int64, []),
        "egr_sizeof": (C.c_int, [i32]),
        "egr_chunk_gather": (C.c_int, [f32p, i32, i64, vp, vp, i32, i32, f32p, vp]),
        "egr_wola_stitch": (C.c_int, [f32p, i32, vp, vp, i32, i32, i64, i32, f32p, f32p, vp]),
        "egr_dfn_mix_workspace_bytes": (C.c_size_t, [i32, i64]),
        "egr_dfn_mix": (C.c_int, [f32p, f32p, f32p, i32, i64, i32, C.c_double, i32, i32, i32, C.c_double, C.c_double, i32,
                                  C.c_double, i32, C.c_double, vp, C.c_size_t, vp]),
        "egr_eval_workspace_bytes": (C.c_size_t, []),
        "egr_eval_null_test": (C.c_int, [f32p, i64, f32p, i64, i32, i64, i32, i32, f32p, vp, vp, C.c_size_t, vp]),
        "egr_eval_lsd_workspace_bytes": (C.c_size_t, [i64, i32, i32]),
        "egr_eval_lsd": (C.c_int, [f32p, i64, f32p, i64, i32, i64, i32, i32, C.c_float, vp, vp, C.c_size_t, vp]),
        "egr_eval_lufs_workspace_bytes": (C.c_size_t, [i32, i64, i32]),
        "egr_eval_lufs": (C.c_int, [f32p, i64, i32, i64, i32, vp, vp, C.c_size_t, vp]),
        "egr_eval_hf_band_workspace_bytes": (C.c_size_t, [vp, i64]),
        "egr_eval_hf_band": (C.c_int, [vp, f32p, i64, i32, i64, i32, C.c_double, vp, vp, C.c_size_t, vp]),
        "egr_noise_fill": (C.c_int, [C.c_uint64, i64, i64, i64, f32p, vp]),
        "egr_resample_poly": (C.c_int, [f32p, i32, i64, i32, i32, f32p, i32, i64, i64, f32p, vp]),
        "egr_plan_create": (C.c_int, [C.POINTER(Op), i32, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(vp)]),
        "egr_plan_run": (C.c_int, [vp, i32, i32, vp]),
        "egr_plan_num_launches": (C.c_int, [vp, i32, i32]),
        "egr_plan_run_code": (C.c_int, [vp, i32, vp]),
        "egr_plan_count_code": (C.c_int, [vp, i32]),
        "egr_plan_destroy": (None, [vp]),
        "egr_fft_plan_create": (C.c_int, [i64, i32, C.POINTER(vp)]),
        "egr_fft_plan_workspace_bytes": (C.c_size_t, [vp]),
        "egr_fft_plan_passes": (C.c_int, [vp]),
        "egr_fft_exec": (C.c_int, [vp, f32p, f32p, i32, i32, vp]),
        "egr_fft_plan_destroy": (None, [vp]),
        "egr_fatllama_workspace_bytes": (C.c_size_t, [i32, i64, i32]),
        "egr_fatllama_run": (C.c_int, [f32p, f32p, i32, i64, i32, i32, C.c_float, C.c_uint32, vp, C.c_size_t, vp]),
        "egr_pcm16_quantize": (C.c_int, [f32p, vp, i64, vp]),
        "egr_pcm16_to_float": (C.c_int, [vp, f32p, i64, C.c_float, vp]),
        "egr_absmax": (C.c_int, [f32p, i64, f32p, vp]),
        "egr_scale_if_above": (C.c_int, [f32p, i64, f32p, C.c_float, C.c_float, vp]),
    }


def load() -> C.CDLL:
    """Load the shared library (no device needed) and declare every prototype of the header."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not p.exists():
        raise RuntimeError(
            f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback."
        )
    lib = C.CDLL(str(p))
    sigs = signatures()
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError here == the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.egr_sizeof(0) != C.sizeof(Tensor) or lib.egr_sizeof(1) != C.sizeof(Op):
        raise RuntimeError(
            f"ABI struct mismatch: C {lib.egr_sizeof(0)}/{lib.egr_sizeof(1)} vs ctypes {C.sizeof(Tensor)}/{C.sizeof(Op)}"
        )
    if lib.egr_abi_version() != K["EGR_ABI_VERSION"]:
        raise RuntimeError("libegregora_b200.so ABI version does not match the header")
    _lib = lib
    return lib


EXPORTS = None  # filled lazily by exported_symbols()


def exported_symbols() -> list:
    """Names declared in the header (used by the CPU-side 'every symbol is exported' test)."""
    for h in _HEADER_CANDIDATES:
        if h.exists():
            txt = h.read_text()
            return sorted(set(re.findall(r"\b(egr_[a-z0-9_]+)\s*\(", txt)))
    return []


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().egr_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what or 'libegregora_b200'} failed ({rc}): {msg}")


_inited_device = None


def init(device: int = 0) -> C.CDLL:
    """egr_init once per process/device.  Raises when no B200-class GPU is visible."""
    global _inited_device
    lib = load()
    if _inited_device != device:
        check(lib.egr_init(int(device)), "egr_init")
        _inited_device = device
    return lib
