"""Builds the CPU emulation of the SIMT part of libegregora_b200 — TEST INFRASTRUCTURE ONLY (see cuda_runtime.h here).

The kernels' real sources (csrc/*.cu) are copied with two textual rewrites and compiled by g++ against the shim:
  kernel<<<grid, block, smem, stream>>>(args);   ->  cusim::launch(dim3(grid), dim3(block), smem, [&]() { kernel(args); });
  extern __shared__ T name[];                     ->  T* name = reinterpret_cast<T*>(cusim::dyn_smem());
gemm_tc.cu (tcgen05 / TMA) is replaced by the host loops of gemm_tc_ref.cpp; clusters larger than one block are refused.  The result,
tests/cusim/_build/libegregora_b200_cusim.so, exports the same C ABI for those entry points with HOST pointers in place
of device pointers; only tests/test_cusim.py loads it.
"""
import hashlib
import os
import re
import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
CSRC = ROOT / "comfyui-egregora-audio-super-resolution_b200" / "csrc"
SOURCES = ["core.cu", "wola.cu", "eval_metrics.cu", "dfn_mix.cu", "fft.cu", "fatllama.cu", "ops.cu", "frontend.cu", "plan.cu"]
HEADERS = ["common.cuh", "select.cuh", "fft_plan.cuh", "fft_device.cuh", "ops.cuh", "mega.cuh"]
SHIM = ["cuda_runtime.h", "cuda_fp16.h", "cooperative_groups.h", "cusim.cpp", "gemm_tc_ref.cpp", "build.py"]
OUT = HERE / "_build"
LIB = OUT / "libegregora_b200_cusim.so"


def _match(text: str, i: int, open_ch: str, close_ch: str) -> int:
    """index just past the bracket that closes the one at text[i]"""
    depth = 0
    while True:
        c = text[i]
        if c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1


def _split_top(s: str):
    parts, depth, cur = [], 0, ""
    for c in s:
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += c
    parts.append(cur.strip())
    return parts


def rewrite(text: str, clusters: bool = False) -> str:
    text = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([A-Za-z_][\w:]*(?:\s+[A-Za-z_]\w*)*?)\s+(\w+)\s*\[\s*\]\s*;",
                  r"\1* \2 = reinterpret_cast<\1*>(cusim::dyn_smem());", text)
    # kernels compiled for a fixed thread-block cluster cannot be emulated: their launches become an error return
    clustered = {m.group(2): m.group(1).split(",")[0].strip()
                 for m in re.finditer(r"__cluster_dims__\(([^)]*)\)\s*(?:__launch_bounds__\([^)]*\)\s*)?(\w+)\s*\(", text)}
    out, pos = "", 0
    while True:
        k = text.find("<<<", pos)
        if k < 0:
            return out + text[pos:]
        # kernel expression: identifier, optionally followed by a balanced <...> template argument list
        j = k
        while j > 0 and text[j - 1].isspace():
            j -= 1
        if text[j - 1] == ">":
            depth, j = 0, j - 1
            while True:
                if text[j] == ">":
                    depth += 1
                elif text[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                j -= 1
        while j > 0 and (text[j - 1].isalnum() or text[j - 1] in "_:"):
            j -= 1
        kernel = text[j:k].strip()
        e = text.index(">>>", k)
        cfg = _split_top(text[k + 3:e])
        a0 = text.index("(", e)
        a1 = _match(text, a0, "(", ")")
        semi = text.index(";", a1)
        assert text[a1:semi].strip() == "", (kernel, text[a1:semi])
        smem = cfg[2] if len(cfg) > 2 else "0"
        if kernel in clustered and not clusters:
            out += text[pos:j] + f'return egr::fail(EGR_ERR_UNSUPPORTED, "cusim: {kernel} needs a thread-block cluster, which this build of the emulator does not provide");'
            pos = semi + 1
            continue
        ncl = clustered.get(kernel, "1")
        out += text[pos:j] + f"cusim::launch(dim3({cfg[0]}), dim3({cfg[1]}), (size_t)({smem}), [&]() {{ {kernel}{text[a0:a1]}; }}, {ncl});"
        pos = semi + 1


def build(force: bool = False, clusters: bool = False) -> Path:
    """clusters=True: the slower variant in which thread-block clusters work (CUSIM_CLUSTERS, see cuda_runtime.h)."""
    global OUT, LIB
    OUT = HERE / ("_build_clusters" if clusters else "_build")
    LIB = OUT / ("libegregora_b200_cusim_clusters.so" if clusters else "libegregora_b200_cusim.so")
    srcs = [CSRC / s for s in SOURCES + HEADERS] + [HERE / n for n in SHIM]
    srcs.append(ROOT / "include" / "egregora_b200.h")
    h = hashlib.sha256()
    for p in srcs:
        h.update(p.read_bytes())
    stamp = OUT / "stamp"
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == h.hexdigest():
        return LIB
    gxx = shutil.which("g++")
    if gxx is None:
        raise RuntimeError("g++ not found")
    # same relative layout as the package so that "../../include/egregora_b200.h" in common.cuh resolves
    work = OUT / "pkg" / "csrc"
    work.mkdir(parents=True, exist_ok=True)
    inc = OUT / "include"
    inc.mkdir(exist_ok=True)
    shutil.copy(ROOT / "include" / "egregora_b200.h", inc / "egregora_b200.h")
    for n in SOURCES + HEADERS:
        (work / (n[:-3] + ".cpp" if n.endswith(".cu") else n)).write_text(rewrite((CSRC / n).read_text(), clusters))
    cpps = [str(work / (n[:-3] + ".cpp")) for n in SOURCES] + [str(HERE / "cusim.cpp"), str(HERE / "gemm_tc_ref.cpp")]
    extra = ["-DCUSIM_CLUSTERS", "-pthread"] if clusters else []
    cmd = [gxx, "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-strict-aliasing", "-w"] + extra + [
           "-I", str(HERE), "-I", str(work), "-o", str(LIB)] + cpps
    subprocess.run(cmd, check=True)
    stamp.write_text(h.hexdigest())
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force=True, clusters="--clusters" in sys.argv))
