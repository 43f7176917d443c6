#!/usr/bin/env python
"""bench.py — real-time factor of the FlashSR hot path (BASELINE.json `metric`) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "c2"): 5.12 s mono chunks @ 48 kHz, 1 diffusion step, lowpass_input=True.
One step = one pass of the whole node path (span gather -> FlashSR plan on every chunk-channel -> [all-gather]
-> Hann WOLA) over a clip of N chunks (N = number of GPUs, weak scaling: one chunk-channel per GPU; at N=1 this is
exactly c2's single chunk).  Random-init weights of the spec'd architecture and synthetic band-limited audio
 — there is no checkpoint and no dataset in this environment.

  value     : sec of 48 kHz audio / sec, inputs already resident in HBM, device-timed (CUDA events, max over ranks)
  e2e       : same through the reference-facing node call EgregoraAudioSuperResolution.run() with HOST (pinned)
              buffers: H2D of the clip and D2H of the result inside the timed region
  roofline  : the tcgen05 tap-GEMM (dominant kernel): algorithmic FLOPs of every GEMM_TC launch of one pass / the
              device time of exactly those launches (egr_plan_run_code, CUDA events on the launching stream)
  cpu_baseline / --impl reference : the fp32 oracle port of the same path on the host cores (the upstream
              FlashSR_Inference package and its weights cannot be installed here) — baseline only.
  c3        : BASELINE.json configs[2] for real — 10 min stereo, 130 spans x 2 channels, 4 diffusion steps, through
              node.run() with a pinned host clip, STRONG scaling over the N ranks (each rank uploads and runs its own span
              block, ONE all_gather_into_tensor, stitch on every rank), with per-phase device times
  chain_c5  : configs[4] — stub "wet" signal -> adaptive DFN mix -> FlashSR (4 steps) -> Fat-Llama (50 it), 5 min stereo,
              node to node with host AUDIO dicts; path B runs on rank 0 after the gather (replicas only)
  gpu_eager_baseline : the same fp32 torch graph as the CPU port, eager on the same GPU (cuDNN / cuBLAS, TF32 allowed)
              — the library-call comparator of SURVEY 8(d), baseline only
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
PKG_DIR = ROOT / "comfyui-egregora-audio-super-resolution_b200"
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "real-time factor (sec 48 kHz audio / sec wall) FlashSR"
UNIT = "x real-time"
WORKLOAD = "c2: FlashSR 5.12 s mono chunk @48 kHz, 1 diffusion step, lowpass_input=True (one chunk-channel per GPU)"


def load_pkg():
    if "egregora_b200" in sys.modules:
        return sys.modules["egregora_b200"]
    spec = importlib.util.spec_from_file_location("egregora_b200", PKG_DIR / "__init__.py",
                                                  submodule_search_locations=[str(PKG_DIR)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["egregora_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def synth_audio(total: int, channels: int = 1, sr: int = 48000, seed: int = 1234):
    """white noise x0.1 low-passed at 4 kHz + 5 partials of 220 Hz, peak 0.5, float32 [C,total]."""
    import torch
    g = torch.Generator().manual_seed(seed)
    x = 0.1 * torch.randn(channels, total, generator=g)
    X = torch.fft.rfft(x)
    f = torch.fft.rfftfreq(total, 1.0 / sr)
    X[:, f > 4000.0] = 0
    x = torch.fft.irfft(X, n=total)
    t = torch.arange(total, dtype=torch.float64) / sr
    for h in (1, 2, 3, 5, 7):
        x = x + 0.05 * torch.sin(2 * math.pi * 220.0 * h * t).float()[None]
    x = x / x.abs().max() * 0.5
    return x.float().contiguous()


class ClockSampler:
    """nvidia-smi clocks line of B200_PROFILING.md, sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            p = [s.strip() for s in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); pw.append(float(p[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic(kernels):
    """Per-launch DRAM bytes of the named kernel(s) from the committed ncu pass of one c2 step (tools/traffic_summary.py
    writes profiles/*_traffic_c2_b1.json: dram__bytes_read.sum + dram__bytes_write.sum over every launch).  The tap-GEMM
    has two entry points (gemm_tc_kernel and its CTA-pair variant): their launches are pooled."""
    cands = sorted((ROOT / "profiles").glob("*_traffic_c2_b1.json"))
    if not cands:
        return None
    js = json.loads(cands[-1].read_text())
    ds = [js[k] for k in ((kernels,) if isinstance(kernels, str) else kernels) if js.get(k) and js[k].get("launches")]
    if not ds:
        return None
    n = sum(d["launches"] for d in ds)
    us = sum(d["us"] for d in ds)
    return {"bytes_per_launch": sum(d["dram_read_bytes"] + d["dram_write_bytes"] for d in ds) / n, "source": cands[-1].name, "launches": n,
            "tensor_pipe_pct": sum(d["us"] * d.get("tensor_pipe_pct_time_weighted", 0.0) for d in ds) / us if us else None}


def peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"bf16_burst": d.get("bf16_tflops"), "bf16_sustained": d.get("bf16_tflops_sustained"), "hbm": d.get("hbm_gbs"),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------ oracle leg
def oracle_chunk_seconds(spec, weights, frac: int, threads: int):
    """Time ONE pass of the fp32 oracle port over a [1, chunk/frac] sample on the host cores -> (sec, audio_sec)."""
    import torch
    from oracle import flashsr_oracle as O
    torch.set_num_threads(threads)
    s = dict(spec)
    s["chunk"] = spec["chunk"] // frac
    wav = synth_audio(s["chunk"], 1)
    fr = s["chunk"] // s["mel"]["hop"]
    noise = torch.randn(1, s["vae"]["embed_dim"], fr // 8, s["mel"]["n_mels"] // 8, generator=torch.Generator().manual_seed(4321))
    t0 = time.perf_counter()
    O.run_flashsr(s, weights, wav, noise, steps=1, lowpass=True)
    return time.perf_counter() - t0, s["chunk"] / spec["sr"]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    load_pkg()
    from egregora_b200 import flashsr_model as M
    import torch
    spec = M.default_spec()
    W = M.init_weights(spec, 0)
    threads = os.cpu_count() or 1
    # bounded sample: shrink the chunk (the graph is length-agnostic) until warmup+steps fit in ~4 minutes
    frac, budget = 1, 240.0
    t1, a1 = oracle_chunk_seconds(spec, W, frac, threads)
    while (args.steps + max(args.warmup - 1, 0)) * t1 > budget and frac < 8:
        frac *= 2
        t1, a1 = oracle_chunk_seconds(spec, W, frac, threads)
    for _ in range(max(args.warmup - 1, 0)):
        oracle_chunk_seconds(spec, W, frac, threads)
    tot = 0.0
    for _ in range(args.steps):
        t, a = oracle_chunk_seconds(spec, W, frac, threads)
        tot += t
    v = args.steps * a1 / tot
    sample = f"{args.steps} x one {a1:.2f} s mono chunk (chunk/{frac}), 1 step, lowpass on, fp32 torch oracle port on {threads} host threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic band-limited audio, random-init weights",
        "config": {"workload": WORKLOAD, "note": "upstream FlashSR_Inference + weights are not installable here; "
                   "this arm is the fp32 oracle port of the same graph on the host cores"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)



# ------------------------------------------------------------------------------------------------ path B leg
def bench_fatllama(dev, pk):
    """BASELINE config c4: Fat-Llama spectral enhance, 3 min 44.1 kHz stereo, 300 iterations, thr 0.6, autoscale on."""
    import numpy as np
    import torch
    from egregora_b200 import _abi, egregora_fat_llama_gpu as G
    from oracle import fat_llama_oracle as O
    C, S, sr, iters, thr = 2, 7938000, 44100, 300, 0.6
    x_host = synth_audio(S, C, sr=sr, seed=77).pin_memory()
    x_dev = x_host.to(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    G.fat_llama_device(x_dev, sr, iters, thr, 1411, True, True)
    torch.cuda.synchronize(dev)
    reps = 3
    lib0 = _abi.init(dev.index or 0)
    n0 = int(lib0.egr_launch_count())
    e0.record()
    for _ in range(reps):
        G.fat_llama_device(x_dev, sr, iters, thr, 1411, True, True)
    e1.record()
    n_launch_b = (int(lib0.egr_launch_count()) - n0) // reps
    torch.cuda.synchronize(dev)
    t_dev = e0.elapsed_time(e1) / 1e3 / reps
    node = G.EgregoraFatLlamaGPU()
    e0.record()
    for _ in range(reps):
        (res,) = node.run("wav", iters, thr, 1411, True, True, AUDIO={"waveform": x_host[None], "sample_rate": sr})
    e1.record()
    torch.cuda.synchronize(dev)
    t_e2e = e0.elapsed_time(e1) / 1e3 / reps
    # the loop alone (egr_fatllama_run): 2 kernels per iteration, 16*N algorithmic bytes per iteration and channel
    lib = _abi.init(dev.index or 0)
    samples = (x_dev * 32767.0).round()
    y = torch.empty_like(samples)
    wb = lib.egr_fatllama_workspace_bytes(C, S, 1)
    w = torch.empty(wb, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    flags = _abi.K["EGR_FL_NORMALIZE"] | _abi.K["EGR_FL_AUTOSCALE"]
    _abi.check(lib.egr_fatllama_run(samples.data_ptr(), y.data_ptr(), C, S, 1, iters, thr, flags, w.data_ptr(), wb, st))
    torch.cuda.synchronize(dev)
    e0.record()
    _abi.check(lib.egr_fatllama_run(samples.data_ptr(), y.data_ptr(), C, S, 1, iters, thr, flags, w.data_ptr(), wb, st))
    e1.record()
    torch.cuda.synchronize(dev)
    t_loop = e0.elapsed_time(e1) / 1e3
    alg_bytes = 16.0 * S * C * iters
    ach = alg_bytes / t_loop / 1e9
    # CPU port on a bounded sample: 1 channel, 4 iterations at the full length, scaled to C x iters
    cs = O.pcm16_write(x_host[:1].numpy().T).astype(np.float32)
    k = 4
    t0 = time.perf_counter()
    O.upscale(cs, 1, k, thr, True, True, dtype=np.float32)
    t_cpu = (time.perf_counter() - t0) / k * iters * C
    audio_s = S / sr
    return {"workload": "c4: Fat-Llama 3 min 44.1 kHz stereo, 300 iterations, thr 0.6, normalize+autoscale on",
            "metric": "sec audio / sec", "value": audio_s / t_dev, "ms": 1e3 * t_dev,
            "e2e": {"value": audio_s / t_e2e, "ms": 1e3 * t_e2e, "h2d_bytes": int(x_host.numel() * 4), "d2h_bytes": int(res["waveform"].numel() * 4)},
            "gpu_launches": n_launch_b,
            "roofline": {"bound": "hbm", "kernel": "fl_row_kernel + fl_col_kernel (2 per iteration)", "achieved": ach, "peak": pk["hbm"],
                         "unit": "GB/s", "frac": ach / pk["hbm"], "algorithmic_bytes": alg_bytes, "loop_ms": 1e3 * t_loop,
                         "note": "16*N bytes per iteration and channel; the 63.5 MB of work arrays are L2-resident, so DRAM traffic is lower"},
            "cpu_baseline": {"value": audio_s / t_cpu, "unit": "sec audio / sec", "cores": 1, "kind": "port",
                             "sample": f"numpy/scipy.fft float32 port, 1 channel x {k} iterations at N={S}, scaled to {C} ch x {iters} it"}}

# ------------------------------------------------------------------------------------------------ c3 / c5 / eager legs
def synth_long(total: int, channels: int, sr: int = 48000, seed: int = 1234):
    """SURVEY 8(d) signal at clip length without a 28.8 M-point host FFT per channel: one 2^21-sample band-limited block per
    channel, tiled with a per-tile gain ramp (timing input; the parity tests use synth_audio itself)."""
    import torch
    blk = 1 << 21
    base = synth_audio(blk, channels, sr, seed)
    reps = -(-total // blk)
    gains = torch.linspace(0.6, 1.0, reps)
    x = (base[:, None, :] * gains[None, :, None]).reshape(channels, reps * blk)[:, :total]
    return x.contiguous()


def _max_over_ranks(val: float, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([val], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bench_c3(N, node_cls, dev, world, sync_all, reps=2):
    """configs[2]: 10 min stereo 48 kHz, 4 diffusion steps, the span batch sharded over the ranks.  Everything through
    node.run() with a pinned HOST clip: H2D of the rank's span block, chunk gather, model, all-gather, stitch, D2H."""
    import torch
    total, C, steps = 28_800_000, 2, 4
    x_host = synth_long(total, C).pin_memory()
    audio = {"waveform": x_host[None], "sample_rate": N.REQ_SR}
    node = node_cls()
    node.NUM_STEPS = steps
    win, hop = N._win_hop()
    n_spans = len(N._iter_chunks(total, win, hop))
    for _ in range(2):   # builds the sub-batch plans, then captures their graphs
        node.run(audio=audio, lowpass_input=False, output_sr="48000")
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        (res,) = node.run(audio=audio, lowpass_input=False, output_sr="48000")
    e1.record()
    sync_all()
    t = _max_over_ranks(e0.elapsed_time(e1) / 1e3 / reps, dev, world)
    assert res["waveform"].shape == (1, C, total)
    # one instrumented pass: CUDA events between the phases of upscale_48k
    node._marks = []
    e0.record()
    (res,) = node.run(audio=audio, lowpass_input=False, output_sr="48000")
    e1.record()
    sync_all()
    phases, prev = {}, e0
    for name, ev in node._marks:
        phases[name + "_ms"] = _max_over_ranks(prev.elapsed_time(ev), dev, world)
        prev = ev
    phases["d2h_ms"] = _max_over_ranks(prev.elapsed_time(e1), dev, world)
    node._marks = None
    per = -(-n_spans // world)
    limiting = max(phases, key=phases.get)
    return {"workload": "c3: FlashSR 10 min stereo 48 kHz, 130 spans x 2 channels = 260 chunk-channels, 4 diffusion steps, "
                        "through node.run() with a pinned host clip", "scaling": "strong", "n_gpus": world,
            "value": total / N.REQ_SR / t, "unit": UNIT, "seconds": t, "spans": n_spans, "spans_per_rank": per,
            "chunk_channels_per_rank": per * C, "phases_max_over_ranks": phases, "limiting_phase": limiting,
            "h2d_bytes_per_rank": int(min(per * hop + win - hop, total) * C * 4), "d2h_bytes_per_rank": int(total * C * 4),
            "all_gather_bytes": int(world * per * C * win * 4) if world > 1 else 0}


def bench_c5(N, node_cls, dev, world, rank, sync_all):
    """configs[4]: DeepFilterNet3 denoise -> FlashSR (4 steps) -> Fat-Llama (50 it), 5 min stereo 48 kHz.  The DFN3 model is
    third-party and absent: `wet` is a stub; its adaptive wet/dry mix, the reference's own part of that node
    (egregora_audio_enhance_extras.py:607-724), runs for real.  Node to node with host AUDIO dicts, as a graph does."""
    import torch
    from egregora_b200 import egregora_dfn_mix as D, egregora_fat_llama_gpu as G
    total, C = 14_400_000, 2
    dry = synth_long(total, C, seed=99).pin_memory()
    wet = (0.85 * dry + 0.01 * synth_long(total, C, seed=7)).pin_memory()
    node = node_cls()
    node.NUM_STEPS = 4
    fl = G.EgregoraFatLlamaGPU()

    def chain():
        mixed = D.adaptive_mix(dry, wet, 48000)                                             # device [C,T]
        (up,) = node.run(audio={"waveform": mixed[None], "sample_rate": 48000}, lowpass_input=False, output_sr="48000")
        if rank == 0:   # path B does not shard (one FFT spans the clip): replicas only, rank 0 carries it
            (out,) = fl.run("wav", 50, 0.6, 1411, True, True, AUDIO=up)
            return out
        return up

    chain()
    chain()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = chain()
    e1.record()
    sync_all()
    t = _max_over_ranks(e0.elapsed_time(e1) / 1e3, dev, world)
    assert out["waveform"].shape == (1, C, total) and bool(torch.isfinite(out["waveform"][0, :, ::997]).all())
    return {"workload": "c5: stub wet -> adaptive DFN mix -> FlashSR (4 steps) -> Fat-Llama (50 it, thr 0.6), 5 min stereo 48 kHz, "
                        "node to node", "n_gpus": world, "value": total / 48000 / t, "unit": "sec audio / sec", "seconds": t,
            "note": "DeepFilterNet3 itself is third-party and absent: wet is a stub, the mix is real; path B on rank 0 after the gather"}


def bench_gpu_eager(engine, dev):
    """the fp32 torch graph (the oracle port — upstream is eager fp32 torch too) run EAGER ON THE SAME GPU through
    cuDNN / cuBLAS with TF32 allowed.  Baseline only; its low-pass is scipy on the host, as upstream's is."""
    import torch
    from oracle import flashsr_oracle as O
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    spec, out = engine.spec, {}
    be = O.TorchBackend(spec, engine.weights, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for B, lowpass, reps in ((1, True, 3), (1, False, 3), (8, False, 2)):
        wav = synth_audio(spec["chunk"], 1).expand(B, -1).contiguous().to(dev)
        noise = engine.make_noise(B, 4321)
        try:
            for _ in range(2):
                O.run_flashsr(spec, None, wav, noise, steps=1, lowpass=lowpass, backend=be)
            torch.cuda.synchronize(dev)
            e0.record()
            for _ in range(reps):
                O.run_flashsr(spec, None, wav, noise, steps=1, lowpass=lowpass, backend=be)
            e1.record()
            torch.cuda.synchronize(dev)
            t = e0.elapsed_time(e1) / 1e3 / reps
            out[f"b{B}_lowpass_{'on' if lowpass else 'off'}"] = {"ms": 1e3 * t, "rtf": B * spec["chunk"] / spec["sr"] / t}
        except Exception as e:  # pragma: no cover
            out[f"b{B}_lowpass_{'on' if lowpass else 'off'}"] = {"error": str(e)[:200]}
    del be
    torch.cuda.empty_cache()
    out["what"] = ("fp32 torch eager (cuDNN/cuBLAS, TF32 allowed, cudnn.benchmark) of the same graph on the same GPU, 1 step, "
                   "output copied to the host; baseline only")
    return out


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    # no FlashSR checkpoint exists here: seeded random weights of the spec'd architecture, said so in `data`.
    # The node itself never does this silently (flashsr_weights.py); a real checkpoint directory wins when present.
    os.environ.setdefault("EGREGORA_FLASHSR_RANDOM_INIT", "1")
    import warnings
    warnings.filterwarnings("ignore", category=RuntimeWarning, message="FlashSR: EGREGORA_FLASHSR_RANDOM_INIT")
    load_pkg()
    from egregora_b200 import _abi, egregora_audio_super_resolution as N
    K = _abi.K

    os.environ["EGREGORA_FLASHSR_STEPS"] = "1"
    N.EgregoraAudioSuperResolution.NUM_STEPS = 1
    engine = N.get_engine(dev)
    node = N.EgregoraAudioSuperResolution()
    win, hop = N._win_hop()
    total = win + (world - 1) * hop
    audio_s = total / N.REQ_SR
    x_host = synth_audio(total, 1).pin_memory()
    x_dev = x_host.to(dev)
    chunk_model = lambda c, row0=0: engine.infer(c, lowpass=True, steps=1, seed=4321, row0=row0)  # noqa: E731

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_dev():
        return N.upscale_48k(x_dev, chunk_model)

    def step_e2e():
        (res,) = node.run(audio={"waveform": x_host[None], "sample_rate": N.REQ_SR}, lowpass_input=True, output_sr="48000")
        return res

    # the clock sampler starts BEFORE the warm-up: nvidia-smi's NVML start-up takes driver locks for tens of ms, which must
    # not land inside the 10-step timed region (round 2 saw a 17 -> 20 ms/step outlier exactly there)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        t_wait = time.time()
        while not clocks.lines and time.time() - t_wait < 5.0:
            time.sleep(0.05)
    for _ in range(max(args.warmup, 3)):
        step_dev()
    # settle: on a box that has just booted the first ~10 passes after the W warm-up steps still ran up to 2x slow (first
    # graph replays, clocks and driver housekeeping: round 2 measured 31 ms/step for the first bench of a fresh box against
    # 16.2 ms for the second and third) — keep stepping, untimed, until two consecutive steps agree within 2 % (at most 40)
    settle, prev_ms = 0, None
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    while settle < 40:
        es0.record()
        step_dev()
        es1.record()
        torch.cuda.synchronize(dev)
        cur = es0.elapsed_time(es1)
        settle += 1
        if prev_ms is not None and abs(cur - prev_ms) <= 0.02 * prev_ms and settle >= 4:
            break
        prev_ms = cur
    be, handle = engine.plan(1, 1, True)

    # ---- device-resident timing
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = int(engine.lib.egr_launch_count())
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e0.record()
    for i in range(args.steps):
        step_dev()
        marks[i].record()
    e1.record()
    sync_all()
    step_ms = [(e0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]) for i in range(args.steps)]
    gpu_launches = int(engine.lib.egr_launch_count()) - launches0  # kernels of libegregora_b200 in the timed region
    t_dev = torch.tensor([e0.elapsed_time(e1) / 1e3], dtype=torch.float64, device=dev)
    # ---- end-to-end timing through the node (host buffers)
    for _ in range(2):
        step_e2e()
    settle_e2e, prev_ms = 0, None       # same settling rule as above, on the host-buffer path (pinned staging blocks, first D2H)
    while settle_e2e < 40:
        es0.record()
        step_e2e()
        es1.record()
        torch.cuda.synchronize(dev)
        cur = es0.elapsed_time(es1)
        settle_e2e += 1
        if prev_ms is not None and abs(cur - prev_ms) <= 0.02 * prev_ms and settle_e2e >= 4:
            break
        prev_ms = cur
    sync_all()
    e0.record()
    for i in range(args.steps):
        res = step_e2e()
        marks[i].record()
    e1.record()
    sync_all()
    e2e_step_ms = [(e0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]) for i in range(args.steps)]
    t_e2e = torch.tensor([e0.elapsed_time(e1) / 1e3], dtype=torch.float64, device=dev)
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = float(t_dev.item()), float(t_e2e.item())
    assert res["waveform"].shape == (1, 1, total) and bool(torch.isfinite(res["waveform"]).all())

    # ---- the BASELINE multi-GPU configs, on every rank (they hold the collective)
    c3 = c5 = None
    log = lambda m: (print(f"[bench rank {rank}] {m}", file=sys.stderr, flush=True) if os.environ.get("EGR_BENCH_VERBOSE") else None)  # noqa: E731
    log("c2 done; c3 leg")
    if os.environ.get("EGR_BENCH_C3", "1") == "1":
        try:
            c3 = bench_c3(N, N.EgregoraAudioSuperResolution, dev, world, sync_all)
        except Exception as e:  # pragma: no cover
            c3 = {"error": repr(e)[:300]}
    log("c5 leg")
    if os.environ.get("EGR_BENCH_C5", "1") == "1":
        try:
            c5 = bench_c5(N, N.EgregoraAudioSuperResolution, dev, world, rank, sync_all)
        except Exception as e:  # pragma: no cover
            c5 = {"error": repr(e)[:300]}

    log("rank-0 legs")
    be, handle = engine.plan(1, 1, True)   # the c3 / c5 legs grew the workspace: every plan handle was rebuilt
    if rank == 0:
        pk = peaks()
        lib = engine.lib
        es = engine.stream  # the engine's launching stream: CUDA events are recorded on it
        st = es.cuda_stream
        # ---- roofline of the dominant kernel: every GEMM_TC launch of one pass, timed alone on its stream
        n_tc = lib.egr_plan_count_code(handle, K["EGR_OP_GEMM_TC"])
        for _ in range(3):
            _abi.check(lib.egr_plan_run_code(handle, K["EGR_OP_GEMM_TC"], st))
        torch.cuda.synchronize(dev)
        reps = 5
        e0.record(es)
        for _ in range(reps):
            _abi.check(lib.egr_plan_run_code(handle, K["EGR_OP_GEMM_TC"], st))
        e1.record(es)
        torch.cuda.synchronize(dev)
        t_tc = e0.elapsed_time(e1) / 1e3 / reps
        # whole plan alone (no host plumbing) for the share of the step (the handle was just rebuilt: an eager pass and
        # the capturing pass come first)
        for _ in range(3):
            _abi.check(lib.egr_plan_run(handle, 0, -1, st))
        torch.cuda.synchronize(dev)
        e0.record(es)
        for _ in range(reps):
            _abi.check(lib.egr_plan_run(handle, 0, -1, st))
        e1.record(es)
        torch.cuda.synchronize(dev)
        t_plan = e0.elapsed_time(e1) / 1e3 / reps
        # FLOPs of exactly the launches timed above: GEMM ops the library runs inside the persistent UNet kernel
        # (flagged in the plan, skipped by egr_plan_run_code) are not gemm_tc_kernel launches
        mega_on = os.environ.get("EGR_NO_MEGA") is None
        tc_layers = [l for l in be.layer_table if l["kind"] == "tc" and not (mega_on and l.get("mega"))]
        tc_flops = sum(l["flops"] for l in tc_layers)
        unet_tc = [l for l in be.layer_table if l["kind"] == "tc" and l.get("mega")]
        assert len(tc_layers) == n_tc or not mega_on, (len(tc_layers), n_tc)
        achieved = tc_flops / t_tc / 1e12
        # the UNet region alone (one persistent launch per diffusion step when EGR_NO_MEGA is unset)
        flagged = [i for i, o in enumerate(be.ops) if o.flags & 1]
        t_unet = None
        if flagged:
            a_, b_ = flagged[0], flagged[-1] + 1
            for _ in range(2):
                _abi.check(lib.egr_plan_run(handle, a_, b_, st))
            torch.cuda.synchronize(dev)
            e0.record(es)
            for _ in range(reps):
                _abi.check(lib.egr_plan_run(handle, a_, b_, st))
            e1.record(es)
            torch.cuda.synchronize(dev)
            t_unet = e0.elapsed_time(e1) / 1e3 / reps
        # per-launch governing roofline: a layer with few output pixels is bound by streaming its weights, not by the
        # tensor pipe.  ideal = sum over launches of max(FLOP / tensor peak, algorithmic bytes / HBM peak), with
        # algorithmic bytes = f16 activations read once + f16 weights once + f32 output (+ f32 residual) per launch.
        ideal_s, n_hbm, alg_bytes_tc = 0.0, 0, 0.0
        for l in tc_layers:
            k1 = l["K"] // max(l["taps"], 1)
            byt = 2.0 * l["M"] * k1 + 2.0 * l["N"] * l["K"] + 8.0 * l["M"] * l["N"]
            t_t, t_h = l["flops"] / (pk["bf16_sustained"] * 1e12), byt / (pk["hbm"] * 1e9)
            ideal_s += max(t_t, t_h)
            alg_bytes_tc += byt
            n_hbm += t_h > t_t
        roof = {"bound": "tensor", "kernel": "gemm_tc_kernel + gemm_tc_pair_kernel (tcgen05 tap-GEMM, f16 operands, f32 TMEM accumulate; the pair variant is cta_group::2)",
                "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"],
                "peak_source": pk["source"] + ", sustained figure (kernel timed inside a long pass)",
                "traffic": (ncu_traffic(("gemm_tc_kernel", "gemm_tc_pair_kernel")) or {}).get("bytes_per_launch"),
                "traffic_note": ncu_traffic(("gemm_tc_kernel", "gemm_tc_pair_kernel")),
                "launches_per_pass": n_tc, "flops_per_pass": tc_flops, "avg_launch_us": 1e6 * t_tc / max(n_tc, 1),
                "gemm_share_of_plan": t_tc / t_plan, "plan_ms": 1e3 * t_plan,
                "algorithmic_bytes_per_launch": alg_bytes_tc / max(n_tc, 1),
                "layer_table": "roofline/flashsr_layers.json (tools/make_layer_table.py: per-op M/N/K/FLOPs/bytes of this plan, "
                               "cross-checked against forward hooks on the fp32 oracle)",
                "composite": {"ideal_ms": 1e3 * ideal_s, "frac": ideal_s / t_tc, "hbm_bound_launches": int(n_hbm),
                              "note": "sum over the GEMM launches of max(FLOP/tensor peak, algorithmic bytes/HBM peak) / measured"},
                "unet_region": None if t_unet is None else {
                    "kernel": "unet_mega_kernel (persistent: one cooperative launch per diffusion step, grid barriers between ops)"
                              if mega_on else "one launch per op (EGR_NO_MEGA=1)",
                    "ms": 1e3 * t_unet, "plan_ops": len(flagged), "gemm_ops": len(unet_tc), "gemm_flops": sum(l["flops"] for l in unet_tc),
                    "weight_bytes": sum(2.0 * l["N"] * l["K"] for l in unet_tc),
                    "bound": "hbm (weight streaming) in the limit; measured time is op-to-op latency (see profiles/r2_mega_trace_*.txt)",
                    "share_of_plan": t_unet / t_plan}}
        # ---- the same GEMM-only measurement on a batch of 8 chunk-channels (one sub-batch of c3 / c5): M grows 8x, so the
        # layers that are short of output tiles at batch 1 fill the 148 SMs; reported beside the c2 figure, not instead of it
        roof_b8 = None
        try:
            be8, h8 = engine.plan(8, 1, True)
            engine.infer(x_dev[:, :win].expand(8, win).contiguous(), lowpass=True, steps=1)
            be8, h8 = engine.plan(8, 1, True)
            for _ in range(2):
                _abi.check(lib.egr_plan_run_code(h8, K["EGR_OP_GEMM_TC"], st))
            torch.cuda.synchronize(dev)
            e0.record(es)
            for _ in range(3):
                _abi.check(lib.egr_plan_run_code(h8, K["EGR_OP_GEMM_TC"], st))
            e1.record(es)
            torch.cuda.synchronize(dev)
            t8 = e0.elapsed_time(e1) / 1e3 / 3
            for _ in range(2):
                _abi.check(lib.egr_plan_run(h8, 0, -1, st))
            torch.cuda.synchronize(dev)
            e0.record(es)
            for _ in range(3):
                _abi.check(lib.egr_plan_run(h8, 0, -1, st))
            e1.record(es)
            torch.cuda.synchronize(dev)
            tp8 = e0.elapsed_time(e1) / 1e3 / 3
            tc8 = [l for l in be8.layer_table if l["kind"] == "tc" and not (mega_on and l.get("mega"))]
            fl8 = sum(l["flops"] for l in tc8)
            ideal8 = 0.0
            for l in tc8:
                k1 = l["K"] // max(l["taps"], 1)
                byt = 2.0 * l["M"] * k1 + 2.0 * l["N"] * l["K"] + 8.0 * l["M"] * l["N"]
                ideal8 += max(l["flops"] / (pk["bf16_sustained"] * 1e12), byt / (pk["hbm"] * 1e9))
            roof_b8 = {"workload": "one pass over 8 chunk-channels (1 diffusion step, low-pass on)", "bound": "tensor",
                       "achieved": fl8 / t8 / 1e12, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": fl8 / t8 / 1e12 / pk["bf16_sustained"],
                       "launches_per_pass": len(tc8), "flops_per_pass": fl8, "gemm_ms": 1e3 * t8, "plan_ms": 1e3 * tp8,
                       "ms_per_chunk_channel": 1e3 * tp8 / 8, "composite_frac": ideal8 / t8}
            be, handle = engine.plan(1, 1, True)
        except Exception as e:  # pragma: no cover
            roof_b8 = {"error": repr(e)[:300]}
        roof["batch8"] = roof_b8
        # ---- batched throughput (c3's per-GPU share: 33 chunk-channels, 4 steps), extra information
        extra = None
        if os.environ.get("EGR_BENCH_BATCHED", "1") == "1":
            try:
                nb = 33
                xb = x_dev[:, :win].expand(nb, win).contiguous()
                engine.infer(xb, lowpass=False, steps=4)   # untimed: builds the sub-batch plans (7,7,7,6,6) and their graphs
                engine.infer(xb, lowpass=False, steps=4)
                torch.cuda.synchronize(dev)
                e0.record()
                engine.infer(xb, lowpass=False, steps=4)
                e1.record()
                torch.cuda.synchronize(dev)
                tb = e0.elapsed_time(e1) / 1e3
                extra = {"workload": f"c3 per-GPU share: 33 chunk-channels x 5.12 s, 4 diffusion steps, sub-batches of <= {engine.max_batch}",
                         "chunk_channels_per_s": nb / tb, "rtf_mono_equiv": nb * win / N.REQ_SR / tb, "seconds": tb}
            except Exception as e:  # pragma: no cover
                extra = {"error": str(e)[:200]}
        # ---- CPU baseline beside it (N=1 only): the oracle port on a bounded sample
        cpu = None
        if world == 1 and os.environ.get("EGR_BENCH_CPU", "1") == "1":
            from egregora_b200 import flashsr_model as M
            threads = os.cpu_count() or 1
            frac = 4
            tc_, a_ = oracle_chunk_seconds(engine.spec, engine.weights, frac, threads)
            if tc_ < 8.0:
                frac = 1
                tc_, a_ = oracle_chunk_seconds(engine.spec, engine.weights, frac, threads)
            cpu = {"value": a_ / tc_, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"one {a_:.2f} s mono chunk (chunk/{frac}), 1 step, lowpass on, fp32 torch oracle port, {tc_:.1f} s"}
        eager = None
        if world == 1 and os.environ.get("EGR_BENCH_EAGER", "1") == "1":
            try:
                eager = bench_gpu_eager(engine, dev)
            except Exception as e:  # pragma: no cover
                eager = {"error": repr(e)[:300]}
        path_b = None
        if world == 1 and os.environ.get("EGR_BENCH_PATHB", "1") == "1":
            try:
                path_b = bench_fatllama(dev, pk)
            except Exception as e:  # pragma: no cover
                path_b = {"error": str(e)[:300]}
        ws_mb = be.ws_bytes / 1e6
        out = {
            "metric": METRIC, "value": args.steps * audio_s / t_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16 operands / f32 accumulate (f32 activations between layers)",
            "data": "synthetic band-limited 48 kHz audio (SURVEY.md 8d); random-init weights of the spec'd architecture "
                    f"(engine weights: {getattr(engine, 'weights_tag', '?')})",
            "config": {"settle_steps_untimed": [settle, settle_e2e], "step_ms_min_median_max": [min(step_ms), sorted(step_ms)[len(step_ms) // 2], max(step_ms)],
                       "e2e_step_ms_min_median_max": [min(e2e_step_ms), sorted(e2e_step_ms)[len(e2e_step_ms) // 2], max(e2e_step_ms)],
                       "workload": WORKLOAD, "clip_samples": total, "chunks": world, "chunk_channels_per_gpu": 1, "steps_diffusion": 1,
                       "lowpass_input": True, "parallelism": f"chunk-sharded dp{world}" + (" + 1 NCCL all-gather" if world > 1 else ""),
                       "l2": f"no explicit flush: per-step working set (weights {engine.d_weights.numel() / 1e6:.0f} MB + "
                             f"workspace {ws_mb:.0f} MB) exceeds the 126 MB L2"},
            "e2e": {"value": args.steps * audio_s / t_e2e, "unit": UNIT, "ms_per_step": 1e3 * t_e2e / args.steps,
                    "h2d_bytes_per_step": int(x_host.numel() * 4),   # the diffusion noise is generated on the device
                    "d2h_bytes_per_step": int(total * 4)},
            "gpu_launches": gpu_launches,
            "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "gpu_eager_baseline": eager, "c3": c3, "chain_c5": c5,
            "batched": extra, "path_b": path_b,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    import faulthandler
    faulthandler.enable()   # a native crash in any leg prints the Python stack to stderr instead of dying silently
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
