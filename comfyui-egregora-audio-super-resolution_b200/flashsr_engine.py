"""FlashSR engine: owns the device weight blob, the workspace and one compiled plan per (batch, steps, lowpass).

Replaces `_FlashSRRunner` of the reference (egregora_audio_super_resolution.py:254-369): `infer` keeps its
contract — [N, 245760] f32 at 48 kHz in, same shape out — but takes every chunk-channel of a clip at once
and runs them through libegregora_b200.so in sub-batches.  Host code is plumbing only.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _abi
from . import flashsr_model as M
from .flashsr_plan import PlanBackend, WeightBlob, build_plan


class FlashSREngine:
    def __init__(self, device: torch.device, spec: Optional[dict] = None, weights: Optional[Dict[str, torch.Tensor]] = None,
                 seed: int = 0, max_batch: Optional[int] = None, debug: bool = False):
        if device.type != "cuda":
            raise RuntimeError("FlashSREngine needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = device
        self.lib = _abi.init(device.index or 0)
        self.spec = spec or M.default_spec()
        # No checkpoint exists in this environment: seeded random weights of the spec'd architecture.
        self.weights = weights if weights is not None else M.init_weights(self.spec, seed)
        # chunk-channels per plan launch.  The UNet's time grows slowly with the batch (latency-bound: 4.3 ms at 1, 6.2 ms at
        # 8, 8.2 at 16, 12.4 at 32, 16.9 at 48 per step) while VAE / vocoder scale linearly, so larger sub-batches amortise it:
        # c3 on one GPU 192 -> 204 -> 208 x real-time at 8 / 12 / 16, and on the final round-2 kernels 237 -> 250 -> 257 x at
        # 16 / 32 / 48 (tools/gpu_r2_nn.sh).  The workspace is ~0.33 GB per chunk-channel (16 GB at 48) of the 180 GB.
        self.max_batch = int(max_batch or os.environ.get("EGREGORA_FLASHSR_BATCH", "48"))
        self.debug = debug
        self.blob = WeightBlob()
        # one dry walk (batch 1, lowpass on) packs every weight/constant the graph can touch
        build_plan(self.spec, self.weights, self.blob, 1, 1, True)
        self.blob.frozen = True
        raw = self.blob.tobytes()
        self.d_weights = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
        # plan runs go to a private stream: the legacy default stream cannot be captured into the plan's CUDA graph;
        # ordering against the caller's stream is kept with events (wait_stream), never with a host sync
        self.stream = torch.cuda.Stream(device=device)
        self.ws: Optional[torch.Tensor] = None
        self.plans: Dict[Tuple[int, int, bool], Tuple[PlanBackend, int]] = {}
        self.launches_last = 0

    # ------------------------------------------------------------------ plans
    def _create(self, be: PlanBackend) -> int:
        ops = be.build_ops()
        handle = C.c_void_p()
        _abi.check(self.lib.egr_plan_create(ops, len(be.ops), self.ws.data_ptr(), self.ws.numel(), self.d_weights.data_ptr(),
                                            self.d_weights.numel(), C.byref(handle)), "egr_plan_create")
        return handle.value

    def plan(self, batch: int, steps: int, lowpass: bool) -> Tuple[PlanBackend, int]:
        key = (int(batch), int(steps), bool(lowpass))
        if key in self.plans:
            return self.plans[key]
        be = build_plan(self.spec, self.weights, self.blob, batch, steps, lowpass, debug=self.debug)
        need = be.ws_bytes + 4096
        if self.ws is None or self.ws.numel() < need:
            # a larger workspace: every existing plan holds absolute addresses into the old one, so all handles are
            # rebuilt.  Forget them BEFORE anything can raise (allocation, plan creation): a failed re-plan must not
            # leave freed handles behind for the next infer() / close() to use.
            torch.cuda.synchronize(self.device)
            old, self.plans = self.plans, {}
            for _, h in old.values():
                self.lib.egr_plan_destroy(h)
            self.ws = None
            self.ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            for k, (obe, _) in old.items():
                self.plans[k] = (obe, self._create(obe))
        self.plans[key] = (be, self._create(be))
        return self.plans[key]

    def view(self, buf, dtype, shape) -> torch.Tensor:
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        return self.ws[buf.offset: buf.offset + n].view(dtype).view(*shape)

    def read(self, be: PlanBackend, name: str) -> torch.Tensor:
        """debug: fetch a named intermediate as NCHW / [B,C,T] float32 on the CPU."""
        t = be.named[name]
        if t.f32 is not None:
            x = self.view(t.f32, torch.float32, (t.B, t.H, t.W, t.C)).float()
        elif t.f16_transposed:
            return self.view(t.f16, torch.float16, (t.B, t.C, t.H, t.W)).float().cpu()
        else:
            x = self.view(t.f16, torch.float16, (t.B, t.H, t.W, t.C)).float()
        return x.permute(0, 3, 1, 2).contiguous().cpu()

    # ------------------------------------------------------------------ inference
    def noise_shape(self, n: int) -> Tuple[int, int, int, int]:
        s = self.spec
        fr = s["chunk"] // s["mel"]["hop"]
        return (n, s["vae"]["embed_dim"], fr // 8, s["mel"]["n_mels"] // 8)

    def make_noise(self, n: int, seed: int, row0: int = 0) -> torch.Tensor:
        """x_T [n, z, T/8, F/8] (NCHW, on the device) for chunk-channel rows row0 .. row0+n-1 of a clip.  Counter-based
        (egr_noise_fill): a row's noise depends on (seed, global row index) only — not on the batch it runs in, not on
        how a clip is split over sub-batches or ranks — and no host RNG sits on the critical path."""
        shape = self.noise_shape(n)
        out = torch.empty(shape, dtype=torch.float32, device=self.device)
        if n:
            _abi.check(self.lib.egr_noise_fill(int(seed) & 0xFFFFFFFFFFFFFFFF, int(row0), n, out[0].numel(), out.data_ptr(),
                                               torch.cuda.current_stream(self.device).cuda_stream), "egr_noise_fill")
        return out

    def infer(self, x: torch.Tensor, lowpass: bool = False, steps: int = 1, seed: int = 4321,
              noise: Optional[torch.Tensor] = None, row0: int = 0) -> torch.Tensor:
        """x [N, chunk] f32 on the engine's device -> [N, chunk].  `row0`: global index of x[0] among the clip's
        chunk-channels (selects its diffusion noise when `noise` is not supplied)."""
        if x.dim() != 2 or x.shape[1] != self.spec["chunk"]:
            raise RuntimeError(f"FlashSR expects [N, {self.spec['chunk']}] chunks, got {tuple(x.shape)}")
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        N = x.shape[0]
        if noise is None:
            noise = self.make_noise(N, seed, row0)
        noise = noise.to(self.device, torch.float32).permute(0, 2, 3, 1).contiguous()  # NHWC
        out = torch.empty_like(x)
        caller = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(caller)
        self.launches_last = 0
        with torch.cuda.stream(self.stream):
            st = self.stream.cuda_stream
            # balanced sub-batches (33 chunk-channels -> 7,7,7,6,6 rather than 8,8,8,8,1: a trailing batch of one
            # runs the small UNet layers at a fraction of the batched efficiency)
            n_sub = -(-N // self.max_batch)
            base_b, extra = divmod(N, n_sub)
            sizes = [base_b + (1 if k < extra else 0) for k in range(n_sub)]
            i = 0
            for b in sizes:
                be, h = self.plan(b, steps, lowpass)
                wav_in, nz_in = be.inputs["wav"], be.inputs["noise"]
                self.view(wav_in.f32, torch.float32, (b, x.shape[1])).copy_(x[i:i + b])
                self.view(nz_in.f32, torch.float32, tuple(noise[i:i + b].shape)).copy_(noise[i:i + b])
                _abi.check(self.lib.egr_plan_run(h, 0, -1, st), "egr_plan_run")
                self.launches_last += len(be.ops)
                out[i:i + b].copy_(self.view(be.output.f32, torch.float32, (b, x.shape[1])))
                i += b
        caller.wait_stream(self.stream)
        for t in (x, noise, out):
            t.record_stream(self.stream)
        return out

    def close(self):
        for _, h in self.plans.values():
            self.lib.egr_plan_destroy(h)
        self.plans = {}
