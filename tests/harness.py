"""Test helper: build a tiny plan out of PlanBackend calls, run it on the GPU (or through the CPU interpreter)
and read tensors back as NCHW float32."""
import ctypes as C

import numpy as np
import torch

from conftest import load_pkg

load_pkg()
from egregora_b200 import _abi, flashsr_model as M, flashsr_plan as P  # noqa: E402


_CUSIM = None


def cusim_lib():
    """ctypes handle of the emulator build (tests/cusim/build.py), prototypes declared from _abi.signatures().
    EGR_TEST_CUSIM=clusters picks the slower variant in which thread-block clusters work."""
    global _CUSIM
    if _CUSIM is None:
        import sys
        from pathlib import Path
        sys.path.insert(0, str(Path(__file__).resolve().parent / "cusim"))
        import build as cusim_build
        import os
        lib = C.CDLL(str(cusim_build.build(clusters=os.environ.get("EGR_TEST_CUSIM") == "clusters")))
        for name, (res, args) in _abi.signatures().items():
            if hasattr(lib, name):
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
        assert lib.egr_init(0) == 0, lib.egr_last_error()
        _CUSIM = lib
    return _CUSIM


class MiniPlan:
    def __init__(self, weights=None, spec=None, batch=1):
        self.blob = P.WeightBlob()
        self.be = P.PlanBackend(spec or M.tiny_spec(), weights or {}, self.blob, batch)
        self.be.debug = True
        self.inputs = []

    def input(self, x_nchw: torch.Tensor, f16=False):
        """x [B,C,H,W] -> persistent channels-last plan tensor holding it (f32 and optionally f16)."""
        B, Cc, H, W = x_nchw.shape
        t = self.be.new(B, H, W, Cc, f32=True, f16=f16, tag=f"in{len(self.inputs)}", persistent=True)
        self.inputs.append((t, x_nchw.permute(0, 2, 3, 1).contiguous().float()))
        return t

    # ---- execution
    def _finish(self):
        self.ws_bytes = self.be.allocate()
        self.ops = self.be.build_ops()

    def run_gpu(self, device="cuda:0", first=0, last=-1):
        import os
        if os.environ.get("EGR_TEST_INTERP"):  # dry-run the GPU tests' references through the CPU interpreter
            return self.run_cpu()
        if os.environ.get("EGR_TEST_CUSIM"):   # run the SIMT kernels' real source under the CPU emulator (tests/cusim)
            return self.run_sim(first, last)
        self._finish()
        dev = torch.device(device)
        lib = _abi.init(dev.index or 0)
        self.ws = torch.zeros(self.ws_bytes + 4096, dtype=torch.uint8, device=dev)
        self.wt = torch.frombuffer(bytearray(self.blob.tobytes() or b"\0" * 256), dtype=torch.uint8).to(dev)
        for t, x in self.inputs:
            self.view(t.f32, torch.float32, x.shape).copy_(x)
            if t.f16 is not None:
                self.view(t.f16, torch.float16, x.shape).copy_(x.half())
        h = C.c_void_p()
        _abi.check(lib.egr_plan_create(self.ops, len(self.be.ops), self.ws.data_ptr(), self.ws.numel(), self.wt.data_ptr(),
                                       self.wt.numel(), C.byref(h)), "egr_plan_create")
        _abi.check(lib.egr_plan_run(h, first, last, torch.cuda.current_stream(dev).cuda_stream), "egr_plan_run")
        torch.cuda.synchronize(dev)
        lib.egr_plan_destroy(h)
        return self

    def run_sim(self, first=0, last=-1):
        """The same plan through tests/cusim's build of the library: host memory stands in for device memory, the SIMT
        kernels execute their real source, tensor-core GEMM ops are evaluated by gemm_tc_ref.cpp's plain loops."""
        lib = cusim_lib()
        self._finish()
        self.ws = torch.zeros(self.ws_bytes + 4096, dtype=torch.uint8)
        self.wt = torch.frombuffer(bytearray(self.blob.tobytes() or b"\0" * 256), dtype=torch.uint8).clone()
        for t, x in self.inputs:
            self.view(t.f32, torch.float32, x.shape).copy_(x)
            if t.f16 is not None:
                self.view(t.f16, torch.float16, x.shape).copy_(x.half())
        h = C.c_void_p()
        rc = lib.egr_plan_create(self.ops, len(self.be.ops), self.ws.data_ptr(), self.ws.numel(), self.wt.data_ptr(), self.wt.numel(), C.byref(h))
        assert rc == 0, lib.egr_last_error().decode()
        rc = lib.egr_plan_run(h, first, last, None)
        assert rc == 0, lib.egr_last_error().decode()
        lib.egr_plan_destroy(h)
        return self

    def run_cpu(self):
        from plan_interp import Interp
        self._finish()
        it = Interp(_abi.K, self.ops, self.ws_bytes, self.blob.tobytes() or b"\0" * 256)
        self.ws = it.ws
        for t, x in self.inputs:
            self.view(t.f32, torch.float32, x.shape).copy_(x)
            if t.f16 is not None:
                self.view(t.f16, torch.float16, x.shape).copy_(x.half())
        it.run()
        return self

    def view(self, buf, dtype, shape):
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        return self.ws[buf.offset: buf.offset + n].view(dtype).view(*shape)

    def read(self, t):
        """plan tensor -> NCHW float32 CPU."""
        if t.f32 is not None:
            return self.view(t.f32, torch.float32, (t.B, t.H, t.W, t.C)).float().permute(0, 3, 1, 2).contiguous().cpu()
        if t.f16_transposed:
            return self.view(t.f16, torch.float16, (t.B, t.C, t.H, t.W)).float().cpu()
        return self.view(t.f16, torch.float16, (t.B, t.H, t.W, t.C)).float().permute(0, 3, 1, 2).contiguous().cpu()


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-30))
