"""Tiny-spec FlashSR plan (one chunk-channel, 1 step, lowpass ON) under the cluster-capable CPU emulator (tests/cusim,
EGR_TEST_CUSIM=clusters): every CUDA-core kernel of the c2 op set — incl. the 8-CTA low-pass and cluster GroupNorm — runs
its real source, tensor-core GEMM ops by host loops; prints the waveform RMS error vs the fp32 oracle and the detected
cutoff bins.  ~1 min, no GPU.
    python tools/cusim_e2e.py"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = str(Path(__file__).resolve().parents[1])
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + '/tests'); sys.path.insert(0, ROOT + '/tests/cusim')
os.environ["EGR_TEST_CUSIM"]="clusters"
from conftest import load_pkg
load_pkg()
import harness
from egregora_b200 import flashsr_model as M, flashsr_plan as P
from oracle import flashsr_oracle
lib=harness.cusim_lib()
spec=M.tiny_spec(); W=M.init_weights(spec,0); blob=P.WeightBlob(); B=1
be=P.build_plan(spec,W,blob,B,1,True)
ws=torch.zeros(be.ws_bytes+4096,dtype=torch.uint8); wt=torch.frombuffer(bytearray(blob.tobytes()),dtype=torch.uint8).clone()
g=torch.Generator().manual_seed(5)
wav=(0.1*torch.randn(B,spec["chunk"],generator=g)).cumsum(1)*0.05; wav=wav-wav.mean(1,keepdim=True); wav=(wav/wav.abs().max()*0.5).float()
fr=spec["chunk"]//spec["mel"]["hop"]
noise=torch.randn((B,spec["vae"]["embed_dim"],fr//8,spec["mel"]["n_mels"]//8),generator=torch.Generator().manual_seed(4321))
def view(buf,dt,shape):
    n=int(np.prod(shape))*torch.empty((),dtype=dt).element_size(); return ws[buf.offset:buf.offset+n].view(dt).view(*shape)
view(be.inputs["wav"].f32,torch.float32,wav.shape).copy_(wav)
nz=noise.permute(0,2,3,1).contiguous(); view(be.inputs["noise"].f32,torch.float32,nz.shape).copy_(nz)
h=C.c_void_p()
assert lib.egr_plan_create(be.build_ops(),len(be.ops),ws.data_ptr(),ws.numel(),wt.data_ptr(),wt.numel(),C.byref(h))==0, lib.egr_last_error()
rc=lib.egr_plan_run(h,0,-1,None); assert rc==0, lib.egr_last_error()
y=view(be.output.f32,torch.float32,wav.shape).clone()
yo,obe=flashsr_oracle.run_flashsr(spec,W,wav,noise,steps=1,lowpass=True)
print('rms',float((y-yo).pow(2).mean().sqrt()), 'cutoff', list(view(be.cutoff_buf,torch.int32,(B,))), obe.cutoff_bins)
