import subprocess, sys, os
"""Sweep of the path-B tuning overrides (env vars read at plan creation): EGR_FFT_R1 / EGR_FFT_CW (split and column
tile), EGR_FFT_MAXR (radix cap), EGR_FL_ROW_T / EGR_FL_COL_T (threads per CTA), EGR_FL_MINB (__launch_bounds__ variant).
    python tools/fl_sweep.py [quick]"""
cfgs = [
    {},
    {"EGR_FFT_MAXR": "12"},
    {"EGR_FFT_MAXR": "10"},
    {"EGR_FFT_MAXR": "9"},
    {"EGR_FFT_MAXR": "10", "EGR_FL_MINB": "3"},
    {"EGR_FL_MINB": "3"},
    {"EGR_FFT_R1": "2100"},
    {"EGR_FFT_R1": "1800"},
    {"EGR_FFT_R1": "2520"},
]
for cfg in cfgs:
    env = dict(os.environ)
    env.update(cfg)
    code = '''
import sys, time, torch
sys.path.insert(0, ".")
import bench
bench.load_pkg()
from egregora_b200 import _abi
S, C, iters = 7938000, 2, 50
dev = torch.device("cuda", 0); lib = _abi.init(0)
x = (bench.synth_audio(S, C, sr=44100, seed=77) * 32767).round().to(dev); y = torch.empty_like(x)
wb = lib.egr_fatllama_workspace_bytes(C, S, 1); w = torch.empty(wb, dtype=torch.uint8, device=dev)
_abi.check(lib.egr_fatllama_run(x.data_ptr(), y.data_ptr(), C, S, 1, 3, 0.6, 3, w.data_ptr(), wb, 0)); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); _abi.check(lib.egr_fatllama_run(x.data_ptr(), y.data_ptr(), C, S, 1, iters, 0.6, 3, w.data_ptr(), wb, 0)); e1.record(); torch.cuda.synchronize()
print("us/iter", e0.elapsed_time(e1) * 1e3 / iters, "checksum", float(y.double().abs().sum()))
'''
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(f"{cfg}:", out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr.strip()[-300:], flush=True)
