"""The SIMT kernels' REAL source run on the CPU under tests/cusim (a fiber-per-thread emulation of blocks, shared memory,
__syncthreads, warp shuffles and atomics — test infrastructure, see tests/cusim/cuda_runtime.h), against the same goldens
and oracles as the `-m gpu` tests.  Covers: egr_chunk_gather / egr_wola_stitch / egr_resample_poly (bit-exact),
egr_pcm16_* / egr_absmax, egr_dfn_mix, egr_eval_null_test / _lsd / _lufs / _hf_band, egr_fft_exec, egr_fatllama_run, and a
whole FlashSR plan (tiny spec: front end, VAE, UNet, sampler, vocoder) through egr_plan_create / egr_plan_run, where every
CUDA-core kernel (GroupNorm, LayerNorm, attention, GEGLU, snake, STFT-mel, element-wise glue, SIMT convs) executes its
real source and the tensor-core GEMM ops are evaluated by plain loops (tests/cusim/gemm_tc_ref.cpp).  The whole
`-m gpu` op suite also runs this way on demand: `EGR_TEST_CUSIM=1 pytest tests/test_ops_gpu.py -m gpu` (60 of 65 pass;
the other five need thread-block clusters, which the default build refuses loudly and the slower
`EGR_TEST_CUSIM=clusters` variant runs: 65 of 65).

What this is for: (1) kernels written without GPU time (egr_eval_lsd) execute their actual code — indexing, barriers,
radix select — before their first hardware run; (2) the emulator reproducing what the B200 already verified for the other
kernels is the check on the emulator itself.  What it is not: the product never loads this library (there is no CPU
fallback, tests/test_abi.py::test_no_gpu_fails_loudly), nothing here says anything about speed, and the tcgen05 / TMA /
kernel (the tap-GEMM) is outside its reach — that one stays GPU-only.
"""
import ctypes as C
import hashlib
import json
import shutil
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_pkg
from lsd_cases import signals as lsd_signals
from lufs_cases import signal as lufs_signal
from hf_cases import signal as hf_signal

sys.path.insert(0, str(ROOT / "tests" / "cusim"))

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ is needed to build the emulator")


POISON = False   # True: scratch buffers start as 0xFF bytes (NaN as float / double), like torch.empty on a used GPU heap


def _dev(nbytes_or_array):
    """256-byte aligned host buffer standing in for device memory; accepts a byte count or an array to copy."""
    if isinstance(nbytes_or_array, (int, np.integer)):
        raw = np.full(int(nbytes_or_array) + 256, 0xFF if POISON else 0, np.uint8)
        off = (-raw.ctypes.data) % 256
        return raw[off:off + int(nbytes_or_array)]
    a = np.ascontiguousarray(nbytes_or_array)
    buf = _dev(max(a.nbytes, 1))
    out = buf[:a.nbytes].view(a.dtype).reshape(a.shape)
    out[...] = a
    return out


@pytest.fixture(scope="session")
def sim():
    import build as cusim_build
    load_pkg()
    from egregora_b200 import _abi
    lib = C.CDLL(str(cusim_build.build()))
    for name, (res, args) in _abi.signatures().items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
    assert lib.egr_init(0) == 0, lib.egr_last_error()

    def ck(rc):
        assert rc == 0, lib.egr_last_error().decode()
    lib.ck = ck
    lib.K = _abi.K
    return lib


# ------------------------------------------------------------------------------------------------ driver kernels
def _wola(sim, preds, spans, total, win, window):
    n, Cc, lp = preds.shape
    p, w = _dev(preds), _dev(window)
    st, ln = _dev(np.array([s for s, _ in spans], np.int64)), _dev(np.array([l for _, l in spans], np.int32))
    out = _dev(np.zeros((Cc, total), np.float32))
    sim.ck(sim.egr_wola_stitch(p.ctypes.data, lp, st.ctypes.data, ln.ctypes.data, n, Cc, total, win, w.ctypes.data, out.ctypes.data, None))
    return out


@pytest.mark.parametrize("case", list("abcdef"))
def test_wola_reference_golden_cases_bit_exact(case, golden, sim):
    total, w, hp, Cc, lpred = (int(v) for v in golden[f"wola_{case}_meta"])
    spans = [(int(s), int(L)) for s, L in golden[f"wola_{case}_spans"]]
    out = _wola(sim, golden[f"wola_{case}_preds"], spans, total, w, np.hanning(w).astype(np.float32))
    assert np.array_equal(out, golden[f"wola_{case}_out"])


@pytest.mark.parametrize("total,Cc", [(300000, 2), (245761, 1), (100, 1)])
def test_gather_and_wola_real_window_bit_exact(total, Cc, sim):
    from oracle import driver_oracle as O
    x = (np.random.default_rng(total).standard_normal((Cc, total)) * 0.1).astype(np.float32)
    win, hop = O.win_hop()
    spans = O.iter_chunks(total, win, hop)
    xd = _dev(x)
    st, ln = _dev(np.array([s for s, _ in spans], np.int64)), _dev(np.array([l for _, l in spans], np.int32))
    chunks = _dev(np.zeros((len(spans), Cc, win), np.float32))
    sim.ck(sim.egr_chunk_gather(xd.ctypes.data, Cc, total, st.ctypes.data, ln.ctypes.data, len(spans), win, chunks.ctypes.data, None))
    assert np.array_equal(chunks, O.gather_chunks(x, spans, win))
    y = (chunks * np.float32(0.5)).astype(np.float32)
    got = _wola(sim, y, spans, total, win, np.hanning(win).astype(np.float32))
    want = O.wola_stitch([(y[k], s, L) for k, (s, L) in enumerate(spans)], total, win)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("src,dst,n", [(16000, 48000, 5000), (44100, 48000, 4410), (48000, 44100, 9999), (48000, 96000, 3001)])
def test_resample_poly_bit_exact_vs_scipy(src, dst, n, sim, pkg):
    from scipy.signal import resample_poly
    from egregora_b200 import egregora_audio_super_resolution as N
    x = (np.random.default_rng(n).standard_normal((2, n)) * 0.3).astype(np.float32)
    up, down, hflip, n_pre_remove, n_out = N._resample_design(dst, src, n)
    xd, bank, y = _dev(x), _dev(hflip), _dev(np.zeros((2, n_out), np.float32))
    sim.ck(sim.egr_resample_poly(xd.ctypes.data, 2, n, up, down, bank.ctypes.data, hflip.shape[1], n_pre_remove, n_out, y.ctypes.data, None))
    want = np.stack([resample_poly(x[c], up, down).astype(np.float32) for c in range(2)])
    assert np.array_equal(y, want)


def test_pcm16_and_absmax(sim):
    from oracle import fat_llama_oracle as FO
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-1.2, 1.2, 20001), [1.0, -1.0, 0.5 / 32768, 1.5 / 32768, 2.5 / 32768, 0.99999]]).astype(np.float32)
    xd, q, f, m = _dev(x), _dev(np.zeros(x.shape, np.int16)), _dev(np.zeros(x.shape, np.float32)), _dev(np.zeros(1, np.float32))
    sim.ck(sim.egr_pcm16_quantize(xd.ctypes.data, q.ctypes.data, x.size, None))
    assert np.array_equal(q, FO.pcm16_write(x))
    sim.ck(sim.egr_pcm16_to_float(q.ctypes.data, f.ctypes.data, x.size, 1.0 / 32768.0, None))
    assert np.array_equal(f, FO.pcm16_read(FO.pcm16_write(x)))
    sim.ck(sim.egr_absmax(xd.ctypes.data, x.size, m.ctypes.data, None))
    assert float(m[0]) == float(np.max(np.abs(x)))


# ------------------------------------------------------------------------------------------------ dfn mix
def test_dfn_mix_reference_golden(sim):
    from test_dfn_mix import DEFAULTS, _check, _signals
    G, cases = np.load(GOLDEN / "dfn_mix_golden.npz"), json.loads((GOLDEN / "dfn_mix_cases.json").read_text())
    K = sim.K
    modes = {"off": "EGR_MIX_OFF", "more_on_noise": "EGR_MIX_MORE_ON_NOISE", "more_on_speech": "EGR_MIX_MORE_ON_SPEECH",
             "gate_on_noise": "EGR_MIX_GATE_ON_NOISE"}
    for name, c in cases.items():
        if c["T"] > 48000 * 30:
            continue  # the emulator is ~1000x slower than the device: clip-scale cases stay GPU-only
        dry, wet = _signals(name, c["C"], c["T"])
        p = dict(DEFAULTS)
        p.update(c["kwargs"])
        d, w, out = _dev(dry), _dev(wet), _dev(np.zeros_like(dry))
        wb = sim.egr_dfn_mix_workspace_bytes(c["C"], c["T"])
        wk = _dev(wb)
        vad = K["EGR_VAD_RMS"] if p.get("adaptive_vad_source", "rms") == "rms" else K["EGR_VAD_NONE"]
        sim.ck(sim.egr_dfn_mix(d.ctypes.data, w.ctypes.data, out.ctypes.data, c["C"], c["T"], 48000, float(p["strength"]),
                               K["EGR_CURVE_EQUAL_POWER"] if p["mix_curve"] == "equal_power" else K["EGR_CURVE_LINEAR"], vad,
                               K[modes.get(p["adaptive_mode"], "EGR_MIX_OFF")], float(p["adaptive_amount"]), float(p["vad_threshold"]),
                               int(p["vad_smooth_ms"]), float(p["post_gain_db"]), 1 if p["limit_ceiling"] else 0, float(p["ceiling"]),
                               wk.ctypes.data, wb, None))
        _check(G, name, out, exact=p["mix_curve"] != "equal_power")


# ------------------------------------------------------------------------------------------------ eval metrics
def test_eval_null_test_reference_golden(sim):
    from test_eval_metrics import _check_null, _close, _signals
    egold = json.loads((GOLDEN / "eval_golden.json").read_text())
    K = sim.K
    for name, c in egold.items():
        if max(c["Na"], c["Nb"]) > 100000:
            continue
        A, B = _signals(name, c)
        n = min(A.shape[1], B.shape[1])
        a, b, null, met = _dev(A), _dev(B), _dev(np.zeros((c["C"], n), np.float32)), _dev(np.zeros(K["EGR_EVAL_NUM"], np.float64))
        wb = sim.egr_eval_workspace_bytes()
        wk = _dev(wb)
        sim.ck(sim.egr_eval_null_test(a.ctypes.data, A.shape[1], b.ctypes.data, B.shape[1], c["C"], n, int(c["invert_b"]),
                                      int(c["least_squares_scale"]), null.ctypes.data, met.ctypes.data, wk.ctypes.data, wb, None))
        _check_null(c, null)
        ref = c["metrics"]
        assert int(met[K["EGR_EVAL_OVERSHOOT"]]) == ref["overshoot_count"]
        assert _close(met[K["EGR_EVAL_SCALE_K"]], ref["scale_k"], 1e-9)
        assert _close(met[K["EGR_EVAL_NULL_RMS_DBFS"]], ref["null_rms_dbfs"], 1e-9)
        assert _close(met[K["EGR_EVAL_CORR"]], ref["corr_coef"], 0.0, 3e-6)
        assert _close(met[K["EGR_EVAL_SI_SDR_DB"]], c["si_sdr_db"], 1e-9, 1e-9)


def test_eval_lsd_reference_golden(sim):
    """egr_eval_lsd's actual kernels (tables, per-frame STFT distance, radix-select percentile) against the reference
    node's numbers: 1e-4 dB wherever every bin holds signal (see test_eval_metrics.py for the noise-floor case)."""
    lgold = json.loads((GOLDEN / "eval_lsd_golden.json").read_text())
    K = sim.K
    for name, c in lgold.items():
        A, B = lsd_signals(name, c)
        n = min(A.shape[1], B.shape[1])
        a, b, met = _dev(A), _dev(B), _dev(np.zeros(K["EGR_LSD_NUM"], np.float64))
        wb = sim.egr_eval_lsd_workspace_bytes(n, c["n_fft"], c["hop"])
        wk = _dev(wb)
        sim.ck(sim.egr_eval_lsd(a.ctypes.data, A.shape[1], b.ctypes.data, B.shape[1], c["C"], n, c["n_fft"], c["hop"], 1.0,
                                met.ctypes.data, wk.ctypes.data, wb, None))
        assert int(met[K["EGR_LSD_FRAMES"]]) == c["frames"]
        if c["band"] >= 1.0:
            assert abs(met[K["EGR_LSD_MEAN_DB"]] - c["lsd_mean_db"]) <= 1e-4, (name, met)
            assert abs(met[K["EGR_LSD_P95_DB"]] - c["lsd_p95_db"]) <= 1e-4, (name, met)
        else:
            assert abs(met[K["EGR_LSD_MEAN_DB"]] - c["lsd_mean_db"]) <= 8.0
    # identical clips: exactly sqrt(1e-12) per frame; unsupported n_fft is an error, not a fallback
    x = _dev((np.random.default_rng(3).standard_normal((2, 20000)) * 0.1).astype(np.float32))
    met = _dev(np.zeros(4, np.float64))
    wb = sim.egr_eval_lsd_workspace_bytes(20000, 2048, 512)
    wk = _dev(wb)
    sim.ck(sim.egr_eval_lsd(x.ctypes.data, 20000, x.ctypes.data, 20000, 2, 20000, 2048, 512, 1.0, met.ctypes.data, wk.ctypes.data, wb, None))
    assert met[0] == float(np.sqrt(np.float32(1e-12))) and met[1] == met[0]
    assert sim.egr_eval_lsd(x.ctypes.data, 20000, x.ctypes.data, 20000, 2, 20000, 640, 160, 1.0, met.ctypes.data, wk.ctypes.data, wb, None) != 0


def _lufs(sim, x, sr):
    K = sim.K
    Cc, n = x.shape
    xd, met = _dev(x), _dev(np.zeros(K["EGR_LUFS_NUM"], np.float64))
    wb = sim.egr_eval_lufs_workspace_bytes(Cc, n, sr)
    wk = _dev(wb)
    sim.ck(sim.egr_eval_lufs(xd.ctypes.data, n, Cc, n, sr, met.ctypes.data, wk.ctypes.data, wb, None))
    # the filtered signal sits at offset 256 of the workspace (layout of egr_eval_lufs): checked bit for bit below
    y = wk[256:256 + 4 * Cc * n].view(np.float32).reshape(Cc, n).copy()
    return met.copy(), y


def test_eval_lufs_reference_golden(sim):
    """egr_eval_lufs: the speculated-and-verified float32 high-pass is bit-identical to the reference's per-sample
    Python loop (sha256 of the filtered signal before the tilt), the loudness agrees to 1e-9 dB."""
    from oracle import eval_oracle as O
    lg = json.loads((GOLDEN / "eval_lufs_golden.json").read_text())
    K = sim.K
    for name, c in lg.items():
        x = lufs_signal(name, c)
        met, y = _lufs(sim, x, c["sr"])
        y[:, 1:] += np.float32(0.02) * (y[:, 1:] - y[:, :-1])   # the tilt the block kernel applies on the fly
        assert hashlib.sha256(np.ascontiguousarray(y).tobytes()).hexdigest() == c["kweight_sha256"], name
        assert abs(met[K["EGR_LUFS_INTEGRATED"]] - c["lufs"]) <= 1e-9, (name, met)
        blk, hop = round(0.4 * c["sr"]), round(0.1 * c["sr"])
        assert int(met[K["EGR_LUFS_BLOCKS"]]) == 1 + max(0, (c["N"] - blk) // hop)
    # a case built to defeat the speculation: a 1e30 burst leaves a filter state that is still ~8 two chunks later
    # (k^4096 ~ 1e-28), far above the 1e-3 signal the guess for that chunk is computed from -> the chunk must be
    # repaired, and the result must stay bit-exact
    x = np.zeros((1, 3 * 2048 + 100), np.float32)
    x[0, :5] = 1e30
    x[0, 2048:] = (np.random.default_rng(1).standard_normal(x.shape[1] - 2048) * 1e-3).astype(np.float32)
    met, y = _lufs(sim, x, 48000)
    y[:, 1:] += np.float32(0.02) * (y[:, 1:] - y[:, :-1])
    assert met[K["EGR_LUFS_REPAIRED"]] >= 1
    assert np.array_equal(y, O.k_weight(48000, x))
    assert abs(met[K["EGR_LUFS_INTEGRATED"]] - O.integrated_lufs(x, 48000)) <= 1e-9


def test_eval_hf_band_reference_golden(sim):
    """egr_eval_hf_band (pack -> path-B FFT, incl. odd and prime lengths -> float64 band sums) against the reference's
    _band_energy_hi_db: 1e-3 dB (the reference sums |X|^2 in float32, both FFTs are float32), bin mask identical."""
    hg = json.loads((GOLDEN / "eval_hf_golden.json").read_text())
    K = sim.K
    for name, c in hg.items():
        x = hf_signal(name, c)
        plan = C.c_void_p()
        sim.ck(sim.egr_fft_plan_create(c["N"], 1, C.byref(plan)))
        try:
            xd, met = _dev(x), _dev(np.zeros(K["EGR_HF_NUM"], np.float64))
            wb = sim.egr_eval_hf_band_workspace_bytes(plan, c["N"])
            wk = _dev(wb)
            sim.ck(sim.egr_eval_hf_band(plan, xd.ctypes.data, c["N"], c["C"], c["N"], c["sr"], float(c["lo_hz"]), met.ctypes.data,
                                        wk.ctypes.data, wb, None))
        finally:
            sim.egr_fft_plan_destroy(plan)
        assert int(met[K["EGR_HF_BINS_HI"]]) == c["bins_hi"], name
        assert abs(met[K["EGR_HF_RESIDUAL_DB"]] - c["hf_db"]) <= 1e-3, (name, met, c["hf_db"])


def test_full_null_test_reference_golden(sim):
    """Audio_Null_Test.execute with every toggle on (goldens from the reference node): the same sequence of C-ABI calls
    `egregora_eval_metrics.audio_null_test` makes — null test, LUFS of the null, LSD of A vs the scaled B, HF band of the
    null — through the emulator."""
    G = json.loads((GOLDEN / "null_full_golden.json").read_text())
    K = sim.K
    for name, c in G.items():
        rng = np.random.default_rng(sum(map(ord, name)))
        t = np.arange(c["N"]) / c["sr"]
        A = (0.3 * np.sin(2 * np.pi * 440 * t) + 0.1 * rng.standard_normal((c["C"], c["N"]))).astype(np.float32)
        B = (c["gain"] * A + c["noise"] * rng.standard_normal((c["C"], c["N"]))).astype(np.float32)
        n, Cc, ref = c["N"], c["C"], c["metrics"]
        a, b, null, met = _dev(A), _dev(B), _dev(np.zeros((Cc, n), np.float32)), _dev(np.zeros(K["EGR_EVAL_NUM"], np.float64))
        wb = sim.egr_eval_workspace_bytes()
        wk = _dev(wb)
        sim.ck(sim.egr_eval_null_test(a.ctypes.data, n, b.ctypes.data, n, Cc, n, int(c["invert_b"]), int(c["least_squares_scale"]),
                                      null.ctypes.data, met.ctypes.data, wk.ctypes.data, wb, None))
        assert abs(met[K["EGR_EVAL_CORR"]] - ref["corr_coef"]) <= 3e-6
        assert abs(met[K["EGR_EVAL_NULL_RMS_DBFS"]] - ref["null_rms_dbfs"]) <= 1e-9 * abs(ref["null_rms_dbfs"])
        assert int(met[K["EGR_EVAL_OVERSHOOT"]]) == ref["overshoot_count"]
        assert abs(met[K["EGR_EVAL_CLIPPED_PCT"]] - ref["clipped_pct"]) <= 1e-12
        k = float(met[K["EGR_EVAL_SCALE_K"]])
        assert abs(k - ref["scale_k"]) <= 1e-9 * abs(ref["scale_k"])
        lufs, _ = _lufs(sim, null, c["sr"])
        assert abs(lufs[K["EGR_LUFS_INTEGRATED"]] - ref["null_lufs"]) <= 1e-9
        lm = _dev(np.zeros(K["EGR_LSD_NUM"], np.float64))
        wb = sim.egr_eval_lsd_workspace_bytes(n, 2048, 512)
        wk = _dev(wb)
        gain = float(np.float32(k)) if c["least_squares_scale"] else 1.0
        sim.ck(sim.egr_eval_lsd(a.ctypes.data, n, b.ctypes.data, n, Cc, n, 2048, 512, gain, lm.ctypes.data, wk.ctypes.data, wb, None))
        assert abs(lm[K["EGR_LSD_MEAN_DB"]] - ref["lsd_mean_db"]) <= 1e-4 and abs(lm[K["EGR_LSD_P95_DB"]] - ref["lsd_p95_db"]) <= 1e-4
        plan = C.c_void_p()
        sim.ck(sim.egr_fft_plan_create(n, 1, C.byref(plan)))
        try:
            hm = _dev(np.zeros(K["EGR_HF_NUM"], np.float64))
            wb = sim.egr_eval_hf_band_workspace_bytes(plan, n)
            wk = _dev(wb)
            sim.ck(sim.egr_eval_hf_band(plan, null.ctypes.data, n, Cc, n, c["sr"], 8000.0, hm.ctypes.data, wk.ctypes.data, wb, None))
        finally:
            sim.egr_fft_plan_destroy(plan)
        assert abs(hm[K["EGR_HF_RESIDUAL_DB"]] - ref["hf_residual_db"]) <= 1e-3


# ------------------------------------------------------------------------------------------------ path B
def _fft(sim, x, inverse=False, scale=True):
    batch, n = x.shape
    plan = C.c_void_p()
    sim.ck(sim.egr_fft_plan_create(n, batch, C.byref(plan)))
    try:
        d = _dev(np.ascontiguousarray(x).view(np.float32).reshape(batch, 2 * n))
        wb = sim.egr_fft_plan_workspace_bytes(plan)
        w = _dev(max(wb, 16))
        sim.ck(sim.egr_fft_exec(plan, d.ctypes.data, w.ctypes.data, int(inverse), int(scale), None))
        return d.view(np.complex64).reshape(batch, n).copy()
    finally:
        sim.egr_fft_plan_destroy(plan)


def _rel(a, b):
    return float(np.sqrt(np.mean(np.abs(a - b) ** 2)) / (np.sqrt(np.mean(np.abs(b) ** 2)) + 1e-30))


@pytest.mark.parametrize("n", [1, 2, 3, 5, 7, 11, 13, 16, 60, 210, 1000, 2100, 4096, 8190, 30030])
def test_fft_smooth_lengths(n, sim):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n))).astype(np.complex64)
    assert _rel(_fft(sim, x), np.fft.fft(x.astype(np.complex128), axis=1)) < 3e-6
    assert _rel(_fft(sim, x, inverse=True), np.fft.ifft(x.astype(np.complex128), axis=1)) < 1e-5


@pytest.mark.parametrize("n", [17, 10007])
def test_fft_bluestein_lengths(n, sim):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal((1, n)) + 1j * rng.standard_normal((1, n))).astype(np.complex64)
    assert _rel(_fft(sim, x), np.fft.fft(x.astype(np.complex128), axis=1)) < 1e-5


@pytest.mark.parametrize("Cc,S,U,iters,thr", [(1, 16000, 1, 5, 0.5), (2, 9000, 2, 3, 0.6), (1, 4410, 1, 4, 300.0), (1, 10007, 1, 3, 0.5), (2, 4801, 1, 2, 0.6)])
def test_fatllama_loop_matches_oracle(Cc, S, U, iters, thr, sim):
    """egr_fatllama_run (fast two-kernel path for plannable even lengths, general / Bluestein path otherwise) against
    the float64 oracle, pre-quantisation, 1e-5 of full scale (north_star tolerance)."""
    from oracle import fat_llama_oracle as O
    from test_fatllama_gpu import _audio
    x = _audio(Cc, S, seed=S)
    samples = O.pcm16_write(x.T).astype(np.float32)            # [S,C] integer-scaled
    want = O.upscale(samples, U, iters, thr, True, True, dtype=np.float64).T
    d_in, d_out = _dev(np.ascontiguousarray(samples.T)), _dev(np.zeros((Cc, S * U), np.float32))
    wb = sim.egr_fatllama_workspace_bytes(Cc, S, U)
    wk = _dev(wb)
    flags = sim.K["EGR_FL_NORMALIZE"] | sim.K["EGR_FL_AUTOSCALE"]
    sim.ck(sim.egr_fatllama_run(d_in.ctypes.data, d_out.ctypes.data, Cc, S, U, iters, thr, flags, wk.ctypes.data, wb, None))
    assert d_out.shape == want.shape
    assert float(np.max(np.abs(d_out - want))) <= 1e-5


# ------------------------------------------------------------------------------------------------ attention, strided q/k/v
@pytest.mark.parametrize("S,heads,hd", [(32, 4, 16), (96, 2, 32)])
def test_attention_reads_fused_qkv_column_blocks(S, heads, hd, sim, pkg):
    """attn_rows_kernel<HD, STRIDED=true> (the opt-in EGR_FUSE_QKV plan variant): q, k, v as column blocks of one
    [B*S, 3C] f16 tensor against torch on the same f16-rounded inputs."""
    import torch
    from harness import MiniPlan, rel_err
    from egregora_b200.flashsr_plan import PT
    Cc, B = heads * hd, 2
    g = torch.Generator().manual_seed(S)
    qkv = torch.randn((B, 3 * Cc, 1, S), generator=g)
    mp = MiniPlan()
    t = mp.input(qkv, f16=True)
    q, k, v = (PT(B, 1, S, Cc, f16=t.f16, ld=3 * Cc, coff=i * Cc) for i in range(3))
    o = mp.be.attention(q, k, v, heads, hd)
    mp.run_sim()

    def sp(i):
        x = qkv[:, i * Cc:(i + 1) * Cc].half().float()[:, :, 0].permute(0, 2, 1)
        return x.reshape(B, S, heads, hd).permute(0, 2, 1, 3)
    ref = torch.softmax(sp(0) @ sp(1).transpose(-1, -2) * hd ** -0.5, -1) @ sp(2)
    ref = ref.permute(0, 2, 1, 3).reshape(B, S, Cc).permute(0, 2, 1)
    assert rel_err(mp.read(o)[:, :, 0], ref) < 2e-3


# ------------------------------------------------------------------------------------------------ whole plan
def test_flashsr_tiny_plan_end_to_end(sim, pkg, monkeypatch):
    """Tiny-spec FlashSR, one chunk-channel, 1 step, lowpass off (the low-pass kernel needs a cluster), through the real
    plan executor: waveform RMS error vs the fp32 oracle < 1e-3 (north_star tolerance, as in the GPU test).  GroupNorm
    runs its one-CTA-per-group form (EGR_GN_NO_CLUSTER, the library's own switch; same sums, different order)."""
    monkeypatch.setenv("EGR_GN_NO_CLUSTER", "1")
    import torch
    from egregora_b200 import flashsr_model as M, flashsr_plan as P
    from oracle import flashsr_oracle
    spec = M.tiny_spec()
    W = M.init_weights(spec, 0)
    blob = P.WeightBlob()
    B = 1   # ~50 s under emulation; the GPU test covers B in {1, 2, 3}
    be = P.build_plan(spec, W, blob, B, 1, False)
    ws = _dev(be.ws_bytes + 4096)
    ws[:] = 0xFF   # the engine hands the plan a torch.empty workspace: every byte a NaN until some op writes it
    wt = _dev(np.frombuffer(blob.tobytes(), np.uint8))
    g = torch.Generator().manual_seed(5)
    wav = (0.1 * torch.randn(B, spec["chunk"], generator=g)).cumsum(1) * 0.05
    wav = wav - wav.mean(1, keepdim=True)
    wav = (wav / wav.abs().max() * 0.5).float()
    fr = spec["chunk"] // spec["mel"]["hop"]
    noise = torch.randn((B, spec["vae"]["embed_dim"], fr // 8, spec["mel"]["n_mels"] // 8), generator=torch.Generator().manual_seed(4321))

    def put(buf, arr):
        a = np.ascontiguousarray(arr, np.float32)
        ws[buf.offset: buf.offset + a.nbytes] = a.view(np.uint8).reshape(-1)

    put(be.inputs["wav"].f32, wav.numpy())
    put(be.inputs["noise"].f32, noise.permute(0, 2, 3, 1).contiguous().numpy())   # NHWC
    ops = be.build_ops()
    h = C.c_void_p()
    sim.ck(sim.egr_plan_create(ops, len(be.ops), ws.ctypes.data, ws.size, wt.ctypes.data, wt.size, C.byref(h)))
    try:
        sim.ck(sim.egr_plan_run(h, 0, -1, None))
    finally:
        sim.egr_plan_destroy(h)
    off = be.output.f32.offset
    y = ws[off: off + 4 * B * spec["chunk"]].view(np.float32).reshape(B, spec["chunk"])
    yo, _ = flashsr_oracle.run_flashsr(spec, W, wav, noise, steps=1, lowpass=False)
    rms = float(np.sqrt(np.mean((y - yo.numpy()) ** 2)))
    assert np.isfinite(y).all() and rms < 1e-3, rms


def test_cluster_kernels_are_refused_not_faked(sim, pkg):
    """A plan with the low-pass front end must fail loudly under the emulator (no silent wrong numbers)."""
    from egregora_b200 import flashsr_model as M, flashsr_plan as P
    spec = M.tiny_spec()
    blob = P.WeightBlob()
    be = P.build_plan(spec, M.init_weights(spec, 0), blob, 1, 1, True)
    ws, wt = _dev(be.ws_bytes + 4096), _dev(np.frombuffer(blob.tobytes(), np.uint8))
    h = C.c_void_p()
    sim.ck(sim.egr_plan_create(be.build_ops(), len(be.ops), ws.ctypes.data, ws.size, wt.ctypes.data, wt.size, C.byref(h)))
    try:
        assert sim.egr_plan_run(h, 0, -1, None) != 0
        assert b"cluster" in sim.egr_last_error()
    finally:
        sim.egr_plan_destroy(h)


def test_metrics_kernels_do_not_depend_on_a_zeroed_workspace(sim, monkeypatch):
    """The host mirrors hand the kernels torch.empty workspaces.  Re-run the metrics tests with every scratch buffer
    poisoned (0xFF bytes = NaN): a kernel that reads workspace it has not written shows up as a NaN or a wrong number."""
    import test_cusim as me
    monkeypatch.setattr(me, "POISON", True)
    assert _dev(16)[0] == 0xFF
    test_eval_lsd_reference_golden(sim)       # the three entry points that have not met a real torch.empty yet
    test_eval_lufs_reference_golden(sim)
    test_eval_hf_band_reference_golden(sim)


def test_cluster_kernels_under_the_cluster_variant():
    """The CUSIM_CLUSTERS build (every CTA of a cluster an OS thread, `__shared__` thread_local, cluster.sync a pthread
    barrier, map_shared_rank the distance between two threads' TLS blocks) runs the two cluster kernels of the hot path:
    GroupNorm on clusters of 2 / 4 / 8 CTAs exchanging moments through distributed shared memory, and the 8-CTA
    zero-phase low-pass with its three-level scan — the GPU op tests for them, through the emulator."""
    import os
    import subprocess
    env = dict(os.environ, EGR_TEST_CUSIM="clusters")
    r = subprocess.run([sys.executable, "-m", "pytest", str(ROOT / "tests" / "test_ops_gpu.py"), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider",
                        "-k", "cluster_sizes or lowpass"], env=env, capture_output=True, text=True, cwd=str(ROOT))
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_new_kernels_do_not_depend_on_thread_order():
    """The scheduler normally resumes a block's threads in ascending order; CUSIM_ORDER=reverse resumes them in
    descending order.  Kernels that miss a barrier tend to change their results between the two — the metrics kernels
    written without GPU time must pass their (partly bit-exact) tests both ways."""
    import os
    import subprocess
    env = dict(os.environ, CUSIM_ORDER="reverse")
    r = subprocess.run([sys.executable, "-m", "pytest", str(Path(__file__)), "-q", "-x", "-p", "no:cacheprovider", "-k",
                        "eval_lsd or eval_lufs or eval_hf_band"], env=env, capture_output=True, text=True, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
