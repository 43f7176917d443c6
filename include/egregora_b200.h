/*
 * egregora_b200.h — C ABI of libegregora_b200.so (sm_100a).
 *
 * The reference pack (lucasgattas/ComfyUI-Egregora-Audio-Super-Resolution) is pure Python and has
 * no FFI of its own; its arithmetic sits behind two Python call sites:
 *
 *   path A  _FlashSRRunner.infer -> FlashSR.forward      egregora_audio_super_resolution.py:361-369
 *           _wola_stitch / chunk slice+pad                egregora_audio_super_resolution.py:227-251,411-418
 *   path B  feed.upscale(...)                              egregora_fat_llama_gpu.py:213-224
 *                                                          egregora_fat_llama_cpu.py:126-134
 *
 * This header declares what a ctypes binding at those call sites needs.
 * Conventions: every entry returns 0 on success and a negative code on failure, with a
 * human-readable message available from egr_last_error() (the Python host turns it into
 * RuntimeError, the reference's only error type).  All buffers are caller-owned DEVICE pointers
 * unless a parameter is named h_*; all work is stream-ordered on the cudaStream_t passed as
 * `void* stream`; no entry synchronises unless documented.  No torch types cross this boundary.
 */
#ifndef EGREGORA_B200_H
#define EGREGORA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EGR_ABI_VERSION 1

/* ------------------------------------------------------------------------------------------ */
/* lifecycle                                                                                    */
/* ------------------------------------------------------------------------------------------ */
int         egr_abi_version(void);
const char* egr_last_error(void);
/* Select `device`, resolve the driver entry points the TMA paths need, set kernel attributes.
 * Fails (no CPU fallback) when no sm_100 device is present. */
int         egr_init(int device);
int         egr_sm_count(void);
/* kernels this library has launched in this process so far (bench.py reports the delta over its timed region) */
int64_t     egr_launch_count(void);
/* size of the C structs below, for the ctypes mirror's self-check: 0=egr_tensor, 1=egr_op */
int         egr_sizeof(int which);

/* ------------------------------------------------------------------------------------------ */
/* path A driver: chunk gather + Hann WOLA stitch                                               */
/* ------------------------------------------------------------------------------------------ */
/* Replaces the slice + right-zero-pad of egregora_audio_super_resolution.py:411-416 for all spans
 * at once.  d_in [C,total] f32 -> d_chunks [n_spans, C, win] f32.  d_starts/d_lens: device int64/int32. */
int egr_chunk_gather(const float* d_in, int C, int64_t total, const int64_t* d_starts,
                     const int32_t* d_lens, int n_spans, int win, float* d_chunks, void* stream);

/* Replaces _wola_stitch (egregora_audio_super_resolution.py:227-251): out[c,t] =
 * sum_k y_k[c,t-s_k]*w[t-s_k] / sum_k w[t-s_k] over spans k with s_k <= t < s_k+min(L_k,L_pred),
 * wsum==0 -> 1; products and sums are rounded separately in span order (bit-exact with numpy).
 * d_chunks [n_spans, C, l_pred] f32, spans sorted by start, d_window [win] f32 (np.hanning(win)). */
int egr_wola_stitch(const float* d_chunks, int l_pred, const int64_t* d_starts, const int32_t* d_lens,
                    int n_spans, int C, int64_t total, int win, const float* d_window,
                    float* d_out, void* stream);

/* Replaces the scipy branch of _resample_hq (egregora_audio_super_resolution.py:181-191):
 * scipy.signal.resample_poly(x, up, down) on float32 data, default Kaiser(5.0) FIR, zero padding.
 * d_x [C,n_in] f32 -> d_y [C,n_out] f32.  d_hflip [up,hpp] f32 is the transposed, flipped polyphase bank
 * of the padded filter (scipy's _pad_h), y_first the number of leading upfirdn outputs that resample_poly
 * trims (n_pre_remove) and n_out = ceil(n_in*up/down); the node module's _resample_design computes all
 * three.  Every output is accumulated in float32 in scipy's tap order with separately rounded multiply
 * and add: bit-identical to the reference's result. */
int egr_resample_poly(const float* d_x, int C, int64_t n_in, int up, int down, const float* d_hflip,
                      int hpp, int64_t y_first, int64_t n_out, float* d_y, void* stream);

/* Diffusion start noise x_T of the FlashSR sampler (upstream draws it inside forward(), call site
 * egregora_audio_super_resolution.py:366-369).  Counter-based Philox4x32-10 + Box-Muller: element e
 * of chunk-channel row (row0 + r) depends on (seed, row0 + r, e) only, so a rank that owns rows [lo, hi) of a clip
 * produces exactly the numbers the single-GPU run uses for those rows.  d_out [n_rows, row_elems] f32, standard normal. */
int egr_noise_fill(uint64_t seed, int64_t row0, int64_t n_rows, int64_t row_elems, float* d_out, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* path A model: a plan is a straight-line list of ops over one workspace + one weight blob     */
/* ------------------------------------------------------------------------------------------ */
/* Addresses inside ops: top 4 bits = space, low 60 bits = byte offset (or absolute pointer). */
#define EGR_SPACE_NULL 0ull
#define EGR_SPACE_WS   1ull   /* byte offset into the plan workspace   */
#define EGR_SPACE_WT   2ull   /* byte offset into the plan weight blob  */
#define EGR_SPACE_ABS  3ull   /* absolute device pointer                */
#define EGR_ADDR(space, off) (((uint64_t)(space) << 60) | ((uint64_t)(off) & 0x0FFFFFFFFFFFFFFFull))

#define EGR_MAX_TAPS 16

/* strided 5-D view, dim[0] innermost (channels), strides in ELEMENTS, stride[0] must be 1 */
typedef struct egr_tensor {
  uint64_t addr;
  int32_t  rank;
  int32_t  elem;        /* 0 = f32, 1 = f16 */
  int64_t  dim[5];
  int64_t  stride[5];
} egr_tensor;

/* egr_op.flags */
#define EGR_FLAG_MEGA   1   /* member of a run of ops the library may execute as ONE persistent kernel (the UNet) */
#define EGR_FLAG_NOSYNC 2   /* inside such a run: the next op does not depend on this one (no grid barrier after it) */

typedef struct egr_op {
  int32_t    code;      /* EGR_OP_* */
  int32_t    flags;     /* EGR_FLAG_* */
  egr_tensor x0;        /* primary input view   */
  egr_tensor x1;        /* secondary input view */
  uint64_t   ptr[10];   /* EGR_P_* */
  int64_t    i[40];     /* EGR_I_* */
  double     f[8];      /* EGR_F_* */
  int16_t    tap[EGR_MAX_TAPS][5];
  char       name[48];  /* debug label */
} egr_op;

/* op codes */
#define EGR_OP_GEMM_TC      1   /* tcgen05 tap-GEMM: out[pix,n] = post*(act(alpha*sum_t sum_k A[pix+tap_t,k]*B[t,n,k] + bias + rowbias) + resid + resid2) */
#define EGR_OP_GEMM_SIMT    2   /* same contract, fp32 CUDA-core path (tiny K/N layers, and the in-library cross-check) */
#define EGR_OP_GN_STATS     3   /* GroupNorm partial sums (f64) over a virtual channel-concat of x0|x1 */
#define EGR_OP_GN_APPLY     4   /* GroupNorm normalise (+SiLU) -> f16 and/or f32 */
#define EGR_OP_LAYERNORM    5   /* per-token LayerNorm -> f16 */
#define EGR_OP_SOFTMAX      6   /* row softmax(scale*x) f32 -> f16 */
#define EGR_OP_ATTN_SMALL   7   /* fused multi-head attention for short sequences, head_dim<=64 */
#define EGR_OP_GEGLU        8   /* a*gelu(g) -> f16 */
#define EGR_OP_ELTWISE      9   /* EGR_ELT_* in i[EGR_I_MODE] */
#define EGR_OP_SNAKE_AA    10   /* anti-aliased SnakeBeta: up2x FIR -> snake -> down2x FIR */
#define EGR_OP_STFT_MEL    11   /* reflect-pad, frame, Hann, 2048 real FFT, |.|, mel, log */
#define EGR_OP_LOWPASS     12   /* cutoff detect + zero-phase SOS filter (f64 recurrences) */
#define EGR_OP_TIME_EMBED  13   /* sinusoidal timestep embedding */
#define EGR_OP_ZERO        14   /* memset a buffer (stats accumulators) */

/* ptr[] slots */
#define EGR_P_W        0   /* weights / B operand: f16 [Z][N][K] for GEMM_TC, f32 same shape for GEMM_SIMT */
#define EGR_P_BIAS     1   /* f32 [N] */
#define EGR_P_ROWBIAS  2   /* f32 [B][N] added per batch item (time embedding) */
#define EGR_P_RESID    3   /* f32, addressed like OUT32 */
#define EGR_P_OUT32    4
#define EGR_P_OUT16    5
#define EGR_P_STATS    6   /* f64 [B][G][2] sums (GN) */
#define EGR_P_GAMMA    7
#define EGR_P_BETA     8
#define EGR_P_AUX      9
#define EGR_P_RESID2   9   /* GEMM ops: second residual, addressed like RESID (the slot is AUX for the other ops) */

/* i[] slots (GEMM family) */
#define EGR_I_DIMW      0  /* which x0 dim the tile's w extent walks */
#define EGR_I_DIMH      1
#define EGR_I_DIMB      2
#define EGR_I_BW        3  /* tile extents, BW*BH*BB == 128 */
#define EGR_I_BH        4
#define EGR_I_BB        5
#define EGR_I_WO        6  /* output logical extents */
#define EGR_I_HO        7
#define EGR_I_BO        8
#define EGR_I_NTAPS     9
#define EGR_I_K        10  /* reduction length per tap */
#define EGR_I_N        11
#define EGR_I_BLOCKN   12
#define EGR_I_WSTRIDE_N 13 /* B operand strides, elements */
#define EGR_I_WSTRIDE_Z 14
#define EGR_I_WZ_BATCH 15  /* 0: z = tap index, 1: z = batch index */
#define EGR_I_ROWBIAS_STRIDE 16
#define EGR_I_OUT_PIX_STRIDE 17
#define EGR_I_OUT_BATCH_STRIDE 18
#define EGR_I_OUT_OFFSET 19
#define EGR_I_OUT_LO   20  /* crop: keep flat in-batch index in [LO,HI) */
#define EGR_I_OUT_HI   21
#define EGR_I_TRANSPOSED 22
#define EGR_I_OUT_N_STRIDE 23
#define EGR_I_ACT      24  /* EGR_ACT_* */
#define EGR_I_KBLOCK   25  /* 64 (default) */
/* i[] slots (other ops) */
#define EGR_I_MODE     26
#define EGR_I_GROUPS   27
#define EGR_I_C0       28  /* channels taken from x0 */
#define EGR_I_C1       29  /* channels taken from x1 (virtual concat) */
#define EGR_I_ROWS     30
#define EGR_I_COLS     31
#define EGR_I_HEADS    32
#define EGR_I_HEADDIM  33
#define EGR_I_SEQ      34
#define EGR_I_BATCH    35
#define EGR_I_AUX0     36
#define EGR_I_OUT_H_STRIDE 36 /* GEMM_TC ops: output stride of one step along H (0 = WO*OUT_PIX_STRIDE, the dense default) */
#define EGR_I_AUX1     37
#define EGR_I_AUX2     38
#define EGR_I_AUX3     39

/* f[] slots */
#define EGR_F_ALPHA    0
#define EGR_F_EPS      1
#define EGR_F_A        2
#define EGR_F_POST     2   /* GEMM ops: out = POST * (act(...) + resid + resid2); 0 means 1 */
#define EGR_F_B        3
#define EGR_F_C        4
#define EGR_F_D        5

#define EGR_ACT_NONE 0
#define EGR_ACT_SILU 1
#define EGR_ACT_TANH 2

/* eltwise modes */
#define EGR_ELT_CAST16      1  /* out16 = f16(concat(x0,x1)) */
#define EGR_ELT_AXPBY       2  /* out32 = A*x0 + B*x1 (+ optional out16) */
#define EGR_ELT_UPSAMPLE2X  3  /* nearest 2x in H and W of x0 [C,W,H,B] -> out16 (and/or out32) */
#define EGR_ELT_COPY32      4  /* out32 = concat(x0,x1) as f32 */
#define EGR_ELT_SCALE_SHIFT 5  /* out32 = A*x0 + B */
#define EGR_ELT_SUM3        6  /* out = A*((x0 + x1) + x2), x2 in ptr[EGR_P_AUX]; f32 and/or f16 out (mean of the vocoder's parallel blocks) */

typedef struct egr_plan egr_plan;

/* Build a plan: validates ops, resolves addresses, encodes one TMA tensor map pair per GEMM_TC op.
 * h_ops is a HOST array.  d_workspace/d_weights are caller-owned device buffers that must outlive
 * the plan. */
int  egr_plan_create(const egr_op* h_ops, int n_ops, void* d_workspace, size_t ws_bytes,
                     const void* d_weights, size_t wt_bytes, egr_plan** out);
/* Enqueue ops [first,last) on `stream`. */
int  egr_plan_run(egr_plan* plan, int first, int last, void* stream);
int  egr_plan_num_launches(const egr_plan* plan, int first, int last);
/* Measurement helpers (bench.py roofline leg): enqueue only the ops whose code == `code` (same shapes and
 * addresses as in a real run; inputs are whatever the workspace holds), and count them.  Ops that the library executes
 * inside a persistent-kernel run (EGR_FLAG_MEGA) are not launches of their own and are skipped by both. */
int  egr_plan_run_code(egr_plan* plan, int code, void* stream);
int  egr_plan_count_code(const egr_plan* plan, int code);
void egr_plan_destroy(egr_plan* plan);

/* ------------------------------------------------------------------------------------------ */
/* path B: Fat-Llama iterative spectral loop (replaces feed.upscale's arithmetic)               */
/* ------------------------------------------------------------------------------------------ */
typedef struct egr_fft_plan egr_fft_plan;
/* complex-to-complex FFT of length n (any n >= 1): mixed radix 2/3/4/5/7/8 multi-pass Stockham with
 * shared-memory sub-transforms; Bluestein for lengths with a prime factor > 7. */
int  egr_fft_plan_create(int64_t n, int batch, egr_fft_plan** out);
size_t egr_fft_plan_workspace_bytes(const egr_fft_plan* p);
int  egr_fft_plan_passes(const egr_fft_plan* p);
/* d_data: interleaved complex64 [batch][n], transformed in place (d_work: workspace_bytes).
 * inverse=1 is unnormalised unless scale_inverse=1 (divide by n, like numpy/cupy ifft). */
int  egr_fft_exec(egr_fft_plan* p, float* d_data, float* d_work, int inverse, int scale_inverse, void* stream);
void egr_fft_plan_destroy(egr_fft_plan* p);

#define EGR_FL_NORMALIZE 1u
#define EGR_FL_AUTOSCALE 2u
/* Whole per-clip loop, channel-parallel: for each channel (d_in [C][n] f32, integer-scaled samples as
 * upstream read_audio yields them): expanded = repeat(x, upscale); ist = IST(expanded, iters, thr);
 * y = expanded + ist; optional autoscale to the channel's input peak; optional global peak normalise.
 * d_out [C][n*upscale] f32.  d_work: egr_fatllama_workspace_bytes(). */
size_t egr_fatllama_workspace_bytes(int C, int64_t n, int upscale);
int  egr_fatllama_run(const float* d_in, float* d_out, int C, int64_t n, int upscale, int iters,
                      float threshold, uint32_t flags, void* d_work, size_t work_bytes, void* stream);

/* PCM-16 wire format emulation (the reference round-trips through 16-bit WAV files:
 * egregora_fat_llama_gpu.py:34-37, :291).  f32 [-1,1) -> int16 as libsndfile does for float input
 * with clipping on, and int16 -> f32 / 32768. */
int  egr_pcm16_quantize(const float* d_in, int16_t* d_out, int64_t n, void* stream);
int  egr_pcm16_to_float(const int16_t* d_in, float* d_out, int64_t n, float scale, void* stream);
/* max |x| over n f32 values -> d_out[0] (f32). */
int  egr_absmax(const float* d_in, int64_t n, float* d_out, void* stream);
/* d_x[i] *= scale for all i iff d_ref[0] > threshold, decided on the device: the patched write_audio of the reference
 * (egregora_fat_llama_gpu.py:195-200) divides integer-scaled data by 2^(8*sample_width-1) when its peak exceeds 1. */
int  egr_scale_if_above(float* d_x, int64_t n, const float* d_ref, float threshold, float scale, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* next row: adaptive wet/dry mix of Egregora_DeepFilterNet_Denoise        */
/* ------------------------------------------------------------------------------------------ */
#define EGR_VAD_NONE 0            /* constant strength */
#define EGR_VAD_RMS  1            /* adaptive_vad_source="rms"  (egregora_audio_enhance_extras.py:548-559) */
#define EGR_MIX_OFF            0  /* adaptive_mode, :575-594 */
#define EGR_MIX_MORE_ON_NOISE  1
#define EGR_MIX_MORE_ON_SPEECH 2
#define EGR_MIX_GATE_ON_NOISE  3
#define EGR_CURVE_EQUAL_POWER 0   /* mix_curve, :596-605 */
#define EGR_CURVE_LINEAR      1
/* Replaces steps 5-6 of Egregora_DeepFilterNet_Denoise.execute (egregora_audio_enhance_extras.py:657-704) and the
 * helpers they call (:548-605): 10 ms frame RMS -> / 95th percentile -> smoothing -> per-frame strength -> dry/wet
 * gains -> y = clip(g_dry*dry + g_wet*wet) -> post gain -> peak limiter -> clamp.  d_dry / d_wet / d_out: [C,T] f32 at
 * 48 kHz (d_wet = the DeepFilterNet output, produced elsewhere).  Arguments carry the node's own parameter values;
 * all float32 arithmetic follows numpy's operation order (frame means by pairwise summation, exact order statistics
 * for the percentile), so the per-frame gains equal the reference's up to the last bit of sinf / cosf. */
size_t egr_dfn_mix_workspace_bytes(int C, int64_t T);
int egr_dfn_mix(const float* d_dry, const float* d_wet, float* d_out, int C, int64_t T, int sample_rate,
                double strength, int mix_curve, int vad_source, int adaptive_mode, double adaptive_amount,
                double vad_threshold, int vad_smooth_ms, double post_gain_db, int limit_ceiling, double ceiling,
                void* d_work, size_t work_bytes, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* next row: clip-scale evaluation reductions                              */
/* ------------------------------------------------------------------------------------------ */
#define EGR_EVAL_SI_SDR_DB      0  /* _si_sdr, egregora_audio_eval_pack.py:414-429 (float64, mono means)       */
#define EGR_EVAL_CORR           1  /* corr_coef, egregora_null_test_suite.py:444-448                           */
#define EGR_EVAL_NULL_RMS_DBFS  2  /* _rms_db(null.mean(axis=0)), :119-122, :449-450                           */
#define EGR_EVAL_OVERSHOOT      3  /* count of |null| > 1, :463                                                */
#define EGR_EVAL_CLIPPED_PCT    4  /* :465                                                                     */
#define EGR_EVAL_SCALE_K        5  /* least-squares scale k, :431-436                                          */
#define EGR_EVAL_NUM            8
/* Replaces the arithmetic of Audio_Null_Test.execute (egregora_null_test_suite.py:421-467, the LSD, LUFS and HF-band options are
 * egr_eval_lsd / egr_eval_lufs / egr_eval_hf_band below) and _si_sdr: d_ref / d_proc [C,N] f32 with row strides ld_* (so the "trim both to the shorter"
 * step needs no copy), d_null [C,N] f32 or NULL, d_metrics [EGR_EVAL_NUM] f64 ON THE DEVICE (read it after a stream
 * sync).  The null signal is bit-identical to numpy's; the reductions are deterministic float64 sums. */
size_t egr_eval_workspace_bytes(void);
int egr_eval_null_test(const float* d_ref, int64_t ld_ref, const float* d_proc, int64_t ld_proc, int C, int64_t N,
                       int invert_b, int least_squares_scale, float* d_null, double* d_metrics, void* d_work,
                       size_t work_bytes, void* stream);


#define EGR_LSD_MEAN_DB  0  /* lsd_mean_db = mean over frames,           egregora_audio_eval_pack.py:405-411 */
#define EGR_LSD_P95_DB   1  /* lsd_p95_db  = np.percentile(per, 95),     same                                */
#define EGR_LSD_FRAMES   2  /* 1 + max(0, (N - n_fft) // hop),           :393                                */
#define EGR_LSD_NUM      4
/* Replaces _stft_mag + _lsd (egregora_audio_eval_pack.py:389-411, identical copies in egregora_null_test_suite.py
 * :167-189; callers Metrics_LSD_SISDR.execute :462-467 and Audio_Null_Test.execute :456-461): mono means of d_ref /
 * d_proc [C,N] f32, frames of n_fft samples every hop samples (no centring, a clip shorter than n_fft is one zero-padded
 * frame), symmetric Hann (np.hanning), float32 real FFT, L = 20*log10(|X| + 1e-12), per-frame sqrt(mean_k dL^2 + 1e-12),
 * then the mean and the 95th percentile over frames into d_metrics [EGR_LSD_NUM] f64 ON THE DEVICE.  n_fft must be a
 * power of two in [64, 8192] (the nodes' default is 2048); other sizes return an error, nothing falls back.  Two float32
 * FFTs agree to rounding, so bins that hold real signal agree to ~1e-5 dB; bins that hold only rounding noise are noise
 * in the reference too (tests state the tolerance per case).  proc_gain: 1, or the null test's least-squares scale k as
 * float32 — every proc sample is multiplied by it (rounded to float32) before the channel mean, as (B * k).astype(f32). */
size_t egr_eval_lsd_workspace_bytes(int64_t N, int n_fft, int hop);
int egr_eval_lsd(const float* d_ref, int64_t ld_ref, const float* d_proc, int64_t ld_proc, int C, int64_t N, int n_fft,
                 int hop, float proc_gain, double* d_metrics, void* d_work, size_t work_bytes, void* stream);

#define EGR_LUFS_INTEGRATED 0  /* integrated_lufs(), egregora_null_test_suite.py:143-164                          */
#define EGR_LUFS_UNGATED    1  /* lufs_ungated of the same function (:158)                                         */
#define EGR_LUFS_BLOCKS     2  /* number of 400 ms blocks (:148-150)                                                */
#define EGR_LUFS_REPAIRED   3  /* chunks whose speculated filter state had to be recomputed (diagnostic)            */
#define EGR_LUFS_NUM        4
/* Replaces integrated_lufs + _k_weight (egregora_null_test_suite.py:125-164, identical in egregora_audio_eval_pack.py
 * :132-167; callers: Audio_Null_Test.execute :455 "null_lufs", the loudness meter :327 and loudness match :373-374):
 * d_x [C,N] f32 with row stride ld.  The reference's per-sample float32 high-pass recurrence is reproduced BIT-EXACTLY
 * (speculated per 2048-sample chunk, verified and repaired where needed), as are its float32 tilt and channel mean; the
 * block mean squares and the gate are float64 (summation order differs from numpy's pairwise sum: ~1e-15 relative).
 * d_metrics [EGR_LUFS_NUM] f64 ON THE DEVICE. */
size_t egr_eval_lufs_workspace_bytes(int C, int64_t N, int sample_rate);
int egr_eval_lufs(const float* d_x, int64_t ld, int C, int64_t N, int sample_rate, double* d_metrics, void* d_work,
                  size_t work_bytes, void* stream);

#define EGR_HF_RESIDUAL_DB 0  /* 10*log10(e_hi / (e_all + 1e-20) + 1e-20), egregora_null_test_suite.py:190-197       */
#define EGR_HF_E_HI        1  /* sum |X_k|^2 over bins with rfftfreq(k) >= lo_hz                                      */
#define EGR_HF_E_ALL       2  /* sum |X_k|^2 over all N/2 + 1 bins                                                    */
#define EGR_HF_BINS_HI     3
#define EGR_HF_NUM         4
/* Replaces _band_energy_hi_db (egregora_null_test_suite.py:190-197; "hf_residual_db" of Audio_Null_Test.execute :462):
 * channel mean of d_x [C,N] f32 -> length-N FFT with the path-B transform (`plan` must come from
 * egr_fft_plan_create(N, 1)) -> energy above lo_hz over total energy, float64 sums.  d_metrics [EGR_HF_NUM] f64 ON THE
 * DEVICE. */
size_t egr_eval_hf_band_workspace_bytes(const egr_fft_plan* plan, int64_t N);
int egr_eval_hf_band(egr_fft_plan* plan, const float* d_x, int64_t ld, int C, int64_t N, int sample_rate, double lo_hz,
                     double* d_metrics, void* d_work, size_t work_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EGREGORA_B200_H */
