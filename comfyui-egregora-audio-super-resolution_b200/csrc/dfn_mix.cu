// dfn_mix.cu — adaptive wet/dry mix behind the reference's Egregora_DeepFilterNet_Denoise node (SURVEY.md §8(f)
// rank 1: the step that follows the third-party DeepFilterNet model in config c5).  Semantics follow
// /root/reference egregora_audio_enhance_extras.py: _vad_probs_rms_48k (:548-559), _smooth_probs (:561-573),
// _strength_per_frame (:575-594), _gains_from_strength (:596-605) and steps 5-6 of execute (:657-704), at 48 kHz
// (the VAD branch resamples through a third-party resampler otherwise).  `wet` (the model output) is an input.
//
// Four launches, every float32 operation in numpy's order (oracle/dfn_mix_oracle.py) so the per-frame gains are
// bit-identical to the reference's up to sinf/cosf:
//   dfn_frame_rms_kernel   rms of every 480-sample frame; np.mean's pairwise float32 summation reproduced exactly
//   dfn_frame_gain_kernel  one CTA per channel: 95th percentile by radix select (exact order statistics + numpy's
//                          float32 lerp) -> probs -> smoothing recurrence (serial, float32) -> strength -> gains
//   dfn_mix_kernel         y = clip(g_dry*dry + g_wet*wet) * post_gain, global peak (atomic max on the bits)
//   dfn_limit_kernel       y = clamp(y * ceiling/peak)   (only rescales when the peak exceeds the ceiling)
// HBM roofline: rms reads 4T, mix reads 8T + writes 4T, limit reads + writes 8T bytes per channel = 24*T.
#include <cmath>
#include "common.cuh"
#include "select.cuh"  // radix_select_nonneg / percentile95_f32_nonneg (shared with eval_metrics.cu)

using namespace egr;

#define DFN_HOP 480

// numpy pairwise float32 sum of squares over a[0..n): blocks of <= 128 use 8 interleaved accumulators combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), larger ranges split at n/2 rounded down to a multiple of 8.
__device__ float dfn_pairwise_sq(const float* __restrict__ a, int n) {
  if (n < 8) {
    float r = 0.f;
    for (int i = 0; i < n; ++i) r = __fadd_rn(r, __fmul_rn(a[i], a[i]));
    return r;
  }
  if (n <= 128) {
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __fmul_rn(a[j], a[j]);
    int i = 8;
    for (; i < n - (n & 7); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], __fmul_rn(a[i + j], a[i + j]));
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, __fmul_rn(a[i], a[i]));
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 & 7;
  return __fadd_rn(dfn_pairwise_sq(a, n2), dfn_pairwise_sq(a + n2, n - n2));
}

// one 120-sample leaf of a full frame, 16-byte loads (a is 16-byte aligned: frames start at multiples of 1920 bytes)
__device__ __forceinline__ float dfn_leaf120(const float4* __restrict__ a) {
  float r[8];
  {
    const float4 u = __ldg(a), v = __ldg(a + 1);
    r[0] = __fmul_rn(u.x, u.x); r[1] = __fmul_rn(u.y, u.y); r[2] = __fmul_rn(u.z, u.z); r[3] = __fmul_rn(u.w, u.w);
    r[4] = __fmul_rn(v.x, v.x); r[5] = __fmul_rn(v.y, v.y); r[6] = __fmul_rn(v.z, v.z); r[7] = __fmul_rn(v.w, v.w);
  }
#pragma unroll
  for (int i = 1; i < 15; ++i) {
    const float4 u = __ldg(a + 2 * i), v = __ldg(a + 2 * i + 1);
    r[0] = __fadd_rn(r[0], __fmul_rn(u.x, u.x)); r[1] = __fadd_rn(r[1], __fmul_rn(u.y, u.y));
    r[2] = __fadd_rn(r[2], __fmul_rn(u.z, u.z)); r[3] = __fadd_rn(r[3], __fmul_rn(u.w, u.w));
    r[4] = __fadd_rn(r[4], __fmul_rn(v.x, v.x)); r[5] = __fadd_rn(r[5], __fmul_rn(v.y, v.y));
    r[6] = __fadd_rn(r[6], __fmul_rn(v.z, v.z)); r[7] = __fadd_rn(r[7], __fmul_rn(v.w, v.w));
  }
  return __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                   __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
}

__global__ void __launch_bounds__(128) dfn_frame_rms_kernel(const float* __restrict__ x, long long T, int nfr,
                                                             float* __restrict__ rms, int aligned) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x, ch = blockIdx.y;
  if (f >= nfr) return;
  const float* a = x + (long long)ch * T + (long long)f * DFN_HOP;
  const long long left = T - (long long)f * DFN_HOP;
  const int n = left < DFN_HOP ? (int)left : DFN_HOP;
  float ssq;
  if (n == DFN_HOP && aligned) {
    const float4* a4 = reinterpret_cast<const float4*>(a);  // 480 = 240 + 240 = (120 + 120) + (120 + 120)
    ssq = __fadd_rn(__fadd_rn(dfn_leaf120(a4), dfn_leaf120(a4 + 30)), __fadd_rn(dfn_leaf120(a4 + 60), dfn_leaf120(a4 + 90)));
  } else {
    ssq = dfn_pairwise_sq(a, n);
  }
  rms[(long long)ch * nfr + f] = __fsqrt_rn(__fdiv_rn(ssq, (float)n));
}

struct DfnGainArgs {
  int nfr, vad, mode, curve, smooth;
  float s0, amt, oms0, s_noise, s_speech, thr;  // float32 images of the python floats (weak scalars)
  float alpha, oma;
};

__global__ void __launch_bounds__(1024) dfn_frame_gain_kernel(DfnGainArgs g, float* __restrict__ rms /* in: rms, scratch */,
                                                               float* __restrict__ g_dry, float* __restrict__ g_wet) {
  __shared__ unsigned hist[4096];
  __shared__ unsigned bc[2];
  __shared__ unsigned wsum[32];
  __shared__ float chunk[2048];
  __shared__ float s_acc;
  const int ch = blockIdx.x, n = g.nfr;
  float* p = rms + (long long)ch * n;
  float* gd = g_dry + (long long)ch * n;
  float* gw = g_wet + (long long)ch * n;
  if (g.vad) {
    // np.percentile(rms, 95) on float32 data (exact order statistics + numpy's float32 lerp, select.cuh)
    float p95 = percentile95_f32_nonneg(p, n, hist, bc, wsum);
    if (p95 == 0.f) p95 = (float)1e-6;
    for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = fminf(fmaxf(__fdiv_rn(p[i], p95), 0.f), 1.f);
    __syncthreads();
  }
  // smoothing recurrence (float32, serial) over shared-memory chunks, then strength and gains in parallel
  for (int base = 0; base < n; base += 2048) {
    const int m = min(2048, n - base);
    const bool smooth = g.vad && g.smooth;
    // the (1-alpha)*p[i] products do not depend on the recurrence: all threads form them, one thread chains the rest
    for (int i = threadIdx.x; i < m; i += blockDim.x) chunk[i] = smooth ? __fmul_rn(g.oma, p[base + i]) : p[base + i];
    if (smooth && base == 0 && threadIdx.x == 0) s_acc = p[0];
    __syncthreads();
    if (smooth && threadIdx.x == 0) {
      float acc = s_acc;
      for (int i = 0; i < m; ++i) {
        acc = __fadd_rn(__fmul_rn(g.alpha, acc), chunk[i]);
        chunk[i] = acc;
      }
      s_acc = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      float s = g.s0;
      if (g.vad) {
        const float v = fminf(fmaxf(chunk[i], 0.f), 1.f);
        if (g.mode == EGR_MIX_MORE_ON_NOISE) s = __fadd_rn(g.s0, __fmul_rn(__fmul_rn(g.amt, __fsub_rn(1.0f, v)), g.oms0));
        else if (g.mode == EGR_MIX_MORE_ON_SPEECH) s = __fadd_rn(g.s0, __fmul_rn(__fmul_rn(g.amt, v), g.oms0));
        else if (g.mode == EGR_MIX_GATE_ON_NOISE) s = v < g.thr ? g.s_noise : g.s_speech;
      }
      s = fminf(fmaxf(s, 0.f), 1.f);
      if (g.curve == EGR_CURVE_EQUAL_POWER) {
        const float ang = __fmul_rn(1.57079637050628662109375f /* float32(pi/2) */, s);
        gd[base + i] = cosf(ang);
        gw[base + i] = sinf(ang);
      } else {
        gd[base + i] = __fsub_rn(1.0f, s);
        gw[base + i] = s;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) dfn_mix_kernel(const float* __restrict__ dry, const float* __restrict__ wet,
                                                       long long T, int nfr, const float* __restrict__ g_dry,
                                                       const float* __restrict__ g_wet, float gain, int apply_gain,
                                                       float* __restrict__ out, unsigned* __restrict__ peak) {
  const int ch = blockIdx.y;
  const float* d = dry + (long long)ch * T;
  const float* w = wet + (long long)ch * T;
  float* o = out + (long long)ch * T;
  const float* gd = g_dry + (long long)ch * nfr;
  const float* gw = g_wet + (long long)ch * nfr;
  float pk = 0.f;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(t / DFN_HOP);
    float y = __fadd_rn(__fmul_rn(__ldg(gd + f), __ldg(d + t)), __fmul_rn(__ldg(gw + f), __ldg(w + t)));
    y = fminf(fmaxf(y, -1.f), 1.f);
    if (apply_gain) y = __fmul_rn(y, gain);
    o[t] = y;
    pk = fmaxf(pk, fabsf(y));
  }
  pk = warp_max(pk);
  if ((threadIdx.x & 31) == 0 && pk > 0.f) atomicMax(peak, __float_as_uint(pk));
}

__global__ void __launch_bounds__(256) dfn_limit_kernel(float* __restrict__ y, long long total, const unsigned* __restrict__ peak,
                                                         int limit, double ceiling) {
  const float pk = __uint_as_float(*peak);
  const bool rescale = limit && (double)pk > ceiling && pk > 0.f;
  const float sc = rescale ? (float)(ceiling / (double)pk) : 1.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float v = y[i];
    if (rescale) v = __fmul_rn(v, sc);
    y[i] = fminf(fmaxf(v, -1.f), 1.f);
  }
}

static long long dfn_frames(long long T) { return (T + DFN_HOP - 1) / DFN_HOP; }

extern "C" size_t egr_dfn_mix_workspace_bytes(int C, int64_t T) {
  if (C < 1 || T < 1) return 256;
  return 256 + 3 * sizeof(float) * (size_t)C * (size_t)dfn_frames(T);
}

extern "C" int egr_dfn_mix(const float* d_dry, const float* d_wet, float* d_out, int C, int64_t T, int sample_rate,
                           double strength, int mix_curve, int vad_source, int adaptive_mode, double adaptive_amount,
                           double vad_threshold, int vad_smooth_ms, double post_gain_db, int limit_ceiling, double ceiling,
                           void* d_work, size_t work_bytes, void* stream) {
  if (!devinfo().inited) return fail(EGR_ERR_STATE, "egr_dfn_mix: call egr_init first");
  if (C < 1 || C > 65535 || T < 0) return fail(EGR_ERR_ARG, "egr_dfn_mix: bad arguments");
  if (T == 0) return EGR_OK;
  if (!d_dry || !d_wet || !d_out || !d_work) return fail(EGR_ERR_ARG, "egr_dfn_mix: null pointer");
  if (sample_rate != 48000)
    return fail(EGR_ERR_UNSUPPORTED, "egr_dfn_mix: only 48 kHz (the reference resamples its VAD branch through df.io otherwise)");
  if (vad_source != EGR_VAD_NONE && vad_source != EGR_VAD_RMS)
    return fail(EGR_ERR_UNSUPPORTED, "egr_dfn_mix: VAD source must be none or rms (rnnoise is a third-party model)");
  if (work_bytes < egr_dfn_mix_workspace_bytes(C, T)) return fail(EGR_ERR_ARG, "egr_dfn_mix: workspace too small");
  if (reinterpret_cast<uintptr_t>(d_work) % 256) return fail(EGR_ERR_ARG, "egr_dfn_mix: workspace must be 256-byte aligned");
  const long long nfr = dfn_frames(T);
  if (nfr > 0x7fffffffLL / 4) return fail(EGR_ERR_ARG, "egr_dfn_mix: clip too long");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned* peak = reinterpret_cast<unsigned*>(d_work);
  float* rms = reinterpret_cast<float*>(reinterpret_cast<char*>(d_work) + 256);
  float* gd = rms + (size_t)C * nfr;
  float* gw = gd + (size_t)C * nfr;
  EGR_CUDA(cudaMemsetAsync(peak, 0, 256, st));
  const int vad = vad_source == EGR_VAD_RMS;
  if (vad) {
    const int aligned = (reinterpret_cast<uintptr_t>(d_dry) % 16 == 0) && (T % 4 == 0);
    dfn_frame_rms_kernel<<<dim3((unsigned)((nfr + 127) / 128), C), 128, 0, st>>>(d_dry, T, (int)nfr, rms, aligned);
    EGR_CHECK_LAUNCH("dfn_frame_rms_kernel");
  }
  DfnGainArgs g;
  g.nfr = (int)nfr; g.vad = vad; g.mode = adaptive_mode; g.curve = mix_curve;
  g.smooth = vad_smooth_ms > 0;
  g.s0 = (float)strength; g.amt = (float)adaptive_amount; g.oms0 = (float)(1.0 - strength);
  g.s_noise = (float)(strength + adaptive_amount * (1.0 - strength));
  g.s_speech = (float)(strength * (1.0 - adaptive_amount));
  g.thr = (float)vad_threshold;
  const double tau = vad_smooth_ms > 1e-3 ? (double)vad_smooth_ms : 1e-3;
  const double alpha = std::exp(-10.0 / tau);
  g.alpha = (float)alpha; g.oma = (float)(1.0 - alpha);
  dfn_frame_gain_kernel<<<C, 1024, 0, st>>>(g, rms, gd, gw);
  EGR_CHECK_LAUNCH("dfn_frame_gain_kernel");
  const int apply_gain = post_gain_db != 0.0;
  const float gain = (float)std::pow(10.0, post_gain_db / 20.0);
  const int sms = devinfo().sm_count ? devinfo().sm_count : 148;
  long long bx = (T + 255) / 256;
  if (bx > (long long)sms * 16) bx = (long long)sms * 16;
  dfn_mix_kernel<<<dim3((unsigned)bx, C), 256, 0, st>>>(d_dry, d_wet, T, (int)nfr, gd, gw, gain, apply_gain, d_out, peak);
  EGR_CHECK_LAUNCH("dfn_mix_kernel");
  long long bl = ((long long)C * T + 255) / 256;
  if (bl > (long long)sms * 16) bl = (long long)sms * 16;
  dfn_limit_kernel<<<(unsigned)bl, 256, 0, st>>>(d_out, (long long)C * T, peak, limit_ceiling, ceiling);
  EGR_CHECK_LAUNCH("dfn_limit_kernel");
  return EGR_OK;
}
