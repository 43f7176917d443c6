// eval_metrics.cu — clip-scale evaluation reductions on the device: the null test of the
// reference's Audio_Null_Test node (/root/reference egregora_null_test_suite.py:421-470: optional least-squares scale,
// inversion, null = A + B, correlation, null RMS, overshoot statistics) and SI-SDR (egregora_audio_eval_pack.py:414-429).
// and the STFT log-spectral distance of the same nodes (_stft_mag / _lsd, egregora_audio_eval_pack.py:389-411 =
// egregora_null_test_suite.py:167-189).
//
// Two streaming passes with deterministic two-stage reductions (per-CTA partials in fixed order, no float atomics):
//   eval_pass1  mono means (float32, as A.mean(axis=0)) and their float64 dot products: k = <a,b>/<b,b>,
//               alpha = <s_hat,s>/<s,s> (float64 channel means, as _si_sdr), sums for the correlation means
//   eval_pass2  null signal (float32, bit-identical: (B*k).astype(f32), negate, add), |null| > 1 count, null energy,
//               centred correlation sums, SI-SDR target / noise energies
// HBM roofline: pass 1 reads 8*C*N bytes, pass 2 reads 8*C*N and writes 4*C*N.
//
// LSD (egr_eval_lsd), three launches:
//   eval_lsd_tables  np.hanning(n_fft) (float64 -> float32) and exp(-2 pi i k / n_fft), built on the device per call
//   eval_lsd_frames  one CTA per STFT frame: mono means of both clips, window, two packed-real radix-2 Stockham FFTs
//                    in shared memory (the front end's validated scheme, frontend.cu), 20*log10(|X| + 1e-12) in
//                    float32, per[frame] = sqrt(mean_k (LA - LB)^2 + 1e-12)
//   eval_lsd_final   one CTA: mean of per[] (fixed-order float64 sum) and np.percentile(per, 95) by exact radix
//                    select + numpy's float32 lerp (select.cuh)
//   Frames overlap 4x (hop = n_fft/4 by default) but stay in L2: HBM reads 8*C*N bytes, writes 4*frames.
//
// Integrated loudness (egr_eval_lufs; integrated_lufs + _k_weight, egregora_null_test_suite.py:125-164 =
// egregora_audio_eval_pack.py:132-167), four launches:
//   lufs_hp_kernel     the reference's one-pole high-pass, z = fl(fl(c1*x) + fl(k*z)), y = fl(x - z), is a float32
//                      recurrence evaluated sample by sample in Python.  Bit-exact AND parallel: every chunk of 2048
//                      samples starts from a GUESSED state (the recurrence run over the 2048 samples before it, from
//                      zero — a contraction with k = 0.984 forgets its start long before that) and records guess and
//                      end state
//   lufs_fix_kernel    one CTA per channel finds the first chunk whose guess is not bit-identical to the true end
//                      state of the previous chunk (none on ordinary audio); from there one thread verifies in order
//                      and recomputes mismatching chunks serially from the true state (the count is reported)
//   lufs_block_kernel  one CTA per 400 ms block (100 ms hop): HF tilt y[n] += fl(0.02*(y[n] - y[n-1])) and the channel
//                      mean in float32 as numpy does them, mean square in float64
//   lufs_final_kernel  ungated mean, -10 LU relative gate, gated mean (fixed-order float64 sums)
//   HBM: read 4*C*N + write 4*C*N (filtered signal) + read 4*C*N x 4 (block overlap, L2).
//
// High-band energy ratio (egr_eval_hf_band; _band_energy_hi_db, egregora_null_test_suite.py:190-197): channel mean ->
// whole-clip FFT through the path-B transform (egr_fft_exec: mixed radix, Bluestein for awkward lengths; no cuFFT) ->
// |X_k|^2 summed in float64 over all bins and over the bins with rfftfreq(k) >= lo_hz.
#include <cmath>
#include "common.cuh"
#include "select.cuh"

using namespace egr;

#define EV_THREADS 256
#define EV_NP1 6
#define EV_NP2 7

struct EvScalars {   // written by the finalize kernels, read by pass 2
  double k, alpha, mean_a, mean_b_raw;
  float kf;
  int pad;
};

template <int NV>
__device__ __forceinline__ void ev_block_reduce(double (&v)[NV], double* __restrict__ out /*[NV] of this CTA*/) {
  __shared__ double red[NV][EV_THREADS / 32];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const double w = warp_sum(v[i]);
    if ((threadIdx.x & 31) == 0) red[i][threadIdx.x >> 5] = w;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double t = 0.0;
    for (int w = 0; w < EV_THREADS / 32; ++w) t += red[threadIdx.x][w];
    out[threadIdx.x] = t;
  }
}

__device__ __forceinline__ float ev_mean32(const float* __restrict__ x, long long ld, int C, long long t) {
  float s = __ldg(x + t);
  for (int c = 1; c < C; ++c) s = __fadd_rn(s, __ldg(x + (long long)c * ld + t));
  return __fdiv_rn(s, (float)C);
}
// channel mean of (x * k) with the product rounded to float32 first, as (B * k).astype(np.float32).mean(axis=0)
__device__ __forceinline__ float ev_mean32_scaled(const float* __restrict__ x, long long ld, int C, long long t, float k) {
  float s = __fmul_rn(__ldg(x + t), k);
  for (int c = 1; c < C; ++c) s = __fadd_rn(s, __fmul_rn(__ldg(x + (long long)c * ld + t), k));
  return __fdiv_rn(s, (float)C);
}
__device__ __forceinline__ double ev_mean64(const float* __restrict__ x, long long ld, int C, long long t) {
  double s = (double)__ldg(x + t);
  for (int c = 1; c < C; ++c) s += (double)__ldg(x + (long long)c * ld + t);
  return s / (double)C;
}

__global__ void __launch_bounds__(EV_THREADS) eval_pass1_kernel(const float* __restrict__ A, long long lda,
                                                                 const float* __restrict__ B, long long ldb, int C,
                                                                 long long N, double* __restrict__ partials) {
  double v[EV_NP1] = {0, 0, 0, 0, 0, 0};  // <a,b>, <b,b>, <s_hat,s>, <s,s>, sum a, sum b
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x) {
    const double a = (double)ev_mean32(A, lda, C, t), b = (double)ev_mean32(B, ldb, C, t);
    const double s = ev_mean64(A, lda, C, t), sh = ev_mean64(B, ldb, C, t);
    v[0] = fma(a, b, v[0]); v[1] = fma(b, b, v[1]);
    v[2] = fma(sh, s, v[2]); v[3] = fma(s, s, v[3]);
    v[4] += a; v[5] += b;
  }
  ev_block_reduce<EV_NP1>(v, partials + (long long)blockIdx.x * EV_NP1);
}

__global__ void eval_fin1_kernel(const double* __restrict__ partials, int nblk, long long N, int ls_scale,
                                 EvScalars* __restrict__ sc) {
  if (threadIdx.x != 0) return;
  double r[EV_NP1] = {0, 0, 0, 0, 0, 0};
  for (int b = 0; b < nblk; ++b)
    for (int i = 0; i < EV_NP1; ++i) r[i] += partials[(long long)b * EV_NP1 + i];
  const double k = ls_scale ? r[0] / (r[1] + 1e-20) : 1.0;
  sc->k = k;
  sc->kf = (float)k;
  sc->alpha = r[2] / (r[3] + 1e-20);
  sc->mean_a = r[4] / (double)N;
  sc->mean_b_raw = r[5] / (double)N;
}

__global__ void __launch_bounds__(EV_THREADS) eval_pass2_kernel(const float* __restrict__ A, long long lda,
                                                                 const float* __restrict__ B, long long ldb, int C,
                                                                 long long N, int invert_b, int ls_scale,
                                                                 const EvScalars* __restrict__ sc, float* __restrict__ null_out,
                                                                 double* __restrict__ partials) {
  const float kf = sc->kf;
  const double alpha = sc->alpha, mean_a = sc->mean_a;
  // b_m = (-B).mean(axis=0) after scaling / inversion: sign * (scaled B) mean
  const double sgn = invert_b ? 1.0 : -1.0;
  const double mean_b = sgn * (ls_scale ? (double)kf : 1.0) * sc->mean_b_raw;
  double v[EV_NP2] = {0, 0, 0, 0, 0, 0, 0};  // overs, null energy, cab, caa, cbb, noise, target
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x) {
    float nsum = 0.f, bsum = 0.f;
    for (int c = 0; c < C; ++c) {
      const float a = __ldg(A + (long long)c * lda + t);
      float b = __ldg(B + (long long)c * ldb + t);
      if (ls_scale) b = __fmul_rn(b, kf);
      if (invert_b) b = -b;
      const float nl = __fadd_rn(a, b);
      if (null_out) null_out[(long long)c * N + t] = nl;
      if (fabsf(nl) > 1.0f) v[0] += 1.0;
      nsum = c ? __fadd_rn(nsum, nl) : nl;
      bsum = c ? __fadd_rn(bsum, -b) : -b;
    }
    const double nm = (double)__fdiv_rn(nsum, (float)C);
    v[1] = fma(nm, nm, v[1]);
    const double am = (double)ev_mean32(A, lda, C, t) - mean_a;
    const double bm = (double)__fdiv_rn(bsum, (float)C) - mean_b;
    v[2] = fma(am, bm, v[2]); v[3] = fma(am, am, v[3]); v[4] = fma(bm, bm, v[4]);
    const double s = ev_mean64(A, lda, C, t), sh = ev_mean64(B, ldb, C, t);
    const double st = alpha * s, e = sh - st;
    v[5] = fma(e, e, v[5]); v[6] = fma(st, st, v[6]);
  }
  ev_block_reduce<EV_NP2>(v, partials + (long long)blockIdx.x * EV_NP2);
}

__global__ void eval_fin2_kernel(const double* __restrict__ partials, int nblk, int C, long long N,
                                 const EvScalars* __restrict__ sc, double* __restrict__ metrics) {
  if (threadIdx.x != 0) return;
  double r[EV_NP2] = {0, 0, 0, 0, 0, 0, 0};
  for (int b = 0; b < nblk; ++b)
    for (int i = 0; i < EV_NP2; ++i) r[i] += partials[(long long)b * EV_NP2 + i];
  metrics[EGR_EVAL_SI_SDR_DB] = 10.0 * log10((r[6] + 1e-20) / (r[5] + 1e-20));
  metrics[EGR_EVAL_CORR] = r[2] / (sqrt(r[3]) * sqrt(r[4]) + 1e-20);
  metrics[EGR_EVAL_NULL_RMS_DBFS] = 10.0 * log10(r[1] / (double)N + 1e-20);
  metrics[EGR_EVAL_OVERSHOOT] = r[0];
  metrics[EGR_EVAL_CLIPPED_PCT] = 100.0 * r[0] / ((double)C * (double)N);
  metrics[EGR_EVAL_SCALE_K] = sc->k;
}

static int ev_blocks(long long N) {
  long long b = (N + EV_THREADS - 1) / EV_THREADS;
  const long long cap = (long long)(devinfo().sm_count ? devinfo().sm_count : 148) * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" size_t egr_eval_workspace_bytes(void) { return 256 + sizeof(double) * 148 * 8 * 2 * (EV_NP1 + EV_NP2); }

extern "C" int egr_eval_null_test(const float* d_ref, int64_t ld_ref, const float* d_proc, int64_t ld_proc, int C, int64_t N,
                                  int invert_b, int least_squares_scale, float* d_null, double* d_metrics, void* d_work,
                                  size_t work_bytes, void* stream) {
  if (!devinfo().inited) return fail(EGR_ERR_STATE, "egr_eval_null_test: call egr_init first");
  if (!d_ref || !d_proc || !d_metrics || !d_work || C < 1 || C > 64 || N < 1 || ld_ref < N || ld_proc < N)
    return fail(EGR_ERR_ARG, "egr_eval_null_test: bad arguments");
  if (reinterpret_cast<uintptr_t>(d_work) % 256) return fail(EGR_ERR_ARG, "egr_eval_null_test: workspace must be 256-byte aligned");
  const int nblk = ev_blocks(N);
  const size_t need = 256 + sizeof(double) * (size_t)nblk * (EV_NP1 + EV_NP2);
  if (work_bytes < need || need > egr_eval_workspace_bytes()) return fail(EGR_ERR_ARG, "egr_eval_null_test: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  EvScalars* sc = reinterpret_cast<EvScalars*>(d_work);
  double* p1 = reinterpret_cast<double*>(reinterpret_cast<char*>(d_work) + 256);
  double* p2 = p1 + (size_t)nblk * EV_NP1;
  eval_pass1_kernel<<<nblk, EV_THREADS, 0, st>>>(d_ref, ld_ref, d_proc, ld_proc, C, N, p1);
  EGR_CHECK_LAUNCH("eval_pass1_kernel");
  eval_fin1_kernel<<<1, 32, 0, st>>>(p1, nblk, N, least_squares_scale, sc);
  EGR_CHECK_LAUNCH("eval_fin1_kernel");
  eval_pass2_kernel<<<nblk, EV_THREADS, 0, st>>>(d_ref, ld_ref, d_proc, ld_proc, C, N, invert_b, least_squares_scale, sc, d_null, p2);
  EGR_CHECK_LAUNCH("eval_pass2_kernel");
  eval_fin2_kernel<<<1, 32, 0, st>>>(p2, nblk, C, N, sc, d_metrics);
  EGR_CHECK_LAUNCH("eval_fin2_kernel");
  return EGR_OK;
}

// ------------------------------------------------------------------------------------------------ LSD
#define LSD_THREADS 256

__global__ void eval_lsd_tables_kernel(int n_fft, float* __restrict__ window, float2* __restrict__ tw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  // np.hanning(M) = 0.5 + 0.5*cos(pi*n/(M-1)), n = 1-M, 3-M, ..., M-1 (float64), then .astype(float32)
  if (i < n_fft) window[i] = (float)(0.5 + 0.5 * cospi((double)(1 - n_fft + 2 * i) / (double)(n_fft - 1)));
  if (i <= n_fft / 2) {
    double sn, cs;
    sincospi(-2.0 * (double)i / (double)n_fft, &sn, &cs);
    tw[i] = make_float2((float)cs, (float)sn);
  }
}

__device__ __forceinline__ float2 lsd_cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// One frame of _stft_mag: mono[start : start + n_fft] (zero padded past N) * window -> rfft -> emit(k, 20*log10(|X_k| + 1e-12))
// for k = 0..n_fft/2.  Real input packed as an M = n_fft/2 point complex transform + Hermitian split.
template <class EMIT>
__device__ __forceinline__ void lsd_frame_logmag(const float* __restrict__ x, long long ld, int C, long long N, long long start,
                                                 int n_fft, const float* __restrict__ window, const float2* __restrict__ tw,
                                                 float2* bufA, float2* bufB, float gain, EMIT emit) {
  const int M = n_fft >> 1;
  for (int m = threadIdx.x; m < M; m += LSD_THREADS) {
    float v[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = 2 * m + u;
      const long long t = start + j;
      v[u] = t < N ? __fmul_rn(gain == 1.0f ? ev_mean32(x, ld, C, t) : ev_mean32_scaled(x, ld, C, t, gain), __ldg(window + j)) : 0.f;
    }
    bufA[m] = make_float2(v[0], v[1]);
  }
  __syncthreads();
  float2* src = bufA;
  float2* dst = bufB;
  for (int Ns = 1; Ns < M; Ns <<= 1) {  // Stockham radix-2, autosort, ping-pong; exp(-2 pi i k/M) = tw[2k]
    const int tstep = M / (2 * Ns);
    for (int j = threadIdx.x; j < (M >> 1); j += LSD_THREADS) {
      const int k = j & (Ns - 1);
      const float2 w = __ldg(tw + 2 * k * tstep);
      const float2 a = src[j];
      const float2 bb = lsd_cmul(src[j + (M >> 1)], w);
      const int j0 = ((j - k) << 1) + k;
      dst[j0] = make_float2(a.x + bb.x, a.y + bb.y);
      dst[j0 + Ns] = make_float2(a.x - bb.x, a.y - bb.y);
    }
    __syncthreads();
    float2* t = src; src = dst; dst = t;
  }
  // X[k] = (Z[k] + conj(Z[M-k]))/2 - i*w_k*(Z[k] - conj(Z[M-k]))/2, w_k = exp(-2 pi i k/n_fft)
  for (int k = threadIdx.x; k <= M; k += LSD_THREADS) {
    const float2 zk = src[k == M ? 0 : k];
    const float2 zm = src[k == 0 ? 0 : M - k];
    const float2 e = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
    const float2 o = make_float2(0.5f * (zk.x - zm.x), 0.5f * (zk.y + zm.y));
    const float2 wo = lsd_cmul(__ldg(tw + k), o);
    const float re = e.x + wo.y, im = e.y - wo.x;
    emit(k, __fmul_rn(20.0f, log10f(__fadd_rn(sqrtf(re * re + im * im), 1e-12f))));
  }
  __syncthreads();  // the buffers are reused by the next call
}

__global__ void __launch_bounds__(LSD_THREADS) eval_lsd_frames_kernel(const float* __restrict__ A, long long lda,
                                                                       const float* __restrict__ B, long long ldb, int C,
                                                                       long long N, int n_fft, int hop,
                                                                       const float* __restrict__ window,
                                                                       const float2* __restrict__ tw, float* __restrict__ per,
                                                                       float proc_gain) {
  extern __shared__ float2 lsd_sm[];
  __shared__ double red[LSD_THREADS / 32];
  const int M = n_fft >> 1;
  float2* bufA = lsd_sm;
  float2* bufB = lsd_sm + M;
  float* LA = reinterpret_cast<float*>(lsd_sm + 2 * M);  // [M+1]
  const long long start = (long long)blockIdx.x * hop;
  lsd_frame_logmag(A, lda, C, N, start, n_fft, window, tw, bufA, bufB, 1.0f, [&](int k, float l) { LA[k] = l; });
  double acc = 0.0;  // every thread meets the bins k it wrote itself (same k -> thread mapping in both calls)
  lsd_frame_logmag(B, ldb, C, N, start, n_fft, window, tw, bufA, bufB, proc_gain, [&](int k, float l) {
    const float d = __fsub_rn(LA[k], l);
    acc += (double)__fmul_rn(d, d);
  });
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < LSD_THREADS / 32; ++w) t += red[w];
    per[blockIdx.x] = __fsqrt_rn(__fadd_rn((float)(t / (double)(M + 1)), 1e-12f));
  }
}

__global__ void __launch_bounds__(1024) eval_lsd_final_kernel(const float* __restrict__ per, int frames, double* __restrict__ metrics) {
  __shared__ unsigned hist[4096];
  __shared__ unsigned bc[2];
  __shared__ unsigned wsum[32];
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < frames; i += 1024) s += (double)per[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  const float p95 = percentile95_f32_nonneg(per, frames, hist, bc, wsum);
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 32; ++w) t += red[w];
    metrics[EGR_LSD_MEAN_DB] = (double)(float)(t / (double)frames);  // np.mean of a float32 array is a float32
    metrics[EGR_LSD_P95_DB] = (double)p95;
    metrics[EGR_LSD_FRAMES] = (double)frames;
  }
}

static long long lsd_frames(long long N, int n_fft, int hop) { return 1 + (N > n_fft ? (N - n_fft) / hop : 0); }

extern "C" size_t egr_eval_lsd_workspace_bytes(int64_t N, int n_fft, int hop) {
  if (N < 1 || n_fft < 2 || hop < 1) return 0;
  const size_t raw = sizeof(float) * (size_t)n_fft + sizeof(float2) * (size_t)(n_fft / 2 + 1) + sizeof(float) * (size_t)lsd_frames(N, n_fft, hop);
  return (raw + 255) / 256 * 256;
}

extern "C" int egr_eval_lsd(const float* d_ref, int64_t ld_ref, const float* d_proc, int64_t ld_proc, int C, int64_t N, int n_fft,
                            int hop, float proc_gain, double* d_metrics, void* d_work, size_t work_bytes, void* stream) {
  if (!devinfo().inited) return fail(EGR_ERR_STATE, "egr_eval_lsd: call egr_init first");
  if (!d_ref || !d_proc || !d_metrics || !d_work || C < 1 || C > 64 || N < 1 || ld_ref < N || ld_proc < N || hop < 1)
    return fail(EGR_ERR_ARG, "egr_eval_lsd: bad arguments");
  if (n_fft < 64 || n_fft > 8192 || (n_fft & (n_fft - 1)))
    return fail(EGR_ERR_UNSUPPORTED, "egr_eval_lsd: n_fft must be a power of two in [64, 8192] (got %d)", n_fft);
  if (reinterpret_cast<uintptr_t>(d_work) % 256) return fail(EGR_ERR_ARG, "egr_eval_lsd: workspace must be 256-byte aligned");
  const long long frames = lsd_frames(N, n_fft, hop);
  if (frames > 0x7fffffffLL) return fail(EGR_ERR_UNSUPPORTED, "egr_eval_lsd: too many frames");
  if (work_bytes < egr_eval_lsd_workspace_bytes(N, n_fft, hop)) return fail(EGR_ERR_ARG, "egr_eval_lsd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* window = reinterpret_cast<float*>(d_work);
  float2* tw = reinterpret_cast<float2*>(window + n_fft);
  float* per = reinterpret_cast<float*>(tw + n_fft / 2 + 1);
  eval_lsd_tables_kernel<<<ceil_div(n_fft, 256), 256, 0, st>>>(n_fft, window, tw);
  EGR_CHECK_LAUNCH("eval_lsd_tables_kernel");
  const size_t smem = sizeof(float2) * (size_t)n_fft + sizeof(float) * (size_t)(n_fft / 2 + 1);
  if (smem > 48 * 1024) EGR_CUDA(cudaFuncSetAttribute(eval_lsd_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  eval_lsd_frames_kernel<<<(unsigned)frames, LSD_THREADS, smem, st>>>(d_ref, ld_ref, d_proc, ld_proc, C, N, n_fft, hop, window, tw, per, proc_gain);
  EGR_CHECK_LAUNCH("eval_lsd_frames_kernel");
  eval_lsd_final_kernel<<<1, 1024, 0, st>>>(per, (int)frames, d_metrics);
  EGR_CHECK_LAUNCH("eval_lsd_final_kernel");
  return EGR_OK;
}

// ------------------------------------------------------------------------------------------------ LUFS
#define LUFS_CHUNK 2048
#define LUFS_THREADS 256

struct LufsCoef { float c1, k; };  // float32 images of the reference's python floats (1 - k) and k

__device__ __forceinline__ float lufs_step(const LufsCoef q, float x, float& z) {
  z = __fadd_rn(__fmul_rn(q.c1, x), __fmul_rn(q.k, z));
  return __fsub_rn(x, z);
}

// grid (ceil(nchunks / 128), C); thread = one chunk of one channel
__global__ void __launch_bounds__(128) lufs_hp_kernel(const float* __restrict__ x, long long ld, long long N, int nchunks,
                                                      LufsCoef q, float* __restrict__ y, float* __restrict__ guess,
                                                      float* __restrict__ endst) {
  const int ci = blockIdx.x * blockDim.x + threadIdx.x, ch = blockIdx.y;
  if (ci >= nchunks) return;
  const float* xc = x + (long long)ch * ld;
  float* yc = y + (long long)ch * N;
  const long long s0 = (long long)ci * LUFS_CHUNK;
  float z = 0.f;
  for (long long n = s0 > LUFS_CHUNK ? s0 - LUFS_CHUNK : 0; n < s0; ++n) lufs_step(q, __ldg(xc + n), z);  // warm-up (exact for chunks 0 and 1)
  guess[(long long)ch * nchunks + ci] = z;
  const long long s1 = s0 + LUFS_CHUNK < N ? s0 + LUFS_CHUNK : N;
  for (long long n = s0; n < s1; ++n) yc[n] = lufs_step(q, __ldg(xc + n), z);
  endst[(long long)ch * nchunks + ci] = z;
}

// grid C; all threads look for the first chunk whose guess is not bit-identical to the previous chunk's end state (none on
// ordinary audio: one parallel pass and out); from there thread 0 verifies in order and repairs — a repaired chunk has a
// new end state, so its successor is judged against that
__global__ void __launch_bounds__(LUFS_THREADS) lufs_fix_kernel(const float* __restrict__ x, long long ld, long long N, int nchunks,
                                                                 LufsCoef q, float* __restrict__ y, const float* __restrict__ guess,
                                                                 float* __restrict__ endst, unsigned int* __restrict__ repaired) {
  __shared__ int first;
  const int ch = blockIdx.x;
  const float* gs = guess + (long long)ch * nchunks;
  float* es = endst + (long long)ch * nchunks;
  if (threadIdx.x == 0) first = nchunks;
  __syncthreads();
  for (int ci = 1 + threadIdx.x; ci < nchunks; ci += LUFS_THREADS)
    if (__float_as_uint(es[ci - 1]) != __float_as_uint(gs[ci])) atomicMin(&first, ci);
  __syncthreads();
  if (threadIdx.x != 0 || first >= nchunks) return;
  const float* xc = x + (long long)ch * ld;
  float* yc = y + (long long)ch * N;
  unsigned int cnt = 0;
  for (int ci = first; ci < nchunks; ++ci) {
    float z = es[ci - 1];
    if (__float_as_uint(z) == __float_as_uint(gs[ci])) continue;
    const long long s0 = (long long)ci * LUFS_CHUNK, s1 = s0 + LUFS_CHUNK < N ? s0 + LUFS_CHUNK : N;
    for (long long n = s0; n < s1; ++n) yc[n] = lufs_step(q, xc[n], z);
    es[ci] = z;
    ++cnt;
  }
  if (cnt) atomicAdd(repaired, cnt);
}

// grid frames; ms[i] = mean over the block of (channel mean of the tilted signal)^2
__global__ void __launch_bounds__(LUFS_THREADS) lufs_block_kernel(const float* __restrict__ y, int C, long long N, int blk, int hop,
                                                                   double* __restrict__ ms) {
  __shared__ double red[LUFS_THREADS / 32];
  const long long s = (long long)blockIdx.x * hop;
  const long long e = s + blk < N ? s + blk : N;
  double acc = 0.0;
  for (long long n = s + threadIdx.x; n < e; n += LUFS_THREADS) {
    float sum = 0.f;
    for (int c = 0; c < C; ++c) {
      const float* yc = y + (long long)c * N;
      float v = yc[n];
      if (n > 0) v = __fadd_rn(v, __fmul_rn(0.02f, __fsub_rn(v, yc[n - 1])));   // y[:,1:] += 0.02*(y[:,1:] - y[:,:-1])
      sum = c ? __fadd_rn(sum, v) : v;
    }
    const double m = (double)__fdiv_rn(sum, (float)C);
    acc = fma(m, m, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < LUFS_THREADS / 32; ++w) t += red[w];
    ms[blockIdx.x] = t / (double)(e - s);
  }
}

__device__ double lufs_block_sum(double v, double* red /*[32] smem*/) {  // fixed-order sum over the 1024 threads, to all
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < 32; ++w) t += red[w];
  return t;
}

__global__ void __launch_bounds__(1024) lufs_final_kernel(const double* __restrict__ ms, int frames, const unsigned int* __restrict__ repaired,
                                                           double* __restrict__ metrics) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < frames; i += 1024) s += ms[i] + 1e-20;
  const double ungated = -0.691 + 10.0 * log10(lufs_block_sum(s, red) / (double)frames);
  const double gate = ungated - 10.0;
  double gs = 0.0, gn = 0.0;
  for (int i = threadIdx.x; i < frames; i += 1024) {
    const double m = ms[i] + 1e-20;
    if (-0.691 + 10.0 * log10(m) >= gate) { gs += m; gn += 1.0; }
  }
  gs = lufs_block_sum(gs, red);
  gn = lufs_block_sum(gn, red);
  if (threadIdx.x == 0) {
    metrics[EGR_LUFS_INTEGRATED] = gn > 0.0 ? -0.691 + 10.0 * log10(gs / gn) : ungated;
    metrics[EGR_LUFS_UNGATED] = ungated;
    metrics[EGR_LUFS_BLOCKS] = (double)frames;
    metrics[EGR_LUFS_REPAIRED] = (double)*repaired;
  }
}

static void lufs_geometry(int64_t N, int sr, int* blk, int* hop, long long* frames, int* nchunks) {
  const long r400 = (long)nearbyint(0.400 * (double)sr), r100 = (long)nearbyint(0.100 * (double)sr);  // python round(): half to even
  *blk = (int)(r400 < 1 ? 1 : r400);
  *hop = (int)(r100 < 1 ? 1 : r100);
  *frames = 1 + (N > *blk ? (N - *blk) / *hop : 0);
  *nchunks = (int)((N + LUFS_CHUNK - 1) / LUFS_CHUNK);
}

extern "C" size_t egr_eval_lufs_workspace_bytes(int C, int64_t N, int sample_rate) {
  if (C < 1 || N < 1 || sample_rate < 1) return 0;
  int blk, hop, nchunks; long long frames;
  lufs_geometry(N, sample_rate, &blk, &hop, &frames, &nchunks);
  const size_t y = ((sizeof(float) * (size_t)C * (size_t)N) + 255) / 256 * 256;
  const size_t st = ((2 * sizeof(float) * (size_t)C * (size_t)nchunks) + 255) / 256 * 256;
  const size_t ms = ((sizeof(double) * (size_t)frames) + 255) / 256 * 256;
  return 256 + y + st + ms;
}

extern "C" int egr_eval_lufs(const float* d_x, int64_t ld, int C, int64_t N, int sample_rate, double* d_metrics, void* d_work,
                             size_t work_bytes, void* stream) {
  if (!devinfo().inited) return fail(EGR_ERR_STATE, "egr_eval_lufs: call egr_init first");
  if (!d_x || !d_metrics || !d_work || C < 1 || C > 64 || N < 1 || ld < N || sample_rate < 1)
    return fail(EGR_ERR_ARG, "egr_eval_lufs: bad arguments");
  if (reinterpret_cast<uintptr_t>(d_work) % 256) return fail(EGR_ERR_ARG, "egr_eval_lufs: workspace must be 256-byte aligned");
  if (work_bytes < egr_eval_lufs_workspace_bytes(C, N, sample_rate)) return fail(EGR_ERR_ARG, "egr_eval_lufs: workspace too small");
  int blk, hop, nchunks; long long frames;
  lufs_geometry(N, sample_rate, &blk, &hop, &frames, &nchunks);
  if (frames > 0x7fffffffLL) return fail(EGR_ERR_UNSUPPORTED, "egr_eval_lufs: too many blocks");
  // k = exp(-2*pi*fc), fc = 60 / (sr/2): python floats in the reference, used as float32 (weak scalars) in its loop
  const double fc = 60.0 / ((double)sample_rate * 0.5);
  const double k = exp(-2.0 * 3.141592653589793 * fc);
  LufsCoef q;
  q.c1 = (float)(1.0 - k);
  q.k = (float)k;
  cudaStream_t st = (cudaStream_t)stream;
  char* w = reinterpret_cast<char*>(d_work);
  unsigned int* repaired = reinterpret_cast<unsigned int*>(w);
  float* y = reinterpret_cast<float*>(w + 256);
  float* guess = reinterpret_cast<float*>(w + 256 + ((sizeof(float) * (size_t)C * (size_t)N) + 255) / 256 * 256);
  float* endst = guess + (size_t)C * nchunks;
  double* ms = reinterpret_cast<double*>(reinterpret_cast<char*>(guess) + ((2 * sizeof(float) * (size_t)C * (size_t)nchunks) + 255) / 256 * 256);
  EGR_CUDA(cudaMemsetAsync(repaired, 0, 256, st));
  lufs_hp_kernel<<<dim3((unsigned)((nchunks + 127) / 128), C), 128, 0, st>>>(d_x, ld, N, nchunks, q, y, guess, endst);
  EGR_CHECK_LAUNCH("lufs_hp_kernel");
  lufs_fix_kernel<<<C, LUFS_THREADS, 0, st>>>(d_x, ld, N, nchunks, q, y, guess, endst, repaired);
  EGR_CHECK_LAUNCH("lufs_fix_kernel");
  lufs_block_kernel<<<(unsigned)frames, LUFS_THREADS, 0, st>>>(y, C, N, blk, hop, ms);
  EGR_CHECK_LAUNCH("lufs_block_kernel");
  lufs_final_kernel<<<1, 1024, 0, st>>>(ms, (int)frames, repaired, d_metrics);
  EGR_CHECK_LAUNCH("lufs_final_kernel");
  return EGR_OK;
}

// ------------------------------------------------------------------------------------------------ high-band energy
__global__ void __launch_bounds__(EV_THREADS) hf_pack_kernel(const float* __restrict__ x, long long ld, int C, long long N,
                                                             float2* __restrict__ z) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x)
    z[t] = make_float2(ev_mean32(x, ld, C, t), 0.f);
}

__global__ void __launch_bounds__(EV_THREADS) hf_energy_kernel(const float2* __restrict__ X, long long nbins, double fstep,
                                                               double lo_hz, double* __restrict__ partials) {
  double v[3] = {0, 0, 0};  // all bins, bins at or above lo_hz, how many of those
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nbins; k += (long long)gridDim.x * blockDim.x) {
    const float2 c = X[k];
    const double p = (double)c.x * (double)c.x + (double)c.y * (double)c.y;
    v[0] += p;
    if ((double)k * fstep >= lo_hz) { v[1] += p; v[2] += 1.0; }   // np.fft.rfftfreq: k * (1 / (n * d)), float64
  }
  ev_block_reduce<3>(v, partials + (long long)blockIdx.x * 3);
}

__global__ void hf_final_kernel(const double* __restrict__ partials, int nblk, double* __restrict__ metrics) {
  if (threadIdx.x != 0) return;
  double r[3] = {0, 0, 0};
  for (int b = 0; b < nblk; ++b)
    for (int i = 0; i < 3; ++i) r[i] += partials[(long long)b * 3 + i];
  metrics[EGR_HF_RESIDUAL_DB] = 10.0 * log10(r[1] / (r[0] + 1e-20) + 1e-20);
  metrics[EGR_HF_E_HI] = r[1];
  metrics[EGR_HF_E_ALL] = r[0];
  metrics[EGR_HF_BINS_HI] = r[2];
}

extern "C" size_t egr_eval_hf_band_workspace_bytes(const egr_fft_plan* plan, int64_t N) {
  if (!plan || N < 1) return 0;
  const size_t z = (sizeof(float2) * (size_t)N + 255) / 256 * 256;
  const size_t fw = (egr_fft_plan_workspace_bytes(plan) + 255) / 256 * 256;
  return z + fw + sizeof(double) * 148 * 8 * 3 + 256;
}

extern "C" int egr_eval_hf_band(egr_fft_plan* plan, const float* d_x, int64_t ld, int C, int64_t N, int sample_rate, double lo_hz,
                                double* d_metrics, void* d_work, size_t work_bytes, void* stream) {
  if (!devinfo().inited) return fail(EGR_ERR_STATE, "egr_eval_hf_band: call egr_init first");
  if (!plan || !d_x || !d_metrics || !d_work || C < 1 || C > 64 || N < 1 || ld < N || sample_rate < 1)
    return fail(EGR_ERR_ARG, "egr_eval_hf_band: bad arguments");
  if (reinterpret_cast<uintptr_t>(d_work) % 256) return fail(EGR_ERR_ARG, "egr_eval_hf_band: workspace must be 256-byte aligned");
  if (work_bytes < egr_eval_hf_band_workspace_bytes(plan, N)) return fail(EGR_ERR_ARG, "egr_eval_hf_band: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* w = reinterpret_cast<char*>(d_work);
  float2* z = reinterpret_cast<float2*>(w);
  float* fw = reinterpret_cast<float*>(w + (sizeof(float2) * (size_t)N + 255) / 256 * 256);
  double* partials = reinterpret_cast<double*>(reinterpret_cast<char*>(fw) + (egr_fft_plan_workspace_bytes(plan) + 255) / 256 * 256);
  const int nblk = ev_blocks(N);
  hf_pack_kernel<<<nblk, EV_THREADS, 0, st>>>(d_x, ld, C, N, z);
  EGR_CHECK_LAUNCH("hf_pack_kernel");
  int rc = egr_fft_exec(plan, reinterpret_cast<float*>(z), fw, 0, 0, stream);   // plan must be egr_fft_plan_create(N, 1)
  if (rc) return rc;
  const long long nbins = N / 2 + 1;
  const double fstep = 1.0 / ((double)N * (1.0 / (double)sample_rate));
  const int nb2 = ev_blocks(nbins);
  hf_energy_kernel<<<nb2, EV_THREADS, 0, st>>>(z, nbins, fstep, lo_hz, partials);
  EGR_CHECK_LAUNCH("hf_energy_kernel");
  hf_final_kernel<<<1, 32, 0, st>>>(partials, nb2, d_metrics);
  EGR_CHECK_LAUNCH("hf_final_kernel");
  return EGR_OK;
}
