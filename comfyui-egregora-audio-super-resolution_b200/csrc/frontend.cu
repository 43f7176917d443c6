// frontend.cu — FlashSR front end: fused STFT -> |.| -> mel -> log kernel, and the optional input low-pass
// (spectral roll-off detection + zero-phase SOS filter).
//
// STFT_MEL: one CTA per (frame, batch item).  The reflect-padded, Hann-windowed frame is staged in shared
// memory, transformed by an in-smem Stockham FFT (real input packed as an n_fft/2 complex transform + split),
// reduced to magnitudes and projected on the (sparse, triangular) mel filters without leaving the SM, so the
// only HBM traffic is 4*hop bytes in (frames overlap in L2) and 4*n_mels bytes out per frame.
#include "ops.cuh"

using namespace egr;

#define STFT_THREADS 256

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// mode 0: log-mel out [B, frames, n_mels]; mode 1: accumulate per-bin magnitude into energy f64 [B, n_freq]
__global__ void __launch_bounds__(STFT_THREADS) stft_mel_kernel(const float* __restrict__ wav, int T, int n_fft, int hop,
                                                                 int frames, const float* __restrict__ window,
                                                                 const float2* __restrict__ tw_half,  // exp(-2pi i k/(n_fft/2)), k < n_fft/4
                                                                 const float2* __restrict__ tw_full,  // exp(-2pi i k/n_fft),     k <= n_fft/2
                                                                 const float* __restrict__ mel_basis,
                                                                 const int* __restrict__ mel_lo, const int* __restrict__ mel_hi,
                                                                 int n_mels, float mag_eps, float log_clamp, int mode,
                                                                 float* __restrict__ out, double* __restrict__ energy) {
  extern __shared__ float2 sm2[];
  const int M = n_fft >> 1;  // complex length
  float2* bufA = sm2;
  float2* bufB = sm2 + M;
  float* mag = reinterpret_cast<float*>(sm2 + 2 * M);  // [M+1]
  const int frame = blockIdx.x, b = blockIdx.y;
  const float* x = wav + (long long)b * T;
  const int pad = (n_fft - hop) >> 1;
  const int start = frame * hop - pad;
  for (int m = threadIdx.x; m < M; m += STFT_THREADS) {
    float v[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = 2 * m + u;
      int i = start + j;
      if (i < 0) i = -i;
      if (i >= T) i = 2 * (T - 1) - i;
      i = max(0, min(T - 1, i));
      v[u] = x[i] * window[j];
    }
    bufA[m] = make_float2(v[0], v[1]);
  }
  __syncthreads();
  // Stockham radix-2, autosort, ping-pong
  float2* src = bufA; float2* dst = bufB;
  for (int Ns = 1; Ns < M; Ns <<= 1) {
    const int tstep = M / (2 * Ns);
    for (int j = threadIdx.x; j < (M >> 1); j += STFT_THREADS) {
      const int k = j & (Ns - 1);
      const float2 w = tw_half[k * tstep];
      const float2 a = src[j];
      const float2 bb = cmul(src[j + (M >> 1)], w);
      const int j0 = ((j - k) << 1) + k;
      dst[j0] = make_float2(a.x + bb.x, a.y + bb.y);
      dst[j0 + Ns] = make_float2(a.x - bb.x, a.y - bb.y);
    }
    __syncthreads();
    float2* t = src; src = dst; dst = t;
  }
  // split: X[k] = (Z[k] + conj(Z[M-k]))/2 - i*w_k*(Z[k] - conj(Z[M-k]))/2, w_k = exp(-2pi i k/n_fft)
  for (int k = threadIdx.x; k <= M; k += STFT_THREADS) {
    const float2 zk = src[k == M ? 0 : k];
    const float2 zm = src[(M - k) == M ? 0 : (M - k)];
    const float2 e = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
    const float2 o = make_float2(0.5f * (zk.x - zm.x), 0.5f * (zk.y + zm.y));
    const float2 w = tw_full[k];
    // -i * w * o
    const float2 wo = cmul(w, o);
    const float re = e.x + wo.y, im = e.y - wo.x;
    mag[k] = sqrtf(re * re + im * im + mag_eps);
  }
  __syncthreads();
  if (mode == 0) {
    const int n_freq = M + 1;
    for (int m = threadIdx.x; m < n_mels; m += STFT_THREADS) {
      float acc = 0.f;
      const float* row = mel_basis + (long long)m * n_freq;
      for (int f = mel_lo[m]; f < mel_hi[m]; ++f) acc = fmaf(row[f], mag[f], acc);
      out[((long long)b * frames + frame) * n_mels + m] = logf(fmaxf(acc, log_clamp));
    }
  } else {
    for (int k = threadIdx.x; k <= M; k += STFT_THREADS) atomicAdd(&energy[(long long)b * (M + 1) + k], (double)mag[k]);
  }
}

int egr::launch_stft_mel(const Spaces& s, const egr_op& op, cudaStream_t st) {
  const float* wav = (const float*)resolve(s, op.x0.addr);
  const int B = (int)op.i[EGR_I_BATCH], T = (int)op.i[EGR_I_ROWS];
  const int n_fft = (int)op.i[EGR_I_AUX0], hop = (int)op.i[EGR_I_AUX1], n_mels = (int)op.i[EGR_I_AUX2];
  const int mode = (int)op.i[EGR_I_MODE];
  const int frames = (int)op.i[EGR_I_SEQ];
  // consts block (weights space): window[n_fft] | tw_half[n_fft/4] float2 | tw_full[n_fft/2+1] float2 | lo[n_mels] | hi[n_mels] | basis
  const char* cst = (const char*)resolve(s, op.ptr[EGR_P_AUX]);
  if (!wav || !cst || B <= 0 || T <= 0 || frames <= 0) return fail(EGR_ERR_ARG, "%s: bad arguments", op.name);
  if (n_fft < 64 || n_fft > 4096 || (n_fft & (n_fft - 1))) return fail(EGR_ERR_UNSUPPORTED, "%s: n_fft must be a power of two in [64,4096]", op.name);
  const float* window = (const float*)cst;
  const float2* tw_half = (const float2*)(window + n_fft);
  const float2* tw_full = tw_half + n_fft / 4;
  const int* lo = (const int*)(tw_full + n_fft / 2 + 1);
  const int* hi = lo + n_mels;
  const float* basis = (const float*)(hi + n_mels);
  float* out = (float*)resolve(s, op.ptr[EGR_P_OUT32]);
  double* energy = (double*)resolve(s, op.ptr[EGR_P_STATS]);
  if (mode == 0 && !out) return fail(EGR_ERR_ARG, "%s: null output", op.name);
  if (mode == 1 && !energy) return fail(EGR_ERR_ARG, "%s: null energy buffer", op.name);
  size_t smem = (size_t)n_fft * sizeof(float2) + (n_fft / 2 + 1) * sizeof(float);
  stft_mel_kernel<<<dim3(frames, B), STFT_THREADS, smem, st>>>(wav, T, n_fft, hop, frames, window, tw_half, tw_full, basis, lo, hi,
                                                             n_mels, (float)op.f[EGR_F_A], (float)op.f[EGR_F_B], mode, out, energy);
  EGR_CHECK_LAUNCH(op.name);
  return EGR_OK;
}

// ------------------------------------------------------------------------------------------------
// Low-pass: cutoff bin = (#bins whose cumulative energy < percentile*total) - 1, then scipy-style
// sosfiltfilt (odd extension of 3*(2*nsec+1) samples, sosfilt_zi initial conditions, forward + backward) in
// f64.  The IIR recurrence is parallelised over time: the whole cascade is one linear system, so each of the
// 256 threads runs a chunk from zero state, a 256-step serial scan stitches chunk boundary states with the
// chunk transition matrix, and a second sweep re-runs every chunk from its true initial state.
// ------------------------------------------------------------------------------------------------
#define LP_THREADS 256
#define LP_MAXSEC 4

struct SosState { double z[LP_MAXSEC][2]; };

__device__ __forceinline__ double sos_step(const double (*sos)[6], int nsec, SosState& s, double x) {
#pragma unroll
  for (int i = 0; i < LP_MAXSEC; ++i) {
    if (i < nsec) {
      const double y = fma(sos[i][0], x, s.z[i][0]);
      s.z[i][0] = fma(sos[i][1], x, fma(-sos[i][4], y, s.z[i][1]));
      s.z[i][1] = fma(sos[i][2], x, -sos[i][5] * y);
      x = y;
    }
  }
  return x;
}

// one direction of filtfilt over the (virtually) extended signal of length Lx; `get(n)` supplies input n.
template <class In, class Out>
__device__ void lp_sweep(const double (*sos)[6], const double (*zi)[2], int nsec, int Lx, In get, Out put,
                         double (*Mx)[2 * LP_MAXSEC], double (*ends)[2 * LP_MAXSEC]) {
  const int tid = threadIdx.x, ns2 = 2 * nsec;
  const int Lc = (Lx + LP_THREADS - 1) / LP_THREADS;
  // transition matrix of a full chunk: column k = response of the state to unit state k, zero input
  if (tid < ns2) {
    SosState s;
    for (int i = 0; i < LP_MAXSEC; ++i) s.z[i][0] = s.z[i][1] = 0.0;
    s.z[tid >> 1][tid & 1] = 1.0;
    for (int n = 0; n < Lc; ++n) sos_step(sos, nsec, s, 0.0);
    for (int r = 0; r < ns2; ++r) Mx[r][tid] = s.z[r >> 1][r & 1];
  }
  const int lo = tid * Lc, hi = min(Lx, lo + Lc);
  {
    SosState s;
    for (int i = 0; i < LP_MAXSEC; ++i) s.z[i][0] = s.z[i][1] = 0.0;
    for (int n = lo; n < hi; ++n) sos_step(sos, nsec, s, get(n));
    for (int r = 0; r < ns2; ++r) ends[tid + 1][r] = s.z[r >> 1][r & 1];  // zero-state end of chunk tid
  }
  __syncthreads();
  if (tid == 0) {  // serial scan: start[i+1] = M*start[i] + end0[i]; `ends[i]` becomes the true start state of chunk i
    const double x0 = get(0);
    double cur[2 * LP_MAXSEC];
    for (int r = 0; r < ns2; ++r) cur[r] = zi[r >> 1][r & 1] * x0;
    for (int i = 0; i < LP_THREADS; ++i) {
      double nxt[2 * LP_MAXSEC];
      for (int r = 0; r < ns2; ++r) {
        double a = ends[i + 1][r];
        for (int c = 0; c < ns2; ++c) a = fma(Mx[r][c], cur[c], a);
        nxt[r] = a;
      }
      for (int r = 0; r < ns2; ++r) { ends[i][r] = cur[r]; cur[r] = nxt[r]; }
    }
  }
  __syncthreads();
  {
    SosState s;
    for (int i = 0; i < LP_MAXSEC; ++i) s.z[i][0] = s.z[i][1] = 0.0;
    for (int r = 0; r < ns2; ++r) s.z[r >> 1][r & 1] = ends[tid][r];
    for (int n = lo; n < hi; ++n) put(n, sos_step(sos, nsec, s, get(n)));
  }
  __syncthreads();
}

__global__ void __launch_bounds__(LP_THREADS) lowpass_kernel(const float* __restrict__ wav, int T, const double* __restrict__ energy,
                                                              int n_freq, double percentile, const double* __restrict__ sos_tab,
                                                              const double* __restrict__ zi_tab, int nsec, double* __restrict__ scratch,
                                                              float* __restrict__ out, int* __restrict__ cutoff_out) {
  __shared__ double sos[LP_MAXSEC][6];
  __shared__ double zi[LP_MAXSEC][2];
  __shared__ double Mx[2 * LP_MAXSEC][2 * LP_MAXSEC];
  __shared__ double ends[LP_THREADS + 1][2 * LP_MAXSEC];
  __shared__ int s_bin;
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    const double* e = energy + (long long)b * n_freq;
    double total = 0.0;
    for (int k = 0; k < n_freq; ++k) total += e[k];
    const double thr = total * percentile;
    double cum = 0.0; int cnt = 0;
    for (int k = 0; k < n_freq; ++k) { cum += e[k]; if (cum < thr) ++cnt; }
    s_bin = max(cnt - 1, 0);
    if (cutoff_out) cutoff_out[b] = s_bin;
  }
  __syncthreads();
  const int bin = s_bin;
  if (threadIdx.x < nsec * 6) sos[threadIdx.x / 6][threadIdx.x % 6] = sos_tab[((long long)bin * nsec) * 6 + threadIdx.x];
  if (threadIdx.x < nsec * 2) zi[threadIdx.x / 2][threadIdx.x % 2] = zi_tab[((long long)bin * nsec) * 2 + threadIdx.x];
  __syncthreads();
  const int edge = 3 * (2 * nsec + 1);
  const int Lx = T + 2 * edge;
  const float* x = wav + (long long)b * T;
  double* y1 = scratch + (long long)b * Lx;
  auto ext = [&](int n) -> double {  // odd extension
    const int i = n - edge;
    if (i < 0) return 2.0 * (double)x[0] - (double)x[-i];
    if (i >= T) return 2.0 * (double)x[T - 1] - (double)x[2 * (T - 1) - i];
    return (double)x[i];
  };
  lp_sweep(sos, zi, nsec, Lx, ext, [&](int n, double v) { y1[n] = v; }, Mx, ends);
  float* o = out + (long long)b * T;
  lp_sweep(sos, zi, nsec, Lx, [&](int n) -> double { return y1[Lx - 1 - n]; },
           [&](int n, double v) {
             const int t = (Lx - 1 - n) - edge;
             if (t >= 0 && t < T) o[t] = (float)v;
           }, Mx, ends);
}

int egr::launch_lowpass(const Spaces& s, const egr_op& op, cudaStream_t st) {
  const float* wav = (const float*)resolve(s, op.x0.addr);
  const double* energy = (const double*)resolve(s, op.ptr[EGR_P_STATS]);
  const double* sos_tab = (const double*)resolve(s, op.ptr[EGR_P_W]);
  const double* zi_tab = (const double*)resolve(s, op.ptr[EGR_P_BIAS]);
  double* scratch = (double*)resolve(s, op.ptr[EGR_P_AUX]);
  float* out = (float*)resolve(s, op.ptr[EGR_P_OUT32]);
  int* cutoff = (int*)resolve(s, op.ptr[EGR_P_OUT16]);
  const int B = (int)op.i[EGR_I_BATCH], T = (int)op.i[EGR_I_ROWS], n_freq = (int)op.i[EGR_I_COLS], nsec = (int)op.i[EGR_I_AUX0];
  if (!wav || !energy || !sos_tab || !zi_tab || !scratch || !out || B <= 0 || T <= 0) return fail(EGR_ERR_ARG, "%s: bad arguments", op.name);
  if (nsec < 1 || nsec > LP_MAXSEC) return fail(EGR_ERR_UNSUPPORTED, "%s: 1..%d second-order sections supported", op.name, LP_MAXSEC);
  if (T <= 3 * (2 * nsec + 1)) return fail(EGR_ERR_ARG, "%s: signal shorter than the filtfilt edge", op.name);
  lowpass_kernel<<<B, LP_THREADS, 0, st>>>(wav, T, energy, n_freq, op.f[EGR_F_A], sos_tab, zi_tab, nsec, scratch, out, cutoff);
  EGR_CHECK_LAUNCH(op.name);
  return EGR_OK;
}
