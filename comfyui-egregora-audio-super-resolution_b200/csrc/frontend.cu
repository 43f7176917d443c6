// frontend.cu — FlashSR front end: fused STFT -> |.| -> mel -> log kernel, and the optional input low-pass
// (spectral roll-off detection + zero-phase SOS filter).
//
// STFT_MEL: one CTA per (frame, batch item).  The reflect-padded, Hann-windowed frame is staged in shared
// memory, transformed by an in-smem Stockham FFT (real input packed as an n_fft/2 complex transform + split),
// reduced to magnitudes and projected on the (sparse, triangular) mel filters without leaving the SM, so the
// only HBM traffic is 4*hop bytes in (frames overlap in L2) and 4*n_mels bytes out per frame.
#include <cstdlib>
#include <cooperative_groups.h>
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
// f64.  The IIR recurrence is parallelised over time: the whole cascade is one linear system with 2*nsec
// states, so every thread runs its chunk from zero state, a hierarchical scan turns the chunk end states into
// true chunk start states, and a second sweep re-runs every chunk from there.
// One thread-block CLUSTER of 8 CTAs x 1024 threads per batch item (the sweeps are FP64-issue bound, and one SM
// has 1/8 of the FP64 lanes a cluster has): 8192 chunks; scan levels = 32 lanes-worth of chunks per warp (M),
// 32 warps per CTA (M^32), 8 CTAs per cluster (M^1024, CTA totals exchanged through distributed shared memory).
// Between the passes the signal lives in a chunk-transposed f64 scratch ([step within chunk][chunk]), so every
// per-step access of a CTA's 1024 threads is one contiguous 8 KB row; the natural-order input / output is
// transposed through 32x32 shared-memory tiles.
// ------------------------------------------------------------------------------------------------
#define LP_NT 1024
#define LP_NC 8
#define LP_NQ (LP_NT * LP_NC)
#define LP_MAXSEC 4
#define LP_NS (2 * LP_MAXSEC)

namespace cg = cooperative_groups;

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

struct LpShared {
  double sos[LP_MAXSEC][6];
  double zi[LP_MAXSEC][2];
  double M1[LP_NS][LP_NS];    // state transition of one chunk
  double M32[LP_NS][LP_NS];   // ... of 32 chunks (one warp)
  double M1k[LP_NS][LP_NS];   // ... of 1024 chunks (one CTA)
  double tmp[LP_NS][LP_NS];
  double ends[LP_NT + 1][LP_NS];
  double wtot[33][LP_NS];
  double mytot[LP_NS];        // zero-start end state of this CTA's 1024 chunks (read by the other CTAs of the cluster)
  float tile[32][32][33];
  int bin, dbg;
};

// dst = src^(2^5) by five squarings (ns2*ns2 threads, one element each); ends with a __syncthreads()
__device__ __forceinline__ void lp_pow32(LpShared& sh, const double (*src0)[LP_NS], double (*dst)[LP_NS], int ns2) {
  const int tid = threadIdx.x;
  for (int it = 0; it < 5; ++it) {
    const double (*src)[LP_NS] = it == 0 ? src0 : dst;
    if (tid < ns2 * ns2) {
      const int r = tid / ns2, c = tid % ns2;
      double a = 0.0;
      for (int k = 0; k < ns2; ++k) a = fma(src[r][k], src[k][c], a);
      sh.tmp[r][c] = a;
    }
    __syncthreads();
    if (tid < ns2 * ns2) dst[tid / ns2][tid % ns2] = sh.tmp[tid / ns2][tid % ns2];
    __syncthreads();
  }
}

// One direction of filtfilt over the transposed scratch `buf` (in place).  Sweep chunk g = rank*1024 + tid covers sweep
// positions m = g*Lc + j; REV: storage position n = Lx-1-m (the time-reversed signal, reading what the forward sweep
// wrote — by other CTAs of the cluster, hence the cluster barrier at the end of every sweep).
template <bool REV>
__device__ void lp_sweep(LpShared& sh, int nsec, int Lx, int Lc, double* __restrict__ buf) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, ns2 = 2 * nsec;
  const int gch = rank * LP_NT + tid;
  // chunk transition matrix: column k = response of the state to unit state k under zero input
  if (tid < ns2) {
    SosState s;
    for (int i = 0; i < LP_MAXSEC; ++i) s.z[i][0] = s.z[i][1] = 0.0;
    s.z[tid >> 1][tid & 1] = 1.0;
    for (int n = 0; n < Lc; ++n) sos_step(sh.sos, nsec, s, 0.0);
    for (int r = 0; r < ns2; ++r) sh.M1[r][tid] = s.z[r >> 1][r & 1];
  }
  const long long m_lo = (long long)gch * Lc, m_hi = min((long long)Lx, m_lo + Lc);
  int q0, l0;  // transposed coordinates of the first position: n = q*Lc + l
  if (!REV) { q0 = gch; l0 = 0; }
  else { const long long n0 = (long long)Lx - 1 - m_lo; q0 = n0 >= 0 ? (int)(n0 / Lc) : 0; l0 = n0 >= 0 ? (int)(n0 - (long long)q0 * Lc) : 0; }
  auto run = [&](SosState& s, bool store) {
    int q = q0, l = l0;
    for (long long m = m_lo; m < m_hi; ++m) {
      double* p = buf + (size_t)l * LP_NQ + q;
      const double y = sos_step(sh.sos, nsec, s, *p);
      if (store) *p = y;
      if (!REV) ++l;
      else if (--l < 0) { l = Lc - 1; --q; }
    }
  };
  {
    SosState s;
    for (int i = 0; i < LP_MAXSEC; ++i) s.z[i][0] = s.z[i][1] = 0.0;
    run(s, false);
    for (int r = 0; r < ns2; ++r) sh.ends[tid][r] = s.z[r >> 1][r & 1];  // zero-state end of chunk tid
  }
  __syncthreads();
  lp_pow32(sh, sh.M1, sh.M32, ns2);
  lp_pow32(sh, sh.M32, sh.M1k, ns2);
  // Scans: one warp per sequence, lane r < ns2 owns state component r (row r of the matrix in registers, the
  // vector exchanged by shuffles), so a step is ns2 dependent DFMAs instead of a serial ns2 x ns2 loop.
  const int lane = tid & 31, wid = tid >> 5;
  const int rr = lane < ns2 ? lane : 0;
  double m1row[LP_NS], m32row[LP_NS];
#pragma unroll
  for (int c = 0; c < LP_NS; ++c) { m1row[c] = c < ns2 ? sh.M1[rr][c] : 0.0; m32row[c] = c < ns2 ? sh.M32[rr][c] : 0.0; }
  auto matvec_row = [&](const double (&row)[LP_NS], double v, double add) {
    double a = add;
#pragma unroll
    for (int c = 0; c < LP_NS; ++c) a = fma(row[c], __shfl_sync(0xffffffffu, v, c), a);
    return a;
  };
  // level 1: every warp scans its 32 chunks from zero: ends[k] <- local prefix (state at the START of chunk k if
  // the warp had started from zero), wtot[w+1] <- state after the warp's last chunk
  {
    double cur = 0.0;
    for (int i = 0; i < 32; ++i) {
      const int k = wid * 32 + i;
      const double e = sh.ends[k][rr];
      const double nxt = matvec_row(m1row, cur, e);
      if (lane < ns2) sh.ends[k][lane] = cur;
      cur = nxt;
    }
    if (lane < ns2) sh.wtot[wid + 1][lane] = cur;
  }
  __syncthreads();
  // level 2: warp 0 scans the 32 warp totals with M^32 from a zero CTA start; wtot[w] <- local prefix of warp w,
  // mytot <- zero-start end state of the whole CTA
  if (wid == 0) {
    double cur = 0.0;
    for (int w = 0; w < 32; ++w) {
      const double e = sh.wtot[w + 1][rr];
      const double nxt = matvec_row(m32row, cur, e);
      if (lane < ns2) sh.wtot[w][lane] = cur;
      cur = nxt;
    }
    if (lane < ns2) sh.mytot[lane] = cur;
  }
  cluster.sync();
  // level 3: true start of this CTA = M^1024 applied over the CTAs before it (totals read through DSMEM), then the
  // warp starts pick it up through powers of M^32
  if (wid == 0) {
    double m1krow[LP_NS];
#pragma unroll
    for (int c = 0; c < LP_NS; ++c) m1krow[c] = c < ns2 ? sh.M1k[rr][c] : 0.0;
    const double x0 = buf[REV ? ((size_t)((Lx - 1) % Lc) * LP_NQ + (Lx - 1) / Lc) : 0];
    double cur = sh.zi[rr >> 1][rr & 1] * x0;
    for (int r = 0; r < rank; ++r) {
      const double* remote = cluster.map_shared_rank(&sh.mytot[0], r);
      cur = matvec_row(m1krow, cur, remote[rr]);
    }
    for (int w = 0; w < 32; ++w) {
      if (lane < ns2) sh.wtot[w][lane] += cur;
      cur = matvec_row(m32row, cur, 0.0);
    }
  }
  __syncthreads();
  // level 4: true start of chunk k = M1^i * start(warp) + local prefix
  {
    double v = sh.wtot[wid][rr];
    for (int i = 0; i < 32; ++i) {
      const int k = wid * 32 + i;
      if (lane < ns2) sh.ends[k][lane] += v;
      v = matvec_row(m1row, v, 0.0);
    }
  }
  __syncthreads();
  {
    SosState s;
    for (int i = 0; i < LP_MAXSEC; ++i) s.z[i][0] = s.z[i][1] = 0.0;
    for (int r = 0; r < ns2; ++r) s.z[r >> 1][r & 1] = sh.ends[tid][r];
    run(s, true);
  }
  __threadfence();
  cluster.sync();  // the scratch written here is read by other CTAs next; mytot may be overwritten after this point
}

__global__ void __cluster_dims__(LP_NC, 1, 1) __launch_bounds__(LP_NT)
lowpass_kernel(const float* __restrict__ wav, int T, const double* __restrict__ energy, int n_freq, double percentile,
               const double* __restrict__ sos_tab, const double* __restrict__ zi_tab, int nsec, double* __restrict__ scratch,
               float* __restrict__ out, int* __restrict__ cutoff_out) {
  extern __shared__ __align__(16) unsigned char lp_raw[];
  LpShared& sh = *reinterpret_cast<LpShared*>(lp_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / LP_NC, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {  // every CTA of the cluster derives the same cutoff (sequential f64 sums, as numpy's cumsum)
    const double* e = energy + (long long)b * n_freq;
    double total = 0.0;
    for (int k = 0; k < n_freq; ++k) total += e[k];
    const double thr = total * percentile;
    double cum = 0.0; int cnt = 0;
    for (int k = 0; k < n_freq; ++k) { cum += e[k]; if (cum < thr) ++cnt; }
    sh.bin = max(cnt - 1, 0);
    if (cutoff_out && rank == 0) cutoff_out[b] = sh.bin;
  }
  __syncthreads();
  const int bin = sh.bin;
  if (tid < nsec * 6) sh.sos[tid / 6][tid % 6] = sos_tab[((long long)bin * nsec) * 6 + tid];
  if (tid < nsec * 2) sh.zi[tid / 2][tid % 2] = zi_tab[((long long)bin * nsec) * 2 + tid];
  const int edge = 3 * (2 * nsec + 1);
  const int Lx = T + 2 * edge;
  const int Lc = (Lx + LP_NQ - 1) / LP_NQ;
  const float* x = wav + (long long)b * T;
  double* buf = scratch + (size_t)b * ((size_t)Lc * LP_NQ);
  const double xl = 2.0 * (double)x[0], xr = 2.0 * (double)x[T - 1];
  // transpose in (this CTA's 1024 chunks): tile = 32 chunks x 32 steps; rows of a chunk are read contiguously,
  // written chunk-contiguous
  const int ltiles = (Lc + 31) / 32;
  const int qbase = rank * LP_NT;
  for (int tile = warp; tile < 32 * ltiles; tile += 32) {
    const int qt = tile / ltiles, lt = tile - qt * ltiles;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const long long n = (long long)(qbase + qt * 32 + r) * Lc + lt * 32 + lane;
      float v = 0.f;
      if (lt * 32 + lane < Lc && n < Lx) {
        long long i = n - edge;
        if (i < 0) i = -i;
        else if (i >= T) i = 2 * ((long long)T - 1) - i;
        v = x[i];
      }
      sh.tile[warp][r][lane] = v;
    }
    __syncwarp();
    for (int l = 0; l < 32; ++l) {
      const int ll = lt * 32 + l, q = qbase + qt * 32 + lane;
      if (ll < Lc) {
        const long long n = (long long)q * Lc + ll, i = n - edge;
        const double v = (double)sh.tile[warp][lane][l];
        buf[(size_t)ll * LP_NQ + q] = i < 0 ? xl - v : (i >= T ? xr - v : v);  // odd extension at both ends
      }
    }
    __syncwarp();
  }
  __threadfence();
  cluster.sync();  // the reverse sweep's first sample (x0) and, later, whole chunks come from other CTAs' columns
  lp_sweep<false>(sh, nsec, Lx, Lc, buf);
  lp_sweep<true>(sh, nsec, Lx, Lc, buf);
  // transpose out
  float* o = out + (long long)b * T;
  for (int tile = warp; tile < 32 * ltiles; tile += 32) {
    const int qt = tile / ltiles, lt = tile - qt * ltiles;
#pragma unroll 8
    for (int l = 0; l < 32; ++l) {
      const int ll = lt * 32 + l, q = qbase + qt * 32 + lane;
      sh.tile[warp][lane][l] = ll < Lc ? (float)buf[(size_t)ll * LP_NQ + q] : 0.f;
    }
    __syncwarp();
    for (int r = 0; r < 32; ++r) {
      const long long n = (long long)(qbase + qt * 32 + r) * Lc + lt * 32 + lane, t = n - edge;
      if (lt * 32 + lane < Lc && t >= 0 && t < T) o[t] = sh.tile[warp][r][lane];
    }
    __syncwarp();
  }
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
  // scratch must hold B * ceil(Lx/8192)*8192 doubles (flashsr_plan.py sizes it as B*(Lx+8192)*8 bytes)
  static bool attr_done = false;
  if (!attr_done) {
    EGR_CUDA(cudaFuncSetAttribute(lowpass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LpShared)));
    attr_done = true;
  }
  lowpass_kernel<<<B * LP_NC, LP_NT, sizeof(LpShared), st>>>(wav, T, energy, n_freq, op.f[EGR_F_A], sos_tab, zi_tab, nsec, scratch, out,
                                                              cutoff);
  EGR_CHECK_LAUNCH(op.name);
  return EGR_OK;
}
