// fatllama.cu — path B: the Fat-Llama iterative spectral loop, replacing the arithmetic behind the reference's
// single call feed.upscale(...) (egregora_fat_llama_gpu.py:213-224 / egregora_fat_llama_cpu.py:126-134):
//
//   expanded = repeat(x, U);  s = where(|expanded| > thr, expanded, 0)
//   repeat iters times:  S = fft(s);  S = where(|S| > thr, S, 0);  s = ifft(S).real
//   y = expanded + s;  [autoscale per channel];  [normalise over all channels]
//
// Fast path (N = n*U even, N/2 plannable): the length-N real transform is a length-M = N/2 complex transform of
// z[m] = s[2m] + i s[2m+1] plus a Hermitian split, and the complex transform is two-level (fft_plan.cuh).  One
// iteration is TWO kernels, each one read + one write of the 8*M-byte work array (16*N bytes per iteration and
// channel — the algorithmic minimum of DESIGN.md "K10"; both channels' arrays stay resident in the 126 MB L2):
//
//   fl_row_kernel : rows k1 and R1-k1 together: twiddle, forward length-R2 FFT, Hermitian split -> |X|>thr gate ->
//                   re-pack, inverse length-R2 FFT, conj twiddle, 1/M — all in shared memory
//   fl_col_kernel : a tile of columns: inverse length-R1 FFT (= the time-domain signal s), then straight away the
//                   next iteration's forward length-R1 FFT.  The first launch packs/thresholds the input instead of
//                   the inverse; the last one writes y = expanded + s and the per-channel peaks instead of the forward.
//
// General path (odd N or N/2 not plannable): complex length-N transforms through egr_fft_exec (Bluestein when
// needed) with a separate gate kernel — slower, same semantics.
#include "fft_plan.cuh"

using namespace egr;
using namespace egrfft;

namespace egr {
int fft2_forward_to_positions(const Fft2Plan* p, const float2* d_in, float2* d_out, int batch, cudaStream_t st);
int fft2_inverse_from_positions(const Fft2Plan* p, float2* d_pos, float2* d_out, int batch, float scale, cudaStream_t st);
}

struct FlDev {
  int R1, R2, cw, cw_shift;
  float inv_m;     // 1/M, rounded once from double on the host
  float inv_m_lo;  // 1/M - inv_m: the loop feeds its own output back 300 times, so a scale that is off by a CONSTANT
                   // relative 3e-8 (the rounding of 1/M) compounds into a gain drift of 1e-5; v*hi + v*lo has no bias
  long long M;
  Radices rd1, rd2;
  const float2 *tw1, *tw2, *twM_lo, *twM_hi, *twN_lo, *twN_hi, *twH;
  const int *perm1, *pos1, *perm2, *pos2, *pairT, *pairZ;
  const float2* twHp;
};

static FlDev fl_dev(const Fft2Plan* p) {
  FlDev d;
  d.R1 = p->R1; d.R2 = p->R2; d.cw = p->cw; d.M = p->M; d.rd1 = p->rd1; d.rd2 = p->rd2;
  d.tw1 = p->tw1; d.tw2 = p->tw2; d.twM_lo = p->twM_lo; d.twM_hi = p->twM_hi; d.twN_lo = p->twN_lo; d.twN_hi = p->twN_hi; d.twH = p->twH;
  d.perm1 = p->perm1; d.pos1 = p->pos1; d.perm2 = p->perm2; d.pos2 = p->pos2;
  d.pairT = p->pairT; d.pairZ = p->pairZ; d.twHp = p->twHp;
  d.cw_shift = 0;
  while ((1 << d.cw_shift) < d.cw) ++d.cw_shift;
  d.inv_m = (float)(1.0 / (double)p->M);
  d.inv_m_lo = (float)(1.0 / (double)p->M - (double)d.inv_m);
  return d;
}

__device__ __forceinline__ float2 tw_split(const float2* __restrict__ hi, const float2* __restrict__ lo, long long idx) {
  return cmulf(__ldg(hi + (idx >> 10)), __ldg(lo + (idx & 1023)));
}

__device__ __forceinline__ void block_atomic_max(float v, unsigned int* slot) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0 && v > 0.f) atomicMax(slot, __float_as_uint(v));
}

// Global -> shared staging with U independent loads in flight per thread.  (ncu source view, profiles/r1_n: written as
// a plain `for (i...) sm[i] = g[i]` loop the compiler keeps ONE load in flight, and 34 % / 43 % of the row / column
// kernel's samples sat on the store that waits for it.)
template <int U, typename T, class LD, class ST>
__device__ __forceinline__ void staged(const int n, LD ld, ST st) {
  for (int base = threadIdx.x; base < n; base += U * blockDim.x) {
    T v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * blockDim.x;
      if (i < n) v[u] = ld(i);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * blockDim.x;
      if (i < n) st(i, v[u]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ column kernel
#define FL_FIRST 0
#define FL_MID 1
#define FL_LAST 2

// d_in [C][n] f32 (integer-scaled samples), work [C][M] float2 in position order, d_out [C][N] f32, peaks [C][2] uint
template <int MODE, int MINB>
__global__ void __launch_bounds__(MINB == 4 ? 256 : 320, MINB) fl_col_kernel(FlDev d, const float* __restrict__ d_in, long long n, int U, float thr,
                                                     float2* __restrict__ work, float* __restrict__ d_out,
                                                     unsigned int* __restrict__ peaks) {
  extern __shared__ float2 sm[];
  const int cw = d.cw, R1 = d.R1, R2 = d.R2, cs = d.cw_shift;
  const int c0 = blockIdx.x * cw, ch = blockIdx.y;
  const long long N = 2 * d.M;
  const float* x = d_in + (long long)ch * n;
  float2* wk = work + (long long)ch * d.M;
  const int tile = R1 * cw;
  // stage twiddles of the length-R1 transform in shared memory: every butterfly reads R-1 of them
  float2* tws = sm + tile;
  staged<8, float2>(R1, [&](int i) { return __ldg(d.tw1 + i); }, [&](int i, float2 v) { tws[i] = v; });
  float pk = 0.f;
  auto expanded = [&](long long m) {  // the pair (2m, 2m+1) of the sample-repeated input
    if (U == 1) return __ldg(reinterpret_cast<const float2*>(x + 2 * m));
    return make_float2(__ldg(x + (2 * m) / U), __ldg(x + (2 * m + 1) / U));
  };
  if (MODE == FL_FIRST) {
    staged<8, float2>(tile,
        [&](int i) {
          const int r = i >> cs, c = i & (cw - 1);
          return (c0 + c < R2) ? expanded((long long)r * R2 + c0 + c) : make_float2(0.f, 0.f);
        },
        [&](int i, float2 v) {
          pk = fmaxf(pk, fmaxf(fabsf(v.x), fabsf(v.y)));
          v.x = fabsf(v.x) > thr ? v.x : 0.f;
          v.y = fabsf(v.y) > thr ? v.y : 0.f;
          sm[i] = v;
        });
    block_atomic_max(pk, peaks + 2 * ch);
  } else {
    staged<8, float2>(tile,
        [&](int i) {
          const int r = i >> cs, c = i & (cw - 1);
          return (c0 + c < R2) ? wk[(long long)r * R2 + c0 + c] : make_float2(0.f, 0.f);
        },
        [&](int i, float2 v) { sm[i] = v; });
  }
  __syncthreads();
  const Tile g{R1, cw, cw, 1};
  if (MODE != FL_FIRST) fft_inverse<true>(sm, g, d.rd1, tws);  // -> s in natural time order
  if (MODE != FL_LAST) {
    fft_forward<true>(sm, g, d.rd1, tws);
    for (int i = threadIdx.x; i < tile; i += blockDim.x) {
      const int r = i >> cs, c = i & (cw - 1);
      if (c0 + c < R2) wk[(long long)r * R2 + c0 + c] = sm[i];
    }
  } else {
    float* y = d_out + (long long)ch * N;
    staged<8, float2>(tile,
        [&](int i) {
          const int r = i >> cs, c = i & (cw - 1);
          return (c0 + c < R2) ? expanded((long long)r * R2 + c0 + c) : make_float2(0.f, 0.f);
        },
        [&](int i, float2 e) {
          const int r = i >> cs, c = i & (cw - 1);
          if (c0 + c < R2) {
            const long long m = (long long)r * R2 + c0 + c;
            const float2 sv = sm[i];
            const float2 o = make_float2(__fadd_rn(e.x, sv.x), __fadd_rn(e.y, sv.y));
            *reinterpret_cast<float2*>(y + 2 * m) = o;
            pk = fmaxf(pk, fmaxf(fabsf(o.x), fabsf(o.y)));
          }
        });
    block_atomic_max(pk, peaks + 2 * ch + 1);
  }
}

// ------------------------------------------------------------------------------------------------ row kernel
// Hermitian split of the packed transform:  X[k] = E + w O,  conj(X[M-k]) = E - w O,  w = exp(-2 pi i k / N),
//   E = (Z[k] + conj(Z[M-k]))/2,  O = -i (Z[k] - conj(Z[M-k]))/2 ;  re-pack: Z' = E' + i O'.
__device__ __forceinline__ void gate_pair(float2& za, float2& zb, const float2 w, const float thr) {
  const float2 E = make_float2(0.5f * (za.x + zb.x), 0.5f * (za.y - zb.y));
  const float2 D = make_float2(0.5f * (za.x - zb.x), 0.5f * (za.y + zb.y));  // = i O
  const float2 O = make_float2(D.y, -D.x);
  const float2 wO = cmulf(w, O);
  const float2 Xk = cadd(E, wO), Xc = csub(E, wO);  // X[k], conj(X[M-k])
  const bool gk = sqrtf(fmaf(Xk.x, Xk.x, Xk.y * Xk.y)) > thr;
  const bool gb = sqrtf(fmaf(Xc.x, Xc.x, Xc.y * Xc.y)) > thr;
  if (gk && gb) return;  // where(mask, X, 0) leaves both bins untouched
  if (!gk && !gb) { za = make_float2(0.f, 0.f); zb = za; return; }
  const float2 P = gk ? Xk : make_float2(0.f, 0.f), Q = gb ? Xc : make_float2(0.f, 0.f);
  const float2 E2 = make_float2(0.5f * (P.x + Q.x), 0.5f * (P.y + Q.y));
  const float2 O2 = cmulc(make_float2(0.5f * (P.x - Q.x), 0.5f * (P.y - Q.y)), w);
  za = make_float2(E2.x - O2.y, E2.y + O2.x);   // E' + i O'
  zb = make_float2(E2.x + O2.y, -E2.y + O2.x);  // conj(E') + i conj(O')
}

template <int MINB>
__global__ void __launch_bounds__(MINB == 4 ? 256 : 320, MINB) fl_row_kernel(FlDev d, float thr, float2* __restrict__ work) {
  extern __shared__ float2 sm[];
  const int R1 = d.R1, R2 = d.R2;
  const int k1a = blockIdx.x, k1b = (R1 - k1a) % R1;
  const bool two = (k1a != k1b);
  const int pa = d.pos1[k1a], pb = d.pos1[k1b];
  float2* wk = work + (long long)blockIdx.y * d.M;
  float2* rowA = wk + (long long)pa * R2;
  float2* rowB = wk + (long long)pb * R2;
  float2* sA = sm;
  float2* sB = sm + R2;
  float2* tws = sm + 2 * R2;  // stage twiddles of the length-R2 transform (see fl_col_kernel)
  // Row twiddles w_M^(i*k1), i < R2, as a two-level product built once per CTA: i = 64 a + b,
  // w^(i k1) = HI[a] * LO[b].  (Looking every element up in the global hi/lo tables cost ~6.5 L1 sectors per
  // element — 26x the data volume of the row itself — and made the kernel long-scoreboard bound.)
  const int NA = (R2 + 63) >> 6;
  float2* hiA = sm + 3 * R2;       // [NA]
  float2* loA = hiA + NA;          // [64]
  float2* hiB = loA + 64;          // [NA]
  float2* loB = hiB + NA;          // [64]
  struct Pair { float2 a, b; };
  constexpr int UB = 4;
  const int T = blockDim.x;
  auto load_pair = [&](int i) {
    Pair v;
    v.a = rowA[i];
    v.b = two ? rowB[i] : make_float2(0.f, 0.f);
    return v;
  };
  auto store_pair = [&](int i, const Pair& v) {
    sA[i] = k1a ? cmulf(v.a, cmulf(hiA[i >> 6], loA[i & 63])) : v.a;
    if (two) sB[i] = cmulf(v.b, cmulf(hiB[i >> 6], loB[i & 63]));
  };
  // first batch of row loads goes out before anything else; the twiddle staging below overlaps their latency
  Pair first[UB];
#pragma unroll
  for (int u = 0; u < UB; ++u) {
    const int i = threadIdx.x + u * T;
    if (i < R2) first[u] = load_pair(i);
  }
  staged<8, float2>(R2, [&](int i) { return __ldg(d.tw2 + i); }, [&](int i, float2 v) { tws[i] = v; });
  for (int i = threadIdx.x; i < 2 * (NA + 64); i += T) {
    const int row = i >= NA + 64, e = row ? i - (NA + 64) : i;
    const long long k1 = row ? k1b : k1a;
    const long long idx = (e < NA ? (long long)(e << 6) : (long long)(e - NA)) * k1;
    hiA[i] = tw_split(d.twM_hi, d.twM_lo, idx);  // hiA/loA/hiB/loB are contiguous
  }
  const float2 wN0 = tw_split(d.twN_hi, d.twN_lo, k1a);  // w_N^k1a; w_N^(k1a + R1 k2) = wN0 * twH[k2]
  __syncthreads();
#pragma unroll
  for (int u = 0; u < UB; ++u) {
    const int i = threadIdx.x + u * T;
    if (i < R2) store_pair(i, first[u]);
  }
  for (int base = threadIdx.x + UB * T; base < R2; base += UB * T) {
    Pair v[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u)
      if (base + u * T < R2) v[u] = load_pair(base + u * T);
#pragma unroll
    for (int u = 0; u < UB; ++u)
      if (base + u * T < R2) store_pair(base + u * T, v[u]);
  }
  __syncthreads();
  const Tile g{R2, two ? 2 : 1, 1, R2};
  fft_forward<false>(sm, g, d.rd2, tws);
  // gate: every table is indexed by the POSITION p2 (coalesced, no dependent gathers):
  //   perm2[p2] = k2,  pairT[p2] = pos2[R2-1-k2],  pairZ[p2] = pos2[(R2-k2) % R2],  twHp[p2] = w_N^(R1 k2)
  struct GateIn { int k2, q2; float2 w; };
  if (two) {
    staged<4, GateIn>(R2,
        [&](int p2) {
          GateIn in;
          in.k2 = 0;
          in.q2 = __ldg(d.pairT + p2);
          in.w = __ldg(d.twHp + p2);
          return in;
        },
        [&](int p2, const GateIn& in) {
          float2 za = sA[p2], zb = sB[in.q2];
          gate_pair(za, zb, cmulf(wN0, in.w), thr);
          sA[p2] = za;
          sB[in.q2] = zb;
        });
  } else {
    const int* pair = k1a == 0 ? d.pairZ : d.pairT;
    staged<4, GateIn>(R2,
        [&](int p2) {
          GateIn in;
          in.k2 = __ldg(d.perm2 + p2);
          in.q2 = __ldg(pair + p2);
          in.w = __ldg(d.twHp + p2);
          return in;
        },
        [&](int p2, const GateIn& in) {
          const long long k = k1a + (long long)R1 * in.k2;
          const long long kb = (d.M - k) % d.M;
          if (k > kb) return;
          float2 za = sA[p2];
          if (k == 0) {  // X[0] = re + im, X[M] = re - im, both real
            const float x0 = za.x + za.y, xm = za.x - za.y;
            const bool g0 = fabsf(x0) > thr, gm = fabsf(xm) > thr;
            if (!(g0 && gm)) {
              const float P = g0 ? x0 : 0.f, Q = gm ? xm : 0.f;
              sA[p2] = make_float2(0.5f * (P + Q), 0.5f * (P - Q));
            }
          } else if (k == kb) {  // k = M/2: X = conj(Z)
            if (!(sqrtf(fmaf(za.x, za.x, za.y * za.y)) > thr)) sA[p2] = make_float2(0.f, 0.f);
          } else {
            float2 zb = sA[in.q2];
            gate_pair(za, zb, cmulf(wN0, in.w), thr);
            sA[p2] = za;
            sA[in.q2] = zb;
          }
        });
  }
  __syncthreads();
  fft_inverse<false>(sm, g, d.rd2, tws);
  const float sc = d.inv_m, sl = d.inv_m_lo;
  for (int i = threadIdx.x; i < R2; i += T) {
    float2 v = k1a ? cmulc(sA[i], cmulf(hiA[i >> 6], loA[i & 63])) : sA[i];
    rowA[i] = make_float2(fmaf(v.x, sc, v.x * sl), fmaf(v.y, sc, v.y * sl));
    if (two) {
      v = cmulc(sB[i], cmulf(hiB[i >> 6], loB[i & 63]));
      rowB[i] = make_float2(fmaf(v.x, sc, v.x * sl), fmaf(v.y, sc, v.y * sl));
    }
  }
}

// ------------------------------------------------------------------------------------------------ general path kernels
__global__ void fl_gen_init_kernel(const float* __restrict__ d_in, long long n, int U, long long N, float thr,
                                   float2* __restrict__ z, unsigned int* __restrict__ peaks) {
  const int ch = blockIdx.y;
  float pk = 0.f;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += (long long)gridDim.x * blockDim.x) {
    const float v = d_in[(long long)ch * n + j / U];
    pk = fmaxf(pk, fabsf(v));
    z[(long long)ch * N + j] = make_float2(fabsf(v) > thr ? v : 0.f, 0.f);
  }
  block_atomic_max(pk, peaks + 2 * ch);
}
__global__ void fl_gen_gate_kernel(float2* __restrict__ z, long long total, float thr) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (long long)gridDim.x * blockDim.x) {
    const float2 v = z[j];
    if (!(sqrtf(fmaf(v.x, v.x, v.y * v.y)) > thr)) z[j] = make_float2(0.f, 0.f);
  }
}
__global__ void fl_gen_real_kernel(float2* __restrict__ z, long long total) {  // .real
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (long long)gridDim.x * blockDim.x) z[j].y = 0.f;
}
__global__ void fl_gen_final_kernel(const float* __restrict__ d_in, long long n, int U, long long N, const float2* __restrict__ z,
                                    float* __restrict__ y, unsigned int* __restrict__ peaks) {
  const int ch = blockIdx.y;
  float pk = 0.f;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += (long long)gridDim.x * blockDim.x) {
    const float o = __fadd_rn(d_in[(long long)ch * n + j / U], z[(long long)ch * N + j].x);
    y[(long long)ch * N + j] = o;
    pk = fmaxf(pk, fabsf(o));
  }
  block_atomic_max(pk, peaks + 2 * ch + 1);
}
// iters == 0: s = thresholded input
__global__ void fl_zero_iter_kernel(const float* __restrict__ d_in, long long n, int U, long long N, float thr,
                                    float* __restrict__ y, unsigned int* __restrict__ peaks) {
  const int ch = blockIdx.y;
  float pin = 0.f, pout = 0.f;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += (long long)gridDim.x * blockDim.x) {
    const float v = d_in[(long long)ch * n + j / U];
    const float o = __fadd_rn(v, fabsf(v) > thr ? v : 0.f);
    y[(long long)ch * N + j] = o;
    pin = fmaxf(pin, fabsf(v));
    pout = fmaxf(pout, fabsf(o));
  }
  block_atomic_max(pin, peaks + 2 * ch);
  block_atomic_max(pout, peaks + 2 * ch + 1);
}

// autoscale: y = (y / max|y_c|) * max|x_c| per channel; normalise: y / max over channels of the result's peak.
// Each step is a separately rounded f32 op, as numpy/cupy evaluate them.
__global__ void fl_scale_kernel(float* __restrict__ y, long long N, int C, const unsigned int* __restrict__ peaks, unsigned flags) {
  const int ch = blockIdx.y;
  const float pin = __uint_as_float(peaks[2 * ch]), pout = __uint_as_float(peaks[2 * ch + 1]);
  float g = 0.f;
  for (int c = 0; c < C; ++c) {
    const float pi_c = __uint_as_float(peaks[2 * c]), po_c = __uint_as_float(peaks[2 * c + 1]);
    // peak of channel c after the optional autoscale: |y|max/pout*pin evaluated at |y| = pout
    const float pc = (flags & EGR_FL_AUTOSCALE) ? __fmul_rn(__fdiv_rn(po_c, po_c), pi_c) : po_c;
    g = fmaxf(g, pc);
  }
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += (long long)gridDim.x * blockDim.x) {
    float v = y[(long long)ch * N + j];
    if (flags & EGR_FL_AUTOSCALE) v = __fmul_rn(__fdiv_rn(v, pout), pin);
    if (flags & EGR_FL_NORMALIZE) v = __fdiv_rn(v, g);
    y[(long long)ch * N + j] = v;
  }
}

// ------------------------------------------------------------------------------------------------ host
static unsigned grid1(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)(devinfo().sm_count ? devinfo().sm_count : 148) * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

static bool fast_path(long long N) { return N >= 2 && (N % 2 == 0) && fft2_plannable(N / 2); }

template <int MINB>
static int fl_fast_loop(const Fft2Plan* p, const FlDev& d, const float* d_in, long long n, int upscale, int iters,
                        float threshold, float2* work, float* d_out, unsigned int* peaks, int C, size_t smc, size_t smr,
                        cudaStream_t st) {
  if (smc > 48 * 1024) {
    EGR_CUDA(cudaFuncSetAttribute(fl_col_kernel<FL_FIRST, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smc));
    EGR_CUDA(cudaFuncSetAttribute(fl_col_kernel<FL_MID, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smc));
    EGR_CUDA(cudaFuncSetAttribute(fl_col_kernel<FL_LAST, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smc));
  }
  if (smr > 48 * 1024) EGR_CUDA(cudaFuncSetAttribute(fl_row_kernel<MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smr));
  const dim3 gc(ceil_div(p->R2, p->cw), C), gr(p->R1 / 2 + 1, C);
  fl_col_kernel<FL_FIRST, MINB><<<gc, p->col_threads, smc, st>>>(d, d_in, n, upscale, threshold, work, d_out, peaks);
  EGR_CHECK_LAUNCH("fl_col_kernel<first>");
  for (int it = 0; it < iters; ++it) {
    fl_row_kernel<MINB><<<gr, p->row_threads, smr, st>>>(d, threshold, work);
    if (it + 1 < iters)
      fl_col_kernel<FL_MID, MINB><<<gc, p->col_threads, smc, st>>>(d, d_in, n, upscale, threshold, work, d_out, peaks);
  }
  launch_count() += (unsigned long long)(2 * iters - 2);  // the loop's launches beyond the one counted below
  EGR_CHECK_LAUNCH("fl_row_kernel / fl_col_kernel<mid>");
  fl_col_kernel<FL_LAST, MINB><<<gc, p->col_threads, smc, st>>>(d, d_in, n, upscale, threshold, work, d_out, peaks);
  EGR_CHECK_LAUNCH("fl_col_kernel<last>");
  return EGR_OK;
}

struct GenPlanCache {
  long long N = 0;
  int batch = 0;
  egr_fft_plan* plan = nullptr;
};
static GenPlanCache g_gen;

extern "C" size_t egr_fatllama_workspace_bytes(int C, int64_t n, int upscale) {
  if (C < 1 || n < 1 || upscale < 1) return 256;
  const long long N = n * upscale;
  size_t head = 256;  // peaks
  if (fast_path(N)) return head + sizeof(float2) * (size_t)(N / 2) * C;
  // general path: complex signal + transform workspace (Bluestein needs 2 P per channel, P < 4N + slack)
  long long P = fft2_plannable(N) ? N : fft2_next_plannable(2 * N - 1);
  if (P < 0) P = 4 * N;
  const size_t wsz = fft2_plannable(N) ? sizeof(float2) * (size_t)N * C : 2 * sizeof(float2) * (size_t)P * C;
  return head + sizeof(float2) * (size_t)N * C + wsz;
}

extern "C" int egr_fatllama_run(const float* d_in, float* d_out, int C, int64_t n, int upscale, int iters, float threshold,
                                uint32_t flags, void* d_work, size_t work_bytes, void* stream) {
  if (!devinfo().inited) return fail(EGR_ERR_STATE, "egr_fatllama_run: call egr_init first");
  if (C < 1 || C > 65535 || n < 0 || upscale < 1 || iters < 0) return fail(EGR_ERR_ARG, "egr_fatllama_run: bad arguments");
  if (n == 0) return EGR_OK;
  if (!d_in || !d_out || !d_work) return fail(EGR_ERR_ARG, "egr_fatllama_run: null pointer");
  const size_t need = egr_fatllama_workspace_bytes(C, n, upscale);
  if (work_bytes < need) return fail(EGR_ERR_ARG, "egr_fatllama_run: workspace %zu < required %zu bytes", work_bytes, need);
  if (reinterpret_cast<uintptr_t>(d_work) % 256) return fail(EGR_ERR_ARG, "egr_fatllama_run: workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const long long N = n * upscale;
  unsigned int* peaks = reinterpret_cast<unsigned int*>(d_work);
  float2* work = reinterpret_cast<float2*>(reinterpret_cast<char*>(d_work) + 256);
  EGR_CUDA(cudaMemsetAsync(peaks, 0, 256, st));
  if (C > 32) return fail(EGR_ERR_ARG, "egr_fatllama_run: at most 32 channels");

  if (iters == 0) {
    fl_zero_iter_kernel<<<dim3(grid1(N), C), 256, 0, st>>>(d_in, n, upscale, N, threshold, d_out, peaks);
    EGR_CHECK_LAUNCH("fl_zero_iter_kernel");
  } else if (fast_path(N)) {
    const Fft2Plan* p = fft2_get_plan(N / 2);
    if (!p) return EGR_ERR_UNSUPPORTED;
    if (upscale == 1 && ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 7))
      return fail(EGR_ERR_ARG, "egr_fatllama_run: buffers must be 8-byte aligned");
    const FlDev d = fl_dev(p);
    const size_t smc = ((size_t)p->R1 * p->cw + p->R1) * sizeof(float2), smr = (3 * (size_t)p->R2 + 2 * (((size_t)p->R2 + 63) / 64 + 64)) * sizeof(float2);
    int rc = 0;
    switch (p->min_blocks) {
      case 3: rc = fl_fast_loop<3>(p, d, d_in, n, upscale, iters, threshold, work, d_out, peaks, C, smc, smr, st); break;
      case 4: rc = fl_fast_loop<4>(p, d, d_in, n, upscale, iters, threshold, work, d_out, peaks, C, smc, smr, st); break;
      default: rc = fl_fast_loop<2>(p, d, d_in, n, upscale, iters, threshold, work, d_out, peaks, C, smc, smr, st); break;
    }
    if (rc) return rc;
  } else {
    if (g_gen.N != N || g_gen.batch != C) {
      if (g_gen.plan) { cudaStreamSynchronize(st); egr_fft_plan_destroy(g_gen.plan); g_gen.plan = nullptr; }
      int rc = egr_fft_plan_create(N, C, &g_gen.plan);
      if (rc) return rc;
      g_gen.N = N; g_gen.batch = C;
    }
    float2* z = work;
    float* fw = reinterpret_cast<float*>(work + (size_t)N * C);
    fl_gen_init_kernel<<<dim3(grid1(N), C), 256, 0, st>>>(d_in, n, upscale, N, threshold, z, peaks);
    EGR_CHECK_LAUNCH("fl_gen_init_kernel");
    for (int it = 0; it < iters; ++it) {
      int rc = egr_fft_exec(g_gen.plan, reinterpret_cast<float*>(z), fw, 0, 0, st);
      if (rc) return rc;
      fl_gen_gate_kernel<<<grid1(N * C), 256, 0, st>>>(z, N * C, threshold);
      rc = egr_fft_exec(g_gen.plan, reinterpret_cast<float*>(z), fw, 1, 1, st);
      if (rc) return rc;
      fl_gen_real_kernel<<<grid1(N * C), 256, 0, st>>>(z, N * C);
    }
    fl_gen_final_kernel<<<dim3(grid1(N), C), 256, 0, st>>>(d_in, n, upscale, N, z, d_out, peaks);
    EGR_CHECK_LAUNCH("fl_gen_final_kernel");
  }
  if (flags & (EGR_FL_AUTOSCALE | EGR_FL_NORMALIZE)) {
    fl_scale_kernel<<<dim3(grid1(N), C), 256, 0, st>>>(d_out, N, C, peaks, flags);
    EGR_CHECK_LAUNCH("fl_scale_kernel");
  }
  return EGR_OK;
}
