// fft_device.cuh — in-register small DFTs (radix 2..16) and in-place shared-memory mixed-radix FFT stages.
//
// A length-L transform lives in shared memory as float2 and is transformed IN PLACE by a list of stages:
//   forward  = decimation in frequency: natural order in  -> mixed-radix digit-reversed order out
//   inverse  = the adjoint stages in reverse order: digit-reversed in -> natural out (unnormalised)
// so a forward followed by an inverse needs no reordering pass and no second buffer.  Each thread owns one
// radix-R butterfly per step: R complex values in registers, a fully unrolled DFT_R whose internal twiddles are
// compile-time constants, then the stage twiddles from a per-length table (computed in f64 on the host).
#pragma once
#include <cuda_runtime.h>

namespace egrfft {

// ------------------------------------------------------------------------------------------------ constexpr trig
constexpr double kPi = 3.14159265358979323846264338327950288;

constexpr double cx_sin_taylor(double x) {  // |x| <= pi/4
  double term = x, sum = x;
  const double x2 = x * x;
  for (int i = 1; i < 12; ++i) {
    term *= -x2 / ((2 * i) * (2 * i + 1));
    sum += term;
  }
  return sum;
}
constexpr double cx_cos_taylor(double x) {  // |x| <= pi/4
  double term = 1.0, sum = 1.0;
  const double x2 = x * x;
  for (int i = 1; i < 12; ++i) {
    term *= -x2 / ((2 * i - 1) * (2 * i));
    sum += term;
  }
  return sum;
}
// cos / sin of 2*pi*m/R with exact octant reduction on the integers
constexpr double cx_cos2pi(int m, int R) {
  m %= R;
  if (m < 0) m += R;
  // work in eighths of a turn: angle = 2*pi*m/R, o = floor(8m/R)
  const int o = (8 * m) / R;
  const double t = 2.0 * kPi * ((double)m / (double)R);
  switch (o) {
    case 0: return cx_cos_taylor(t);
    case 1: return cx_sin_taylor(kPi / 2 - t);
    case 2: return -cx_sin_taylor(t - kPi / 2);
    case 3: return -cx_cos_taylor(kPi - t);
    case 4: return -cx_cos_taylor(t - kPi);
    case 5: return -cx_sin_taylor(3 * kPi / 2 - t);
    case 6: return cx_sin_taylor(t - 3 * kPi / 2);
    default: return cx_cos_taylor(2 * kPi - t);
  }
}
constexpr double cx_sin2pi(int m, int R) {
  m %= R;
  if (m < 0) m += R;
  const int o = (8 * m) / R;
  const double t = 2.0 * kPi * ((double)m / (double)R);
  switch (o) {
    case 0: return cx_sin_taylor(t);
    case 1: return cx_cos_taylor(kPi / 2 - t);
    case 2: return cx_cos_taylor(t - kPi / 2);
    case 3: return cx_sin_taylor(kPi - t);
    case 4: return -cx_sin_taylor(t - kPi);
    case 5: return -cx_cos_taylor(3 * kPi / 2 - t);
    case 6: return -cx_cos_taylor(t - 3 * kPi / 2);
    default: return -cx_sin_taylor(2 * kPi - t);
  }
}
constexpr float tw_re(int m, int R) {
  m = ((m % R) + R) % R;
  if (m == 0) return 1.f;
  if (2 * m == R) return -1.f;
  if (4 * m == R || 4 * m == 3 * R) return 0.f;
  return (float)cx_cos2pi(m, R);
}
// imaginary part of exp(-2*pi*i*m/R) (forward sign)
constexpr float tw_im(int m, int R) {
  m = ((m % R) + R) % R;
  if (m == 0 || 2 * m == R) return 0.f;
  if (4 * m == R) return -1.f;
  if (4 * m == 3 * R) return 1.f;
  return (float)(-cx_sin2pi(m, R));
}

// ------------------------------------------------------------------------------------------------ complex helpers
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {  // a * conj(b)
  return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}
// multiply by the compile-time constant exp(-/+ 2*pi*i*m/R)
template <int M_, int R, bool INV>
__device__ __forceinline__ float2 ctw(float2 a) {
  constexpr float c = tw_re(M_, R);
  constexpr float s = INV ? -tw_im(M_, R) : tw_im(M_, R);
  if constexpr (s == 0.f && c == 1.f) return a;
  else if constexpr (s == 0.f && c == -1.f) return make_float2(-a.x, -a.y);
  else if constexpr (c == 0.f && s == 1.f) return make_float2(-a.y, a.x);
  else if constexpr (c == 0.f && s == -1.f) return make_float2(a.y, -a.x);
  else return make_float2(fmaf(a.x, c, -a.y * s), fmaf(a.x, s, a.y * c));
}

constexpr int smallest_factor(int r) {
  for (int p = 2; p * p <= r; ++p)
    if (r % p == 0) return p;
  return r;
}
// split used for composite radices: prefer 4 as the inner radix of 8/12/16
constexpr int split_a(int r) {
  if (r == 8) return 2;
  if (r == 16 || r == 12) return 4;
  return smallest_factor(r);
}

// ------------------------------------------------------------------------------------------------ DFT_R in registers
template <int R, bool INV>
struct Dft;

template <bool INV>
struct Dft<1, INV> {
  static __device__ __forceinline__ void run(float2 (&v)[1]) {}
};
template <bool INV>
struct Dft<2, INV> {
  static __device__ __forceinline__ void run(float2 (&v)[2]) {
    const float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  }
};
template <bool INV>
struct Dft<4, INV> {
  static __device__ __forceinline__ void run(float2 (&v)[4]) {
    const float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
    const float2 c = cadd(v[1], v[3]), d = csub(v[1], v[3]);
    // forward: X1 = b - i d, X3 = b + i d ; inverse swaps
    const float2 id = INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);  // (-/+ i) * d
    v[0] = cadd(a, c);
    v[2] = csub(a, c);
    v[1] = cadd(b, id);
    v[3] = csub(b, id);
  }
};

// odd primes: symmetric form.  a_k = v_k + v_{P-k}, b_k = v_k - v_{P-k};
//   X_m = v0 + sum_k cos(2 pi m k/P) a_k  -/+ i sum_k sin(2 pi m k/P) b_k   (forward: minus)
template <int P, bool INV>
struct DftOddPrime {
  static constexpr int H = (P - 1) / 2;
  template <int MM, int KK>
  static __device__ __forceinline__ void acc(const float2 (&a)[H], const float2 (&b)[H], float2& p, float2& q) {
    if constexpr (KK <= H) {
      constexpr float c = tw_re(MM * KK, P);
      constexpr float s = -tw_im(MM * KK, P);  // +sin(2 pi m k / P)
      p.x = fmaf(c, a[KK - 1].x, p.x);
      p.y = fmaf(c, a[KK - 1].y, p.y);
      q.x = fmaf(s, b[KK - 1].x, q.x);
      q.y = fmaf(s, b[KK - 1].y, q.y);
      acc<MM, KK + 1>(a, b, p, q);
    }
  }
  template <int MM>
  static __device__ __forceinline__ void outs(float2 (&v)[P], const float2 v0, const float2 (&a)[H], const float2 (&b)[H]) {
    if constexpr (MM <= H) {
      float2 p = v0, q = make_float2(0.f, 0.f);
      acc<MM, 1>(a, b, p, q);
      // forward: X_m = p - i q = (p.x + q.y, p.y - q.x); X_{P-m} = p + i q
      const float2 lo = make_float2(p.x + q.y, p.y - q.x), hi = make_float2(p.x - q.y, p.y + q.x);
      v[MM] = INV ? hi : lo;
      v[P - MM] = INV ? lo : hi;
      outs<MM + 1>(v, v0, a, b);
    }
  }
  static __device__ __forceinline__ void run(float2 (&v)[P]) {
    float2 a[H], b[H];
    float2 s0 = v[0];
#pragma unroll
    for (int k = 1; k <= H; ++k) {
      a[k - 1] = cadd(v[k], v[P - k]);
      b[k - 1] = csub(v[k], v[P - k]);
      s0 = cadd(s0, a[k - 1]);
    }
    const float2 v0 = v[0];
    outs<1>(v, v0, a, b);
    v[0] = s0;
  }
};
template <bool INV> struct Dft<3, INV> : DftOddPrime<3, INV> {};
template <bool INV> struct Dft<5, INV> : DftOddPrime<5, INV> {};
template <bool INV> struct Dft<7, INV> : DftOddPrime<7, INV> {};
template <bool INV> struct Dft<11, INV> : DftOddPrime<11, INV> {};
template <bool INV> struct Dft<13, INV> : DftOddPrime<13, INV> {};

// composite R = A*B, Cooley-Tukey in registers: n = n1*B + n2, k = k1 + A*k2
template <int R, bool INV>
struct Dft {
  static constexpr int A = split_a(R), B = R / A;
  template <int N2, int K1>
  static __device__ __forceinline__ void twrow(float2 (&y)[R], const float2 (&t)[A]) {
    if constexpr (K1 < A) {
      y[K1 * B + N2] = ctw<N2 * K1, R, INV>(t[K1]);
      twrow<N2, K1 + 1>(y, t);
    }
  }
  template <int N2>
  static __device__ __forceinline__ void cols(const float2 (&v)[R], float2 (&y)[R]) {
    if constexpr (N2 < B) {
      float2 t[A];
#pragma unroll
      for (int n1 = 0; n1 < A; ++n1) t[n1] = v[n1 * B + N2];
      Dft<A, INV>::run(t);
      twrow<N2, 0>(y, t);
      cols<N2 + 1>(v, y);
    }
  }
  static __device__ __forceinline__ void run(float2 (&v)[R]) {
    float2 y[R];
    cols<0>(v, y);
#pragma unroll
    for (int k1 = 0; k1 < A; ++k1) {
      float2 u[B];
#pragma unroll
      for (int n2 = 0; n2 < B; ++n2) u[n2] = y[k1 * B + n2];
      Dft<B, INV>::run(u);
#pragma unroll
      for (int k2 = 0; k2 < B; ++k2) v[k1 + A * k2] = u[k2];
    }
  }
};

// ------------------------------------------------------------------------------------------------ smem stages
// Geometry of the `nt` parallel transforms inside one CTA's shared-memory tile:
//   element e of transform o lives at data[o*ostride + e*estride]
// COLFAST = true  -> consecutive threads walk o first (column tiles: ostride 1, estride = tile width)
// COLFAST = false -> consecutive threads walk the butterfly index first (row tiles: estride 1)
struct Tile {
  int L;        // transform length
  int nt;       // transforms in the tile
  int estride;  // smem stride between consecutive elements of a transform
  int ostride;  // smem stride between transforms
};

template <int R, bool INV, bool COLFAST>
__device__ __forceinline__ void fft_stage(float2* __restrict__ data, const Tile g, const int Lb,
                                          const float2* __restrict__ tw /* exp(-2 pi i t / L), t < L; global or shared */) {
  const int Ls = Lb / R;
  const int nb = g.L / R;
  const int tstep = g.L / Lb;
  const int total = nb * g.nt;
  for (int t = threadIdx.x; t < total; t += blockDim.x) {
    int o, bf;
    if (COLFAST) { bf = t / g.nt; o = t - bf * g.nt; }
    else { o = t / nb; bf = t - o * nb; }
    const int blk = bf / Ls, j = bf - blk * Ls;
    float2* p = data + o * g.ostride + (blk * Lb + j) * g.estride;
    const int step = Ls * g.estride;
    float2 v[R];
#pragma unroll
    for (int q = 0; q < R; ++q) v[q] = p[q * step];
    if (!INV) {
      Dft<R, false>::run(v);
      if (Ls > 1) {
#pragma unroll
        for (int q = 1; q < R; ++q) v[q] = cmulf(v[q], tw[j * q * tstep]);
      }
    } else {
      if (Ls > 1) {
#pragma unroll
        for (int q = 1; q < R; ++q) v[q] = cmulc(v[q], tw[j * q * tstep]);
      }
      Dft<R, true>::run(v);
    }
#pragma unroll
    for (int q = 0; q < R; ++q) p[q * step] = v[q];
  }
}

template <bool INV, bool COLFAST>
__device__ __forceinline__ void fft_stage_dispatch(int radix, float2* data, const Tile g, int Lb, const float2* tw) {
  switch (radix) {
    case 2: fft_stage<2, INV, COLFAST>(data, g, Lb, tw); break;
    case 3: fft_stage<3, INV, COLFAST>(data, g, Lb, tw); break;
    case 4: fft_stage<4, INV, COLFAST>(data, g, Lb, tw); break;
    case 5: fft_stage<5, INV, COLFAST>(data, g, Lb, tw); break;
    case 6: fft_stage<6, INV, COLFAST>(data, g, Lb, tw); break;
    case 7: fft_stage<7, INV, COLFAST>(data, g, Lb, tw); break;
    case 8: fft_stage<8, INV, COLFAST>(data, g, Lb, tw); break;
    case 9: fft_stage<9, INV, COLFAST>(data, g, Lb, tw); break;
    case 10: fft_stage<10, INV, COLFAST>(data, g, Lb, tw); break;
    case 11: fft_stage<11, INV, COLFAST>(data, g, Lb, tw); break;
    case 12: fft_stage<12, INV, COLFAST>(data, g, Lb, tw); break;
    case 13: fft_stage<13, INV, COLFAST>(data, g, Lb, tw); break;
    case 14: fft_stage<14, INV, COLFAST>(data, g, Lb, tw); break;
    case 15: fft_stage<15, INV, COLFAST>(data, g, Lb, tw); break;
    case 16: fft_stage<16, INV, COLFAST>(data, g, Lb, tw); break;
    default: break;
  }
}

#define EGR_FFT_MAX_STAGES 12
struct Radices {
  int n;
  int r[EGR_FFT_MAX_STAGES];
};

// forward DIF over the whole tile: natural -> digit-reversed.  Ends with a __syncthreads().
template <bool COLFAST>
__device__ __forceinline__ void fft_forward(float2* data, const Tile g, const Radices rd, const float2* tw) {
  int Lb = g.L;
  for (int s = 0; s < rd.n; ++s) {
    fft_stage_dispatch<false, COLFAST>(rd.r[s], data, g, Lb, tw);
    Lb /= rd.r[s];
    __syncthreads();
  }
}
// inverse (adjoint, unnormalised): digit-reversed -> natural.  Ends with a __syncthreads().
template <bool COLFAST>
__device__ __forceinline__ void fft_inverse(float2* data, const Tile g, const Radices rd, const float2* tw) {
  int Lb = 1;
  for (int s = rd.n - 1; s >= 0; --s) {
    Lb *= rd.r[s];
    fft_stage_dispatch<true, COLFAST>(rd.r[s], data, g, Lb, tw);
    __syncthreads();
  }
}

}  // namespace egrfft
