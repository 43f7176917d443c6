// select.cuh — exact order statistics of non-negative float32 data inside ONE CTA (1024 threads), shared by the
// DeepFilterNet mix (dfn_mix.cu: 95th percentile of the frame RMS) and the LSD metric (eval_metrics.cu: 95th
// percentile of the per-frame distances).  Both reproduce np.percentile(x, q) on a float32 array.
#pragma once
#include "common.cuh"

namespace egr {

// k-th smallest (0-based) of v[0..n) for non-negative floats: three radix passes over the bit pattern
__device__ inline unsigned radix_select_nonneg(const float* __restrict__ v, int n, int k, unsigned* hist /*[4096] smem*/,
                                               unsigned* bc /*[2] smem*/, unsigned* wsum /*[32] smem*/) {
  unsigned prefix = 0, mask = 0;
  const int shifts[3] = {20, 8, 0}, widths[3] = {12, 12, 8};
  for (int pass = 0; pass < 3; ++pass) {
    const int sh = shifts[pass], nb = 1 << widths[pass];
    for (int i = threadIdx.x; i < nb; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned b = __float_as_uint(v[i]);
      if ((b & mask) == prefix) atomicAdd(&hist[(b >> sh) & (nb - 1)], 1u);
    }
    __syncthreads();
    // block-wide exclusive scan of the histogram (each thread owns `per` consecutive bins), then the one thread whose
    // range contains rank k walks its own bins
    {
      const int per = nb >= 1024 ? nb / 1024 : 1;
      const int first = threadIdx.x * per;
      unsigned mine = 0;
      if (first < nb)
        for (int j = 0; j < per; ++j) mine += hist[first + j];
      unsigned incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += up;
      }
      if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
      __syncthreads();
      unsigned woff = 0;
      for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += wsum[w];
      const unsigned excl = woff + incl - mine;
      if (first < nb && (int)excl <= k && k < (int)(excl + mine)) {
        int acc = (int)excl, bin = first;
        for (; bin < first + per - 1; ++bin) {
          if (acc + (int)hist[bin] > k) break;
          acc += (int)hist[bin];
        }
        bc[0] = (unsigned)bin;
        bc[1] = (unsigned)acc;
      }
    }
    __syncthreads();
    prefix |= bc[0] << sh;
    mask |= (unsigned)(nb - 1) << sh;
    k -= (int)bc[1];
    __syncthreads();
  }
  return prefix;
}

// np.percentile(v, 95) on float32 data: q/100, the virtual index (n-1)*q and its fractional part are float32, then
// numpy's _lerp (incl. its t >= 0.5 form) between the two neighbouring order statistics.  All threads of the CTA call
// it and all receive the result.
__device__ inline float percentile95_f32_nonneg(const float* __restrict__ v, int n, unsigned* hist, unsigned* bc, unsigned* wsum) {
  const float q32 = __fdiv_rn(95.0f, 100.0f);
  const float pos = __fmul_rn((float)(n - 1), q32);
  const int lo = min((int)floorf(pos), n - 1), hi = min(lo + 1, n - 1);
  const float a = __uint_as_float(radix_select_nonneg(v, n, lo, hist, bc, wsum));
  const float b = __uint_as_float(radix_select_nonneg(v, n, hi, hist, bc, wsum));
  const float t = __fsub_rn(pos, (float)lo);
  const float d = __fsub_rn(b, a);
  return t >= 0.5f ? __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.0f, t))) : __fadd_rn(a, __fmul_rn(d, t));
}

}  // namespace egr
