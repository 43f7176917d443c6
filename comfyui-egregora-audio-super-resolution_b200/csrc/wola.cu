// wola.cu — path A driver kernels: span gather (slice + right zero-pad) and Hann weighted overlap-add.
// Semantics follow /root/reference egregora_audio_super_resolution.py:411-416 (gather) and :227-251
// (_wola_stitch).  Both are pure HBM streaming kernels: every chunk sample is read once, every output
// sample written once (algorithmic bytes = 4*(n*C*L + C*total), DESIGN.md "K9").
#include "common.cuh"

using namespace egr;

// ------------------------------------------------------------------------------------------------
// gather: chunks[k][c][j] = j < L_k ? in[c][s_k + j] : 0
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) chunk_gather_kernel(const float* __restrict__ in, int C, int64_t total,
                                                            const int64_t* __restrict__ starts,
                                                            const int32_t* __restrict__ lens, int win,
                                                            float* __restrict__ chunks) {
  const int k = blockIdx.z, c = blockIdx.y;
  const int64_t s = starts[k];
  const int L = min(lens[k], win);
  const float* src = in + (int64_t)c * total + s;
  float* dst = chunks + ((int64_t)k * C + c) * win;
  const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int jstride = gridDim.x * blockDim.x * 4;
  for (int j = j0; j < win; j += jstride) {
    if (vec && j + 4 <= L) {
      *reinterpret_cast<float4*>(dst + j) = __ldg(reinterpret_cast<const float4*>(src + j));
    } else if (vec && j >= L && j + 4 <= win) {
      *reinterpret_cast<float4*>(dst + j) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j + u < win) dst[j + u] = (j + u < L) ? src[j + u] : 0.f;
    }
  }
}

extern "C" int egr_chunk_gather(const float* d_in, int C, int64_t total, const int64_t* d_starts,
                                const int32_t* d_lens, int n_spans, int win, float* d_chunks, void* stream) {
  if (n_spans == 0) return EGR_OK;
  if (!d_in || !d_starts || !d_lens || !d_chunks || C <= 0 || win <= 0 || n_spans < 0 || total < 0)
    return fail(EGR_ERR_ARG, "egr_chunk_gather: bad arguments");
  if (C > 65535 || n_spans > 65535) return fail(EGR_ERR_ARG, "egr_chunk_gather: C and n_spans must be <= 65535");
  int bx = ceil_div(win, 256 * 4 * 4);
  if (bx < 1) bx = 1;
  dim3 grid(bx, C, n_spans);
  chunk_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_in, C, total, d_starts, d_lens, win, d_chunks);
  EGR_CHECK_LAUNCH("chunk_gather_kernel");
  return EGR_OK;
}

// ------------------------------------------------------------------------------------------------
// WOLA stitch, output-centric (no atomics): each thread owns 4 consecutive output samples of every
// channel, finds the spans that cover them by binary search over the sorted starts, and accumulates in
// span order with separately rounded multiply and add so the result is bit-identical to numpy's
//   acc[:, s:s+L] += y[:, :L] * w[None, :];  wsum[s:s+L] += w;  out = acc / where(wsum==0, 1, wsum)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wola_stitch_kernel(const float* __restrict__ chunks, int l_pred,
                                                           const int64_t* __restrict__ starts,
                                                           const int32_t* __restrict__ lens, int n_spans, int C,
                                                           int64_t total, int win,
                                                           const float* __restrict__ window,
                                                           float* __restrict__ out) {
  const int64_t t0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (t0 >= total) return;
  const int64_t tlast = min(t0 + 3, total - 1);
  // khi = last span with start <= tlast
  int lo = 0, hi = n_spans;  // first index with start > tlast
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (starts[mid] <= tlast) lo = mid + 1; else hi = mid;
  }
  const int khi = lo - 1;
  int klo = khi;
  while (klo > 0 && starts[klo - 1] > t0 - win) --klo;
  if (klo < 0) klo = 0;

  float wsum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = klo; k <= khi; ++k) {
    const int64_t s = starts[k];
    const int L = min(min(lens[k], l_pred), win);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t j = t0 + u - s;
      if (j >= 0 && j < L) wsum[u] = __fadd_rn(wsum[u], window[j]);
    }
  }
  const bool out_vec = (t0 + 4 <= total) && ((total & 3) == 0) &&
                       ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  for (int c = 0; c < C; ++c) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = klo; k <= khi; ++k) {
      const int64_t s = starts[k];
      const int L = min(min(lens[k], l_pred), win);
      const float* y = chunks + ((int64_t)k * C + c) * l_pred;
      const int64_t j0 = t0 - s;
      if (j0 >= 0 && j0 + 4 <= L && ((j0 & 3) == 0) && ((l_pred & 3) == 0) &&
          ((reinterpret_cast<uintptr_t>(chunks) | reinterpret_cast<uintptr_t>(window)) & 15) == 0) {
        const float4 yv = __ldg(reinterpret_cast<const float4*>(y + j0));
        const float4 wv = __ldg(reinterpret_cast<const float4*>(window + j0));
        acc[0] = __fadd_rn(acc[0], __fmul_rn(yv.x, wv.x));
        acc[1] = __fadd_rn(acc[1], __fmul_rn(yv.y, wv.y));
        acc[2] = __fadd_rn(acc[2], __fmul_rn(yv.z, wv.z));
        acc[3] = __fadd_rn(acc[3], __fmul_rn(yv.w, wv.w));
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t j = j0 + u;
          if (j >= 0 && j < L) acc[u] = __fadd_rn(acc[u], __fmul_rn(y[j], window[j]));
        }
      }
    }
    float r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) r[u] = __fdiv_rn(acc[u], wsum[u] == 0.f ? 1.0f : wsum[u]);
    float* o = out + (int64_t)c * total + t0;
    if (out_vec) {
      *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (t0 + u < total) o[u] = r[u];
    }
  }
}

extern "C" int egr_wola_stitch(const float* d_chunks, int l_pred, const int64_t* d_starts, const int32_t* d_lens,
                               int n_spans, int C, int64_t total, int win, const float* d_window, float* d_out,
                               void* stream) {
  if (total <= 0) return EGR_OK;
  if (!d_out || C <= 0 || win <= 0 || n_spans < 0) return fail(EGR_ERR_ARG, "egr_wola_stitch: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_spans == 0) {  // reference returns zeros for an empty span list
    EGR_CUDA(cudaMemsetAsync(d_out, 0, sizeof(float) * (size_t)C * (size_t)total, st));
    return EGR_OK;
  }
  if (!d_chunks || !d_starts || !d_lens || !d_window || l_pred <= 0)
    return fail(EGR_ERR_ARG, "egr_wola_stitch: null pointer / bad l_pred");
  int64_t threads = (total + 3) / 4;
  int64_t blocks = (threads + 255) / 256;
  if (blocks > 0x7fffffffLL) return fail(EGR_ERR_ARG, "egr_wola_stitch: total too large");
  wola_stitch_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_chunks, l_pred, d_starts, d_lens, n_spans, C, total, win,
                                                       d_window, d_out);
  EGR_CHECK_LAUNCH("wola_stitch_kernel");
  return EGR_OK;
}

// ------------------------------------------------------------------------------------------------
// Polyphase rational resampler either side of path A.
// Replaces the scipy branch of _resample_hq (egregora_audio_super_resolution.py:181-191):
// scipy.signal.resample_poly(x, up, down) on float32 data = upfirdn with a float32 Kaiser FIR.
// Output sample y (index into the un-trimmed upfirdn result) has phase t = (y*down) % up and newest
// input xi = (y*down) / up; it is  sum_{j=0}^{hpp-1} x[xi-hpp+1+j] * hflip[t][j]  accumulated in float32
// from a zero start in ascending j with separately rounded multiply and add — the order of scipy's
// upfirdn inner loop, so the result is bit-identical (zeros stand for samples outside [0, n_in)).
// hflip is scipy's transposed+flipped polyphase bank (host side: _resample_design in the node module).
//
// Persistent CTAs: the coefficient bank ([hpp][up] in smem, j-major so lanes with different phases hit
// different banks) is staged once per CTA, then the CTA walks output tiles; every tile stages its input
// window with coalesced loads.  HBM roofline: 4*n_in read + 4*n_out written per channel.
// ------------------------------------------------------------------------------------------------
template <bool BANK_SMEM>
__global__ void __launch_bounds__(256) resample_poly_kernel(const float* __restrict__ x, int64_t n_in, int up,
                                                             int down, const float* __restrict__ hflip, int hpp,
                                                             int64_t y_first, int64_t n_out,
                                                             float* __restrict__ y, int blk, int x_tile,
                                                             int64_t n_tiles) {
  extern __shared__ float rs_smem[];
  float* xs = rs_smem;                 // [x_tile]
  float* hs = rs_smem + x_tile;        // [hpp][up] when BANK_SMEM
  const int c = blockIdx.y;
  const float* xc = x + (int64_t)c * n_in;
  float* yc = y + (int64_t)c * n_out;
  if (BANK_SMEM) {
    for (int i = threadIdx.x; i < up * hpp; i += blockDim.x) {
      const int t = i / hpp, j = i - t * hpp;
      hs[j * up + t] = __ldg(hflip + i);
    }
  }
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t o0 = tile * blk;                       // first output of this tile (trimmed index)
    const int64_t p0 = (y_first + o0) * (int64_t)down;   // position on the up-sampled grid
    const int64_t xi0 = p0 / up;
    const int t0 = (int)(p0 - xi0 * up);
    const int64_t xlo = xi0 - hpp + 1;                   // global input index of xs[0]
    const int nb = (int)min((int64_t)blk, n_out - o0);
    const int need = (int)(((int64_t)t0 + (int64_t)(nb - 1) * down) / up) + hpp;
    __syncthreads();                                     // previous tile's readers are done (and hs is staged)
    for (int i = threadIdx.x; i < need; i += blockDim.x) {
      const int64_t g = xlo + i;
      xs[i] = (g >= 0 && g < n_in) ? __ldg(xc + g) : 0.f;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < nb; o += blockDim.x) {
      const int q = t0 + o * down;
      const int xr = q / up;
      const int t = q - xr * up;
      float acc = 0.f;
      if (BANK_SMEM) {
        for (int j = 0; j < hpp; ++j) acc = __fadd_rn(acc, __fmul_rn(xs[xr + j], hs[j * up + t]));
      } else {
        const float* hr = hflip + (int64_t)t * hpp;
        for (int j = 0; j < hpp; ++j) acc = __fadd_rn(acc, __fmul_rn(xs[xr + j], __ldg(hr + j)));
      }
      yc[o0 + o] = acc;
    }
  }
}

extern "C" int egr_resample_poly(const float* d_x, int C, int64_t n_in, int up, int down, const float* d_hflip,
                                 int hpp, int64_t y_first, int64_t n_out, float* d_y, void* stream) {
  if (n_out == 0 || C == 0) return EGR_OK;
  if (!d_x || !d_hflip || !d_y || C < 0 || C > 65535 || n_in <= 0 || up < 1 || down < 1 || hpp < 1 ||
      y_first < 0 || n_out < 0)
    return fail(EGR_ERR_ARG, "egr_resample_poly: bad arguments");
  if ((int64_t)up * hpp > (1 << 24) || down > (1 << 16) || up > (1 << 16))
    return fail(EGR_ERR_ARG, "egr_resample_poly: up/down/filter too large");
  // every produced sample must exist in the un-trimmed upfirdn output: ((n_in-1)*up + hpp*up - 1)/down + 1
  const int64_t full = (((n_in - 1) * (int64_t)up + (int64_t)hpp * up) - 1) / down + 1;
  if (y_first + n_out > full) return fail(EGR_ERR_ARG, "egr_resample_poly: output range beyond the filtered signal");
  const size_t bank_bytes = (size_t)up * hpp * sizeof(float);
  const bool bank_smem = bank_bytes <= 96 * 1024;
  // outputs per tile: the staged input window (blk*down/up + hpp + 2 floats) stays under 64 KB
  int64_t blk = 4096;
  const int64_t cap = (int64_t)(16384 - hpp - 2) * up / down;
  if (cap < 32) return fail(EGR_ERR_ARG, "egr_resample_poly: decimation factor too large");
  if (blk > cap) blk = cap;
  blk = (blk / 32) * 32;
  const int x_tile = (int)((blk * down) / up) + hpp + 2;
  const int64_t n_tiles = (n_out + blk - 1) / blk;
  const size_t smem = sizeof(float) * (size_t)x_tile + (bank_smem ? bank_bytes : 0);
  int sms = 148;
  {
    int dev = 0;
    EGR_CUDA(cudaGetDevice(&dev));
    EGR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  int64_t gx = (2 * (int64_t)sms + C - 1) / C;
  if (gx > n_tiles) gx = n_tiles;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)C);
  cudaStream_t st = (cudaStream_t)stream;
  if (bank_smem) {
    EGR_CUDA(cudaFuncSetAttribute(resample_poly_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    resample_poly_kernel<true><<<grid, 256, smem, st>>>(d_x, n_in, up, down, d_hflip, hpp, y_first, n_out, d_y,
                                                         (int)blk, x_tile, n_tiles);
  } else {
    EGR_CUDA(cudaFuncSetAttribute(resample_poly_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    resample_poly_kernel<false><<<grid, 256, smem, st>>>(d_x, n_in, up, down, d_hflip, hpp, y_first, n_out, d_y,
                                                          (int)blk, x_tile, n_tiles);
  }
  EGR_CHECK_LAUNCH("resample_poly_kernel");
  return EGR_OK;
}
