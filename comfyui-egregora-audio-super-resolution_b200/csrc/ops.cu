// ops.cu — CUDA-core ops of the FlashSR plan: SIMT tap-GEMM (tiny-K / tiny-N layers and the in-library
// cross-check of the tcgen05 path), GroupNorm, LayerNorm, softmax, short-sequence attention, GEGLU,
// element-wise glue, anti-aliased SnakeBeta, time embedding.  All are HBM-bound streaming kernels:
// coalesced float4 / half2 accesses with channels innermost, warp-shuffle reductions, f64 accumulators
// where cancellation matters (GroupNorm moments).
#include <cstdlib>
#include <cooperative_groups.h>
#include "ops.cuh"

namespace egr {

int gemm_args_from_op(const Spaces& s, const egr_op& op, GemmArgs* g, Taps* taps, View* a) {
  *a = make_view(s, op.x0);
  if (!a->p) return fail(EGR_ERR_ARG, "%s: null A operand", op.name);
  if (a->stride[0] != 1) return fail(EGR_ERR_ARG, "%s: A stride[0] must be 1", op.name);
  g->dimW = (int)op.i[EGR_I_DIMW]; g->dimH = (int)op.i[EGR_I_DIMH]; g->dimB = (int)op.i[EGR_I_DIMB];
  g->bw = (int)op.i[EGR_I_BW]; g->bh = (int)op.i[EGR_I_BH]; g->bb = (int)op.i[EGR_I_BB];
  g->Wo = (int)op.i[EGR_I_WO]; g->Ho = (int)op.i[EGR_I_HO]; g->Bo = (int)op.i[EGR_I_BO];
  g->ntaps = (int)op.i[EGR_I_NTAPS]; g->K = (int)op.i[EGR_I_K]; g->N = (int)op.i[EGR_I_N];
  g->block_n = (int)op.i[EGR_I_BLOCKN];
  g->wstride_n = op.i[EGR_I_WSTRIDE_N]; g->wstride_z = op.i[EGR_I_WSTRIDE_Z];
  g->wz_batch = (int)op.i[EGR_I_WZ_BATCH];
  g->W = resolve(s, op.ptr[EGR_P_W]);
  g->bias = (const float*)resolve(s, op.ptr[EGR_P_BIAS]);
  g->rowbias = (const float*)resolve(s, op.ptr[EGR_P_ROWBIAS]);
  g->rowbias_stride = op.i[EGR_I_ROWBIAS_STRIDE];
  g->resid = (const float*)resolve(s, op.ptr[EGR_P_RESID]);
  g->resid2 = (const float*)resolve(s, op.ptr[EGR_P_RESID2]);
  g->post = op.f[EGR_F_POST] != 0.0 ? (float)op.f[EGR_F_POST] : 1.0f;
  g->out32 = (float*)resolve(s, op.ptr[EGR_P_OUT32]);
  g->out16 = (__half*)resolve(s, op.ptr[EGR_P_OUT16]);
  g->out_pix_stride = op.i[EGR_I_OUT_PIX_STRIDE]; g->out_batch_stride = op.i[EGR_I_OUT_BATCH_STRIDE];
  g->out_offset = op.i[EGR_I_OUT_OFFSET]; g->out_lo = op.i[EGR_I_OUT_LO]; g->out_hi = op.i[EGR_I_OUT_HI];
  g->out_n_stride = op.i[EGR_I_OUT_N_STRIDE];
  g->out_h_stride = (op.code == EGR_OP_GEMM_TC && op.i[EGR_I_OUT_H_STRIDE] > 0) ? op.i[EGR_I_OUT_H_STRIDE] : (long long)g->Wo * g->out_pix_stride;
  g->transposed = (int)op.i[EGR_I_TRANSPOSED]; g->act = (int)op.i[EGR_I_ACT];
  g->alpha = (float)op.f[EGR_F_ALPHA];
  if (g->ntaps < 1 || g->ntaps > EGR_MAX_TAPS) return fail(EGR_ERR_ARG, "%s: ntaps=%d out of range", op.name, g->ntaps);
  if (g->bw * g->bh * g->bb != 128) return fail(EGR_ERR_ARG, "%s: tile %dx%dx%d != 128 rows", op.name, g->bw, g->bh, g->bb);
  if (!g->W || (!g->out32 && !g->out16)) return fail(EGR_ERR_ARG, "%s: missing weights or output", op.name);
  if (g->K < 1 || g->N < 1) return fail(EGR_ERR_ARG, "%s: bad K/N", op.name);
  for (int d : {g->dimW, g->dimH, g->dimB})
    if (d < 1 || d > 4) return fail(EGR_ERR_ARG, "%s: tile dims must index A dims 1..4", op.name);
  memcpy(taps->t, op.tap, sizeof(taps->t));
  return EGR_OK;
}

}  // namespace egr

using namespace egr;

__device__ __forceinline__ float load_view(const View& a, const long long c[5]) {
  long long off = 0;
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    if (c[d] < 0 || c[d] >= a.dim[d]) return 0.f;
    off += c[d] * a.stride[d];
  }
  return a.elem ? __half2float(reinterpret_cast<const __half*>(a.p)[off]) : reinterpret_cast<const float*>(a.p)[off];
}

// ------------------------------------------------------------------------------------------------
// SIMT tap-GEMM, general tile: 32 pixels x 32 outputs per block, reduction index r = tap*K + k in chunks
// of 32 through shared memory.  W is f32 [Z][N][K].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gemm_simt_kernel(View a, GemmArgs g, Taps taps) {
  __shared__ float As[32][33];
  __shared__ float Ws[32][33];
  __shared__ int pw[32], ph[32], pb[32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long npix = (long long)g.Wo * g.Ho * g.Bo;
  const long long p0 = (long long)blockIdx.x * 32;
  const int n0 = blockIdx.y * 32;
  if (threadIdx.x < 32) {
    long long p = p0 + threadIdx.x;
    if (p < npix) {
      pw[threadIdx.x] = (int)(p % g.Wo);
      ph[threadIdx.x] = (int)((p / g.Wo) % g.Ho);
      pb[threadIdx.x] = (int)(p / ((long long)g.Wo * g.Ho));
    } else {
      pw[threadIdx.x] = -1; ph[threadIdx.x] = 0; pb[threadIdx.x] = 0;
    }
  }
  __syncthreads();
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int R = g.ntaps * g.K;
  const float* W = reinterpret_cast<const float*>(g.W);
  for (int r0 = 0; r0 < R; r0 += 32) {
    const int r = r0 + tx;
    const int tap = r < R ? r / g.K : 0, k = r < R ? r % g.K : 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int px = ty + 8 * i;
      float av = 0.f;
      if (r < R && pw[px] >= 0) {
        long long c[5];
#pragma unroll
        for (int d = 0; d < 5; ++d) c[d] = taps.t[tap][d];
        c[0] += k;
        c[g.dimW] += pw[px];
        c[g.dimH] += ph[px];
        c[g.dimB] += pb[px];
        av = load_view(a, c);
      }
      As[px][tx] = av;
      const int n = n0 + px;  // reuse the same (row, col) pattern for the weight tile: row = output n
      float wv = 0.f;
      if (r < R && n < g.N) {
        // z = tap for weights; batch-indexed B operands (attention) are not supported on this path
        wv = W[(long long)tap * g.wstride_z + (long long)n * g.wstride_n + k];
      }
      Ws[px][tx] = wv;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      const float w = Ws[tx][kk];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(As[ty + 8 * i][kk], w, acc[i]);
    }
    __syncthreads();
  }
  const int n = n0 + tx;
  if (n < g.N) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int px = ty + 8 * i;
      if (pw[px] >= 0) epilogue_store(g, pb[px], (long long)ph[px] * g.Wo + pw[px], n, acc[i]);
    }
  }
}

// SIMT tap-GEMM for N <= 4 (the 1-channel output convs): one warp per pixel, lanes stride the channel dimension.
// The bounds test and the address of a tap are computed once per (pixel, tap); with VEC the lanes read 4 channels
// per load (uint2 of f16 or float4) and the weights as float4, so a 128-channel tap is one load instruction.
template <bool VEC>
__global__ void __launch_bounds__(256) gemm_simt_smalln_kernel(View a, GemmArgs g, Taps taps) {
  const int lane = threadIdx.x & 31;
  const long long p = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long npix = (long long)g.Wo * g.Ho * g.Bo;
  if (p >= npix) return;
  const int w = (int)(p % g.Wo), h = (int)((p / g.Wo) % g.Ho), b = (int)(p / ((long long)g.Wo * g.Ho));
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float* W = reinterpret_cast<const float*>(g.W);
  for (int tap = 0; tap < g.ntaps; ++tap) {
    long long c[5];
#pragma unroll
    for (int d = 0; d < 5; ++d) c[d] = taps.t[tap][d];
    c[g.dimW] += w; c[g.dimH] += h; c[g.dimB] += b;
    bool inb = true;
    long long off = 0;
#pragma unroll
    for (int d = 1; d < 5; ++d) {
      inb = inb && c[d] >= 0 && c[d] < a.dim[d];
      off += c[d] * a.stride[d];
    }
    if (!inb) continue;  // warp-uniform: the whole pixel row of this tap is padding
    const long long c0 = c[0];
    const float* wt = W + (long long)tap * g.wstride_z;
    if (VEC) {
      for (int k = lane * 4; k < g.K; k += 128) {
        const long long ch = c0 + k;  // c0 is a multiple of 4 and dim[0] too (checked on the host)
        if (ch < 0 || ch >= a.dim[0]) continue;
        float av[4];
        if (a.elem) {
          const uint2 raw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(a.p) + off + ch));
          const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
          const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
          av[0] = f0.x; av[1] = f0.y; av[2] = f1.x; av[3] = f1.y;
        } else {
          const float4 raw = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.p) + off + ch));
          av[0] = raw.x; av[1] = raw.y; av[2] = raw.z; av[3] = raw.w;
        }
        for (int n = 0; n < g.N; ++n) {
          const float4 wv = __ldg(reinterpret_cast<const float4*>(wt + (long long)n * g.wstride_n + k));
          acc[n] = fmaf(av[0], wv.x, acc[n]); acc[n] = fmaf(av[1], wv.y, acc[n]);
          acc[n] = fmaf(av[2], wv.z, acc[n]); acc[n] = fmaf(av[3], wv.w, acc[n]);
        }
      }
    } else {
      for (int k = lane; k < g.K; k += 32) {
        const long long ch = c0 + k;
        if (ch < 0 || ch >= a.dim[0]) continue;
        const float av = a.elem ? __half2float(reinterpret_cast<const __half*>(a.p)[off + ch]) : reinterpret_cast<const float*>(a.p)[off + ch];
        for (int n = 0; n < g.N; ++n) acc[n] = fmaf(av, wt[(long long)n * g.wstride_n + k], acc[n]);
      }
    }
  }
  for (int n = 0; n < g.N; ++n) {
    const float v = warp_sum(acc[n]);
    if (lane == 0) epilogue_store(g, b, (long long)h * g.Wo + w, n, v);
  }
}

// N <= 4 convolutions with K a multiple of 8 (vae.decoder.conv_out 128->1, vocoder.conv_post 48->1): 8 lanes per
// pixel, each lane owns 8-channel slices (one 16-byte f16 load or two float4 per slice), the packed weights
// [taps][N][K] sit in shared memory, partial sums meet through three xor-shuffles.
__global__ void __launch_bounds__(256) conv_smalln8_kernel(View a, GemmArgs g, Taps taps) {
  extern __shared__ float wsm[];
  const int WN = g.ntaps * g.N * g.K;
  const float* W = reinterpret_cast<const float*>(g.W);
  for (int i = threadIdx.x; i < WN; i += blockDim.x) {
    const int k = i % g.K, tn = i / g.K, n = tn % g.N, tap = tn / g.N;
    wsm[i] = W[(long long)tap * g.wstride_z + (long long)n * g.wstride_n + k];
  }
  __syncthreads();
  const int sub = threadIdx.x & 7;
  const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const long long npix = (long long)g.Wo * g.Ho * g.Bo;
  const bool live = p < npix;
  const long long pp = live ? p : 0;
  const int w = (int)(pp % g.Wo), h = (int)((pp / g.Wo) % g.Ho), b = (int)(pp / ((long long)g.Wo * g.Ho));
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int tap = 0; tap < g.ntaps; ++tap) {
    bool inb = live;
    long long off = 0;
#pragma unroll
    for (int d = 1; d < 5; ++d) {
      const long long c = taps.t[tap][d] + (d == g.dimW ? w : 0) + (d == g.dimH ? h : 0) + (d == g.dimB ? b : 0);
      inb = inb && c >= 0 && c < a.dim[d];
      off += c * a.stride[d];
    }
    if (!inb) continue;
    const int c0 = taps.t[tap][0];
    const float* wt = wsm + (size_t)tap * g.N * g.K;
    for (int k = sub * 8; k < g.K; k += 64) {
      const long long ch = c0 + k;
      if (ch < 0 || ch + 8 > a.dim[0]) continue;  // slices are 8-aligned (checked on the host)
      float x[8];
      if (a.elem) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(a.p) + off + ch));
        const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
        const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
        const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&raw.z));
        const float2 f3 = __half22float2(*reinterpret_cast<const __half2*>(&raw.w));
        x[0] = f0.x; x[1] = f0.y; x[2] = f1.x; x[3] = f1.y; x[4] = f2.x; x[5] = f2.y; x[6] = f3.x; x[7] = f3.y;
      } else {
        const float4 r0 = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.p) + off + ch));
        const float4 r1 = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.p) + off + ch + 4));
        x[0] = r0.x; x[1] = r0.y; x[2] = r0.z; x[3] = r0.w; x[4] = r1.x; x[5] = r1.y; x[6] = r1.z; x[7] = r1.w;
      }
      for (int n = 0; n < g.N; ++n) {
        const float4 w0 = *reinterpret_cast<const float4*>(wt + (size_t)n * g.K + k);
        const float4 w1 = *reinterpret_cast<const float4*>(wt + (size_t)n * g.K + k + 4);
        float s = acc[n];
        s = fmaf(x[0], w0.x, s); s = fmaf(x[1], w0.y, s); s = fmaf(x[2], w0.z, s); s = fmaf(x[3], w0.w, s);
        s = fmaf(x[4], w1.x, s); s = fmaf(x[5], w1.y, s); s = fmaf(x[6], w1.z, s); s = fmaf(x[7], w1.w, s);
        acc[n] = s;
      }
    }
  }
  for (int n = 0; n < g.N; ++n) {
    float v = acc[n];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    if (sub == 0 && live) epilogue_store(g, b, (long long)h * g.Wo + w, n, v);
  }
}

// One-output-channel convolutions over an f16 map (vae.decoder.conv_out 128 -> 1 3x3, vocoder.conv_post 48 -> 1 k=7),
// input-stationary: a CTA owns TW output pixels of one row.  Phase 1: a thread loads ONE input pixel (all K channels, once,
// into registers) and forms its dot product with every tap that reads that input row, into shared memory; phase 2: an
// output is the sum over taps of the dot products of the inputs it sees.  The output-stationary kernel above spends one
// bounds test + address computation + load per (pixel, tap, 8-channel slice) — ~60 instructions around 8 FMAs; here every
// input element is loaded and converted once per CTA and the FMAs dominate (98 -> ~12 us on the c2 pass per layer).
struct RowTaps {
  int ntaps, nrows, min_dw, span;   // span = max_dw - min_dw
  int dh[4];                        // input-row offsets (distinct values of the taps' H offset)
  signed char row_of[EGR_MAX_TAPS]; // tap -> index into dh
  short dw[EGR_MAX_TAPS];
};

template <int CH>   // 8-channel groups per register chunk (CH * 8 channels live at a time: 165 registers and one CTA per SM
                    // with all 128 channels of conv_out in registers, 63 and four CTAs with 32)
__global__ void __launch_bounds__(256) conv_n1_rows_kernel(View a, GemmArgs g, RowTaps rt, int TW) {
  egr_pdl_sync();
  extern __shared__ float wsm[];  // weights [ntaps][K], then dots [ntaps][TW + span]
  constexpr int KC = CH * 8;
  const int K = g.K, nchunks = K / KC;
  const int IW = TW + rt.span;
  float* dots = wsm + rt.ntaps * K;
  const float* W = reinterpret_cast<const float*>(g.W);
  for (int i = threadIdx.x; i < rt.ntaps * K; i += blockDim.x) {
    const int tap = i / K, k = i - tap * K;
    wsm[i] = W[(long long)tap * g.wstride_z + k];
  }
  __syncthreads();
  const int w0 = blockIdx.x * TW, h = blockIdx.y, b = blockIdx.z;
  for (int it = threadIdx.x; it < rt.nrows * IW; it += blockDim.x) {
    const int r = it / IW, j = it - r * IW;
    const long long wi = (long long)w0 + rt.min_dw + j, hi = (long long)h + rt.dh[r];
    const bool inb = wi >= 0 && wi < a.dim[g.dimW] && hi >= 0 && hi < a.dim[g.dimH];
    if (!inb) {   // zero padding
      for (int t = 0; t < rt.ntaps; ++t)
        if (rt.row_of[t] == r) dots[t * IW + j] = 0.f;
      continue;
    }
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(a.p) + wi * a.stride[g.dimW] +
                                                      hi * a.stride[g.dimH] + (long long)b * a.stride[g.dimB]);
    for (int c = 0; c < nchunks; ++c) {
      uint4 raw[CH];
#pragma unroll
      for (int i = 0; i < CH; ++i) raw[i] = __ldg(src + c * CH + i);
      float x[KC];
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&raw[i].x));
        const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&raw[i].y));
        const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&raw[i].z));
        const float2 f3 = __half22float2(*reinterpret_cast<const __half2*>(&raw[i].w));
        x[8 * i] = f0.x; x[8 * i + 1] = f0.y; x[8 * i + 2] = f1.x; x[8 * i + 3] = f1.y;
        x[8 * i + 4] = f2.x; x[8 * i + 5] = f2.y; x[8 * i + 6] = f3.x; x[8 * i + 7] = f3.y;
      }
      for (int t = 0; t < rt.ntaps; ++t) {
        if (rt.row_of[t] != r) continue;
        const float* wt = wsm + t * K + c * KC;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};   // four interleaved partial sums (k mod 4), added pairwise at the end
#pragma unroll
        for (int k = 0; k < KC; k += 4) {
          const float4 wv = *reinterpret_cast<const float4*>(wt + k);
          acc[0] = fmaf(x[k], wv.x, acc[0]); acc[1] = fmaf(x[k + 1], wv.y, acc[1]);
          acc[2] = fmaf(x[k + 2], wv.z, acc[2]); acc[3] = fmaf(x[k + 3], wv.w, acc[3]);
        }
        const float part = (acc[0] + acc[1]) + (acc[2] + acc[3]);
        float* d = dots + t * IW + j;          // chunk partials are added in chunk order (this thread owns the slot)
        *d = c == 0 ? part : *d + part;
      }
    }
  }
  __syncthreads();
  for (int jo = threadIdx.x; jo < TW; jo += blockDim.x) {
    const int w = w0 + jo;
    if (w >= g.Wo) break;
    float sum = 0.f;
    for (int t = 0; t < rt.ntaps; ++t) sum += dots[t * IW + jo + rt.dw[t] - rt.min_dw];   // tap order
    epilogue_store(g, b, (long long)h * g.Wo + w, 0, sum);
  }
}

// host side: does the op have the shape conv_n1_rows_kernel handles?  Fills rt / TW / K8.
static bool conv_n1_rows_match(const View& a, const GemmArgs& g, const Taps& taps, RowTaps* rt, int* TW) {
  if (g.N != 1 || a.elem != 1 || a.stride[0] != 1 || g.K != a.dim[0] || (g.K & 7) != 0 || g.K > 1024) return false;
  if (g.Wo != a.dim[g.dimW] || g.Ho != a.dim[g.dimH] || g.Bo != a.dim[g.dimB]) return false;   // stride-1, same size
  if (g.dimW == g.dimH || g.dimW == g.dimB || g.dimH == g.dimB) return false;
  if (((uintptr_t)a.p & 15) != 0) return false;
  for (int d = 1; d < 5; ++d)
    if ((a.stride[d] & 7) != 0) return false;   // 16-byte pixel rows
  if (g.Ho > 65535 || g.Bo > 65535) return false;
  rt->ntaps = g.ntaps; rt->nrows = 0;
  int lo = 1 << 30, hi = -(1 << 30);
  for (int t = 0; t < g.ntaps; ++t) {
    for (int d = 0; d < 5; ++d)
      if (d != g.dimW && d != g.dimH && taps.t[t][d] != 0) return false;   // channel / batch offsets: not this kernel
    const int dh = taps.t[t][g.dimH], dw = taps.t[t][g.dimW];
    int r = -1;
    for (int i = 0; i < rt->nrows; ++i)
      if (rt->dh[i] == dh) r = i;
    if (r < 0) {
      if (rt->nrows == 4) return false;
      r = rt->nrows++;
      rt->dh[r] = dh;
    }
    rt->row_of[t] = (signed char)r;
    rt->dw[t] = (short)dw;
    lo = dw < lo ? dw : lo; hi = dw > hi ? dw : hi;
  }
  rt->min_dw = lo; rt->span = hi - lo;
  if (rt->span > 64) return false;
  *TW = rt->nrows == 1 ? 256 - rt->span : 128;   // one row: exactly one 256-thread pass over the inputs
  return true;
}

// Convolutions with a tiny reduction (1-channel inputs: vae.encoder.conv_in, vocoder.wave_pre; taps*K <= 64): a CTA
// owns 64 pixels.  Phase 1 gathers the 64 x R input patch into shared memory (the tap bounds / address math runs
// once per patch value, not once per output); phase 2: 4 threads per pixel sweep the output channels as float4
// against the packed weights [R][N] in shared memory (4 threads x 16 B = 64 contiguous bytes per store).
#define SMK_PIX 64
__global__ void __launch_bounds__(256) conv_smallk_kernel(View a, GemmArgs g, Taps taps) {
  extern __shared__ float wsm[];  // [R][N] weights, then [SMK_PIX][R] patch
  const int R = g.ntaps * g.K;
  float* xs = wsm + (size_t)R * g.N;
  const float* W = reinterpret_cast<const float*>(g.W);
  for (int i = threadIdx.x; i < R * g.N; i += blockDim.x) {
    const int n = i % g.N, r = i / g.N, tap = r / g.K, k = r % g.K;
    wsm[i] = W[(long long)tap * g.wstride_z + (long long)n * g.wstride_n + k];
  }
  const long long npix = (long long)g.Wo * g.Ho * g.Bo;
  const long long p0 = (long long)blockIdx.x * SMK_PIX;
  for (int i = threadIdx.x; i < SMK_PIX * R; i += blockDim.x) {
    const int pl = i / R, r = i - pl * R, tap = r / g.K, k = r - tap * g.K;
    const long long p = p0 + pl;
    float x = 0.f;
    if (p < npix) {
      const int w = (int)(p % g.Wo), h = (int)((p / g.Wo) % g.Ho), b = (int)(p / ((long long)g.Wo * g.Ho));
      bool inb = true;
      long long off = 0;
#pragma unroll
      for (int d = 1; d < 5; ++d) {
        const long long c = taps.t[tap][d] + (d == g.dimW ? w : 0) + (d == g.dimH ? h : 0) + (d == g.dimB ? b : 0);
        inb = inb && c >= 0 && c < a.dim[d];
        off += c * a.stride[d];
      }
      const long long ch = taps.t[tap][0] + k;
      if (inb && ch >= 0 && ch < a.dim[0])
        x = a.elem ? __half2float(reinterpret_cast<const __half*>(a.p)[off + ch]) : reinterpret_cast<const float*>(a.p)[off + ch];
    }
    xs[i] = x;
  }
  __syncthreads();
  const int pl = threadIdx.x >> 2, q = threadIdx.x & 3;
  const long long p = p0 + pl;
  if (p >= npix) return;
  const int w = (int)(p % g.Wo), h = (int)((p / g.Wo) % g.Ho), b = (int)(p / ((long long)g.Wo * g.Ho));
  const long long pix = (long long)h * g.Wo + w;
  const float* xp = xs + (size_t)pl * R;
  for (int n = q * 4; n < g.N; n += 16) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < R; ++r) {
      const float x = xp[r];
      const float4 wv = *reinterpret_cast<const float4*>(wsm + (size_t)r * g.N + n);
      acc[0] = fmaf(x, wv.x, acc[0]); acc[1] = fmaf(x, wv.y, acc[1]); acc[2] = fmaf(x, wv.z, acc[2]); acc[3] = fmaf(x, wv.w, acc[3]);
    }
    epilogue_store4(g, b, pix, n, acc);
  }
}

// GEMV for a handful of rows (time-embedding MLP and the per-block embedding projections: M = 1): one warp per
// output feature, lanes stride K with float4 loads of the f32 weight row.  A must be f32 with unit channel stride.
__global__ void __launch_bounds__(256) gemm_simt_gemv_kernel(View a, GemmArgs g, int npix) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= g.N) return;
  const float* wrow = reinterpret_cast<const float*>(g.W) + (long long)n * g.wstride_n;
  for (int p = 0; p < npix; ++p) {
    const int w = p % g.Wo, h = (p / g.Wo) % g.Ho, b = p / (g.Wo * g.Ho);
    const float* x = reinterpret_cast<const float*>(a.p) + (long long)w * a.stride[g.dimW] + (long long)h * a.stride[g.dimH] +
                     (long long)b * a.stride[g.dimB];
    float acc = 0.f;
    if ((g.K & 3) == 0) {
      for (int k = lane * 4; k < g.K; k += 128) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + k));
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wrow + k));
        acc = fmaf(xv.x, wv.x, acc); acc = fmaf(xv.y, wv.y, acc); acc = fmaf(xv.z, wv.z, acc); acc = fmaf(xv.w, wv.w, acc);
      }
    } else {
      for (int k = lane; k < g.K; k += 32) acc = fmaf(x[k], wrow[k], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) epilogue_store(g, b, (long long)h * g.Wo + w, n, acc);
  }
}

int egr::launch_gemm_simt(const Spaces& s, const egr_op& op, cudaStream_t st) {
  GemmArgs g; Taps taps; View a;
  int rc = gemm_args_from_op(s, op, &g, &taps, &a);
  if (rc) return rc;
  if (g.wz_batch) return fail(EGR_ERR_UNSUPPORTED, "%s: batch-indexed B operand needs the tensor-core path", op.name);
  const long long npix = (long long)g.Wo * g.Ho * g.Bo;
  auto al16 = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
  bool zero_taps = true;
  for (int t = 0; t < g.ntaps; ++t)
    for (int d = 0; d < 5; ++d) zero_taps = zero_taps && taps.t[t][d] == 0;
  RowTaps rt;
  int TW = 0;
  if (npix <= 8 && g.ntaps == 1 && zero_taps && a.elem == 0 && a.stride[0] == 1 && g.K <= a.dim[0] && al16(a.p) && al16(g.W) &&
      (g.wstride_n & 3) == 0 && (a.stride[g.dimW] & 3) == 0 && (a.stride[g.dimH] & 3) == 0 && (a.stride[g.dimB] & 3) == 0) {
    gemm_simt_gemv_kernel<<<(unsigned)((g.N + 7) / 8), 256, 0, st>>>(a, g, (int)npix);
  } else if (conv_n1_rows_match(a, g, taps, &rt, &TW) && getenv("EGR_NO_CONV_N1_ROWS") == nullptr &&
             ((size_t)g.ntaps * g.K + (size_t)g.ntaps * (TW + rt.span)) * sizeof(float) <= 48 * 1024) {
    const dim3 grid((unsigned)((g.Wo + TW - 1) / TW), (unsigned)g.Ho, (unsigned)g.Bo);
    const size_t smem = ((size_t)g.ntaps * g.K + (size_t)g.ntaps * (TW + rt.span)) * sizeof(float);
    const int k8 = g.K / 8;   // channels per register chunk: the largest of 32 / 24 / 16 / 8 that divides K
    if (k8 % 4 == 0) conv_n1_rows_kernel<4><<<grid, 256, smem, st>>>(a, g, rt, TW);
    else if (k8 % 3 == 0) conv_n1_rows_kernel<3><<<grid, 256, smem, st>>>(a, g, rt, TW);
    else if (k8 % 2 == 0) conv_n1_rows_kernel<2><<<grid, 256, smem, st>>>(a, g, rt, TW);
    else conv_n1_rows_kernel<1><<<grid, 256, smem, st>>>(a, g, rt, TW);
  } else if (g.K <= 4 && g.ntaps * g.K <= 64 && (g.N & 3) == 0 && (size_t)g.ntaps * g.K * (g.N + SMK_PIX) * sizeof(float) <= 48 * 1024) {
    conv_smallk_kernel<<<(unsigned)((npix + SMK_PIX - 1) / SMK_PIX), 256, (size_t)g.ntaps * g.K * (g.N + SMK_PIX) * sizeof(float), st>>>(a, g, taps);
  } else if (g.N <= 4 && (g.K & 7) == 0 && (a.dim[0] & 7) == 0 && a.stride[0] == 1 && al16(a.p) &&
             (size_t)g.ntaps * g.N * g.K * sizeof(float) <= 48 * 1024 && [&] {
               bool ok = true;
               for (int d = 1; d < 5; ++d) ok = ok && (a.stride[d] & 7) == 0;
               for (int t = 0; t < g.ntaps; ++t) ok = ok && (taps.t[t][0] & 7) == 0;
               return ok;
             }()) {
    const long long threads = npix * 8;
    conv_smalln8_kernel<<<(unsigned)((threads + 255) / 256), 256, (size_t)g.ntaps * g.N * g.K * sizeof(float), st>>>(a, g, taps);
  } else if (g.N <= 4) {
    bool vec = (g.K & 3) == 0 && (a.dim[0] & 3) == 0 && al16(a.p) && al16(g.W) && (g.wstride_n & 3) == 0 && (g.wstride_z & 3) == 0;
    for (int d = 1; d < 5; ++d) vec = vec && (a.stride[d] & 3) == 0;
    for (int t = 0; t < g.ntaps; ++t) vec = vec && (taps.t[t][0] & 3) == 0;
    if (vec) gemm_simt_smalln_kernel<true><<<(unsigned)((npix + 7) / 8), 256, 0, st>>>(a, g, taps);
    else gemm_simt_smalln_kernel<false><<<(unsigned)((npix + 7) / 8), 256, 0, st>>>(a, g, taps);
  } else {
    dim3 grid((unsigned)((npix + 31) / 32), (unsigned)((g.N + 31) / 32));
    gemm_simt_kernel<<<grid, 256, 0, st>>>(a, g, taps);
  }
  EGR_CHECK_LAUNCH(op.name);
  return EGR_OK;
}

// ------------------------------------------------------------------------------------------------
// GroupNorm.  Input is a virtual channel-concat of x0 (C0 channels) and x1 (C1 channels), f32, channels
// innermost, pixels contiguous per batch item: element (b,p,c) at base + (b*P + p)*Cx + c.
// stats: f64 [B][G][2] = (sum, sumsq), must be zeroed first (EGR_OP_ZERO).
// ------------------------------------------------------------------------------------------------

// Deterministic (no atomics, fixed summation order) and independent of the batch size: block (slab, b) reduces its
// slab of pixels to per-group f64 partial sums; gn_finalize_kernel adds the slabs in index order.
// Thread mapping: QW threads across the channel quads of a pixel row (coalesced float4 loads), RP = 256 / QW row phases;
// a thread sums rows p_lo + rp, + RP, ... in ascending order, eight loads in flight.  The phases meet through shared
// memory after ONE barrier and are added in phase order.  (The first version put all eight warps on the same 32 quads and
// walked wider rows in serial 32-quad passes with eight barriers each: 22 us for a 16 MB map of 8192 x 512.)
__global__ void __launch_bounds__(256) gn_stats_kernel(CatArgs a, double* __restrict__ partials, int slab, int nslabs) {
  egr_pdl_sync();
  extern __shared__ double sh[];  // [RP][C][2]
  const int C = a.C0 + a.C1, Q = C >> 2, cpg = C / a.G;
  int QW = 1;
  while (QW * 2 <= Q && QW * 2 <= 256) QW *= 2;
  const int RP = 256 / QW;
  const int qi = threadIdx.x & (QW - 1), rp = threadIdx.x / QW;
  const int b = blockIdx.y;
  const long long p_lo = (long long)blockIdx.x * slab, p_hi = min(p_lo + slab, a.P);
  for (int q = qi; q < Q; q += QW) {
    float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
    const int c = q << 2;
    const float* base; int cc, Cx;
    if (c < a.C0) { base = a.x0; cc = c; Cx = a.C0; } else { base = a.x1; cc = c - a.C0; Cx = a.C1; }
    const float* rowp = base + ((long long)b * a.P + p_lo + rp) * Cx + cc;
    const long long rstride = (long long)RP * Cx;
    for (long long p = p_lo + rp; p < p_hi; p += 8 * RP, rowp += 8 * rstride) {
      float4 v[8];
      // unconditional loads (a row past the slab re-reads the first one and is zeroed afterwards, which leaves the sums
      // unchanged): predicated loads were scheduled two at a time by ptxas
#pragma unroll
      for (int k = 0; k < 8; ++k)
        v[k] = __ldg(reinterpret_cast<const float4*>(p + (long long)k * RP < p_hi ? rowp + k * rstride : rowp));
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (p + (long long)k * RP >= p_hi) v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        s[0] += v[k].x; ss[0] = fmaf(v[k].x, v[k].x, ss[0]);
        s[1] += v[k].y; ss[1] = fmaf(v[k].y, v[k].y, ss[1]);
        s[2] += v[k].z; ss[2] = fmaf(v[k].z, v[k].z, ss[2]);
        s[3] += v[k].w; ss[3] = fmaf(v[k].w, v[k].w, ss[3]);
      }
    }
    double* dst = sh + ((size_t)rp * C + c) * 2;
#pragma unroll
    for (int u = 0; u < 4; ++u) { dst[2 * u] = (double)s[u]; dst[2 * u + 1] = (double)ss[u]; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * a.G; i += blockDim.x) {
    const int gi = i >> 1, st = i & 1;
    double acc = 0.0;
    for (int r = 0; r < RP; ++r)
      for (int cc = 0; cc < cpg; ++cc) acc += sh[((size_t)r * C + gi * cpg + cc) * 2 + st];
    partials[(((long long)b * nslabs + blockIdx.x) * a.G + gi) * 2 + st] = acc;
  }
}

// stats[b][g][2] = sum over slabs in a fixed order (16 strided rows x 4 interleaved accumulators, then row order):
// 64 slab loads per thread would be one long L2-latency chain, so four independent chains run at once
__global__ void gn_finalize_kernel(const double* __restrict__ partials, double* __restrict__ stats, int G2, int nslabs) {
  egr_pdl_sync();
  extern __shared__ double shf64[];  // [16][G2]
  const int b = blockIdx.x, i = threadIdx.x, y = threadIdx.y;
  const double* p = partials + (long long)b * nslabs * G2 + i;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int sl = y;
  // sixteen loads in flight, added in the order of the four-load loop below (same bits): at ~1184 slabs the four-load
  // form was 18 dependent L2 round trips, 11 us for 600 KB
  for (; sl + 240 < nslabs; sl += 256) {
    double v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = p[(long long)(sl + 16 * k) * G2];
#pragma unroll
    for (int k = 0; k < 16; k += 4) { a0 += v[k]; a1 += v[k + 1]; a2 += v[k + 2]; a3 += v[k + 3]; }
  }
  for (; sl + 48 < nslabs; sl += 64) {
    a0 += p[(long long)sl * G2]; a1 += p[(long long)(sl + 16) * G2];
    a2 += p[(long long)(sl + 32) * G2]; a3 += p[(long long)(sl + 48) * G2];
  }
  for (; sl < nslabs; sl += 16) a0 += p[(long long)sl * G2];
  shf64[y * G2 + i] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (y == 0) {
    double t = 0.0;
    for (int k = 0; k < 16; ++k) t += shf64[k * G2 + i];
    stats[(long long)b * G2 + i] = t;
  }
}

template <bool CAT>   // CAT: the input is a virtual concat of two tensors (per-unit source select, 4 units in flight)
__global__ void __launch_bounds__(256) gn_apply_kernel(CatArgs a, const double* __restrict__ stats,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, int silu, float* __restrict__ out32,
                                                        __half* __restrict__ out16, int slab) {
  egr_pdl_sync();
  extern __shared__ float shf[];  // scale[C], shift[C]
  const int C = a.C0 + a.C1, C4 = C >> 2, cpg = C / a.G;
  const int b = blockIdx.y;
  float* scale = shf; float* shift = shf + C;
  const double cnt = (double)cpg * (double)a.P;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int gi = c / cpg;
    const double sum = stats[((long long)b * a.G + gi) * 2], sq = stats[((long long)b * a.G + gi) * 2 + 1];
    const double mean = sum / cnt;
    double var = sq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = rstd * gamma[c];
    scale[c] = sc;
    shift[c] = beta[c] - (float)mean * sc;
  }
  __syncthreads();
  const long long p_lo = (long long)blockIdx.x * slab, p_hi = min(p_lo + slab, a.P);
  if (p_lo >= p_hi) return;
  const int total = (int)(p_hi - p_lo) * C4;  // float4 units of this slab (slab <= a few thousand pixels)
  const long long row0 = (long long)b * a.P + p_lo;
  // (pixel, channel-quad) of unit i advance by a fixed step per iteration: one division per thread, none per element.
  const int step_p = (int)blockDim.x / C4, step_q = (int)blockDim.x - step_p * C4;
  int pl = (int)threadIdx.x / C4, q = (int)threadIdx.x - pl * C4;
  auto finish = [&](float4 v, int c, long long o) {
    const float4 sc = *reinterpret_cast<const float4*>(scale + c), sh = *reinterpret_cast<const float4*>(shift + c);
    float r[4] = {fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w)};
    if (silu) {
#pragma unroll
      for (int u = 0; u < 4; ++u) r[u] = egr_silu(r[u]);
    }
    if (out32) *reinterpret_cast<float4*>(out32 + o) = make_float4(r[0], r[1], r[2], r[3]);
    if (out16) {
      __half2 h0 = __floats2half2_rn(r[0], r[1]), h1 = __floats2half2_rn(r[2], r[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<unsigned*>(&h0);
      pk.y = *reinterpret_cast<unsigned*>(&h1);
      *reinterpret_cast<uint2*>(out16 + o) = pk;
    }
  };
  if (!CAT) {
    // one tensor: the slab is one contiguous run of float4 units, unit i at base + 4 i; only the channel quad needs
    // tracking.  Eight loads per thread are issued before the first use (one load in flight per thread left this kernel
    // at ~40 % of the HBM rate on the 131072 x 128 maps, four at ~55 %).
    const float4* src = reinterpret_cast<const float4*>(a.x0 + row0 * C);
    const long long obase = row0 * C;
    for (int i = threadIdx.x; i < total; i += 8 * blockDim.x) {
      float4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int u = i + k * (int)blockDim.x;
        v[k] = __ldg(src + (u < total ? u : i));   // unconditional: ptxas keeps all eight in flight
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int u = i + k * (int)blockDim.x;
        if (u < total) finish(v[k], q << 2, obase + 4ll * u);
        q += step_q;
        if (q >= C4) q -= C4;
      }
    }
    return;
  }
  for (int i = threadIdx.x; i < total; i += 4 * blockDim.x) {
    float4 v[4];
    int cq[4];
    long long pp[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      cq[k] = q << 2;
      pp[k] = row0 + pl;
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i + k * (int)blockDim.x < total) {
        const int c = cq[k];
        const float* src = c < a.C0 ? a.x0 + pp[k] * a.C0 + c : a.x1 + pp[k] * a.C1 + (c - a.C0);
        v[k] = __ldg(reinterpret_cast<const float4*>(src));
      }
      pl += step_p; q += step_q;
      if (q >= C4) { q -= C4; ++pl; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i + k * (int)blockDim.x < total) finish(v[k], cq[k], pp[k] * C + cq[k]);
    }
  }
}

// Small feature maps (the UNet, the VAE mid block): ONE kernel, one CTA per (group, batch item).  The group's
// P x cpg elements (L2-resident: the producing GEMM has just written them) are read twice — f32 partials per
// thread combined in f64 in a fixed order (deterministic, independent of the batch size), then normalise (+SiLU)
// and store.  Replaces the stats + finalize + apply launches whose grids of 1..32 CTAs were pure latency.
#define GN_FUSED_THREADS 512
// `ncl` CTAs (one thread-block cluster) share a (group, item): each takes a contiguous range of pixels, the partial
// moments are exchanged through distributed shared memory and combined in rank order by every CTA (same bits
// everywhere, independent of the batch size).  ncl is a function of the op's geometry only (gn_cluster_size).
__global__ void __launch_bounds__(GN_FUSED_THREADS, 2) gn_fused_kernel(CatArgs a, const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta, float eps, int silu,
                                                                    float* __restrict__ out32, __half* __restrict__ out16,
                                                                    int ncl) {
  __shared__ double red[2][GN_FUSED_THREADS / 32];
  __shared__ double part[2];
  __shared__ float s_mean, s_rstd;
  const int C = a.C0 + a.C1, cpg = C / a.G, q4 = cpg >> 2;
  const int gi = blockIdx.x / ncl, rank = blockIdx.x - gi * ncl, b = blockIdx.y;
  const int c_lo = gi * cpg;
  const int p_lo = (int)((long long)a.P * rank / ncl), p_hi = (int)((long long)a.P * (rank + 1) / ncl);
  const int units = (p_hi - p_lo) * q4;  // float4 units of this CTA's share of the group (P <= 4096)
  auto src = [&](long long p, int c) -> const float* {
    return c < a.C0 ? a.x0 + ((long long)b * a.P + p) * a.C0 + c : a.x1 + ((long long)b * a.P + p) * a.C1 + (c - a.C0);
  };
  // (pixel, channel quad) of unit u advance by a fixed step per thread: one division per thread, none per unit.  Four units
  // per iteration, their loads issued together (a unit past the end re-reads the first one and is zeroed: the sums do not
  // change) and added in ascending u — the same additions in the same order as a one-unit loop, which left one load in
  // flight per thread: eight dependent round trips per pass, 92 us for the 2048-pixel x 1024-channel maps of a batch of 8.
  const int step_p = GN_FUSED_THREADS / q4, step_q = GN_FUSED_THREADS - step_p * q4;
  const int pl0 = (int)threadIdx.x / q4, qq0 = (int)threadIdx.x - pl0 * q4;
  float s = 0.f, ss = 0.f;
  {
    int pl = pl0, qq = qq0;
    for (int u = threadIdx.x; u < units; u += 4 * GN_FUSED_THREADS) {
      float4 v[4];
      const float* first = src(p_lo + pl, c_lo + qq * 4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[k] = __ldg(reinterpret_cast<const float4*>(u + k * GN_FUSED_THREADS < units ? src(p_lo + pl, c_lo + qq * 4) : first));
        pl += step_p; qq += step_q;
        if (qq >= q4) { qq -= q4; ++pl; }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (u + k * GN_FUSED_THREADS >= units) v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
        ss = fmaf(v[k].x, v[k].x, ss); ss = fmaf(v[k].y, v[k].y, ss); ss = fmaf(v[k].z, v[k].z, ss); ss = fmaf(v[k].w, v[k].w, ss);
      }
    }
  }
  double ds = warp_sum((double)s), dss = warp_sum((double)ss);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ds; red[1][threadIdx.x >> 5] = dss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0, tt = 0.0;
    for (int w = 0; w < GN_FUSED_THREADS / 32; ++w) { t += red[0][w]; tt += red[1][w]; }
    part[0] = t; part[1] = tt;
  }
  if (ncl > 1) cooperative_groups::this_cluster().sync();
  if (threadIdx.x == 0) {
    double t = part[0], tt = part[1];
    if (ncl > 1) {
      cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
      t = 0.0; tt = 0.0;
      for (int r = 0; r < ncl; ++r) {
        const double* rp = cluster.map_shared_rank(&part[0], r);
        t += rp[0]; tt += rp[1];
      }
    }
    const double cnt = (double)cpg * (double)a.P;
    const double mean = t / cnt;
    double var = tt / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean = (float)mean;
    s_rstd = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const float mean = s_mean, rstd = s_rstd;
  {
    int pl = pl0, qq = qq0;
    for (int u = threadIdx.x; u < units; u += 4 * GN_FUSED_THREADS) {
      float4 v[4];
      int pk_[4], ck_[4];
      const float* first = src(p_lo + pl, c_lo + qq * 4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        pk_[k] = p_lo + pl; ck_[k] = c_lo + qq * 4;
        v[k] = __ldg(reinterpret_cast<const float4*>(u + k * GN_FUSED_THREADS < units ? src(pk_[k], ck_[k]) : first));
        pl += step_p; qq += step_q;
        if (qq >= q4) { qq -= q4; ++pl; }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (u + k * GN_FUSED_THREADS >= units) continue;
        const int p = pk_[k], c = ck_[k];
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + c));
        float r[4] = {fmaf(v[k].x, rstd * g4.x, b4.x - mean * (rstd * g4.x)), fmaf(v[k].y, rstd * g4.y, b4.y - mean * (rstd * g4.y)),
                      fmaf(v[k].z, rstd * g4.z, b4.z - mean * (rstd * g4.z)), fmaf(v[k].w, rstd * g4.w, b4.w - mean * (rstd * g4.w))};
        if (silu) {
#pragma unroll
          for (int j = 0; j < 4; ++j) r[j] = egr_silu(r[j]);
        }
        const long long o = ((long long)b * a.P + p) * C + c;
        if (out32) *reinterpret_cast<float4*>(out32 + o) = make_float4(r[0], r[1], r[2], r[3]);
        if (out16) {
          __half2 h0 = __floats2half2_rn(r[0], r[1]), h1 = __floats2half2_rn(r[2], r[3]);
          uint2 pk;
          pk.x = *reinterpret_cast<unsigned*>(&h0);
          pk.y = *reinterpret_cast<unsigned*>(&h1);
          *reinterpret_cast<uint2*>(out16 + o) = pk;
        }
      }
    }
  }
  if (ncl > 1) cooperative_groups::this_cluster().sync();  // keep `part` alive until every CTA of the cluster has read it
}

// CTAs per (group, item): at least 8192 elements per CTA, at most 8 (portable cluster size)
static int gn_cluster_size(const CatArgs& a) {
  const long long elems = (long long)a.P * ((a.C0 + a.C1) / a.G);
  int ncl = 1;
  while (ncl < 8 && elems / (ncl * 2) >= 8192) ncl *= 2;
  if (getenv("EGR_GN_NO_CLUSTER")) ncl = 1;
  return ncl;
}

// single-kernel path: small maps whose groups are whole float4 columns (a function of the op's geometry only)
static bool gn_use_fused(const CatArgs& a) {
  const int C = a.C0 + a.C1, cpg = C / a.G;
  return a.P <= 4096 && (cpg & 3) == 0 && getenv("EGR_GN_NO_FUSED") == nullptr;
}

static int slab_for(long long P, int* nslabs) {
  // ~592 slabs for the largest maps (one wave of 4 CTAs per SM at batch 1; the finalize pass reads every slab's partials
  // in one CTA per item, so fewer, longer slabs are cheaper); a function of P only, so results do not depend on B
  long long slab = (P + 591) / 592;
  if (slab < 16) slab = 16;
  if (slab > P) slab = P;
  *nslabs = (int)((P + slab - 1) / slab);
  return (int)slab;
}

int egr::launch_gn_stats(const Spaces& s, const egr_op& op, cudaStream_t st) {
  CatArgs a;
  int rc = cat_args(s, op, &a);
  if (rc) return rc;
  double* stats = (double*)resolve(s, op.ptr[EGR_P_STATS]);
  if (!stats) return fail(EGR_ERR_ARG, "%s: null stats", op.name);
  if (gn_use_fused(a)) return EGR_OK;  // the apply op computes the moments itself
  int ns; int slab = slab_for(a.P, &ns);
  if (op.i[EGR_I_AUX1] != ns) return fail(EGR_ERR_ARG, "%s: plan was built for %lld slabs, kernel wants %d", op.name, (long long)op.i[EGR_I_AUX1], ns);
  const int C = a.C0 + a.C1, G2 = 2 * a.G;
  double* partials = stats + (long long)a.B * G2;
  int QW = 1;
  while (QW * 2 <= (C >> 2) && QW * 2 <= 256) QW *= 2;
  const size_t smem = (size_t)(256 / QW) * C * 2 * sizeof(double);   // [RP][C][2] as the kernel lays it out (< 32 KB up to C = 1024)
  if (smem > 48 * 1024) return fail(EGR_ERR_UNSUPPORTED, "%s: C=%d too wide for the GroupNorm reduction", op.name, C);
  if (G2 > 64) return fail(EGR_ERR_UNSUPPORTED, "%s: at most 32 groups", op.name);
  if (egr::pdl_enabled()) {
#ifdef __CUDACC__
    EGR_CUDA(egr::launch_pdl(gn_stats_kernel, dim3(ns, a.B), dim3(256), smem, st, a, partials, slab, ns));
#endif
  } else gn_stats_kernel<<<dim3(ns, a.B), 256, smem, st>>>(a, partials, slab, ns);
  EGR_CHECK_LAUNCH(op.name);
  if (egr::pdl_enabled()) {
#ifdef __CUDACC__
    EGR_CUDA(egr::launch_pdl(gn_finalize_kernel, dim3(a.B), dim3(G2, 16), (size_t)16 * G2 * sizeof(double), st, partials, stats, G2, ns));
#endif
  } else gn_finalize_kernel<<<a.B, dim3(G2, 16), (size_t)16 * G2 * sizeof(double), st>>>(partials, stats, G2, ns);
  EGR_CHECK_LAUNCH(op.name);
  return EGR_OK;
}

int egr::launch_gn_apply(const Spaces& s, const egr_op& op, cudaStream_t st) {
  CatArgs a;
  int rc = cat_args(s, op, &a);
  if (rc) return rc;
  const double* stats = (const double*)resolve(s, op.ptr[EGR_P_STATS]);
  const float* gamma = (const float*)resolve(s, op.ptr[EGR_P_GAMMA]);
  const float* beta = (const float*)resolve(s, op.ptr[EGR_P_BETA]);
  float* o32 = (float*)resolve(s, op.ptr[EGR_P_OUT32]);
  __half* o16 = (__half*)resolve(s, op.ptr[EGR_P_OUT16]);
  if (!stats || !gamma || !beta || (!o32 && !o16)) return fail(EGR_ERR_ARG, "%s: null pointer", op.name);
  if (gn_use_fused(a)) {
    const int ncl = gn_cluster_size(a);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(a.G * ncl, a.B);
    cfg.blockDim = dim3(GN_FUSED_THREADS);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = ncl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    EGR_CUDA(cudaLaunchKernelEx(&cfg, gn_fused_kernel, a, gamma, beta, (float)op.f[EGR_F_EPS], (int)op.i[EGR_I_MODE], o32, o16, ncl));
    EGR_CHECK_LAUNCH(op.name);
    return EGR_OK;
  }
  // The apply pass is element-wise, so its partition is free (the statistics keep theirs: it fixes their summation
  // order): one wave of row blocks per batch item, sized from the kernel's real occupancy.
  const int C = a.C0 + a.C1;
  const bool cat = a.C1 > 0;
  static int resident[2] = {0, 0};
  if (!resident[cat]) {
    int per_sm = 0;
    if (cat) EGR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gn_apply_kernel<true>, 256, 2 * C * sizeof(float)));
    else EGR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gn_apply_kernel<false>, 256, 2 * C * sizeof(float)));
    resident[cat] = (per_sm > 0 ? per_sm : 4) * (devinfo().sm_count ? devinfo().sm_count : 148);
  }
  long long slab = (a.P + resident[cat] - 1) / resident[cat];
  if (slab < 16) slab = 16;
  if (slab > a.P) slab = a.P;
  const int ns = (int)((a.P + slab - 1) / slab);
  if (egr::pdl_enabled()) {
#ifdef __CUDACC__
    if (cat) EGR_CUDA(egr::launch_pdl(gn_apply_kernel<true>, dim3(ns, a.B), dim3(256), 2 * C * sizeof(float), st, a, stats, gamma, beta,
                                      (float)op.f[EGR_F_EPS], (int)op.i[EGR_I_MODE], o32, o16, (int)slab));
    else EGR_CUDA(egr::launch_pdl(gn_apply_kernel<false>, dim3(ns, a.B), dim3(256), 2 * C * sizeof(float), st, a, stats, gamma, beta,
                                  (float)op.f[EGR_F_EPS], (int)op.i[EGR_I_MODE], o32, o16, (int)slab));
#endif
  } else if (cat)
    gn_apply_kernel<true><<<dim3(ns, a.B), 256, 2 * C * sizeof(float), st>>>(a, stats, gamma, beta, (float)op.f[EGR_F_EPS],
                                                                             (int)op.i[EGR_I_MODE], o32, o16, (int)slab);
  else
    gn_apply_kernel<false><<<dim3(ns, a.B), 256, 2 * C * sizeof(float), st>>>(a, stats, gamma, beta, (float)op.f[EGR_F_EPS],
                                                                              (int)op.i[EGR_I_MODE], o32, o16, (int)slab);
  EGR_CHECK_LAUNCH(op.name);
  return EGR_OK;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the channel dimension: one warp per token row, f32 in -> f16 out.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, long long rows, int C,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         float eps, __half* __restrict__ out16, float* __restrict__ out32) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
  const float mean = warp_sum(s) / C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) { const float d = xr[c] - mean; v = fmaf(d, d, v); }
  const float rstd = rsqrtf(warp_sum(v) / C + eps);
  for (int c = lane; c < C; c += 32) {
    const float y = (xr[c] - mean) * rstd * gamma[c] + beta[c];
    if (out16) out16[row * C + c] = __float2half_rn(y);
    if (out32) out32[row * C + c] = y;
  }
}

int egr::launch_layernorm(const Spaces& s, const egr_op& op, cudaStream_t st) {
  const float* x = (const float*)resolve(s, op.x0.addr);
  const long long rows = op.i[EGR_I_ROWS];
  const int C = (int)op.i[EGR_I_COLS];
  const float* gamma = (const float*)resolve(s, op.ptr[EGR_P_GAMMA]);
  const float* beta = (const float*)resolve(s, op.ptr[EGR_P_BETA]);
  __half* o16 = (__half*)resolve(s, op.ptr[EGR_P_OUT16]);
  float* o32 = (float*)resolve(s, op.ptr[EGR_P_OUT32]);
  if (!x || !gamma || !beta || (!o16 && !o32) || rows <= 0 || C <= 0) return fail(EGR_ERR_ARG, "%s: bad arguments", op.name);
  layernorm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, rows, C, gamma, beta, (float)op.f[EGR_F_EPS], o16, o32);
  EGR_CHECK_LAUNCH(op.name);
  return EGR_OK;
}

// ------------------------------------------------------------------------------------------------
// row softmax(scale * x): f32 [rows, cols] -> f16, one warp per row (cols up to a few thousand).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_kernel(const float* __restrict__ x, long long rows, int cols, float scale,
                                                       __half* __restrict__ out) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + row * cols;
  float m = -INFINITY;
  for (int c = lane; c < cols; c += 32) m = fmaxf(m, xr[c] * scale);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += __expf(xr[c] * scale - m);
  const float inv = 1.0f / warp_sum(s);
  for (int c = lane; c < cols; c += 32) out[row * cols + c] = __float2half_rn(__expf(xr[c] * scale - m) * inv);
}

int egr::launch_softmax(const Spaces& s, const egr_op& op, cudaStream_t st) {
  const float* x = (const float*)resolve(s, op.x0.addr);
  __half* o = (__half*)resolve(s, op.ptr[EGR_P_OUT16]);
  const long long rows = op.i[EGR_I_ROWS];
  const int cols = (int)op.i[EGR_I_COLS];
  if (!x || !o || rows <= 0 || cols <= 0) return fail(EGR_ERR_ARG, "%s: bad arguments", op.name);
  softmax_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, rows, cols, (float)op.f[EGR_F_ALPHA], o);
  EGR_CHECK_LAUNCH(op.name);
  return EGR_OK;
}

// ------------------------------------------------------------------------------------------------
// Short-sequence multi-head self-attention (UNet transformers: S <= 512, head_dim <= 64).
// q,k,v f16 [B,S,heads*hd]; K and V of one (b, head) are staged in shared memory, each warp owns query
// rows: lanes split the keys for QK^T + softmax, then split head_dim for PV.  f32 math throughout.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_small_kernel(const __half* __restrict__ q, const __half* __restrict__ k,
                                                          const __half* __restrict__ v, __half* __restrict__ out, int S,
                                                          int heads, int hd, float scale) {
  extern __shared__ unsigned char smraw[];
  const int C = heads * hd;
  const int kst = hd + 2;  // padded row stride (halves) -> conflict-free column walks
  __half* Ks = reinterpret_cast<__half*>(smraw);
  __half* Vs = Ks + (size_t)S * kst;
  float* probs = reinterpret_cast<float*>(Vs + (size_t)S * kst);  // [8 warps][S]
  float* qrow = probs + 8 * S;                                     // [8 warps][hd]
  const int b = blockIdx.z, h = blockIdx.y;
  const long long base = (long long)b * S * C + (long long)h * hd;
  for (int i = threadIdx.x; i < S * hd; i += blockDim.x) {
    const int j = i / hd, d = i % hd;
    Ks[j * kst + d] = k[base + (long long)j * C + d];
    Vs[j * kst + d] = v[base + (long long)j * C + d];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* pr = probs + warp * S;
  float* qr = qrow + warp * hd;
  const int rows_per_block = (S + gridDim.x - 1) / gridDim.x;
  const int r_lo = blockIdx.x * rows_per_block, r_hi = min(S, r_lo + rows_per_block);
  for (int r = r_lo + warp; r < r_hi; r += 8) {
    for (int d = lane; d < hd; d += 32) qr[d] = __half2float(q[base + (long long)r * C + d]) * scale;
    __syncwarp();
    float m = -INFINITY;
    for (int j = lane; j < S; j += 32) {
      float dot = 0.f;
      for (int d = 0; d < hd; ++d) dot = fmaf(qr[d], __half2float(Ks[j * kst + d]), dot);
      pr[j] = dot;
      m = fmaxf(m, dot);
    }
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) { const float e = __expf(pr[j] - m); pr[j] = e; sum += e; }
    const float inv = 1.0f / warp_sum(sum);
    __syncwarp();
    for (int d = lane; d < hd; d += 32) {
      float o = 0.f;
      for (int j = 0; j < S; ++j) o = fmaf(pr[j], __half2float(Vs[j * kst + d]), o);
      out[base + (long long)r * C + d] = __float2half_rn(o * inv);
    }
    __syncwarp();
  }
}

// Same contract, head_dim 16 / 32 and S <= 512: K and V of one (b, head) are staged as f16 with a 16-byte aligned,
// conflict-free row stride; one warp owns a query row at a time and its LANES SPLIT THE KEYS in both phases: a lane
// holds the scaled query in registers, reads whole K / V rows with 16-byte loads (32 FMAs per 4 loads instead of one
// FMA per 2-byte load), keeps its scores in registers (no probability buffer), and the per-lane partial outputs are
// combined by a reduce-scatter of 31 shuffles.
// STRIDED: q, k, v are column blocks of one wider [B*S, ld] tensor (the fused q/k/v projection, EGR_FUSE_QKV); the
// output is always dense [B*S, C].
template <int HD, bool STRIDED>
__global__ void __launch_bounds__(256) attn_rows_kernel(const __half* __restrict__ q, const __half* __restrict__ k,
                                                        const __half* __restrict__ v, __half* __restrict__ out, int S,
                                                        int heads, float scale, int ld) {
  extern __shared__ unsigned char smraw[];
  constexpr int KST = HD + 8;  // halves per staged row
  __half* Ks = reinterpret_cast<__half*>(smraw);
  __half* Vs = Ks + (size_t)S * KST;
  const int C = heads * HD;
  const int b = blockIdx.z, h = blockIdx.y;
  const long long base = (long long)b * S * C + (long long)h * HD;
  const int L = STRIDED ? ld : C;                                                        // input row stride
  const long long ibase = STRIDED ? (long long)b * S * L + (long long)h * HD : base;   // input offset of (b, head)
  for (int i = threadIdx.x; i < S * (HD / 8); i += blockDim.x) {
    const int j = i / (HD / 8), c8 = (i % (HD / 8)) * 8;
    *reinterpret_cast<uint4*>(Ks + j * KST + c8) = __ldg(reinterpret_cast<const uint4*>(k + ibase + (long long)j * L + c8));
    *reinterpret_cast<uint4*>(Vs + j * KST + c8) = __ldg(reinterpret_cast<const uint4*>(v + ibase + (long long)j * L + c8));
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_per_block = (S + gridDim.x - 1) / gridDim.x;
  const int r_lo = blockIdx.x * rows_per_block, r_hi = min(S, r_lo + rows_per_block);
  const int njj = (S + 31) >> 5;  // keys per lane (<= 16)
  auto row8 = [](const __half* p, float* x) {
    const uint4 raw = *reinterpret_cast<const uint4*>(p);
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
    const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
    const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&raw.z));
    const float2 f3 = __half22float2(*reinterpret_cast<const __half2*>(&raw.w));
    x[0] = f0.x; x[1] = f0.y; x[2] = f1.x; x[3] = f1.y; x[4] = f2.x; x[5] = f2.y; x[6] = f3.x; x[7] = f3.y;
  };
  for (int r = r_lo + warp; r < r_hi; r += 8) {
    float qv[HD];
#pragma unroll
    for (int c8 = 0; c8 < HD; c8 += 8) {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(q + ibase + (long long)r * L + c8));
      const __half2* hp = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 f = __half22float2(hp[u]);
        qv[c8 + 2 * u] = f.x * scale; qv[c8 + 2 * u + 1] = f.y * scale;
      }
    }
    float sc[16];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
      sc[jj] = -INFINITY;
      const int j = jj * 32 + lane;
      if (jj < njj && j < S) {
        float dot = 0.f;
#pragma unroll
        for (int c8 = 0; c8 < HD; c8 += 8) {
          float x[8];
          row8(Ks + j * KST + c8, x);
#pragma unroll
          for (int u = 0; u < 8; ++u) dot = fmaf(qv[c8 + u], x[u], dot);
        }
        sc[jj] = dot;
        mx = fmaxf(mx, dot);
      }
    }
    mx = warp_max(mx);
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
      const int j = jj * 32 + lane;
      if (jj < njj && j < S) {
        const float p = __expf(sc[jj] - mx);
        sum += p;
#pragma unroll
        for (int c8 = 0; c8 < HD; c8 += 8) {
          float x[8];
          row8(Vs + j * KST + c8, x);
#pragma unroll
          for (int u = 0; u < 8; ++u) o[c8 + u] = fmaf(p, x[u], o[c8 + u]);
        }
      }
    }
    const float inv = 1.0f / warp_sum(sum);
    // reduce-scatter: after the step with mask m a lane keeps the half of its values selected by (lane & m)
    if (HD == 32) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float mine = (lane & 16) ? o[16 + i] : o[i], send = (lane & 16) ? o[i] : o[16 + i];
        o[i] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
      }
    } else {  // HD == 16: fold the two half-warps first, every lane keeps all 16 values
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] += __shfl_xor_sync(0xffffffffu, o[i], 16);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float mine = (lane & 8) ? o[8 + i] : o[i], send = (lane & 8) ? o[i] : o[8 + i];
      o[i] = mine + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float mine = (lane & 4) ? o[4 + i] : o[i], send = (lane & 4) ? o[i] : o[4 + i];
      o[i] = mine + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float mine = (lane & 2) ? o[2 + i] : o[i], send = (lane & 2) ? o[i] : o[2 + i];
      o[i] = mine + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    {
      const float mine = (lane & 1) ? o[1] : o[0], send = (lane & 1) ? o[0] : o[1];
      o[0] = mine + __shfl_xor_sync(0xffffffffu, send, 1);
    }
    // lane now holds output dim d = bits of lane: for HD 32 d = lane; for HD 16 d = lane & 15 (both half-warps equal)
    const int d = HD == 32 ? lane : (lane & 15);
    if (HD == 32 || lane < 16) out[base + (long long)r * C + d] = __float2half_rn(o[0] * inv);
  }
}

int egr::launch_attn_small(const Spaces& s, const egr_op& op, cudaStream_t st) {
  const __half* q = (const __half*)resolve(s, op.x0.addr);
  const __half* k = (const __half*)resolve(s, op.x1.addr);
  const __half* v = (const __half*)resolve(s, op.ptr[EGR_P_AUX]);
  __half* o = (__half*)resolve(s, op.ptr[EGR_P_OUT16]);
  const int S = (int)op.i[EGR_I_SEQ], heads = (int)op.i[EGR_I_HEADS], hd = (int)op.i[EGR_I_HEADDIM], B = (int)op.i[EGR_I_BATCH];
  if (!q || !k || !v || !o || S <= 0 || heads <= 0 || hd <= 0 || B <= 0) return fail(EGR_ERR_ARG, "%s: bad arguments", op.name);
  const int C_ = heads * hd;
  auto al16 = [](const void* p_) { return reinterpret_cast<uintptr_t>(p_) % 16 == 0; };
  const int ld = (int)op.i[EGR_I_AUX0];   // 0: dense rows of C halves; otherwise the row stride of a fused q/k/v tensor
  if (ld != 0 && (ld < C_ || ld % 8 != 0)) return fail(EGR_ERR_ARG, "%s: bad q/k/v row stride %d", op.name, ld);
  if ((hd == 32 || hd == 16) && S <= 512 && C_ % 8 == 0 && al16(q) && al16(k) && al16(v) && getenv("EGR_ATTN_OLD") == nullptr) {
    const size_t sm = (size_t)2 * S * (hd + 8) * sizeof(__half);
    static bool attr2 = false;
    if (!attr2) {
      EGR_CUDA(cudaFuncSetAttribute(attn_rows_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      EGR_CUDA(cudaFuncSetAttribute(attn_rows_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      EGR_CUDA(cudaFuncSetAttribute(attn_rows_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      EGR_CUDA(cudaFuncSetAttribute(attn_rows_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      attr2 = true;
    }
    // 16 query rows per block (2 per warp): K/V re-staged per block from L2, many blocks for the short sequences
    const int qblocks = (S + 15) / 16;
    const float sc_ = (float)op.f[EGR_F_ALPHA];
    if (ld == 0) {
      if (hd == 32) attn_rows_kernel<32, false><<<dim3(qblocks, heads, B), 256, sm, st>>>(q, k, v, o, S, heads, sc_, 0);
      else attn_rows_kernel<16, false><<<dim3(qblocks, heads, B), 256, sm, st>>>(q, k, v, o, S, heads, sc_, 0);
    } else {
      if (hd == 32) attn_rows_kernel<32, true><<<dim3(qblocks, heads, B), 256, sm, st>>>(q, k, v, o, S, heads, sc_, ld);
      else attn_rows_kernel<16, true><<<dim3(qblocks, heads, B), 256, sm, st>>>(q, k, v, o, S, heads, sc_, ld);
    }
    EGR_CHECK_LAUNCH(op.name);
    return EGR_OK;
  }
  if (ld != 0) return fail(EGR_ERR_UNSUPPORTED, "%s: strided q/k/v needs the 16/32-dim head kernel", op.name);
  size_t smem = (size_t)2 * S * (hd + 2) * sizeof(__half) + (size_t)8 * S * sizeof(float) + 8 * hd * sizeof(float);
  if (smem > 200 * 1024) return fail(EGR_ERR_UNSUPPORTED, "%s: S=%d hd=%d needs %zu B smem; use the GEMM attention path", op.name, S, hd, smem);
  static bool attr_done = false;
  if (!attr_done) {
    EGR_CUDA(cudaFuncSetAttribute(attn_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  int qblocks = (S + 63) / 64;  // 64 query rows per block: K/V re-staged per block, fine for S <= 512
  attn_small_kernel<<<dim3(qblocks, heads, B), 256, smem, st>>>(q, k, v, o, S, heads, hd, (float)op.f[EGR_F_ALPHA]);
  EGR_CHECK_LAUNCH(op.name);
  return EGR_OK;
}

// ------------------------------------------------------------------------------------------------
// GEGLU: x f32 [rows, 2D] -> f16 [rows, D] = x[:, :D] * gelu(x[:, D:])   (exact erf gelu, torch default)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) geglu_kernel(const float* __restrict__ x, long long rows, int D, __half* __restrict__ out) {
  const long long total = rows * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / D; const int d = (int)(i % D);
    const float a = x[r * 2 * D + d], g = x[r * 2 * D + D + d];
    out[i] = __float2half_rn(a * (0.5f * g * (1.0f + erff(g * 0.70710678118654752f))));
  }
}

static unsigned grid1d(long long n, int per_block = 256) {
  int sms = devinfo().sm_count ? devinfo().sm_count : 148;
  long long b = (n + per_block - 1) / per_block;
  if (b > (long long)sms * 32) b = (long long)sms * 32;
  return (unsigned)(b < 1 ? 1 : b);
}

int egr::launch_geglu(const Spaces& s, const egr_op& op, cudaStream_t st) {
  const float* x = (const float*)resolve(s, op.x0.addr);
  __half* o = (__half*)resolve(s, op.ptr[EGR_P_OUT16]);
  const long long rows = op.i[EGR_I_ROWS]; const int D = (int)op.i[EGR_I_COLS];
  if (!x || !o || rows <= 0 || D <= 0) return fail(EGR_ERR_ARG, "%s: bad arguments", op.name);
  geglu_kernel<<<grid1d(rows * D), 256, 0, st>>>(x, rows, D, o);
  EGR_CHECK_LAUNCH(op.name);
  return EGR_OK;
}

// ------------------------------------------------------------------------------------------------
// element-wise glue.  All operate on channels-innermost f32 tensors of `rows` pixels.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) elt_cat_kernel(const float* __restrict__ x0, const float* __restrict__ x1, int C0,
                                                       int C1, long long rows, long long ld0, long long ld1,
                                                       float* __restrict__ o32, __half* __restrict__ o16) {
  const int C = C0 + C1;
  const long long total = rows * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C; const int c = (int)(i % C);
    const float v = c < C0 ? x0[r * ld0 + c] : x1[r * ld1 + (c - C0)];
    if (o32) o32[i] = v;
    if (o16) o16[i] = __float2half_rn(v);
  }
}

__global__ void __launch_bounds__(256) elt_axpby_kernel(const float* __restrict__ x, const float* __restrict__ y, float a,
                                                         float b, long long n, float* __restrict__ o32, __half* __restrict__ o16) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = y ? fmaf(a, x[i], b * y[i]) : fmaf(a, x[i], b);
    if (o32) o32[i] = v;
    if (o16) o16[i] = __float2half_rn(v);
  }
}

// out = a*((x + y) + z): one pass instead of two adds, a scale and a cast (n is a multiple of 4, 16-byte aligned)
__global__ void __launch_bounds__(256) elt_sum3_kernel(const float4* __restrict__ x, const float4* __restrict__ y,
                                                        const float4* __restrict__ z, float a, long long n4,
                                                        float4* __restrict__ o32, uint2* __restrict__ o16) {
  egr_pdl_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 p = __ldg(x + i), q = __ldg(y + i), r = __ldg(z + i);
    const float4 v = make_float4(((p.x + q.x) + r.x) * a, ((p.y + q.y) + r.y) * a, ((p.z + q.z) + r.z) * a, ((p.w + q.w) + r.w) * a);
    if (o32) o32[i] = v;
    if (o16) {
      __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
      uint2 pk;
      pk.x = *reinterpret_cast<unsigned*>(&h0);
      pk.y = *reinterpret_cast<unsigned*>(&h1);
      o16[i] = pk;
    }
  }
}

// nearest 2x upsample of [B,H,W,C] f32 -> [B,2H,2W,C] f16 (and/or f32)
__global__ void __launch_bounds__(256) elt_up2x_kernel(const float* __restrict__ x, int B, int H, int W, int C,
                                                        float* __restrict__ o32, __half* __restrict__ o16) {
  const long long total = (long long)B * 2 * H * 2 * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long p = i / C;
    const int w2 = (int)(p % (2 * W)); p /= 2 * W;
    const int h2 = (int)(p % (2 * H)); const int b = (int)(p / (2 * H));
    const float v = x[(((long long)b * H + (h2 >> 1)) * W + (w2 >> 1)) * C + c];
    if (o32) o32[i] = v;
    if (o16) o16[i] = __float2half_rn(v);
  }
}

int egr::launch_eltwise(const Spaces& s, const egr_op& op, cudaStream_t st) {
  const float* x0 = (const float*)resolve(s, op.x0.addr);
  const float* x1 = (const float*)resolve(s, op.x1.addr);
  float* o32 = (float*)resolve(s, op.ptr[EGR_P_OUT32]);
  __half* o16 = (__half*)resolve(s, op.ptr[EGR_P_OUT16]);
  const int mode = (int)op.i[EGR_I_MODE];
  if (!x0 || (!o32 && !o16)) return fail(EGR_ERR_ARG, "%s: bad arguments", op.name);
  switch (mode) {
    case EGR_ELT_CAST16:
    case EGR_ELT_COPY32: {
      const int C0 = (int)op.i[EGR_I_C0], C1 = (int)op.i[EGR_I_C1];
      const long long rows = op.i[EGR_I_ROWS];
      const long long ld0 = op.i[EGR_I_AUX0] ? op.i[EGR_I_AUX0] : C0, ld1 = op.i[EGR_I_AUX1] ? op.i[EGR_I_AUX1] : C1;
      if (C1 > 0 && !x1) return fail(EGR_ERR_ARG, "%s: null x1", op.name);
      elt_cat_kernel<<<grid1d(rows * (C0 + C1)), 256, 0, st>>>(x0, x1, C0, C1, rows, ld0, ld1, o32, o16);
      break;
    }
    case EGR_ELT_AXPBY:
      if (!x1) return fail(EGR_ERR_ARG, "%s: null x1", op.name);
      elt_axpby_kernel<<<grid1d(op.i[EGR_I_ROWS]), 256, 0, st>>>(x0, x1, (float)op.f[EGR_F_A], (float)op.f[EGR_F_B],
                                                                 op.i[EGR_I_ROWS], o32, o16);
      break;
    case EGR_ELT_SCALE_SHIFT:
      elt_axpby_kernel<<<grid1d(op.i[EGR_I_ROWS]), 256, 0, st>>>(x0, nullptr, (float)op.f[EGR_F_A], (float)op.f[EGR_F_B],
                                                                 op.i[EGR_I_ROWS], o32, o16);
      break;
    case EGR_ELT_SUM3: {
      const float* x2 = (const float*)resolve(s, op.ptr[EGR_P_AUX]);
      const long long n = op.i[EGR_I_ROWS];
      auto al = [](const void* q) { return q == nullptr || reinterpret_cast<uintptr_t>(q) % 16 == 0; };
      if (!x1 || !x2 || (n & 3) || !al(x0) || !al(x1) || !al(x2) || !al(o32) || !al(o16))
        return fail(EGR_ERR_ARG, "%s: SUM3 needs three 16-byte aligned inputs and a multiple of 4 elements", op.name);
      elt_sum3_kernel<<<grid1d(n / 4), 256, 0, st>>>(reinterpret_cast<const float4*>(x0), reinterpret_cast<const float4*>(x1),
                                                    reinterpret_cast<const float4*>(x2), (float)op.f[EGR_F_A], n / 4,
                                                    reinterpret_cast<float4*>(o32), reinterpret_cast<uint2*>(o16));
      break;
    }
    case EGR_ELT_UPSAMPLE2X: {
      const int B = (int)op.i[EGR_I_BATCH], H = (int)op.i[EGR_I_AUX0], W = (int)op.i[EGR_I_AUX1], C = (int)op.i[EGR_I_C0];
      elt_up2x_kernel<<<grid1d((long long)B * 4 * H * W * C), 256, 0, st>>>(x0, B, H, W, C, o32, o16);
      break;
    }
    default:
      return fail(EGR_ERR_ARG, "%s: unknown eltwise mode %d", op.name, mode);
  }
  EGR_CHECK_LAUNCH(op.name);
  return EGR_OK;
}

// ------------------------------------------------------------------------------------------------
// Anti-aliased SnakeBeta (BigVGAN Activation1d): y = down2(snake(up2(x))), 12-tap Kaiser-sinc FIRs with
// replicate padding.  x f32 [B,T,C] -> f16 (and/or f32).  Thread = (channel, run of TT outputs): lanes walk
// channels (coalesced 128 B per time step), each thread slides a register window along time so every
// up-sampled/activated sample is computed exactly once: 1 load, 2 snake evaluations, 24 FMAs per output.
//   u[2m]   = 2*sum_r x[m-3+r]*f[11-2r],  u[2m+1] = 2*sum_r x[m-2+r]*f[10-2r]        (r = 0..5)
//   s[n]    = u[n] + inv_beta*sin^2(alpha*u[n])
//   y[t]    = sum_j s[clamp(2t+j-5, 0, 2T-1)]*f[j]                                    (j = 0..11)
// ------------------------------------------------------------------------------------------------
#define SNAKE_TT 64

// The kernel is written once over a lane type V: float (one channel per thread) or float2 (two adjacent channels per
// thread, every arithmetic instruction a packed FFMA2 / FMUL2 / FADD2 — same IEEE result per lane, half the issue slots;
// the scalar kernel sat at 74 cycles per warp-output = its 38 FMA-pipe instructions at one per two cycles).
template <typename V> struct SnakeLane;
template <> struct SnakeLane<float> {
  static constexpr int W = 1;
  static __device__ __forceinline__ float ld(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ float bc(float a) { return a; }
  static __device__ __forceinline__ float fma(float a, float b, float c) { return fmaf(a, b, c); }
  static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
  static __device__ __forceinline__ float add(float a, float b) { return a + b; }
  static __device__ __forceinline__ float sfu_sin(float a) { return __sinf(a); }
  static __device__ __forceinline__ float expv(float a) { return __expf(a); }
  static __device__ __forceinline__ float inv_eps(float a) { return 1.0f / (a + 1e-9f); }
  static __device__ __forceinline__ void st32(float* p, float v) { *p = v; }
  static __device__ __forceinline__ void st16(__half* p, float v) { *p = __float2half_rn(v); }
};
template <> struct SnakeLane<float2> {
  static constexpr int W = 2;
  static __device__ __forceinline__ float2 ld(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
  static __device__ __forceinline__ float2 bc(float a) { return make_float2(a, a); }
  static __device__ __forceinline__ float2 fma(float2 a, float2 b, float2 c) { return egr_fma2(a, b, c); }
  static __device__ __forceinline__ float2 mul(float2 a, float2 b) { return egr_mul2(a, b); }
  static __device__ __forceinline__ float2 add(float2 a, float2 b) { return egr_add2(a, b); }
  static __device__ __forceinline__ float2 sfu_sin(float2 a) { return make_float2(__sinf(a.x), __sinf(a.y)); }
  static __device__ __forceinline__ float2 expv(float2 a) { return make_float2(__expf(a.x), __expf(a.y)); }
  static __device__ __forceinline__ float2 inv_eps(float2 a) { return make_float2(1.0f / (a.x + 1e-9f), 1.0f / (a.y + 1e-9f)); }
  static __device__ __forceinline__ void st32(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
  static __device__ __forceinline__ void st16(__half* p, float2 v) { *reinterpret_cast<__half2*>(p) = __floats2half2_rn(v.x, v.y); }
};

// s = U + sin^2(alpha U) / beta, U = the up-sampled sample (the up-sampler's gain of 2 is folded into its taps: doubling
// is exact, so U carries the bits of 2 * sum).  sin is the SFU approximation on the raw phase: MUFU.SIN works on
// phase / 2 pi, whose float32 rounding is |phase| * 1e-8 rad — 3e-6 rad at 50 rad, two orders of magnitude below the f16
// rounding of the value this feeds — so the explicit reduction to [-pi, pi] this kernel used to carry (four more FMA-pipe
// instructions per evaluation, in a kernel that is bound by that pipe) bought nothing measurable.
template <typename V>
__device__ __forceinline__ V snake_eval(V U, V alpha, V inv_beta) {
  using L = SnakeLane<V>;
  const V sn = L::sfu_sin(L::mul(U, alpha));
  return L::fma(L::mul(inv_beta, sn), sn, U);
}

// Thread = (W adjacent channels, run of tt outputs), linear over (run, channel group) so every lane is busy for any C
// and consecutive lanes read consecutive channels (coalesced).  Both the 6-sample input window and the 12-sample
// activated window slide in registers: per output 1 load, 12 FMAs (two up-sampling phases), 2 snake evaluations,
// 12 FMAs (down-sampling), 1 store.
// MB = resident CTAs per SM the register allocation aims at.
template <typename V, int MB, bool O32, bool O16>   // which outputs exist: compile-time, so the per-sample stores carry no branches
__global__ void __launch_bounds__(128, MB) snake_aa_kernel(const float* __restrict__ x, int T, int C, int nruns, int tt,
                                                        const float* __restrict__ log_alpha,
                                                        const float* __restrict__ log_beta,
                                                        const float* __restrict__ filt, float* __restrict__ o32,
                                                        __half* __restrict__ o16) {
  egr_pdl_sync();
  using L = SnakeLane<V>;
  const int CG = C / L::W;   // channel groups
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)nruns * CG) return;
  const int c = (int)(idx % CG) * L::W;
  const int t0 = (int)(idx / CG) * tt;
  const int b = blockIdx.y;
  const V alpha = L::expv(L::ld(log_alpha + c));
  const V inv_beta = L::inv_eps(L::expv(L::ld(log_beta + c)));
  const float* xb = x + (long long)b * T * C + c;
  V f[12], f2[12];   // down-sampling taps, up-sampling taps (x2); uniform registers in the compiled kernel
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    const float fj = __ldg(filt + j);
    f[j] = L::bc(fj);
    f2[j] = L::bc(2.0f * fj);
  }
  const V zero = L::bc(0.f);
  const int n_last = 2 * T - 1;
  auto xat = [&](int m) { m = m < 0 ? 0 : (m >= T ? T - 1 : m); return L::ld(xb + (long long)m * C); };
  auto s_at = [&](int n) {  // activated up-sampled sample n (clamped = replicate padding of s); warm-up only
    n = n < 0 ? 0 : (n > n_last ? n_last : n);
    const int m = n >> 1;
    V u = zero;
    if (n & 1) {
#pragma unroll
      for (int r = 0; r < 6; ++r) u = L::fma(xat(m - 2 + r), f2[10 - 2 * r], u);
    } else {
#pragma unroll
      for (int r = 0; r < 6; ++r) u = L::fma(xat(m - 3 + r), f2[11 - 2 * r], u);
    }
    return snake_eval<V>(u, alpha, inv_beta);
  };
  V sw[12];
  V xw[6];
  if (t0 >= 5 && 2 * t0 + 4 <= n_last && t0 + 4 < T) {
    // interior run: the ten warm-up samples s[2t0-5 .. 2t0+4] only need x[t0-5 .. t0+4] — ten independent loads
    // issued together instead of sixty dependent-latency ones
    V xr[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) xr[i] = L::ld(xb + (long long)(t0 - 5 + i) * C);
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      // n = 2 t0 - 5 + j ; odd n (j even): m = t0 - 3 + j/2, taps x[m-2 .. m+3] ; even n (j odd): m = t0 - 2 + (j-1)/2, x[m-3 .. m+2]
      V u = zero;
      if ((j & 1) == 0) {
#pragma unroll
        for (int r = 0; r < 6; ++r) u = L::fma(xr[(j >> 1) + r], f2[10 - 2 * r], u);
      } else {
#pragma unroll
        for (int r = 0; r < 6; ++r) u = L::fma(xr[((j - 1) >> 1) + r], f2[11 - 2 * r], u);
      }
      sw[j + 2] = snake_eval<V>(u, alpha, inv_beta);
    }
#pragma unroll
    for (int r = 0; r < 5; ++r) xw[r + 1] = xr[5 + r];  // slots 1..5 hold x[t0 .. t0+4]
  } else {
#pragma unroll
    for (int j = 0; j < 10; ++j) sw[j + 2] = s_at(2 * t0 - 5 + j);  // slots 2..11 hold s[2t0-5 .. 2t0+4]
#pragma unroll
    for (int r = 0; r < 5; ++r) xw[r + 1] = xat(t0 + r);
  }
  const int t_end = min(T, t0 + tt);
  if (t0 >= 5 && t0 + tt + 13 <= T && tt % (L::W == 2 ? 6 : 8) == 0) {
    // Interior run (all but the first / last runs of a signal): every index is in range, so the clamps, the end-of-
    // signal selects and the per-output bounds test disappear and addresses advance by pointer increments — the
    // clamped form spends more instructions on integer index math than on the filter itself.
    const float* px = xb + (long long)(t0 + 5) * C;   // next input sample to enter the window
    constexpr int PF = L::W == 2 ? 6 : 8;             // outputs per load group; 6 = one full rotation of both register windows
    float* p32 = O32 ? o32 + ((long long)b * T + t0) * C + c : nullptr;
    __half* p16 = O16 ? o16 + ((long long)b * T + t0) * C + c : nullptr;
    V xn[PF];
#pragma unroll
    for (int i = 0; i < PF; ++i) { xn[i] = L::ld(px); px += C; }
    for (int tb = t0; tb < t_end; tb += PF) {
      V xnext[PF];
#pragma unroll
      for (int i = 0; i < PF; ++i) { xnext[i] = L::ld(px); px += C; }  // one group ahead (in range by the run test)
#pragma unroll
      for (int i = 0; i < PF; ++i) {
#pragma unroll
        for (int j = 0; j < 10; ++j) sw[j] = sw[j + 2];
#pragma unroll
        for (int r = 0; r < 5; ++r) xw[r] = xw[r + 1];
        xw[5] = xn[i];
        V uo = zero, ue = zero;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          uo = L::fma(xw[r], f2[10 - 2 * r], uo);
          ue = L::fma(xw[r], f2[11 - 2 * r], ue);
        }
        sw[10] = snake_eval<V>(uo, alpha, inv_beta);
        sw[11] = snake_eval<V>(ue, alpha, inv_beta);
        V y = zero;
#pragma unroll
        for (int j = 0; j < 12; ++j) y = L::fma(sw[j], f[j], y);
        if (O32) { L::st32(p32, y); p32 += C; }
        if (O16) { L::st16(p16, y); p16 += C; }
      }
#pragma unroll
      for (int i = 0; i < PF; ++i) xn[i] = xnext[i];
    }
    return;
  }
  // boundary runs: clamped indices (replicate padding of the input and of the activated signal)
  V xn[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) xn[i] = xat(t0 + i + 5);
  for (int tb = t0; tb < t_end; tb += 8) {
    V xnext[8];
    if (tb + 8 < t_end) {
#pragma unroll
      for (int i = 0; i < 8; ++i) xnext[i] = xat(tb + 8 + i + 5);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int t = tb + i;
      if (t >= t_end) break;
#pragma unroll
      for (int j = 0; j < 10; ++j) sw[j] = sw[j + 2];
#pragma unroll
      for (int r = 0; r < 5; ++r) xw[r] = xw[r + 1];
      xw[5] = xn[i];
      // s[2t+5] (odd phase of m = t+2) and s[2t+6] (even phase of m = t+3) both read x[t .. t+5]
      V uo = zero, ue = zero;
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        uo = L::fma(xw[r], f2[10 - 2 * r], uo);
        ue = L::fma(xw[r], f2[11 - 2 * r], ue);
      }
      const V so = snake_eval<V>(uo, alpha, inv_beta), se = snake_eval<V>(ue, alpha, inv_beta);
      sw[10] = (2 * t + 5 <= n_last) ? so : sw[9];
      sw[11] = (2 * t + 6 <= n_last) ? se : sw[10];
      V y = zero;
#pragma unroll
      for (int j = 0; j < 12; ++j) y = L::fma(sw[j], f[j], y);
      const long long o = ((long long)b * T + t) * C + c;
      if (O32) L::st32(o32 + o, y);
      if (O16) L::st16(o16 + o, y);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) xn[i] = xnext[i];
  }
}

template <typename V, int MB>
static int snake_aa_launch(const char* name, const float* x, int B, int T, int C, const float* la, const float* lb,
                           const float* filt, float* o32, __half* o16, cudaStream_t st) {
  constexpr int W = SnakeLane<V>::W;
  // Outputs per thread: every run pays ~10 warm-up evaluations of the activation, and a grid slightly larger than
  // what the GPU holds at once costs a whole extra wave — pick the run length (multiple of the unroll) that minimises
  // waves x (run + warm-up) for this tensor.
  static int resident = 0;
  if (!resident) {
    int per_sm = 0;
    EGR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, snake_aa_kernel<V, MB, false, true>, 128, 0));
    resident = (per_sm > 0 ? per_sm : 8) * (egr::devinfo().sm_count ? egr::devinfo().sm_count : 148);
  }
  const int CG = C / W;
  int tt = SNAKE_TT;
  {
    double best = 1e30;
    constexpr int unit = W == 2 ? 6 : 8;   // the interior loop's unroll
    for (int cand = 2 * unit; cand <= 128; cand += unit) {
      const long long blocks = (((long long)((T + cand - 1) / cand) * CG + 127) / 128) * B;
      const long long waves = (blocks + resident - 1) / resident;
      const double cost = (double)waves * (cand + 10);
      if (cost < best) { best = cost; tt = cand; }
    }
  }
  const int nruns = (T + tt - 1) / tt;
  dim3 grid((unsigned)(((long long)nruns * CG + 127) / 128), B);
  if (egr::pdl_enabled()) {
#ifdef __CUDACC__
    if (o32 && o16) EGR_CUDA(egr::launch_pdl(snake_aa_kernel<V, MB, true, true>, grid, dim3(128), 0, st, x, T, C, nruns, tt, la, lb, filt, o32, o16));
    else if (o16) EGR_CUDA(egr::launch_pdl(snake_aa_kernel<V, MB, false, true>, grid, dim3(128), 0, st, x, T, C, nruns, tt, la, lb, filt, o32, o16));
    else EGR_CUDA(egr::launch_pdl(snake_aa_kernel<V, MB, true, false>, grid, dim3(128), 0, st, x, T, C, nruns, tt, la, lb, filt, o32, o16));
#endif
  } else if (o32 && o16) snake_aa_kernel<V, MB, true, true><<<grid, 128, 0, st>>>(x, T, C, nruns, tt, la, lb, filt, o32, o16);
  else if (o16) snake_aa_kernel<V, MB, false, true><<<grid, 128, 0, st>>>(x, T, C, nruns, tt, la, lb, filt, o32, o16);
  else snake_aa_kernel<V, MB, true, false><<<grid, 128, 0, st>>>(x, T, C, nruns, tt, la, lb, filt, o32, o16);
  EGR_CHECK_LAUNCH(name);
  return EGR_OK;
}

int egr::launch_snake_aa(const Spaces& s, const egr_op& op, cudaStream_t st) {
  const float* x = (const float*)resolve(s, op.x0.addr);
  const float* la = (const float*)resolve(s, op.ptr[EGR_P_GAMMA]);
  const float* lb = (const float*)resolve(s, op.ptr[EGR_P_BETA]);
  const float* filt = (const float*)resolve(s, op.ptr[EGR_P_AUX]);
  float* o32 = (float*)resolve(s, op.ptr[EGR_P_OUT32]);
  __half* o16 = (__half*)resolve(s, op.ptr[EGR_P_OUT16]);
  const int B = (int)op.i[EGR_I_BATCH], T = (int)op.i[EGR_I_ROWS], C = (int)op.i[EGR_I_COLS];
  if (!x || !la || !lb || !filt || (!o32 && !o16) || B <= 0 || T <= 0 || C <= 0) return fail(EGR_ERR_ARG, "%s: bad arguments", op.name);
  if (op.i[EGR_I_AUX0] != 12) return fail(EGR_ERR_UNSUPPORTED, "%s: only the 12-tap anti-alias filter is built", op.name);
  // two channels per thread (packed f32x2 arithmetic) whenever the channel pairs are 8-byte aligned everywhere
  const bool no_pack = getenv("EGR_SNAKE_SCALAR") != nullptr;   // read per launch: the op test compares the two paths bit for bit
  const bool pack = !no_pack && (C & 1) == 0 && (((uintptr_t)x | (uintptr_t)la | (uintptr_t)lb | (uintptr_t)o32) & 7) == 0 &&
                    ((uintptr_t)o16 & 3) == 0;
  // register targets swept on B200 (tools/snake_sweep.py, c2 pass, 91 ops): packed at 4 / 5 / 6 / 7 / 8 CTAs per SM
  // 2031 / 2224 / 2223 / 2437 / 2347 us, scalar 2352 us — the kernel is bound by the FMA pipe, not by latency
  if (pack) return snake_aa_launch<float2, 4>(op.name, x, B, T, C, la, lb, filt, o32, o16, st);
  return snake_aa_launch<float, 6>(op.name, x, B, T, C, la, lb, filt, o32, o16, st);
}

// ------------------------------------------------------------------------------------------------
// sinusoidal timestep embedding: out[0:half] = cos(t*f_i), out[half:] = sin(t*f_i), f_i = 10000^(-i/half)
// ------------------------------------------------------------------------------------------------
__global__ void time_embed_kernel(float t, int dim, float* __restrict__ out) {
  const int half = dim / 2;
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float fr = expf(-9.210340371976184f * (float)i / (float)half);
    const float a = t * fr;
    out[i] = cosf(a);
    out[half + i] = sinf(a);
  }
}

int egr::launch_time_embed(const Spaces& s, const egr_op& op, cudaStream_t st) {
  float* o = (float*)resolve(s, op.ptr[EGR_P_OUT32]);
  const int dim = (int)op.i[EGR_I_COLS];
  if (!o || dim <= 0 || (dim & 1)) return fail(EGR_ERR_ARG, "%s: bad arguments", op.name);
  time_embed_kernel<<<1, 128, 0, st>>>((float)op.f[EGR_F_A], dim, o);
  EGR_CHECK_LAUNCH(op.name);
  return EGR_OK;
}

int egr::launch_zero(const Spaces& s, const egr_op& op, cudaStream_t st) {
  void* p = resolve(s, op.ptr[EGR_P_OUT32]);
  if (!p || op.i[EGR_I_ROWS] < 0) return fail(EGR_ERR_ARG, "%s: bad arguments", op.name);
  EGR_CUDA(cudaMemsetAsync(p, 0, (size_t)op.i[EGR_I_ROWS], st));
  return EGR_OK;
}
