// ops.cuh — launch entry points of the plan's ops (one function per EGR_OP_* code).
#pragma once
#include "common.cuh"

namespace egr {

// device-side strided 5-D view (resolved pointer)
struct View {
  const void* p;
  int rank, elem;  // elem: 0 f32, 1 f16
  long long dim[5];
  long long stride[5];
};

inline View make_view(const Spaces& s, const egr_tensor& t) {
  View v;
  v.p = resolve(s, t.addr);
  v.rank = t.rank;
  v.elem = t.elem;
  for (int i = 0; i < 5; ++i) {
    v.dim[i] = i < t.rank ? t.dim[i] : 1;
    v.stride[i] = i < t.rank ? t.stride[i] : 0;
  }
  return v;
}

struct Taps {
  short t[EGR_MAX_TAPS][5];
};

// Everything the GEMM-family kernels need besides the operand views.
struct GemmArgs {
  int dimW, dimH, dimB, bw, bh, bb, Wo, Ho, Bo, ntaps, K, N, block_n;
  long long wstride_n, wstride_z;
  int wz_batch;
  const void* W;
  const float* bias;
  const float* rowbias;
  long long rowbias_stride;
  const float* resid;
  const float* resid2;  // second residual (vocoder: running sum of the parallel residual blocks)
  float post;           // scale applied after the residual adds
  float* out32;
  __half* out16;
  long long out_pix_stride, out_batch_stride, out_offset, out_lo, out_hi, out_n_stride;
  long long out_h_stride;  // stride of one step along H (tensor-core path; Wo*out_pix_stride when dense)
  int transposed, act;
  float alpha;
};

int gemm_args_from_op(const Spaces& s, const egr_op& op, GemmArgs* g, Taps* taps, View* a);

int launch_gemm_simt(const Spaces& s, const egr_op& op, cudaStream_t st);
int launch_gn_stats(const Spaces& s, const egr_op& op, cudaStream_t st);
int launch_gn_apply(const Spaces& s, const egr_op& op, cudaStream_t st);
int launch_layernorm(const Spaces& s, const egr_op& op, cudaStream_t st);
int launch_softmax(const Spaces& s, const egr_op& op, cudaStream_t st);
int launch_attn_small(const Spaces& s, const egr_op& op, cudaStream_t st);
int launch_geglu(const Spaces& s, const egr_op& op, cudaStream_t st);
int launch_eltwise(const Spaces& s, const egr_op& op, cudaStream_t st);
int launch_snake_aa(const Spaces& s, const egr_op& op, cudaStream_t st);
int launch_stft_mel(const Spaces& s, const egr_op& op, cudaStream_t st);
int launch_lowpass(const Spaces& s, const egr_op& op, cudaStream_t st);
int launch_time_embed(const Spaces& s, const egr_op& op, cudaStream_t st);
int launch_zero(const Spaces& s, const egr_op& op, cudaStream_t st);

// tcgen05 path (gemm_tc.cu)
struct TcPrepared;  // tensor maps + launch geometry, built once per op at plan creation
int  tc_prepare(const Spaces& s, const egr_op& op, TcPrepared** out);
int  tc_launch(const TcPrepared* p, cudaStream_t st);
void tc_free(TcPrepared* p);
// split-K scratch (library-owned, shared by all ops of a plan: they run in stream order)
size_t tc_partial_bytes(const TcPrepared* p);
int    tc_num_counters(const TcPrepared* p);
void   tc_bind_scratch(TcPrepared* p, float* partial, unsigned int* counters);
void   tc_describe(const TcPrepared* p, int* out8);  // block_n, mt, splits, halo, n_work, grid, SA, SB
int  tc_global_init();

}  // namespace egr

// shared epilogue math
__device__ __forceinline__ float egr_apply_act(float v, int act) {
  if (act == EGR_ACT_SILU) return v / (1.0f + __expf(-v));
  if (act == EGR_ACT_TANH) return tanhf(v);
  return v;
}

// ------------------------------------------------------------------------------------------------
// shared scalar epilogue: out = act(alpha*acc + bias + rowbias) + resid, with the transposed-conv crop
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void epilogue_store(const egr::GemmArgs& g, int b, long long pix, int n, float acc) {
  float v = acc * g.alpha;
  if (g.bias) v += g.bias[n];
  if (g.rowbias) v += g.rowbias[(long long)b * g.rowbias_stride + n];
  v = egr_apply_act(v, g.act);
  long long idx;
  if (g.transposed) {
    idx = (long long)b * g.out_batch_stride + (long long)n * g.out_n_stride + pix + g.out_offset;
  } else {
    long long flat = pix * g.out_pix_stride + g.out_offset + n;
    if (flat < g.out_lo || flat >= g.out_hi) return;
    idx = (long long)b * g.out_batch_stride + flat;
  }
  if (g.resid) v += g.resid[idx];
  if (g.resid2) v += g.resid2[idx];
  v *= g.post;
  if (g.out32) g.out32[idx] = v;
  if (g.out16) g.out16[idx] = __float2half_rn(v);
}

// four consecutive output channels n .. n+3 of one pixel: one 16-byte store (and one 8-byte f16 store) when the output is
// dense in n, aligned and inside the crop; the scalar path otherwise.  Same arithmetic per element as epilogue_store.
__device__ __forceinline__ void epilogue_store4(const egr::GemmArgs& g, int b, long long pix, int n, const float (&acc)[4]) {
  const long long flat = pix * g.out_pix_stride + g.out_offset + n;
  const long long idx = (long long)b * g.out_batch_stride + flat;
  const bool vec = !g.transposed && (idx & 3) == 0 && flat >= g.out_lo && flat + 3 < g.out_hi && n + 3 < g.N &&
                   (reinterpret_cast<uintptr_t>(g.out32) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.out16) & 7) == 0 &&
                   (!g.resid || (reinterpret_cast<uintptr_t>(g.resid) & 15) == 0) &&
                   (!g.resid2 || (reinterpret_cast<uintptr_t>(g.resid2) & 15) == 0);
  if (!vec) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (n + u < g.N) epilogue_store(g, b, pix, n + u, acc[u]);
    return;
  }
  float v[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    v[u] = acc[u] * g.alpha;
    if (g.bias) v[u] += g.bias[n + u];
    if (g.rowbias) v[u] += g.rowbias[(long long)b * g.rowbias_stride + n + u];
    v[u] = egr_apply_act(v[u], g.act);
  }
  if (g.resid) {
    const float4 r = *reinterpret_cast<const float4*>(g.resid + idx);
    v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
  }
  if (g.resid2) {
    const float4 r = *reinterpret_cast<const float4*>(g.resid2 + idx);
    v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) v[u] *= g.post;
  if (g.out32) *reinterpret_cast<float4*>(g.out32 + idx) = make_float4(v[0], v[1], v[2], v[3]);
  if (g.out16) {
    __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
    uint2 pk;
    pk.x = *reinterpret_cast<unsigned*>(&h0);
    pk.y = *reinterpret_cast<unsigned*>(&h1);
    *reinterpret_cast<uint2*>(g.out16 + idx) = pk;
  }
}

// GroupNorm input: a virtual channel-concat of x0 (C0 channels) and x1 (C1 channels), f32, channels innermost
struct CatArgs {
  const float* x0; const float* x1;
  int C0, C1, G, B;
  long long P;  // pixels per batch item
};

static inline int cat_args(const egr::Spaces& s, const egr_op& op, CatArgs* a) {
  a->x0 = (const float*)egr::resolve(s, op.x0.addr);
  a->x1 = (const float*)egr::resolve(s, op.x1.addr);
  a->C0 = (int)op.i[EGR_I_C0]; a->C1 = (int)op.i[EGR_I_C1];
  a->G = (int)op.i[EGR_I_GROUPS]; a->B = (int)op.i[EGR_I_BATCH];
  a->P = op.i[EGR_I_ROWS];
  const int C = a->C0 + a->C1;
  if (!a->x0 || a->C0 <= 0 || (a->C1 > 0 && !a->x1)) return egr::fail(EGR_ERR_ARG, "%s: bad inputs", op.name);
  if ((a->C0 & 3) || (a->C1 & 3)) return egr::fail(EGR_ERR_ARG, "%s: channel counts must be multiples of 4", op.name);
  if (a->G <= 0 || C % a->G) return egr::fail(EGR_ERR_ARG, "%s: C=%d not divisible by groups=%d", op.name, C, a->G);
  if (a->B <= 0 || a->B > 65535 || a->P <= 0) return egr::fail(EGR_ERR_ARG, "%s: bad batch/pixels", op.name);
  return EGR_OK;
}


