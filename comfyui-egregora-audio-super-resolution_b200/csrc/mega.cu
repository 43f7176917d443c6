// mega.cu — the UNet of the FlashSR plan as ONE persistent kernel.
//
// Why.  A denoising step is ~440 ops over feature maps of 32 .. 2048 pixels per chunk-channel: 1 % of the pass's FLOPs,
// yet 4.7 ms of a 16.5 ms pass at batch 1 and 9.7 ms per step at batch 8 (round-2 measurements) — every op is a kernel
// whose useful work is a few microseconds behind a launch, a prologue (barrier set-up, TMEM allocation, descriptor
// fetch) and a drain.  Its weights (360 MB of f16) stream in 55 us at HBM speed; nothing else bounds the section.
//
// What.  One CTA per SM (cooperative launch, 384 threads) walks an op table in global memory.  Between dependent ops
// the CTAs meet at a grid barrier (one atomic round trip through L2, ~1.5 us) instead of at a kernel boundary; the
// planner marks the ops whose successor does not depend on them (the q/k/v projections, independent casts), which skip
// it.  TMEM is allocated once per launch, tensor maps and per-op arguments are prefetched from the table.
//   GEMM ops run the SAME warp roles as gemm_tc_kernel (gemm_tc.cuh: TMA producers, tcgen05 issuers, TMEM epilogue,
//   split-K with per-layer constant splits) — only the barriers are re-initialised per op;
//   the CUDA-core ops (GroupNorm over a virtual concat, LayerNorm, short-sequence attention, GEGLU, casts, the DDIM
//   update, the time-embedding GEMVs) are grid-strided bodies on warps 0-7 with the arithmetic of their stand-alone
//   kernels in ops.cu.
// Results are deterministic and independent of the batch a chunk-channel runs in (fixed reduction orders, geometry-only
// work splits), like the per-op path.
#include <vector>
#include "gemm_tc.cuh"
#include "mega.cuh"

namespace egr {

enum { MOP_GEMM_TC = 1, MOP_GEMV, MOP_GN_STATS, MOP_GN_APPLY, MOP_LAYERNORM, MOP_ATTN, MOP_GEGLU, MOP_CAT, MOP_AXPBY, MOP_TIME_EMBED, MOP_SPLITK_REDUCE, MOP_GN_FUSED };

struct __align__(16) MegaOp {
  int code, sync_after, tc, pad;
  // parameters of the ops that FOLLOW — [0] the weights of the next weight-reading op, [1] [2] the small vectors (bias,
  // gamma, beta) of the next op — prefetched into L2 while this op runs (static data, so always legal): the layer's
  // first TMA loads and the epilogue's parameter loads then hit L2 instead of paying HBM latency in their critical path
  const void* pf_ptr[3];
  int pf_bytes[3];
  int pad2;
  union {
    struct { View a; GemmArgs g; int npix; } gemv;
    struct { CatArgs a; const float* gamma; const float* beta; float eps; int silu; float* out32; __half* out16; double* part; int nsl; } gn;
    struct { const float* x; long long rows; int C; float eps; const float* gamma; const float* beta; __half* out16; float* out32; } ln;
    struct { const __half* q; const __half* k; const __half* v; __half* out; int S, heads, hd, B, ld, qblocks; float scale; int mma; } attn;
    struct { const float* x; long long rows; int D; __half* out; } geglu;
    struct { const float* x0; const float* x1; int C0, C1; long long rows, ld0, ld1; float* o32; __half* o16; } cat;
    struct { const float* x; const float* y; float a, b; long long n; float* o32; __half* o16; } axpby;
    struct { float t; int dim; float* out; } temb;
  } u;
};
static_assert(sizeof(MegaOp) % 16 == 0 && sizeof(MegaOp) <= 512, "MegaOp is copied to shared memory in 16-byte words");
static_assert(sizeof(TcKernelArgs) % 16 == 0, "TcKernelArgs is copied to shared memory in 16-byte words");

struct MegaRun {
  int first = 0, last = 0, n_ops = 0, n_tc = 0, n_sync = 0;
  MegaOp* d_ops = nullptr;
  TcKernelArgs* d_kas = nullptr;
  CUtensorMap* d_maps = nullptr;   // [n_tc][2], 128-byte entries
  unsigned int* d_bar = nullptr;   // [0] arrival counter, [2] watchdog flag
  int* d_tc_of = nullptr;          // per op: index of its GEMM argument block, -1 for a CUDA-core op
  unsigned long long* d_trace = nullptr;   // EGR_MEGA_TRACE=1: clock64 stamps of CTA 0, 4 per op
  int smem_bytes = 0;
  int grid = 0;
};

static constexpr int MEGA_THREADS = 384;
static constexpr int MEGA_SIMT = 384;        // CUDA-core ops run on all 12 warps
static constexpr int MEGA_NW = MEGA_SIMT / 32;
static constexpr int MEGA_MAX_STAGES = 10;   // ring depth bound of tc_prepare
static constexpr int MEGA_BAR_BYTES = 1024;  // 4 x 10 ring barriers + 4 accumulator barriers, padded
static constexpr int MEGA_DYN_SMEM_MAX = 227 * 1024 - 4096;   // 227 KB per CTA minus the kernel's static shared memory

}  // namespace egr

using namespace egr;

// ------------------------------------------------------------------------------------------------ synchronisation
__device__ __forceinline__ void simt_sync() { asm volatile("bar.sync 2, 384;" ::: "memory"); }

// Grid barrier: one monotonically increasing arrival counter per launch (zeroed by a memset node before the launch).
// A CTA arrives with a fire-and-forget release reduction and waits until the counter reaches epoch * ncta: one L2 round
// trip after the last arrival, no returned atomics, no second word.
// Memory: every thread's earlier global writes are ordered before the arrival by bar.sync + the release of the arriving
// thread; the acquire load on the way out (gpu scope: it also drops the SM's L1 lines, so data other SMs wrote is read
// from L2) + bar.sync orders every thread's later reads after it; both sides fence the generic -> async proxy edge,
// because the next op's TMA loads read what plain stores of the previous op produced.
// Watchdog: a CTA that never arrives (a bug, never a legal state) must not hang the GPU — after ~2^25 polls (seconds) the
// waiter raises bar[2] and bumps the counter past every later target: the launch drains (with garbage) and the host
// reports it (egr_debug_mega_aborted).
__device__ __forceinline__ void grid_sync(unsigned int* bar, unsigned int target) {
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
    unsigned int now, spins = 0;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(now) : "l"(bar) : "memory");
      if (now >= target) break;
      if (++spins > (1u << 25)) {
        atomicExch(bar + 2, 1u);
        atomicAdd(bar, 0x40000000u);
        break;
      }
    }
  }
  __syncthreads();
  asm volatile("fence.proxy.async;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ CUDA-core op bodies
// (threads 0..255 of every CTA; activations are read with ld.global.cg — they were written by other SMs during this launch)
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

__device__ __noinline__ void mega_gemv(const MegaOp& op, int cta, int ncta) {
  const GemmArgs& g = op.u.gemv.g;
  const View& a = op.u.gemv.a;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int npix = op.u.gemv.npix;
  for (int n = cta * MEGA_NW + warp; n < g.N; n += ncta * MEGA_NW) {
    const float* wrow = reinterpret_cast<const float*>(g.W) + (long long)n * g.wstride_n;
    for (int p = 0; p < npix; ++p) {
      const int w = p % g.Wo, h = (p / g.Wo) % g.Ho, b = p / (g.Wo * g.Ho);
      const float* x = reinterpret_cast<const float*>(a.p) + (long long)w * a.stride[g.dimW] + (long long)h * a.stride[g.dimH] +
                       (long long)b * a.stride[g.dimB];
      float acc = 0.f;
      for (int k = lane * 4; k < g.K; k += 128) {
        const float4 xv = ldcg4(x + k);
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wrow + k));
        acc = fmaf(xv.x, wv.x, acc); acc = fmaf(xv.y, wv.y, acc); acc = fmaf(xv.z, wv.z, acc); acc = fmaf(xv.w, wv.w, acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) epilogue_store(g, b, (long long)h * g.Wo + w, n, acc);
    }
  }
}

__device__ __forceinline__ const float* gn_src(const CatArgs& a, int b, long long p, int c) {
  return c < a.C0 ? a.x0 + ((long long)b * a.P + p) * a.C0 + c : a.x1 + ((long long)b * a.P + p) * a.C1 + (c - a.C0);
}

// unit = (item b, group gi, pixel slice sl): f32 partial moments per thread, combined in f64 in a fixed order
__device__ __noinline__ void mega_gn_stats(const MegaOp& op, int cta, int ncta, double (*red)[MEGA_NW]) {
  const CatArgs& a = op.u.gn.a;
  const int C = a.C0 + a.C1, cpg = C / a.G, q4 = cpg >> 2, nsl = op.u.gn.nsl;
  const int units = a.B * a.G * nsl;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int unit = cta; unit < units; unit += ncta) {
    const int sl = unit % nsl, gi = (unit / nsl) % a.G, b = unit / (nsl * a.G);
    const int p_lo = (int)(a.P * sl / nsl), p_hi = (int)(a.P * (sl + 1) / nsl);
    const int n4 = (p_hi - p_lo) * q4;
    float s = 0.f, ss = 0.f;
    for (int u = tid; u < n4; u += MEGA_SIMT) {
      const int pl = u / q4;
      const int c = gi * cpg + (u - pl * q4) * 4;
      const float4 v = ldcg4(gn_src(a, b, p_lo + pl, c));
      s += (v.x + v.y) + (v.z + v.w);
      ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
    }
    const double ds = warp_sum((double)s), dss = warp_sum((double)ss);
    if (lane == 0) { red[0][warp] = ds; red[1][warp] = dss; }
    simt_sync();
    if (tid == 0) {
      double t = 0.0, tt = 0.0;
      for (int w = 0; w < MEGA_NW; ++w) { t += red[0][w]; tt += red[1][w]; }
      double* dst = op.u.gn.part + (((long long)b * a.G + gi) * nsl + sl) * 2;
      dst[0] = t; dst[1] = tt;
    }
    simt_sync();
  }
}

__device__ __noinline__ void mega_gn_apply(const MegaOp& op, int cta, int ncta, float* stat) {
  const CatArgs& a = op.u.gn.a;
  const int C = a.C0 + a.C1, cpg = C / a.G, q4 = cpg >> 2, nsl = op.u.gn.nsl;
  const int units = a.B * a.G * nsl;
  const int tid = threadIdx.x;
  const int silu = op.u.gn.silu;
  float* out32 = op.u.gn.out32;
  __half* out16 = op.u.gn.out16;
  for (int unit = cta; unit < units; unit += ncta) {
    const int sl = unit % nsl, gi = (unit / nsl) % a.G, b = unit / (nsl * a.G);
    if (tid == 0) {
      const double* src = op.u.gn.part + ((long long)b * a.G + gi) * nsl * 2;
      double t = 0.0, tt = 0.0;
      for (int r = 0; r < nsl; ++r) { t += __ldcg(src + 2 * r); tt += __ldcg(src + 2 * r + 1); }
      const double cnt = (double)cpg * (double)a.P;
      const double mean = t / cnt;
      double var = tt / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      stat[0] = (float)mean;
      stat[1] = (float)(1.0 / sqrt(var + (double)op.u.gn.eps));
    }
    simt_sync();
    const float mean = stat[0], rstd = stat[1];
    const int p_lo = (int)(a.P * sl / nsl), p_hi = (int)(a.P * (sl + 1) / nsl);
    const int n4 = (p_hi - p_lo) * q4;
    for (int u = tid; u < n4; u += MEGA_SIMT) {
      const int pl = u / q4;
      const int c = gi * cpg + (u - pl * q4) * 4;
      const int p = p_lo + pl;
      const float4 v = ldcg4(gn_src(a, b, p, c));
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(op.u.gn.gamma + c));
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(op.u.gn.beta + c));
      float r[4] = {fmaf(v.x, rstd * g4.x, b4.x - mean * (rstd * g4.x)), fmaf(v.y, rstd * g4.y, b4.y - mean * (rstd * g4.y)),
                    fmaf(v.z, rstd * g4.z, b4.z - mean * (rstd * g4.z)), fmaf(v.w, rstd * g4.w, b4.w - mean * (rstd * g4.w))};
      if (silu) {
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = egr_silu(r[k]);
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
    simt_sync();   // stat[] is rewritten by the next unit
  }
}

// GroupNorm as ONE op: a CTA owns a whole (item, group) — moments, then normalise (+SiLU) — so no grid barrier sits between
// the two halves.  Groups of up to 384 x 12 float4 stay in registers between the passes (one trip to L2); larger ones are
// read twice.  Per-thread partition and reduction order depend on the op's geometry only (batch-invariant, deterministic).
__device__ __noinline__ void mega_gn_fused(const MegaOp& op, int cta, int ncta, double (*red)[MEGA_NW], float* stat, float (*gb)[64]) {
  constexpr int REG = 12;
  const CatArgs& a = op.u.gn.a;
  const int C = a.C0 + a.C1, cpg = C / a.G, q4 = cpg >> 2;
  const int units = a.B * a.G;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n4 = (int)a.P * q4;
  const bool inreg = n4 <= MEGA_SIMT * REG;
  const int silu = op.u.gn.silu;
  const float* gamma = op.u.gn.gamma;
  const float* beta = op.u.gn.beta;
  float* out32 = op.u.gn.out32;
  __half* out16 = op.u.gn.out16;
  for (int unit = cta; unit < units; unit += ncta) {
    const int gi = unit % a.G, b = unit / a.G;
    float4 xv[REG];
    float s = 0.f, ss = 0.f;
    if (inreg) {
#pragma unroll
      for (int j = 0; j < REG; ++j) {
        const int u = tid + j * MEGA_SIMT;
        if (u < n4) {
          const int pl = u / q4;
          xv[j] = ldcg4(gn_src(a, b, pl, gi * cpg + (u - pl * q4) * 4));
        }
      }
#pragma unroll
      for (int j = 0; j < REG; ++j) {
        if (tid + j * MEGA_SIMT < n4) {
          const float4 v = xv[j];
          s += (v.x + v.y) + (v.z + v.w);
          ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
        }
      }
    } else {
      for (int u = tid; u < n4; u += MEGA_SIMT) {
        const int pl = u / q4;
        const float4 v = ldcg4(gn_src(a, b, pl, gi * cpg + (u - pl * q4) * 4));
        s += (v.x + v.y) + (v.z + v.w);
        ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
      }
    }
    const double ds = warp_sum((double)s), dss = warp_sum((double)ss);
    if (lane == 0) { red[0][warp] = ds; red[1][warp] = dss; }
    // the group's gamma / beta go through shared memory: parameter loads inside the store loop below would be
    // serialised behind the stores (the compiler keeps them in program order), one L2 trip per element
    if (tid < cpg) { gb[0][tid] = __ldg(gamma + gi * cpg + tid); gb[1][tid] = __ldg(beta + gi * cpg + tid); }
    simt_sync();
    if (tid == 0) {
      double t = 0.0, tt = 0.0;
      for (int w = 0; w < MEGA_NW; ++w) { t += red[0][w]; tt += red[1][w]; }
      const double cnt = (double)cpg * (double)a.P;
      const double mean = t / cnt;
      double var = tt / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      stat[0] = (float)mean;
      stat[1] = (float)(1.0 / sqrt(var + (double)op.u.gn.eps));
    }
    simt_sync();
    const float mean = stat[0], rstd = stat[1];
    auto emit = [&](int u, const float4 v) {
      const int pl = u / q4;
      const int cl = (u - pl * q4) * 4;
      const int c = gi * cpg + cl;
      const float4 g4 = *reinterpret_cast<const float4*>(&gb[0][cl]);
      const float4 b4 = *reinterpret_cast<const float4*>(&gb[1][cl]);
      float r[4] = {fmaf(v.x, rstd * g4.x, b4.x - mean * (rstd * g4.x)), fmaf(v.y, rstd * g4.y, b4.y - mean * (rstd * g4.y)),
                    fmaf(v.z, rstd * g4.z, b4.z - mean * (rstd * g4.z)), fmaf(v.w, rstd * g4.w, b4.w - mean * (rstd * g4.w))};
      if (silu) {
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = egr_silu(r[k]);
      }
      const long long o = ((long long)b * a.P + pl) * C + c;
      if (out32) *reinterpret_cast<float4*>(out32 + o) = make_float4(r[0], r[1], r[2], r[3]);
      if (out16) {
        __half2 h0 = __floats2half2_rn(r[0], r[1]), h1 = __floats2half2_rn(r[2], r[3]);
        uint2 pk;
        pk.x = *reinterpret_cast<unsigned*>(&h0);
        pk.y = *reinterpret_cast<unsigned*>(&h1);
        *reinterpret_cast<uint2*>(out16 + o) = pk;
      }
    };
    if (inreg) {
#pragma unroll
      for (int j = 0; j < REG; ++j) {
        const int u = tid + j * MEGA_SIMT;
        if (u < n4) emit(u, xv[j]);
      }
    } else {
      for (int u = tid; u < n4; u += MEGA_SIMT) {
        const int pl = u / q4;
        emit(u, ldcg4(gn_src(a, b, pl, gi * cpg + (u - pl * q4) * 4)));
      }
    }
    simt_sync();   // red[] / stat[] are rewritten by the next unit
  }
}

__device__ __noinline__ void mega_layernorm(const MegaOp& op, int cta, int ncta) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = op.u.ln.C;
  const float* gamma = op.u.ln.gamma;
  const float* beta = op.u.ln.beta;
  constexpr int NJ = 20;   // a row of up to 640 channels lives in registers: ONE trip to L2 instead of three
  for (long long row = (long long)cta * MEGA_NW + warp; row < op.u.ln.rows; row += (long long)ncta * MEGA_NW) {
    const float* xr = op.u.ln.x + row * C;
    if (C <= 32 * NJ) {
      // every load of the row — x, gamma, beta — is issued before the first use: the stores below would otherwise
      // serialise the parameter loads behind them (one cold miss per element: 8.7 us per row in the round-2 trace)
      float xv[NJ], gv[NJ], bv[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int c = lane + 32 * j;
        const bool in = c < C;
        xv[j] = in ? __ldcg(xr + c) : 0.f;
        gv[j] = in ? __ldg(gamma + c) : 0.f;
        bv[j] = in ? __ldg(beta + c) : 0.f;
      }
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) if (lane + 32 * j < C) s += xv[j];
      const float mean = warp_sum(s) / C;
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) if (lane + 32 * j < C) { const float d = xv[j] - mean; v = fmaf(d, d, v); }
      const float rstd = rsqrtf(warp_sum(v) / C + op.u.ln.eps);
      __half* const o16 = op.u.ln.out16;
      float* const o32 = op.u.ln.out32;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int c = lane + 32 * j;
        if (c < C) {
          const float y = (xv[j] - mean) * rstd * gv[j] + bv[j];
          if (o16) o16[row * C + c] = __float2half_rn(y);
          if (o32) o32[row * C + c] = y;
        }
      }
      continue;
    }
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += __ldcg(xr + c);
    const float mean = warp_sum(s) / C;
    float v = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = __ldcg(xr + c) - mean; v = fmaf(d, d, v); }
    const float rstd = rsqrtf(warp_sum(v) / C + op.u.ln.eps);
    for (int c = lane; c < C; c += 32) {
      const float y = (__ldcg(xr + c) - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
      if (op.u.ln.out16) op.u.ln.out16[row * C + c] = __float2half_rn(y);
      if (op.u.ln.out32) op.u.ln.out32[row * C + c] = y;
    }
  }
}

// attn_rows_kernel of ops.cu with (q block, head, item) taken from a virtual block index: K / V of one (item, head) staged
// in shared memory as f16, one warp per query row, lanes split the keys in both phases, reduce-scatter of the outputs
template <int HD>
__device__ __noinline__ void mega_attn(const MegaOp& op, int cta, int ncta, unsigned char* smraw) {
  constexpr int KST = HD + 8;
  const int S = op.u.attn.S, heads = op.u.attn.heads, B = op.u.attn.B, qblocks = op.u.attn.qblocks;
  const __half* q = op.u.attn.q;
  const __half* k = op.u.attn.k;
  const __half* v = op.u.attn.v;
  __half* out = op.u.attn.out;
  const float scale = op.u.attn.scale;
  __half* Ks = reinterpret_cast<__half*>(smraw);
  __half* Vs = Ks + (size_t)S * KST;
  const int C = heads * HD;
  const int L = op.u.attn.ld ? op.u.attn.ld : C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_per_block = (S + qblocks - 1) / qblocks;
  const int njj = (S + 31) >> 5;
  const int nvb = qblocks * heads * B;
  auto row8 = [](const __half* p, float* x) {
    const uint4 raw = *reinterpret_cast<const uint4*>(p);
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
    const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
    const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&raw.z));
    const float2 f3 = __half22float2(*reinterpret_cast<const __half2*>(&raw.w));
    x[0] = f0.x; x[1] = f0.y; x[2] = f1.x; x[3] = f1.y; x[4] = f2.x; x[5] = f2.y; x[6] = f3.x; x[7] = f3.y;
  };
  int staged = -1;   // (item, head) whose K / V currently sit in shared memory
  for (int vb = cta; vb < nvb; vb += ncta) {
    const int qb = vb % qblocks, h = (vb / qblocks) % heads, b = vb / (qblocks * heads);
    const long long base = (long long)b * S * C + (long long)h * HD;
    const long long ibase = (long long)b * S * L + (long long)h * HD;
    if (staged != b * heads + h) {
      simt_sync();   // the previous virtual block's readers are done with Ks / Vs
      for (int i = threadIdx.x; i < S * (HD / 8); i += MEGA_SIMT) {
        const int j = i / (HD / 8), c8 = (i % (HD / 8)) * 8;
        *reinterpret_cast<uint4*>(Ks + j * KST + c8) = __ldcg(reinterpret_cast<const uint4*>(k + ibase + (long long)j * L + c8));
        *reinterpret_cast<uint4*>(Vs + j * KST + c8) = __ldcg(reinterpret_cast<const uint4*>(v + ibase + (long long)j * L + c8));
      }
      staged = b * heads + h;
      simt_sync();
    }
    const int r_lo = qb * rows_per_block, r_hi = min(S, r_lo + rows_per_block);
    for (int r = r_lo + warp; r < r_hi; r += MEGA_NW) {
      float qv[HD];
#pragma unroll
      for (int c8 = 0; c8 < HD; c8 += 8) {
        const uint4 raw = __ldcg(reinterpret_cast<const uint4*>(q + ibase + (long long)r * L + c8));
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
      if (HD == 32) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float mine = (lane & 16) ? o[16 + i] : o[i], send = (lane & 16) ? o[i] : o[16 + i];
          o[i] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
        }
      } else {
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
      const int d = HD == 32 ? lane : (lane & 15);
      if (HD == 32 || lane < 16) out[base + (long long)r * C + d] = __float2half_rn(o[0] * inv);
    }
  }
}

// ------------------------------------------------------------------------------------------------ attention on HMMA
// Flash-style self-attention for the UNet's short sequences (S <= 512, 16/32-dim heads) inside the persistent kernel: K / V
// of one (item, head) are staged in shared memory as f16 (rows past S zero-filled up to a multiple of 16), every warp owns
// 16 query rows: S = Q K^T by mma.sync m16n8k16 (Q fragments straight from global memory, K fragments are 32-bit shared
// loads), online softmax over 16-key chunks in f32, O += P V with P re-packed from the score fragments and V fragments
// fetched by ldmatrix.trans.  Replaces the CUDA-core row kernel here: 2.1 GFLOP per attention at batch 8 took 90-230 us
// on the FMA pipe (round-2 trace), a quarter of the UNet's time.
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2h(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int HD>
__device__ __noinline__ void mega_attn_mma(const MegaOp& op, int cta, int ncta, unsigned char* smraw) {
  constexpr int KST = HD + 8;          // halves per staged row: 16-byte aligned, ldmatrix rows hit distinct bank groups
  constexpr int KS = HD / 16;          // k-steps of Q K^T
  constexpr int NT = HD / 8;           // n-tiles of P V
  const int S = op.u.attn.S, heads = op.u.attn.heads, B = op.u.attn.B, qblocks = op.u.attn.qblocks;
  const __half* q = op.u.attn.q;
  const __half* k = op.u.attn.k;
  const __half* v = op.u.attn.v;
  __half* out = op.u.attn.out;
  const float scale = op.u.attn.scale * 1.4426950408889634f;   // scores in log2 units: exp2f below
  const int S16 = (S + 15) & ~15;
  __half* Ks = reinterpret_cast<__half*>(smraw);
  __half* Vs = Ks + (size_t)S16 * KST;
  const int C = heads * HD;
  const int L = op.u.attn.ld ? op.u.attn.ld : C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qr = lane >> 2, qc = (lane & 3) * 2;   // fragment row / column pair of this lane
  const int tiles = S16 >> 4;                       // 16-row query tiles per (item, head)
  const int tiles_per_vb = (tiles + qblocks - 1) / qblocks;
  const int nvb = qblocks * heads * B;
  int staged = -1;
  for (int vb = cta; vb < nvb; vb += ncta) {
    const int qb = vb % qblocks, h = (vb / qblocks) % heads, b = vb / (qblocks * heads);
    const long long base = (long long)b * S * C + (long long)h * HD;
    const long long ibase = (long long)b * S * L + (long long)h * HD;
    if (staged != b * heads + h) {
      simt_sync();
      for (int i = threadIdx.x; i < S16 * (HD / 8); i += MEGA_SIMT) {
        const int j = i / (HD / 8), c8 = (i % (HD / 8)) * 8;
        uint4 kv = make_uint4(0u, 0u, 0u, 0u), vv = kv;
        if (j < S) {
          kv = __ldcg(reinterpret_cast<const uint4*>(k + ibase + (long long)j * L + c8));
          vv = __ldcg(reinterpret_cast<const uint4*>(v + ibase + (long long)j * L + c8));
        }
        *reinterpret_cast<uint4*>(Ks + j * KST + c8) = kv;
        *reinterpret_cast<uint4*>(Vs + j * KST + c8) = vv;
      }
      staged = b * heads + h;
      simt_sync();
    }
    const int t_lo = qb * tiles_per_vb, t_hi = min(tiles, t_lo + tiles_per_vb);
    for (int t = t_lo + warp; t < t_hi; t += MEGA_NW) {
      const int r0 = t * 16 + qr, r1 = r0 + 8;      // the two query rows this lane holds fragments of
      uint32_t qa[KS][4];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int c0 = ks * 16 + qc;
        qa[ks][0] = r0 < S ? __ldcg(reinterpret_cast<const uint32_t*>(q + ibase + (long long)r0 * L + c0)) : 0u;
        qa[ks][1] = r1 < S ? __ldcg(reinterpret_cast<const uint32_t*>(q + ibase + (long long)r1 * L + c0)) : 0u;
        qa[ks][2] = r0 < S ? __ldcg(reinterpret_cast<const uint32_t*>(q + ibase + (long long)r0 * L + c0 + 8)) : 0u;
        qa[ks][3] = r1 < S ? __ldcg(reinterpret_cast<const uint32_t*>(q + ibase + (long long)r1 * L + c0 + 8)) : 0u;
      }
      float o[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
      float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
      for (int j0 = 0; j0 < S16; j0 += 16) {
        // scores of 16 keys: two n8 tiles
        float s[2][4];
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
          s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
          const __half* kr = Ks + (j0 + nb * 8 + qr) * KST + qc;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks)
            mma16816(s[nb], qa[ks], *reinterpret_cast<const uint32_t*>(kr + ks * 16), *reinterpret_cast<const uint32_t*>(kr + ks * 16 + 8));
        }
        // scale, mask the zero-filled keys, running maximum per row (a row lives in the 4 lanes of a quad)
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const bool ok = j0 + nb * 8 + qc + e < S;
            s[nb][e] = ok ? s[nb][e] * scale : -INFINITY;
            s[nb][2 + e] = ok ? s[nb][2 + e] * scale : -INFINITY;
            mx0 = fmaxf(mx0, s[nb][e]); mx1 = fmaxf(mx1, s[nb][2 + e]);
          }
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float n0 = fmaxf(m0, mx0), n1 = fmaxf(m1, mx1);   // finite from the first chunk on (key 0 is never masked)
        const float f0 = exp2f(m0 - n0), f1 = exp2f(m1 - n1);
        m0 = n0; m1 = n1;
        float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            s[nb][e] = exp2f(s[nb][e] - n0); ps0 += s[nb][e];
            s[nb][2 + e] = exp2f(s[nb][2 + e] - n1); ps1 += s[nb][2 + e];
          }
        }
        l0 = l0 * f0 + ps0; l1 = l1 * f1 + ps1;   // per-lane partial row sums; combined across the quad at the end
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { o[nt][0] *= f0; o[nt][1] *= f0; o[nt][2] *= f1; o[nt][3] *= f1; }
        // P (16 x 16 keys) as the A operand: score fragments re-packed to f16
        uint32_t pa[4];
        pa[0] = pack2h(s[0][0], s[0][1]); pa[1] = pack2h(s[0][2], s[0][3]);
        pa[2] = pack2h(s[1][0], s[1][1]); pa[3] = pack2h(s[1][2], s[1][3]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          uint32_t b0, b1;
          const uint32_t addr = smem_u32(Vs + (j0 + (lane & 15)) * KST + nt * 8);
          asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(addr));
          mma16816(o[nt], pa, b0, b1);
        }
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        if (r0 < S) *reinterpret_cast<uint32_t*>(out + base + (long long)r0 * C + nt * 8 + qc) = pack2h(o[nt][0] * i0, o[nt][1] * i0);
        if (r1 < S) *reinterpret_cast<uint32_t*>(out + base + (long long)r1 * C + nt * 8 + qc) = pack2h(o[nt][2] * i1, o[nt][3] * i1);
      }
    }
  }
}

__device__ __forceinline__ uint2 pack4h(float a, float b, float c, float d) {
  __half2 h0 = __floats2half2_rn(a, b), h1 = __floats2half2_rn(c, d);
  uint2 pk;
  pk.x = *reinterpret_cast<unsigned*>(&h0);
  pk.y = *reinterpret_cast<unsigned*>(&h1);
  return pk;
}
__device__ __forceinline__ float gelu_erf(float g) { return 0.5f * g * (1.0f + erff(g * 0.70710678118654752f)); }

// the element-wise bodies move 16 bytes per access (the planner's channel counts are multiples of 4; build checks it)
__device__ __noinline__ void mega_geglu(const MegaOp& op, int cta, int ncta) {
  const int D = op.u.geglu.D, D4 = D >> 2;
  const long long total4 = op.u.geglu.rows * D4;
  const float* x = op.u.geglu.x;
  __half* out = op.u.geglu.out;
  for (long long i = (long long)cta * MEGA_SIMT + threadIdx.x; i < total4; i += (long long)ncta * MEGA_SIMT) {
    const long long r = i / D4; const int d = (int)(i - r * D4) * 4;
    const float4 a = ldcg4(x + r * 2 * D + d), g = ldcg4(x + r * 2 * D + D + d);
    *reinterpret_cast<uint2*>(out + r * D + d) = pack4h(a.x * gelu_erf(g.x), a.y * gelu_erf(g.y), a.z * gelu_erf(g.z), a.w * gelu_erf(g.w));
  }
}

__device__ __noinline__ void mega_cat(const MegaOp& op, int cta, int ncta) {
  const int C0 = op.u.cat.C0, C = C0 + op.u.cat.C1, C4 = C >> 2;
  const long long total4 = op.u.cat.rows * C4;
  const float* x0 = op.u.cat.x0; const float* x1 = op.u.cat.x1;
  const long long ld0 = op.u.cat.ld0, ld1 = op.u.cat.ld1;
  float* o32 = op.u.cat.o32; __half* o16 = op.u.cat.o16;
  for (long long i = (long long)cta * MEGA_SIMT + threadIdx.x; i < total4; i += (long long)ncta * MEGA_SIMT) {
    const long long r = i / C4; const int c = (int)(i - r * C4) * 4;
    const float4 v = c < C0 ? ldcg4(x0 + r * ld0 + c) : ldcg4(x1 + r * ld1 + (c - C0));
    if (o32) *reinterpret_cast<float4*>(o32 + r * C + c) = v;
    if (o16) *reinterpret_cast<uint2*>(o16 + r * C + c) = pack4h(v.x, v.y, v.z, v.w);
  }
}

__device__ __noinline__ void mega_axpby(const MegaOp& op, int cta, int ncta) {
  const float a = op.u.axpby.a, b = op.u.axpby.b;
  const float* x = op.u.axpby.x;
  const float* y = op.u.axpby.y;
  float* o32 = op.u.axpby.o32; __half* o16 = op.u.axpby.o16;
  const long long n4 = op.u.axpby.n >> 2;
  for (long long i = (long long)cta * MEGA_SIMT + threadIdx.x; i < n4; i += (long long)ncta * MEGA_SIMT) {
    const float4 xv = ldcg4(x + 4 * i);
    float4 v;
    if (y) {
      const float4 yv = ldcg4(y + 4 * i);
      v = make_float4(fmaf(a, xv.x, b * yv.x), fmaf(a, xv.y, b * yv.y), fmaf(a, xv.z, b * yv.z), fmaf(a, xv.w, b * yv.w));
    } else {
      v = make_float4(fmaf(a, xv.x, b), fmaf(a, xv.y, b), fmaf(a, xv.z, b), fmaf(a, xv.w, b));
    }
    if (o32) *reinterpret_cast<float4*>(o32 + 4 * i) = v;
    if (o16) *reinterpret_cast<uint2*>(o16 + 4 * i) = pack4h(v.x, v.y, v.z, v.w);
  }
}

__device__ __noinline__ void mega_time_embed(const MegaOp& op, int cta) {
  if (cta != 0) return;
  const int half = op.u.temb.dim / 2;
  for (int i = threadIdx.x; i < half; i += MEGA_SIMT) {
    const float fr = expf(-9.210340371976184f * (float)i / (float)half);
    const float a = op.u.temb.t * fr;
    op.u.temb.out[i] = cosf(a);
    op.u.temb.out[half + i] = sinf(a);
  }
}

__device__ __noinline__ void mega_splitk_reduce(const TcKernelArgs& ka, int cta, int ncta) {
  tc_reduce_distributed(ka, (long long)cta * MEGA_SIMT + threadIdx.x, (long long)ncta * MEGA_SIMT);
}

// ------------------------------------------------------------------------------------------------ the kernel
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

// op i (and, for a GEMM, its argument block) -> shared-memory slot, asynchronously: issued while op i-1 executes
__device__ __forceinline__ void fetch_op(const MegaOp* ops, const TcKernelArgs* kas, const int* tc_of, int i, MegaOp* s_op, TcKernelArgs* s_ka) {
  constexpr int NOP = (int)(sizeof(MegaOp) / 16), NKA = (int)(sizeof(TcKernelArgs) / 16);
  const int t = threadIdx.x;
  if (t < NOP) cp_async16(reinterpret_cast<uint4*>(s_op) + t, reinterpret_cast<const uint4*>(ops + i) + t);
  const int tc = __ldg(tc_of + i);   // -1 for a CUDA-core op
  if (tc >= 0 && t >= 64 && t < 64 + NKA) cp_async16(reinterpret_cast<uint4*>(s_ka) + (t - 64), reinterpret_cast<const uint4*>(kas + tc) + (t - 64));
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(MEGA_THREADS, 1) unet_mega_kernel(const MegaOp* __restrict__ ops, int n_ops,
                                                                    const TcKernelArgs* __restrict__ kas,
                                                                    const CUtensorMap* __restrict__ maps,
                                                                    const int* __restrict__ tc_of,
                                                                    unsigned int* __restrict__ bar,
                                                                    unsigned long long* __restrict__ trace) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(16) MegaOp s_ops[2];
  __shared__ __align__(16) TcKernelArgs s_kas[2];
  __shared__ double s_red[2][MEGA_NW];
  __shared__ float s_stat[2];
  __shared__ __align__(16) float s_gb[2][64];
  __shared__ uint32_t s_tmem, s_last;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int cta = (int)blockIdx.x, ncta = (int)gridDim.x;

  // barriers live at a FIXED place (the ring geometry changes from op to op); the rings follow, 1024-byte aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint8_t* rings = smem + MEGA_BAR_BYTES;
  TcSmemView sv;
  sv.fullA = bars;
  sv.emptyA = bars + MEGA_MAX_STAGES;
  sv.fullB = bars + 2 * MEGA_MAX_STAGES;
  sv.emptyB = bars + 3 * MEGA_MAX_STAGES;
  sv.acc_full = bars + 4 * MEGA_MAX_STAGES;
  sv.acc_empty = sv.acc_full + 2;
  sv.last_flag = &s_last;

  fetch_op(ops, kas, tc_of, 0, &s_ops[0], &s_kas[0]);
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  int bars_live = 0, live_SA = 0, live_SB = 0;   // the barriers currently initialised (0 = none yet)
  unsigned int epoch = 0;                         // grid barriers passed so far
  unsigned long long* tr = (trace && cta == 0 && tid == 0) ? trace : nullptr;

  for (int i = 0; i < n_ops; ++i) {
    const MegaOp& op = s_ops[i & 1];
    const TcKernelArgs& ka = s_kas[i & 1];
    if (i + 1 < n_ops) fetch_op(ops, kas, tc_of, i + 1, &s_ops[(i + 1) & 1], &s_kas[(i + 1) & 1]);   // lands while this op runs
    if (tr) tr[4 * i] = clock64();
    const int code = op.code;
#pragma unroll
    for (int pr = 0; pr < 3; ++pr) {
      if (op.pf_bytes[pr] > 0) {
        const int lines = (op.pf_bytes[pr] + 127) >> 7;
        const char* base = reinterpret_cast<const char*>(op.pf_ptr[pr]);
        for (int l = cta * MEGA_THREADS + tid; l < lines; l += ncta * MEGA_THREADS)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + ((long long)l << 7)) : "memory");
      }
    }
    if (code == MOP_GEMM_TC) {
      if (tid == 64) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(maps + 2 * op.tc) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(maps + 2 * op.tc + 1) : "memory");
      }
      sv.ringA = rings;
      sv.ringB = sv.ringA + (size_t)ka.SA * ka.a_stage_bytes;
      sv.stage_all = reinterpret_cast<float*>(sv.ringB + (size_t)ka.SB * ka.b_stage_bytes);
      {
        // one barrier per thread: arm this op's set (counts depend on the op).  The previous GEMM's barriers are idle by
        // now — every TMA transaction and every tcgen05.commit arrival was consumed before its epilogue finished — and
        // mbarrier.init simply rewrites the 64-bit state word.
        const int nnew = 2 * ka.SA + 2 * ka.SB + 4;
        if (tid < nnew) {
          tc_init_barrier_k(ka, sv, tid);
          asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
      }
      bars_live = 1; live_SA = ka.SA; live_SB = ka.SB;
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      if (tr) tr[4 * i + 1] = clock64();
      tc_roles(maps + 2 * op.tc, maps + 2 * op.tc + 1, ka, sv, tmem_base, cta, ncta, nullptr);
      tc_fence_before();
    } else {
      if (tr) tr[4 * i + 1] = clock64();
      switch (code) {
        case MOP_GEMV: mega_gemv(op, cta, ncta); break;
        case MOP_GN_STATS: mega_gn_stats(op, cta, ncta, s_red); break;
        case MOP_GN_APPLY: mega_gn_apply(op, cta, ncta, s_stat); break;
        case MOP_GN_FUSED: mega_gn_fused(op, cta, ncta, s_red, s_stat, s_gb); break;
        case MOP_LAYERNORM: mega_layernorm(op, cta, ncta); break;
        case MOP_ATTN:
          if (op.u.attn.mma) {
            if (op.u.attn.hd == 32) mega_attn_mma<32>(op, cta, ncta, rings);
            else mega_attn_mma<16>(op, cta, ncta, rings);
          } else {
            if (op.u.attn.hd == 32) mega_attn<32>(op, cta, ncta, rings);
            else mega_attn<16>(op, cta, ncta, rings);
          }
          break;
        case MOP_GEGLU: mega_geglu(op, cta, ncta); break;
        case MOP_CAT: mega_cat(op, cta, ncta); break;
        case MOP_AXPBY: mega_axpby(op, cta, ncta); break;
        case MOP_TIME_EMBED: mega_time_embed(op, cta); break;
        case MOP_SPLITK_REDUCE: mega_splitk_reduce(ka, cta, ncta); break;
        default: break;
      }
    }
    if (tr) tr[4 * i + 2] = clock64();
    asm volatile("cp.async.wait_all;" ::: "memory");   // the next op's descriptor has landed (this thread's part)
    if (op.sync_after) {
      epoch += 1;
      grid_sync(bar, epoch * (unsigned int)ncta);
    } else {
      __syncthreads();
    }
    if (tr) tr[4 * i + 3] = clock64();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host side
static int mega_nsl(const CatArgs& a) {
  // pixel slices per (item, group): ~2048 elements each (two 16-byte loads per thread, all in flight at once), at most
  // 16 — a function of the op's geometry only
  const long long elems = a.P * ((a.C0 + a.C1) / a.G);
  long long n = (elems + 2047) / 2048;
  if (n > 16) n = 16;
  if (n > a.P) n = a.P;
  return (int)(n < 1 ? 1 : n);
}

bool egr::mega_supports(const egr_op& op) {
  switch (op.code) {
    case EGR_OP_GEMM_TC: case EGR_OP_LAYERNORM: case EGR_OP_GEGLU: case EGR_OP_TIME_EMBED:
      return true;
    case EGR_OP_GEMM_SIMT:   // only the M <= 8 GEMV form (time-embedding MLP, per-block embedding projections)
      return op.i[EGR_I_WO] * op.i[EGR_I_HO] * op.i[EGR_I_BO] <= 8 && op.i[EGR_I_NTAPS] == 1 && op.x0.elem == 0 && (op.i[EGR_I_K] & 3) == 0 &&
             op.i[EGR_I_WZ_BATCH] == 0;
    case EGR_OP_GN_STATS: case EGR_OP_GN_APPLY: {
      const long long C = op.i[EGR_I_C0] + op.i[EGR_I_C1], G = op.i[EGR_I_GROUPS];
      return G > 0 && C % G == 0 && ((C / G) & 3) == 0;
    }
    case EGR_OP_ATTN_SMALL: {
      const long long hd = op.i[EGR_I_HEADDIM];
      return (hd == 16 || hd == 32) && op.i[EGR_I_SEQ] <= 512;
    }
    case EGR_OP_ELTWISE: {
      const int m = (int)op.i[EGR_I_MODE];
      return m == EGR_ELT_CAST16 || m == EGR_ELT_COPY32 || m == EGR_ELT_AXPBY || m == EGR_ELT_SCALE_SHIFT;
    }
    default: return false;
  }
}

int egr::mega_build(const Spaces& sp, const egr_op* ops, TcPrepared* const* tc, int first, int last, MegaRun** out) {
  *out = nullptr;
  int rc = tc_global_init();
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    // static shared memory (op / argument copies, reduction scratch) counts against the same 227 KB
    EGR_CUDA(cudaFuncSetAttribute(unet_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MEGA_DYN_SMEM_MAX));
    attr = true;
  }
  std::vector<MegaOp> mops;
  std::vector<TcKernelArgs> kas;
  std::vector<CUtensorMap> maps;
  size_t ring_bytes = 0;
  for (int i = first; i < last; ++i) {
    const egr_op& op = ops[i];
    MegaOp m;
    memset(&m, 0, sizeof(m));
    m.sync_after = (op.flags & EGR_FLAG_NOSYNC) ? 0 : 1;
    auto bad = [&](const char* why) { return fail(EGR_ERR_UNSUPPORTED, "%s: not supported inside the persistent UNet kernel (%s)", op.name, why); };
    switch (op.code) {
      case EGR_OP_GEMM_TC: {
        const TcPrepared* p = tc[i];
        if (!p) return bad("unprepared GEMM");
        if (p->ka.halo) return bad("halo mode");
        m.code = MOP_GEMM_TC;
        m.tc = (int)kas.size();
        kas.push_back(p->ka);
        if (p->ka.splits > 1 && !p->ka.fold && getenv("EGR_MEGA_NO_DEFER") == nullptr) kas.back().defer = 1;
        maps.push_back(p->tmA);
        maps.push_back(p->tmB);
        const size_t need = (size_t)p->ka.SA * p->ka.a_stage_bytes + (size_t)p->ka.SB * p->ka.b_stage_bytes;
        ring_bytes = need > ring_bytes ? need : ring_bytes;
        if (p->ka.SA > MEGA_MAX_STAGES || p->ka.SB > MEGA_MAX_STAGES) return bad("ring deeper than the barrier block");
        break;
      }
      case EGR_OP_GEMM_SIMT: {
        Taps taps;
        rc = gemm_args_from_op(sp, op, &m.u.gemv.g, &taps, &m.u.gemv.a);
        if (rc) return rc;
        const GemmArgs& g = m.u.gemv.g;
        const View& a = m.u.gemv.a;
        const long long npix = (long long)g.Wo * g.Ho * g.Bo;
        bool zero_taps = true;
        for (int t = 0; t < g.ntaps; ++t)
          for (int d = 0; d < 5; ++d) zero_taps = zero_taps && taps.t[t][d] == 0;
        auto al16 = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
        if (!(npix <= 8 && g.ntaps == 1 && zero_taps && a.elem == 0 && a.stride[0] == 1 && g.K <= a.dim[0] && (g.K & 3) == 0 && al16(a.p) &&
              al16(g.W) && (g.wstride_n & 3) == 0 && (a.stride[g.dimW] & 3) == 0 && (a.stride[g.dimH] & 3) == 0 && (a.stride[g.dimB] & 3) == 0 &&
              !g.wz_batch))
          return bad("only the M <= 8 GEMV form of the CUDA-core GEMM");
        m.code = MOP_GEMV;
        m.u.gemv.npix = (int)npix;
        break;
      }
      case EGR_OP_GN_STATS:
      case EGR_OP_GN_APPLY: {
        rc = cat_args(sp, op, &m.u.gn.a);
        if (rc) return rc;
        const CatArgs& a = m.u.gn.a;
        if (((a.C0 + a.C1) / a.G) & 3) return bad("channels per group not a multiple of 4");
        m.u.gn.part = (double*)resolve(sp, op.ptr[EGR_P_STATS]);
        m.u.gn.nsl = mega_nsl(a);
        if (!m.u.gn.part) return bad("null stats buffer");
        if (m.u.gn.nsl > 1 + (int)op.i[EGR_I_AUX1]) return bad("stats buffer smaller than the slice partials");
        static const bool gn_one_op = getenv("EGR_MEGA_GN_TWO_OPS") == nullptr;
        if (op.code == EGR_OP_GN_STATS) {
          // one-op GroupNorm: the moments are computed by the op that applies them (the next plan op, same tensor)
          if (gn_one_op && i + 1 < last && ops[i + 1].code == EGR_OP_GN_APPLY && a.P * ((a.C0 + a.C1) / a.G) <= 65536 && (a.C0 + a.C1) / a.G <= 64) { m.code = 0; break; }
          m.code = MOP_GN_STATS;
          break;
        }
        m.code = (gn_one_op && i > first && ops[i - 1].code == EGR_OP_GN_STATS && a.P * ((a.C0 + a.C1) / a.G) <= 65536 && (a.C0 + a.C1) / a.G <= 64)
                     ? MOP_GN_FUSED : MOP_GN_APPLY;
        m.u.gn.gamma = (const float*)resolve(sp, op.ptr[EGR_P_GAMMA]);
        m.u.gn.beta = (const float*)resolve(sp, op.ptr[EGR_P_BETA]);
        m.u.gn.out32 = (float*)resolve(sp, op.ptr[EGR_P_OUT32]);
        m.u.gn.out16 = (__half*)resolve(sp, op.ptr[EGR_P_OUT16]);
        m.u.gn.eps = (float)op.f[EGR_F_EPS];
        m.u.gn.silu = (int)op.i[EGR_I_MODE];
        if (!m.u.gn.gamma || !m.u.gn.beta || (!m.u.gn.out32 && !m.u.gn.out16)) return bad("null pointer");
        break;
      }
      case EGR_OP_LAYERNORM: {
        m.code = MOP_LAYERNORM;
        m.u.ln.x = (const float*)resolve(sp, op.x0.addr);
        m.u.ln.rows = op.i[EGR_I_ROWS];
        m.u.ln.C = (int)op.i[EGR_I_COLS];
        m.u.ln.eps = (float)op.f[EGR_F_EPS];
        m.u.ln.gamma = (const float*)resolve(sp, op.ptr[EGR_P_GAMMA]);
        m.u.ln.beta = (const float*)resolve(sp, op.ptr[EGR_P_BETA]);
        m.u.ln.out16 = (__half*)resolve(sp, op.ptr[EGR_P_OUT16]);
        m.u.ln.out32 = (float*)resolve(sp, op.ptr[EGR_P_OUT32]);
        if (!m.u.ln.x || !m.u.ln.gamma || !m.u.ln.beta || (!m.u.ln.out16 && !m.u.ln.out32) || m.u.ln.rows <= 0 || m.u.ln.C <= 0) return bad("bad arguments");
        break;
      }
      case EGR_OP_ATTN_SMALL: {
        m.code = MOP_ATTN;
        m.u.attn.q = (const __half*)resolve(sp, op.x0.addr);
        m.u.attn.k = (const __half*)resolve(sp, op.x1.addr);
        m.u.attn.v = (const __half*)resolve(sp, op.ptr[EGR_P_AUX]);
        m.u.attn.out = (__half*)resolve(sp, op.ptr[EGR_P_OUT16]);
        m.u.attn.S = (int)op.i[EGR_I_SEQ]; m.u.attn.heads = (int)op.i[EGR_I_HEADS]; m.u.attn.hd = (int)op.i[EGR_I_HEADDIM];
        m.u.attn.B = (int)op.i[EGR_I_BATCH]; m.u.attn.ld = (int)op.i[EGR_I_AUX0];
        m.u.attn.scale = (float)op.f[EGR_F_ALPHA];
        m.u.attn.mma = getenv("EGR_MEGA_ATTN_SIMT") == nullptr ? 1 : 0;
        if (m.u.attn.mma) {
          // query tiles of 16 rows per (item, head), split over enough virtual blocks to occupy the SMs; K / V are staged
          // once per virtual block, so no more splits than needed (a function of the op's geometry and the SM count)
          const int tiles = (m.u.attn.S + 15) / 16, pairs = m.u.attn.B * m.u.attn.heads;
          const int sms_ = devinfo().sm_count ? devinfo().sm_count : 148;
          int qb = (sms_ + pairs - 1) / pairs;
          const int max_qb = (tiles + MEGA_NW - 1) / MEGA_NW;   // one tile per warp at least
          if (qb > max_qb) qb = max_qb;
          m.u.attn.qblocks = qb < 1 ? 1 : qb;
        } else {
          m.u.attn.qblocks = (m.u.attn.S + 15) / 16;
        }
        const int C_ = m.u.attn.heads * m.u.attn.hd;
        auto al16 = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
        if (!m.u.attn.q || !m.u.attn.k || !m.u.attn.v || !m.u.attn.out) return bad("null pointer");
        if (!((m.u.attn.hd == 32 || m.u.attn.hd == 16) && m.u.attn.S <= 512 && C_ % 8 == 0 && al16(m.u.attn.q) && al16(m.u.attn.k) && al16(m.u.attn.v)))
          return bad("attention shape outside the 16/32-dim head, S <= 512 kernel");
        if (m.u.attn.ld != 0 && (m.u.attn.ld < C_ || m.u.attn.ld % 8 != 0)) return bad("bad q/k/v row stride");
        const size_t need = (size_t)2 * ((m.u.attn.S + 15) / 16 * 16) * (m.u.attn.hd + 8) * sizeof(__half);
        ring_bytes = need > ring_bytes ? need : ring_bytes;
        break;
      }
      case EGR_OP_GEGLU: {
        m.code = MOP_GEGLU;
        m.u.geglu.x = (const float*)resolve(sp, op.x0.addr);
        m.u.geglu.out = (__half*)resolve(sp, op.ptr[EGR_P_OUT16]);
        m.u.geglu.rows = op.i[EGR_I_ROWS]; m.u.geglu.D = (int)op.i[EGR_I_COLS];
        if (!m.u.geglu.x || !m.u.geglu.out || m.u.geglu.rows <= 0 || m.u.geglu.D <= 0) return bad("bad arguments");
        if ((m.u.geglu.D & 3) || reinterpret_cast<uintptr_t>(m.u.geglu.x) % 16 || reinterpret_cast<uintptr_t>(m.u.geglu.out) % 8) return bad("GEGLU width / alignment");
        break;
      }
      case EGR_OP_TIME_EMBED: {
        m.code = MOP_TIME_EMBED;
        m.u.temb.out = (float*)resolve(sp, op.ptr[EGR_P_OUT32]);
        m.u.temb.dim = (int)op.i[EGR_I_COLS];
        m.u.temb.t = (float)op.f[EGR_F_A];
        if (!m.u.temb.out || m.u.temb.dim <= 0 || (m.u.temb.dim & 1)) return bad("bad arguments");
        break;
      }
      case EGR_OP_ELTWISE: {
        const int mode = (int)op.i[EGR_I_MODE];
        const float* x0 = (const float*)resolve(sp, op.x0.addr);
        const float* x1 = (const float*)resolve(sp, op.x1.addr);
        float* o32 = (float*)resolve(sp, op.ptr[EGR_P_OUT32]);
        __half* o16 = (__half*)resolve(sp, op.ptr[EGR_P_OUT16]);
        if (!x0 || (!o32 && !o16)) return bad("bad arguments");
        if (mode == EGR_ELT_CAST16 || mode == EGR_ELT_COPY32) {
          m.code = MOP_CAT;
          m.u.cat.x0 = x0; m.u.cat.x1 = x1;
          m.u.cat.C0 = (int)op.i[EGR_I_C0]; m.u.cat.C1 = (int)op.i[EGR_I_C1];
          m.u.cat.rows = op.i[EGR_I_ROWS];
          m.u.cat.ld0 = op.i[EGR_I_AUX0] ? op.i[EGR_I_AUX0] : m.u.cat.C0;
          m.u.cat.ld1 = op.i[EGR_I_AUX1] ? op.i[EGR_I_AUX1] : m.u.cat.C1;
          m.u.cat.o32 = o32; m.u.cat.o16 = o16;
          if (m.u.cat.C1 > 0 && !x1) return bad("null x1");
          if ((m.u.cat.C0 & 3) || (m.u.cat.C1 & 3) || (m.u.cat.ld0 & 3) || (m.u.cat.ld1 & 3) || reinterpret_cast<uintptr_t>(x0) % 16 ||
              reinterpret_cast<uintptr_t>(x1) % 16 || reinterpret_cast<uintptr_t>(o32) % 16 || reinterpret_cast<uintptr_t>(o16) % 8)
            return bad("cast / concat widths or alignment");
        } else if (mode == EGR_ELT_AXPBY || mode == EGR_ELT_SCALE_SHIFT) {
          m.code = MOP_AXPBY;
          m.u.axpby.x = x0; m.u.axpby.y = mode == EGR_ELT_AXPBY ? x1 : nullptr;
          if (mode == EGR_ELT_AXPBY && !x1) return bad("null x1");
          m.u.axpby.a = (float)op.f[EGR_F_A]; m.u.axpby.b = (float)op.f[EGR_F_B];
          m.u.axpby.n = op.i[EGR_I_ROWS];
          if ((m.u.axpby.n & 3) || reinterpret_cast<uintptr_t>(x0) % 16 || reinterpret_cast<uintptr_t>(x1) % 16 ||
              reinterpret_cast<uintptr_t>(o32) % 16 || reinterpret_cast<uintptr_t>(o16) % 8)
            return bad("axpby length / alignment");
          m.u.axpby.o32 = o32; m.u.axpby.o16 = o16;
        } else {
          return bad("eltwise mode");
        }
        break;
      }
      default: return bad("op code");
    }
    if (m.code == 0) continue;   // folded into the next op
    mops.push_back(m);
    if (m.code == MOP_GEMM_TC && kas[m.tc].defer) {   // the grid-wide reduction + epilogue of a split-K layer
      MegaOp rd;
      memset(&rd, 0, sizeof(rd));
      rd.code = MOP_SPLITK_REDUCE;
      rd.tc = m.tc;
      rd.sync_after = m.sync_after;
      mops.back().sync_after = 1;
      mops.push_back(rd);
    }
  }
  if (mops.empty()) return EGR_OK;
  // prefetch chain: op k prefetches [0] the weights of the next weight-reading op after it, [1] [2] the small parameter
  // vectors of op k+1
  if (getenv("EGR_MEGA_NO_PREFETCH") == nullptr) {
    const void* nxt_ptr = nullptr; long long nxt_bytes = 0;
    for (int k = (int)mops.size() - 1; k >= 0; --k) {
      MegaOp& m = mops[k];
      m.pf_ptr[0] = nxt_ptr; m.pf_bytes[0] = nxt_ptr ? (int)(nxt_bytes > (1ll << 30) ? (1ll << 30) : nxt_bytes) : 0;
      if (k + 1 < (int)mops.size()) {
        const MegaOp& nx = mops[k + 1];
        const void* a = nullptr; const void* b = nullptr; long long na = 0, nb = 0;
        if (nx.code == MOP_GEMM_TC && !kas[nx.tc].defer) { a = kas[nx.tc].g.bias; na = 4ll * kas[nx.tc].g.N; }
        else if (nx.code == MOP_SPLITK_REDUCE) { a = kas[nx.tc].g.bias; na = 4ll * kas[nx.tc].g.N; }
        else if (nx.code == MOP_GEMV) { a = nx.u.gemv.g.bias; na = 4ll * nx.u.gemv.g.N; }
        else if (nx.code == MOP_GN_APPLY || nx.code == MOP_GN_FUSED) { a = nx.u.gn.gamma; b = nx.u.gn.beta; na = nb = 4ll * (nx.u.gn.a.C0 + nx.u.gn.a.C1); }
        else if (nx.code == MOP_LAYERNORM) { a = nx.u.ln.gamma; b = nx.u.ln.beta; na = nb = 4ll * nx.u.ln.C; }
        m.pf_ptr[1] = a; m.pf_bytes[1] = a ? (int)na : 0;
        m.pf_ptr[2] = b; m.pf_bytes[2] = b ? (int)nb : 0;
      }
      if (m.code == MOP_GEMM_TC) {
        const GemmArgs& g = kas[m.tc].g;
        if (!g.wz_batch) { nxt_ptr = g.W; nxt_bytes = (long long)g.ntaps * (g.wstride_z > 0 ? g.wstride_z : g.wstride_n * g.N) * 2; }
      } else if (m.code == MOP_GEMV) {
        const GemmArgs& g = m.u.gemv.g;
        nxt_ptr = g.W; nxt_bytes = (long long)g.N * g.wstride_n * 4;
      }
    }
  }
  MegaRun* r = new MegaRun();
  r->first = first; r->last = last; r->n_ops = (int)mops.size(); r->n_tc = (int)kas.size();
  for (const MegaOp& m : mops) r->n_sync += m.sync_after;
  auto bail = [&](int code) { mega_free(r); return code; };
  if (cudaMalloc(&r->d_ops, mops.size() * sizeof(MegaOp)) != cudaSuccess) return bail(fail(EGR_ERR_CUDA, "mega: cudaMalloc of the op table failed"));
  if (cudaMemcpy(r->d_ops, mops.data(), mops.size() * sizeof(MegaOp), cudaMemcpyHostToDevice) != cudaSuccess) return bail(fail(EGR_ERR_CUDA, "mega: op table copy failed"));
  if (!kas.empty()) {
    if (cudaMalloc(&r->d_kas, kas.size() * sizeof(TcKernelArgs)) != cudaSuccess || cudaMalloc(&r->d_maps, maps.size() * sizeof(CUtensorMap)) != cudaSuccess)
      return bail(fail(EGR_ERR_CUDA, "mega: cudaMalloc of the GEMM tables failed"));
    if (cudaMemcpy(r->d_kas, kas.data(), kas.size() * sizeof(TcKernelArgs), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(r->d_maps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice) != cudaSuccess)
      return bail(fail(EGR_ERR_CUDA, "mega: GEMM table copy failed"));
  }
  if (cudaMalloc(&r->d_bar, 256) != cudaSuccess || cudaMemset(r->d_bar, 0, 256) != cudaSuccess) return bail(fail(EGR_ERR_CUDA, "mega: barrier allocation failed"));
  {
    std::vector<int> tc_of(mops.size());
    for (size_t k = 0; k < mops.size(); ++k) tc_of[k] = (mops[k].code == MOP_GEMM_TC || mops[k].code == MOP_SPLITK_REDUCE) ? mops[k].tc : -1;
    if (cudaMalloc(&r->d_tc_of, tc_of.size() * sizeof(int)) != cudaSuccess ||
        cudaMemcpy(r->d_tc_of, tc_of.data(), tc_of.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess)
      return bail(fail(EGR_ERR_CUDA, "mega: op index table allocation failed"));
  }
  if (getenv("EGR_MEGA_TRACE")) {
    if (cudaMalloc(&r->d_trace, mops.size() * 4 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(r->d_trace, 0, mops.size() * 4 * sizeof(unsigned long long)) != cudaSuccess)
      return bail(fail(EGR_ERR_CUDA, "mega: trace allocation failed"));
  }
  r->smem_bytes = 1024 + MEGA_BAR_BYTES + (int)ring_bytes + 8 * STAGE_BYTES_PER_WARP;
  if (r->smem_bytes > MEGA_DYN_SMEM_MAX) return bail(fail(EGR_ERR_UNSUPPORTED, "mega: %d B of shared memory needed", r->smem_bytes));
  const int sms = devinfo().sm_count ? devinfo().sm_count : 148;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, unet_mega_kernel, MEGA_THREADS, r->smem_bytes) != cudaSuccess || per_sm < 1)
    return bail(fail(EGR_ERR_CUDA, "mega: the persistent kernel does not fit an SM (%d B shared memory)", r->smem_bytes));
  r->grid = sms;
  *out = r;
  return EGR_OK;
}

int egr::mega_launch(const MegaRun* r, cudaStream_t st) {
  // barrier state is re-armed by a memset node, so an aborted launch cannot poison the next one
  EGR_CUDA(cudaMemsetAsync(r->d_bar, 0, 8, st));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(r->grid);
  cfg.blockDim = dim3(MEGA_THREADS);
  cfg.dynamicSmemBytes = (size_t)r->smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: they meet at grid barriers
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  const MegaOp* ops = r->d_ops;
  const TcKernelArgs* kas = r->d_kas;
  const CUtensorMap* maps = r->d_maps;
  unsigned int* bar = r->d_bar;
  const int* tc_of = r->d_tc_of;
  unsigned long long* trace = r->d_trace;
  int n = r->n_ops;
  EGR_CUDA(cudaLaunchKernelEx(&cfg, unet_mega_kernel, ops, n, kas, maps, tc_of, bar, trace));
  EGR_CHECK_LAUNCH("unet_mega_kernel");
  return EGR_OK;
}

void egr::mega_describe(const MegaRun* r, int* o) {
  o[0] = r->first; o[1] = r->last; o[2] = r->n_ops; o[3] = r->n_tc; o[4] = r->n_sync; o[5] = r->smem_bytes; o[6] = r->grid;
}

// debug: copy the clock trace of the last launch (4 stamps per op: start, body start, body end, after the barrier) and
// the op codes; returns the number of ops, 0 when tracing is off
int egr::mega_trace(const MegaRun* r, unsigned long long* h_stamps, int* h_codes, int max_ops) {
  if (!r->d_trace) return 0;
  const int n = r->n_ops < max_ops ? r->n_ops : max_ops;
  cudaDeviceSynchronize();
  if (cudaMemcpy(h_stamps, r->d_trace, (size_t)n * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  std::vector<MegaOp> mo(n);
  if (cudaMemcpy(mo.data(), r->d_ops, (size_t)n * sizeof(MegaOp), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  for (int k = 0; k < n; ++k) h_codes[k] = mo[k].code;
  return n;
}

int egr::mega_aborted(const MegaRun* r) {
  unsigned int v = 0;
  if (cudaMemcpy(&v, r->d_bar + 2, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int)v;
}

void egr::mega_free(MegaRun* r) {
  if (!r) return;
  if (r->d_ops) cudaFree(r->d_ops);
  if (r->d_kas) cudaFree(r->d_kas);
  if (r->d_maps) cudaFree(r->d_maps);
  if (r->d_bar) cudaFree(r->d_bar);
  if (r->d_tc_of) cudaFree(r->d_tc_of);
  if (r->d_trace) cudaFree(r->d_trace);
  delete r;
}
