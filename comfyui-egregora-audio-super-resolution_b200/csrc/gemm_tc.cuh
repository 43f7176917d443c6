// gemm_tc.cuh — device side of the tcgen05 tap-GEMM, shared by its two hosts: gemm_tc.cu (one launch per layer) and
// mega.cu (the persistent UNet kernel, which runs the same warp roles once per GEMM op of its op table).
// See gemm_tc.cu for the description of the kernel.
#pragma once
#include <cuda.h>
#include <cstdlib>
#include "ops.cuh"

namespace egr {

static constexpr int TILE_M = 128;
static constexpr int KBLK = 64;  // f16 elements per smem row = 128 B = one swizzle span
static constexpr int A_BOX_BYTES = TILE_M * KBLK * 2;
static constexpr int HALO_BOX_ROWS = 64;
static constexpr int TC_GMAX = 4;    // weight slots one barrier round of the MMA issuers may cover (TcKernelArgs::gmax <= this)
static constexpr int STAGE_LD = 36;  // floats per row of the epilogue staging tile (16-byte aligned, conflict-free)
static constexpr int STAGE_BYTES_PER_WARP = 32 * STAGE_LD * 4;
static constexpr int MAX_SPLITS = 16;

// division by a launch-time constant as multiply-high + shift (the scalar loops and the per-row index math run
// on single threads: a 32-bit hardware-less division costs ~100+ cycles there)
struct FastDiv {
  unsigned int mul, shr, d;
};
static inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = (unsigned int)(d < 1 ? 1 : d);
  if (f.d == 1) { f.mul = 0; f.shr = 0; return f; }
  unsigned int l = 0;
  while ((1ull << l) < f.d) ++l;  // ceil(log2 d)
  f.shr = l;
  f.mul = (unsigned int)(((1ull << 32) * ((1ull << l) - f.d)) / f.d + 1);
  return f;
}

struct alignas(16) TcKernelArgs {   // (16-byte multiple: the persistent kernel copies it with 16-byte cp.async)
  GemmArgs g;
  Taps taps;
  short tapw[EGR_MAX_TAPS];  // tap offset along dimW (halo mode)
  int mt, halo, n_iss, kchunks, n_outer, n_inner, nboxA, boxA_bytes, a_stage_bytes, b_stage_bytes, SA, SB, tmin;
  int tiles1, tiles_w, tiles_h, tiles_m, tiles_n, splits, outer_per_split, n_work;
  FastDiv d_splits, d_tiles_n, d_tiles_w, d_tiles_h, d_kchunks, d_bw, d_bh, d_c4n;
  int acc_cols;  // TMEM columns of one accumulator buffer = mt * block_n (x splits when folded)
  int pair;      // CTA-pair mode (cta_group::2): a cluster of two CTAs owns a 256-row x BLOCK_N tile — each loads its own
                 // 128 rows of A and HALF of the weight tile, the leader issues M=256 MMAs that read both halves
  int defer;     // split-K partials are reduced by a separate, grid-wide pass (tc_reduce_distributed: the persistent kernel)
  int fold;      // split-K folded into ONE work item: each split accumulates into its own TMEM region, summed in the epilogue
  int nbuf;      // accumulator buffers in TMEM: 2 when 2 * acc_cols <= 512, else 1
  int vec_ok, need_crop, epi_plain;
  int hgroup;    // halo mode: taps per barrier round of the MMA issuers (1 = the per-tap loop)
  int gmax;      // weight slots per barrier round of the MMA issuers (1 = one round per slot; <= TC_GMAX, <= SB)
  int dbg;       // timing experiments only (EGR_TC_DBG_SKIP): bit 0 = no A loads, bit 1 = no B loads (results are garbage)
  float* partial;          // split-K workspace: [tile][split][mt*128][block_n] f32
  unsigned int* counters;  // one per output tile, zero between launches
  unsigned long long* trace;  // debug: clock64() stamps of CTA 0 when non-null
};

struct TcPrepared {
  CUtensorMap tmA, tmB;
  TcKernelArgs ka;
  int smem_bytes;
  int grid;
  size_t partial_bytes;
  int n_counters;
  char name[48];
};


}  // namespace egr

using namespace egr;


// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#ifdef EGR_TC_GUARD   // bring-up build (make GUARD=1): a wait that never completes traps instead of hanging the GPU
  for (uint32_t spins = 0;; ++spins) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) break;
    if (spins > (1u << 24)) __trap();
  }
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
#endif
}
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {   // non-blocking: has this phase completed?
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// ---- CTA-pair (cta_group::2) variants: the barrier that collects a stage's transaction bytes lives in the LEADER CTA
// (cluster rank 0); both CTAs' producers arrive on it remotely, both CTAs' TMA loads complete on it
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar) {   // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((unsigned short)3) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld16b(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld16x256_x4(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld16x256_x2(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// one lane of a fully converged warp; the surrounding code stays warp-uniform so that descriptors and addresses live
// in uniform registers (a lane==0 guard around the whole loop makes ptxas wrap every UTCHMMA / UTMALDG in a
// per-lane "waterfall" loop of R2UR moves, ~100 cycles per instruction)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// K-major, SWIZZLE_128B canonical layout: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (=1),
// descriptor version 1 (Blackwell), layout type 2 (SWIZZLE_128B).  cf. cute::UMMA::SmemDescriptor.
// The 128-byte swizzle XOR is applied by the hardware to the ABSOLUTE shared-memory address (bits 4-6 ^= bits 7-9),
// exactly as TMA wrote it, so a start address moved by whole 128-byte rows (halo mode) or by 32 bytes inside the
// span (k advance) needs no "base offset" field — measured on B200: setting it breaks the row-shifted reads.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// n / d for 0 <= n < 2^31 (Granlund-Montgomery round-up variant: q = (mulhi(n, mul) + n) >> shr, 33-bit safe form)
__device__ __forceinline__ int fdiv(int n, const FastDiv& f) {
  if (f.d == 1) return n;
  const unsigned int t = __umulhi((unsigned int)n, f.mul);
  return (int)((t + (((unsigned int)n - t) >> 1)) >> (f.shr - 1));
}

// ------------------------------------------------------------------------------------------------ work decoding
struct WorkItem {
  int tm, tn, ks;
  int o_begin, o_end;  // outer-step range of this split
  int mt_eff;          // valid sub-tiles
  int w0[2], h0[2], b0[2];
};

__device__ __forceinline__ void decode_work(const TcKernelArgs& ka, int w, WorkItem& wi, int rank = 0) {
  const GemmArgs& g = ka.g;
  int t = w;
  wi.ks = 0;
  if (!ka.fold) { t = fdiv(w, ka.d_splits); wi.ks = w - t * ka.splits; }
  wi.tm = fdiv(t, ka.d_tiles_n);
  wi.tn = t - wi.tm * ka.tiles_n;
  if (ka.fold) { wi.o_begin = 0; wi.o_end = ka.n_outer; }   // all splits of the tile, streamed back to back
  else { wi.o_begin = wi.ks * ka.outer_per_split; wi.o_end = min(ka.n_outer, wi.o_begin + ka.outer_per_split); }
  wi.mt_eff = 0;
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    wi.w0[m] = wi.h0[m] = wi.b0[m] = 0;
    if (m >= ka.mt) continue;
    bool valid;
    int q;
    if (ka.halo) {  // CTA tile = mt*128 consecutive positions along W inside one (h, b) row
      q = wi.tm;
      const int q1 = fdiv(q, ka.d_tiles_w);
      wi.w0[m] = (q - q1 * ka.tiles_w) * (TILE_M * ka.mt) + TILE_M * m; q = q1;
      valid = wi.w0[m] < g.Wo;
    } else {
      q = ka.pair ? wi.tm * 2 + rank : wi.tm * ka.mt + m;   // pair mode: one 128-row tile per CTA of the pair
      valid = q < ka.tiles1;
      const int q1 = fdiv(q, ka.d_tiles_w);
      wi.w0[m] = (q - q1 * ka.tiles_w) * g.bw; q = q1;
    }
    const int q2 = fdiv(q, ka.d_tiles_h);
    wi.h0[m] = (q - q2 * ka.tiles_h) * g.bh;
    wi.b0[m] = q2 * g.bb;
    if (valid) wi.mt_eff = m + 1;
  }
}

// ------------------------------------------------------------------------------------------------ epilogue math
struct RowInfo {
  long long base;   // output index of column 0 of this row (non-transposed) / of n = 0 (transposed)
  long long flat0;  // in-batch flat index of column 0 (crop test)
  int b;
  int ok;
};

__device__ __forceinline__ RowInfo row_info(const TcKernelArgs& ka, const WorkItem& wi, int m, int row) {
  const GemmArgs& g = ka.g;
  int w, h, b;
  if (ka.halo) { w = wi.w0[m] + row; h = wi.h0[m]; b = wi.b0[m]; }
  else {
    const int t = fdiv(row, ka.d_bw), wl = row - t * g.bw;
    const int t2 = fdiv(t, ka.d_bh);
    w = wi.w0[m] + wl; h = wi.h0[m] + (t - t2 * g.bh); b = wi.b0[m] + t2;
  }
  RowInfo r;
  r.ok = (w < g.Wo) && (h < g.Ho) && (b < g.Bo);
  r.b = b;
  const long long pix = (long long)h * g.Wo + w;
  if (g.transposed) {
    r.flat0 = 0;
    r.base = (long long)b * g.out_batch_stride + pix + g.out_offset;
  } else {
    r.flat0 = (long long)h * g.out_h_stride + (long long)w * g.out_pix_stride + g.out_offset;
    r.base = (long long)b * g.out_batch_stride + r.flat0;
  }
  return r;
}

// final value of 4 consecutive columns n..n+3 of one row (vector path: alignment checked on the host)
__device__ __forceinline__ void finish4(const GemmArgs& g, float4 acc, long long idx, int n, const float* rb) {
  float v[4] = {acc.x * g.alpha, acc.y * g.alpha, acc.z * g.alpha, acc.w * g.alpha};
  if (g.bias) {
    const float4 bv = __ldg(reinterpret_cast<const float4*>(g.bias + n));
    v[0] += bv.x; v[1] += bv.y; v[2] += bv.z; v[3] += bv.w;
  }
  if (rb) {
    const float4 bv = __ldg(reinterpret_cast<const float4*>(rb + n));
    v[0] += bv.x; v[1] += bv.y; v[2] += bv.z; v[3] += bv.w;
  }
  if (g.act) {
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = egr_apply_act(v[u], g.act);
  }
  if (g.resid) {
    const float4 rv = __ldg(reinterpret_cast<const float4*>(g.resid + idx));
    v[0] += rv.x; v[1] += rv.y; v[2] += rv.z; v[3] += rv.w;
  }
  if (g.resid2) {
    const float4 rv = __ldg(reinterpret_cast<const float4*>(g.resid2 + idx));
    v[0] += rv.x; v[1] += rv.y; v[2] += rv.z; v[3] += rv.w;
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
__device__ __forceinline__ void finish1(const GemmArgs& g, float acc, long long idx, int n, const float* rb) {
  float v = acc * g.alpha;
  if (g.bias) v += g.bias[n];
  if (rb) v += rb[n];
  v = egr_apply_act(v, g.act);
  if (g.resid) v += g.resid[idx];
  if (g.resid2) v += g.resid2[idx];
  v *= g.post;
  if (g.out32) g.out32[idx] = v;
  if (g.out16) g.out16[idx] = __float2half_rn(v);
}

// ------------------------------------------------------------------------------------------------ deferred split-K
// The last-arriver reduction above runs on ONE CTA per output tile: 16 partial tiles of 64 KB pulled through a single
// SM, a few loads in flight per thread — 30-45 us for the UNet's big-K layers, most of their time (round-2 trace of the
// persistent kernel).  With a grid barrier at hand the reduction is a separate pass over ALL threads of the grid: every
// thread owns at most a few float4 of the output, issues the loads of all its splits at once and adds them in split
// order (the same additions in the same order: identical bits), then applies the epilogue.
__device__ __forceinline__ void tc_reduce_distributed(const TcKernelArgs& ka, long long gtid, long long gthreads) {
  const GemmArgs& g = ka.g;
  const int BN = g.block_n, c4n = BN >> 2;
  const size_t pstride = (size_t)ka.mt * TILE_M * BN;
  const long long per_tile = (long long)ka.mt * TILE_M * c4n;
  const long long total = (long long)ka.tiles_m * ka.tiles_n * per_tile;
  int cur_tile = -1;
  WorkItem wi;
  for (long long e = gtid; e < total; e += gthreads) {
    const int tile_id = (int)(e / per_tile);
    const int r = (int)(e - (long long)tile_id * per_tile);
    if (tile_id != cur_tile) { decode_work(ka, tile_id * ka.splits, wi); cur_tile = tile_id; }
    const int row = fdiv(r, ka.d_c4n), c = (r - row * c4n) * 4;
    const int m = row >> 7;
    if (m >= wi.mt_eff) continue;
    const RowInfo ri = row_info(ka, wi, m, row & 127);
    const int n = wi.tn * BN + c;
    if (!ri.ok || n >= g.N) continue;
    const float* pe = ka.partial + (size_t)tile_id * ka.splits * pstride + (size_t)row * BN + c;
    const float* rb = g.rowbias ? g.rowbias + (long long)ri.b * g.rowbias_stride : nullptr;
    const bool vec = !g.transposed && ka.vec_ok;
    const long long fl = ri.flat0 + n;
    const bool keep = vec && fl >= g.out_lo && fl < g.out_hi;
    // every load this element needs is in flight before the first addition: the partials of all splits, and (vector
    // path) bias / row bias / residuals — one trip to L2 / HBM instead of a chain of them
    float4 v[MAX_SPLITS];
#pragma unroll
    for (int sp = 0; sp < MAX_SPLITS; ++sp)
      if (sp < ka.splits) v[sp] = __ldcg(reinterpret_cast<const float4*>(pe + (size_t)sp * pstride));
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), r4 = b4, s4 = b4, t4 = b4;
    if (keep) {
      if (g.bias) b4 = __ldg(reinterpret_cast<const float4*>(g.bias + n));
      if (rb) r4 = __ldcg(reinterpret_cast<const float4*>(rb + n));
      if (g.resid) s4 = __ldcg(reinterpret_cast<const float4*>(g.resid + ri.base + n));
      if (g.resid2) t4 = __ldcg(reinterpret_cast<const float4*>(g.resid2 + ri.base + n));
    }
    float4 acc = v[0];
#pragma unroll
    for (int sp = 1; sp < MAX_SPLITS; ++sp)
      if (sp < ka.splits) { acc.x += v[sp].x; acc.y += v[sp].y; acc.z += v[sp].z; acc.w += v[sp].w; }
    const float a4[4] = {acc.x, acc.y, acc.z, acc.w};
    if (vec) {
      if (keep) {   // finish4's arithmetic, operand for operand
        float o[4] = {acc.x * g.alpha, acc.y * g.alpha, acc.z * g.alpha, acc.w * g.alpha};
        if (g.bias) { o[0] += b4.x; o[1] += b4.y; o[2] += b4.z; o[3] += b4.w; }
        if (rb) { o[0] += r4.x; o[1] += r4.y; o[2] += r4.z; o[3] += r4.w; }
        if (g.act) {
#pragma unroll
          for (int u = 0; u < 4; ++u) o[u] = egr_apply_act(o[u], g.act);
        }
        if (g.resid) { o[0] += s4.x; o[1] += s4.y; o[2] += s4.z; o[3] += s4.w; }
        if (g.resid2) { o[0] += t4.x; o[1] += t4.y; o[2] += t4.z; o[3] += t4.w; }
#pragma unroll
        for (int u = 0; u < 4; ++u) o[u] *= g.post;
        const long long idx = ri.base + n;
        if (g.out32) *reinterpret_cast<float4*>(g.out32 + idx) = make_float4(o[0], o[1], o[2], o[3]);
        if (g.out16) {
          __half2 h0 = __floats2half2_rn(o[0], o[1]), h1 = __floats2half2_rn(o[2], o[3]);
          uint2 pk;
          pk.x = *reinterpret_cast<unsigned*>(&h0);
          pk.y = *reinterpret_cast<unsigned*>(&h1);
          *reinterpret_cast<uint2*>(g.out16 + idx) = pk;
        }
      }
    } else if (g.transposed) {
      for (int u = 0; u < 4 && n + u < g.N; ++u) finish1(g, a4[u], ri.base + (long long)(n + u) * g.out_n_stride, n + u, rb);
    } else {
      for (int u = 0; u < 4 && n + u < g.N; ++u) {
        const long long fl1 = ri.flat0 + n + u;
        if (fl1 >= g.out_lo && fl1 < g.out_hi) finish1(g, a4[u], ri.base + n + u, n + u, rb);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ the warp roles
// Shared-memory objects of one CTA (the two hosts lay them out differently).
struct TcSmemView {
  uint8_t* ringA;
  uint8_t* ringB;
  float* stage_all;      // 8 x STAGE_BYTES_PER_WARP of epilogue staging
  uint64_t *fullA, *emptyA, *fullB, *emptyB, *acc_full, *acc_empty;
  uint32_t* last_flag;
};

// barriers of one GEMM (counts depend on the op: issuers, halo mode); one thread calls this, then fences + a CTA barrier
__device__ __forceinline__ void tc_init_barriers(const TcKernelArgs& ka, const TcSmemView& sv) {
  // halo mode: A and B rings advance at different rates and have their own barriers; otherwise the A stage rides
  // on the B barriers (both producers arrive on fullB, one wait and one commit per k-step for the issuers)
  for (int s = 0; s < ka.SA; ++s) { mbar_init(&sv.fullA[s], 1); mbar_init(&sv.emptyA[s], (uint32_t)ka.n_iss); }
  // pair mode: the leader's two producers arrive on its stage barrier (expecting both CTAs' bytes), sixteen epilogue warps hand a buffer back
  for (int s = 0; s < ka.SB; ++s) { mbar_init(&sv.fullB[s], ka.halo ? 1u : 2u); mbar_init(&sv.emptyB[s], (uint32_t)ka.n_iss); }
  for (int s = 0; s < 2; ++s) { mbar_init(&sv.acc_full[s], (uint32_t)ka.n_iss); mbar_init(&sv.acc_empty[s], ka.pair ? 16u : 8u); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// the same, one barrier per calling thread (k = 0 .. 2*SA + 2*SB + 3); fences + CTA barrier are the caller's business
__device__ __forceinline__ void tc_init_barrier_k(const TcKernelArgs& ka, const TcSmemView& sv, int k) {
  if (k < ka.SA) mbar_init(&sv.fullA[k], 1);
  else if ((k -= ka.SA) < ka.SA) mbar_init(&sv.emptyA[k], (uint32_t)ka.n_iss);
  else if ((k -= ka.SA) < ka.SB) mbar_init(&sv.fullB[k], ka.halo ? 1u : 2u);
  else if ((k -= ka.SB) < ka.SB) mbar_init(&sv.emptyB[k], (uint32_t)ka.n_iss);
  else if ((k -= ka.SB) < 2) mbar_init(&sv.acc_full[k], (uint32_t)ka.n_iss);
  else if ((k -= 2) < 2) mbar_init(&sv.acc_empty[k], 8);
}

// All 12 warps of the CTA call this with initialised barriers and allocated TMEM; CTA `cta` of `ncta` takes the work items
// cta, cta + ncta, ...  Returns when every role has finished its items (no trailing CTA barrier).
template <bool PAIR = false>
__device__ __forceinline__ void tc_roles(const CUtensorMap* tmAp, const CUtensorMap* tmBp, const TcKernelArgs& ka,
                                         const TcSmemView& sv, uint32_t tmem_base, int cta, int ncta, unsigned long long* tr,
                                         int rank = 0) {
  const GemmArgs& g = ka.g;
  const int BN = g.block_n;
  uint8_t* const ringA = sv.ringA;
  uint8_t* const ringB = sv.ringB;
  float* const stage_all = sv.stage_all;
  uint64_t* const fullA = sv.fullA; uint64_t* const emptyA = sv.emptyA;
  uint64_t* const fullB = sv.fullB; uint64_t* const emptyB = sv.emptyB;
  uint64_t* const acc_full = sv.acc_full; uint64_t* const acc_empty = sv.acc_empty;
  uint32_t* const last_flag = sv.last_flag;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    // ============================================================ A producer (whole warp, one elected lane issues)
    if (tr && lane == 0) tr[5] = clock64();
    const uint32_t ringA_u = smem_u32(ringA);
    // pair mode: the "full" barrier of a stage is the LEADER's (both CTAs' loads complete on it); "empty" stays local
    const uint32_t fullA_l = smem_u32(ka.halo ? fullA : fullB);
    const uint32_t fullA_u = PAIR ? mapa_u32(fullA_l, 0u) : fullA_l, emptyA_u = smem_u32(ka.halo ? emptyA : emptyB);
    int sa = 0;
    uint32_t pa = 0;  // ring phase
    int tcount = 0;
    for (int w = cta; w < ka.n_work; w += ncta) {
      WorkItem wi;
      decode_work(ka, w, wi, rank);
      const int nA = ka.halo ? ka.nboxA : wi.mt_eff;
      // per-sub-tile base coordinates, selected without dynamic register indexing
      int cb[2][5];
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int d = 0; d < 5; ++d)
          cb[m][d] = (g.dimW == d ? wi.w0[m] : 0) + (g.dimH == d ? wi.h0[m] : 0) + (g.dimB == d ? wi.b0[m] : 0);
      int tap = 0, kc = wi.o_begin;
      if (!ka.halo) { tap = fdiv(wi.o_begin, ka.d_kchunks); kc = wi.o_begin - tap * ka.kchunks; }
      if (tr && lane == 0 && tcount == 0) tr[6] = clock64();
      for (int io = wi.o_begin; io < wi.o_end; ++io) {
        mbar_wait(emptyA_u + 8 * sa, pa ^ 1u);
        if (tr && lane == 0 && tcount == 0) tr[7] = clock64();
        const uint32_t dstA = ringA_u + (uint32_t)sa * (uint32_t)ka.a_stage_bytes;
        if (elect_one()) {
          if (tr && tcount < 250) tr[16 + 2 * tcount] = clock64();
          if (PAIR) {
            // only the LEADER arrives, expecting the bytes of both CTAs (the peer's loads complete on the leader's barrier
            // too).  A remote release-arrive per stage from the peer throttled the whole pair to one k-step per ~1500
            // cycles whatever the tile width or ring depth (round-2 probe: 474 / 777 / 1517 us at BLOCK_N 256 / 128 / 64).
            if (rank == 0) {
              const int peer_ok = (wi.tm * 2 + 1 < ka.tiles1) ? 1 : 0;
              mbar_expect_tx(fullA_l + 8 * sa, (ka.dbg & 1) ? 0u : (uint32_t)((nA + peer_ok) * ka.boxA_bytes));
            }
          } else {
            mbar_expect_tx(fullA_u + 8 * sa, (ka.dbg & 1) ? 0u : (uint32_t)(nA * ka.boxA_bytes));
          }
          if (ka.dbg & 1) {
          } else if (ka.halo) {
            int c[5];
#pragma unroll
            for (int d = 0; d < 5; ++d) c[d] = cb[0][d] + (g.dimW == d ? ka.tmin : 0);
            c[0] = kc * KBLK;
            for (int bx = 0; bx < nA; ++bx) {
              const int sh = bx * HALO_BOX_ROWS;
              tma_load_5d(dstA + (uint32_t)bx * (uint32_t)ka.boxA_bytes, tmAp, fullA_u + 8 * sa, c[0],
                          c[1] + (g.dimW == 1 ? sh : 0), c[2] + (g.dimW == 2 ? sh : 0), c[3] + (g.dimW == 3 ? sh : 0),
                          c[4] + (g.dimW == 4 ? sh : 0));
            }
          } else {
            int t[5];
#pragma unroll
            for (int d = 0; d < 5; ++d) t[d] = ka.taps.t[tap][d];
            t[0] += kc * KBLK;
#pragma unroll
            for (int m = 0; m < 2; ++m) {
              if (m < nA) {
                if (PAIR)
                  tma_load_5d_2sm(dstA + (uint32_t)m * A_BOX_BYTES, tmAp, fullA_u + 8 * sa, t[0] + cb[m][0], t[1] + cb[m][1],
                                  t[2] + cb[m][2], t[3] + cb[m][3], t[4] + cb[m][4]);
                else
                  tma_load_5d(dstA + (uint32_t)m * A_BOX_BYTES, tmAp, fullA_u + 8 * sa, t[0] + cb[m][0], t[1] + cb[m][1],
                              t[2] + cb[m][2], t[3] + cb[m][3], t[4] + cb[m][4]);
              }
            }
          }
          if (tr && tcount < 250) tr[16 + 2 * tcount + 1] = clock64();
        }
        __syncwarp();
        ++tcount;
        if (++sa == ka.SA) { sa = 0; pa ^= 1u; }
        if (++kc == ka.kchunks && !ka.halo) { kc = 0; ++tap; }
      }
    }
  } else if (warp == 6) {
    // ============================================================ B producer (weights / batch-indexed operand)
    const uint32_t ringB_u = smem_u32(ringB), emptyB_u = smem_u32(emptyB);
    const uint32_t fullB_u = PAIR ? mapa_u32(smem_u32(fullB), 0u) : smem_u32(fullB);
    int sb = 0;
    uint32_t pb = 0;
    for (int w = cta; w < ka.n_work; w += ncta) {
      WorkItem wi;
      decode_work(ka, w, wi, rank);
      const int n0 = wi.tn * BN + (PAIR ? rank * (BN >> 1) : 0);   // pair mode: this CTA stages its half of the weight rows
      int tap = 0, kc = wi.o_begin;
      if (!ka.halo) { tap = fdiv(wi.o_begin, ka.d_kchunks); kc = wi.o_begin - tap * ka.kchunks; }
      for (int io = wi.o_begin; io < wi.o_end; ++io) {
        for (int ii = 0; ii < ka.n_inner; ++ii) {
          mbar_wait(emptyB_u + 8 * sb, pb ^ 1u);
          if (elect_one()) {
            const int z = g.wz_batch ? wi.b0[0] : (ka.halo ? ii : tap);
            if (PAIR) {
              if (rank == 0) mbar_expect_tx(smem_u32(fullB) + 8 * sb, (ka.dbg & 2) ? 0u : 2u * (uint32_t)ka.b_stage_bytes);   // both halves
              if (!(ka.dbg & 2)) tma_load_3d_2sm(ringB_u + (uint32_t)sb * (uint32_t)ka.b_stage_bytes, tmBp, fullB_u + 8 * sb, kc * KBLK, n0, z);
            } else {
              mbar_expect_tx(fullB_u + 8 * sb, (ka.dbg & 2) ? 0u : (uint32_t)ka.b_stage_bytes);
              if (!(ka.dbg & 2)) tma_load_3d(ringB_u + (uint32_t)sb * (uint32_t)ka.b_stage_bytes, tmBp, fullB_u + 8 * sb, kc * KBLK, n0, z);
            }
          }
          __syncwarp();
          if (++sb == ka.SB) { sb = 0; pb ^= 1u; }
        }
        if (++kc == ka.kchunks && !ka.halo) { kc = 0; ++tap; }
      }
    }
  } else if (warp == 1 || warp == 7) {
    // ============================================================ MMA issuers (whole warp, one elected lane issues)
    // A UTCHMMA costs ~55 issue cycles whatever its N, and every k-step adds a barrier wait and a commit on top, so
    // one issuing thread cannot keep the tensor pipe busy at N <= 128.  With two issuers each owns half of the CTA
    // tile (a sub-tile when MT = 2, a BLOCK_N/2 column half when MT = 1): disjoint TMEM accumulators, so no ordering
    // between them is needed; stages are released when both have committed (empty barriers count n_iss arrivals).
    const int u = warp == 1 ? 0 : 1;
    if (u < ka.n_iss && (!PAIR || rank == 0)) {   // pair mode: only the leader CTA issues (M = 256 across both CTAs)
      const bool split_n = (ka.n_iss == 2 && ka.mt == 1);
      const int NI = split_n ? (BN >> 1) : BN;  // N of one instruction
      const int m_lo = (ka.n_iss == 2 && ka.mt == 2) ? u : 0;
      const int m_hi = (ka.n_iss == 2 && ka.mt == 2) ? u + 1 : ka.mt;
      // rows of the B stage owned by this issuer.  Pair mode: an instruction of width NI takes NI/2 weight rows from EACH
      // CTA's stage, so issuer u starts at local row u*NI/2 (its accumulator columns then hold the leader's rows
      // u*NI/2.. first and the peer's rows u*NI/2.. second: the epilogue maps columns back, see `nb` there)
      const uint32_t b_off = split_n ? (uint32_t)((PAIR ? u * (NI >> 1) : u * NI) * 128) : 0u;
      const uint32_t c_off = split_n ? (uint32_t)(u * NI) : 0u;          // accumulator columns owned by this issuer
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=f16, K-major both, N>>3 @17, M>>4 @24
      const uint32_t idesc = (1u << 4) | ((uint32_t)(NI >> 3) << 17) | ((uint32_t)((PAIR ? 2 * TILE_M : TILE_M) >> 4) << 24);
      const uint32_t ringA_u = smem_u32(ringA), ringB_u = smem_u32(ringB);
      const uint32_t fullA_u = smem_u32(fullA), emptyA_u = smem_u32(emptyA), fullB_u = smem_u32(fullB), emptyB_u = smem_u32(emptyB);
      const uint32_t accF_u = smem_u32(acc_full), accE_u = smem_u32(acc_empty);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int it = 0, tcount = 0;
      // One barrier ROUND covers up to ka.gmax consecutive weight slots.  Measured on B200 (tools/micro/mma_rate.cu): the
      // tensor pipe retires a 128 x 256 x 16 MMA every 128 cycles from a free-running issuer, but every wait -> elect ->
      // descriptor -> issue sequence leaves it idle for 170-480 cycles (the issuing thread cannot run ahead of a barrier
      // it has not seen), so one round per 64-deep k-step ran the pipe at ~50 %.  A round blocks on its first slot only,
      // takes the following ones if their barriers have ALREADY completed (test_wait, never blocking: a starved ring
      // degrades to one slot per round).
      const int gcap = ka.gmax;
      for (int w = cta; w < ka.n_work; w += ncta, ++it) {
        WorkItem wi;
        decode_work(ka, w, wi, rank);
        const int buf = ka.nbuf == 2 ? (it & 1) : 0;
        const int use = ka.nbuf == 2 ? (it >> 1) : it;
        uint32_t acc = tmem_base + (uint32_t)(buf * ka.acc_cols) + c_off;
        mbar_wait(accE_u + 8 * buf, (((uint32_t)use) & 1u) ^ 1u);  // epilogue has drained this buffer
        tc_fence_after();
        uint32_t first = 0;  // 0 until the first MMA of this item (of this split when folded) has been issued
        int in_split = 0;    // folded split-K: outer steps issued into the current split's accumulator
        if (ka.halo && ka.hgroup > 1) {
          // halo layers, taps in fixed groups: every weight slot of the group is waited for (blocking, in ring order), then
          // the elected lane issues the MMAs of all of them back to back — one wait -> elect -> issue sequence per group
          for (int io = wi.o_begin; io < wi.o_end; ++io) {
            if (ka.fold && in_split == ka.outer_per_split) { in_split = 0; first = 0; acc += (uint32_t)(ka.mt * BN); }
            ++in_split;
            mbar_wait(fullA_u + 8 * sa, pa);
            const uint32_t aBase = ringA_u + (uint32_t)sa * (uint32_t)ka.a_stage_bytes;
            for (int ii = 0; ii < ka.n_inner; ii += ka.hgroup) {
              const int g = min(ka.hgroup, ka.n_inner - ii);
              uint64_t bdesc[TC_GMAX];
              uint32_t shift[TC_GMAX], eB[TC_GMAX];
              {
                int s = sb;
                uint32_t par = pb;
#pragma unroll
                for (int j = 0; j < TC_GMAX; ++j) {
                  bdesc[j] = 0; shift[j] = 0; eB[j] = 0;
                  if (j < g) {
                    mbar_wait(fullB_u + 8 * s, par);
                    bdesc[j] = make_smem_desc(ringB_u + (uint32_t)s * (uint32_t)ka.b_stage_bytes + b_off);
                    shift[j] = (uint32_t)((ka.tapw[ii + j] - ka.tmin) * 128);
                    eB[j] = emptyB_u + 8 * s;
                    if (++s == ka.SB) { s = 0; par ^= 1u; }
                  }
                }
                sb = s; pb = par;
              }
              tc_fence_after();
              const bool lastG = (ii + g == ka.n_inner);
              if (elect_one()) {
                if (tr && u == 0 && tcount < 250 && ii == 0) tr[528 + 2 * tcount] = clock64();
#pragma unroll
                for (int j = 0; j < TC_GMAX; ++j) {
                  if (j < g) {
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                      if (m >= m_lo && m < m_hi && m < wi.mt_eff) {
                        const uint64_t adesc = make_smem_desc(aBase + shift[j] + (uint32_t)m * A_BOX_BYTES);
#pragma unroll
                        for (int k = 0; k < KBLK / 16; ++k)
                          tc_mma_f16(acc + (uint32_t)(m * BN), adesc + (uint64_t)(2 * k), bdesc[j] + (uint64_t)(2 * k), idesc, (j ? 1u : first) | (uint32_t)k);
                      }
                    }
                    tc_commit(eB[j]);
                  }
                }
                if (lastG) {
                  tc_commit(emptyA_u + 8 * sa);
                  if (io == wi.o_end - 1) tc_commit(accF_u + 8 * buf);
                  if (tr && u == 0 && tcount < 250) tr[528 + 2 * tcount + 1] = clock64();
                }
              }
              __syncwarp();
              first = 1;
            }
            ++tcount;
            if (++sa == ka.SA) { sa = 0; pa ^= 1u; }
          }
        } else if (gcap <= 1) {
          // one round per weight slot: the tap loop of the halo layers (ptxas unrolls it; multi-slot rounds measured slower there)
          for (int io = wi.o_begin; io < wi.o_end; ++io) {
            if (ka.fold && in_split == ka.outer_per_split) {   // next split: its own TMEM region, accumulation restarts
              in_split = 0; first = 0; acc += (uint32_t)(ka.mt * BN);
            }
            ++in_split;
            if (ka.halo) mbar_wait(fullA_u + 8 * sa, pa);  // otherwise A rides on the B barriers (same stage index)
            const uint32_t aBase = ringA_u + (uint32_t)sa * (uint32_t)ka.a_stage_bytes;
            for (int ii = 0; ii < ka.n_inner; ++ii) {
              mbar_wait(fullB_u + 8 * sb, pb);
              tc_fence_after();
              const uint64_t bdesc = make_smem_desc(ringB_u + (uint32_t)sb * (uint32_t)ka.b_stage_bytes + b_off);
              const uint32_t shift = ka.halo ? (uint32_t)((ka.tapw[ii] - ka.tmin) * 128) : 0u;
              const bool lastB = (ii == ka.n_inner - 1);
              if (elect_one()) {
                if (tr && u == 0 && tcount < 250 && ii == 0) tr[528 + 2 * tcount] = clock64();
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                  if (m >= m_lo && m < m_hi && (PAIR ? m == 0 : m < wi.mt_eff)) {
                    const uint64_t adesc = make_smem_desc(aBase + shift + (uint32_t)m * A_BOX_BYTES);
#pragma unroll
                    for (int k = 0; k < KBLK / 16; ++k) {
                      // advance 16 f16 = 32 B inside the 128 B swizzle span: +2 in the (addr >> 4) field
                      if (PAIR) tc_mma_f16_2sm(acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, first | (uint32_t)k);
                      else tc_mma_f16(acc + (uint32_t)(m * BN), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, first | (uint32_t)k);
                    }
                  }
                }
                if (PAIR) tc_commit_2sm(emptyB_u + 8 * sb); else tc_commit(emptyB_u + 8 * sb);
                if (lastB) {
                  if (ka.halo) tc_commit(emptyA_u + 8 * sa);
                  if (io == wi.o_end - 1) { if (PAIR) tc_commit_2sm(accF_u + 8 * buf); else tc_commit(accF_u + 8 * buf); }
                  if (tr && u == 0 && tcount < 250) tr[528 + 2 * tcount + 1] = clock64();
                }
              }
              __syncwarp();
              first = 1;
              if (++sb == ka.SB) { sb = 0; pb ^= 1u; }
            }
            ++tcount;
            if (++sa == ka.SA) { sa = 0; pa ^= 1u; }
          }
        } else {
          int io = wi.o_begin, ii = 0;
          while (io < wi.o_end) {
            if (ka.fold && (!ka.halo || ii == 0) && in_split == ka.outer_per_split) {   // next split: its own TMEM region
              in_split = 0; first = 0; acc += (uint32_t)(ka.mt * BN);
            }
            int gmax;
            if (ka.halo) {
              if (ii == 0) ++in_split;
              gmax = min(gcap, ka.n_inner - ii);                 // taps of one channel chunk share its A stage
            } else {
              gmax = min(gcap, wi.o_end - io);
              if (ka.fold) gmax = min(gmax, ka.outer_per_split - in_split);
            }
            if (ka.halo && ii == 0) mbar_wait(fullA_u + 8 * sa, pa);
            mbar_wait(fullB_u + 8 * sb, pb);
            int n = 1;
#pragma unroll
            for (int j = 1; j < TC_GMAX; ++j) {
              if (j == n && j < gmax) {
                int s = sb + j;
                uint32_t par = pb;
                if (s >= ka.SB) { s -= ka.SB; par ^= 1u; }
                if (__all_sync(0xffffffffu, mbar_test(fullB_u + 8 * s, par))) n = j + 1;
              }
            }
            tc_fence_after();
            if (elect_one()) {
              if (tr && u == 0 && tcount < 250) tr[528 + 2 * tcount] = clock64();
#pragma unroll
              for (int j = 0; j < TC_GMAX; ++j) {
                if (j < n) {
                  int s = sb + j;
                  if (s >= ka.SB) s -= ka.SB;
                  // descriptors of slot j are formed here, behind the MMAs of slot j-1 already in the pipe's queue
                  const uint64_t bdj = make_smem_desc(ringB_u + (uint32_t)s * (uint32_t)ka.b_stage_bytes + b_off);
                  const uint32_t aB = ka.halo ? ringA_u + (uint32_t)sa * (uint32_t)ka.a_stage_bytes + (uint32_t)((ka.tapw[ii + j] - ka.tmin) * 128)
                                              : ringA_u + (uint32_t)s * (uint32_t)ka.a_stage_bytes;   // A rides on the B barriers (same slot)
#pragma unroll
                  for (int m = 0; m < 2; ++m) {
                    if (m >= m_lo && m < m_hi && (PAIR ? m == 0 : m < wi.mt_eff)) {
                      const uint64_t adj = make_smem_desc(aB + (uint32_t)m * A_BOX_BYTES);
#pragma unroll
                      for (int k = 0; k < KBLK / 16; ++k) {
                        // advance 16 f16 = 32 B inside the 128 B swizzle span: +2 in the (addr >> 4) field
                        const uint32_t accum = (j ? 1u : first) | (uint32_t)k;
                        if (PAIR) tc_mma_f16_2sm(acc, adj + (uint64_t)(2 * k), bdj + (uint64_t)(2 * k), idesc, accum);
                        else tc_mma_f16(acc + (uint32_t)(m * BN), adj + (uint64_t)(2 * k), bdj + (uint64_t)(2 * k), idesc, accum);
                      }
                    }
                  }
                  if (PAIR) tc_commit_2sm(emptyB_u + 8 * s); else tc_commit(emptyB_u + 8 * s);
                  const bool last_of_io = ka.halo ? (ii + j == ka.n_inner - 1) : true;
                  if (last_of_io) {
                    if (ka.halo) tc_commit(emptyA_u + 8 * sa);
                    if ((ka.halo ? io : io + j) == wi.o_end - 1) { if (PAIR) tc_commit_2sm(accF_u + 8 * buf); else tc_commit(accF_u + 8 * buf); }
                  }
                }
              }
              if (tr && u == 0 && tcount < 250) tr[528 + 2 * tcount + 1] = clock64();
            }
            __syncwarp();
            first = 1;
            ++tcount;
            sb += n;
            if (sb >= ka.SB) { sb -= ka.SB; pb ^= 1u; }
            if (ka.halo) {
              ii += n;
              if (ii == ka.n_inner) { ii = 0; ++io; if (++sa == ka.SA) { sa = 0; pa ^= 1u; } }
            } else {
              io += n;
              in_split += n;
            }
        }
        }
      }
    }
  } else if ((warp >= 2 && warp <= 5) || warp >= 8) {
    // ============================================================ epilogue: warps 2..5 (set 0) and 8..11 (set 1)
    // A warp may only read TMEM lanes 32*(warp%4)..+31, so each lane group has one warp per set; the two warps
    // take alternate 32-column blocks of the tile.  The store pass is ALU-latency bound inside a single warp
    // (address math, predicates, FMAs on a dependent chain): a second warp per scheduler hides that latency.
    const int lg = warp & 3;
    const int eset = warp >= 8 ? 1 : 0;
    const int ew = eset * 4 + (eset ? warp - 8 : warp - 2);   // 0..7
    const int et = ew * 32 + lane;                             // epilogue thread id 0..255
    float* stg = stage_all + (size_t)ew * (32 * STAGE_LD);
    // pair mode: the accumulator-drained barrier is the leader's (its MMA warp waits for the epilogues of BOTH CTAs)
    const uint32_t accF_u = smem_u32(acc_full), accE_u = (PAIR && rank != 0) ? mapa_u32(smem_u32(acc_empty), 0u) : smem_u32(acc_empty);
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;  // vector pass: 4 rows x 8 float4 per instruction
    const int nblk = (BN + 31) >> 5;
    const bool has_resid = g.resid != nullptr, has_bias = g.bias != nullptr, has_out16 = g.out16 != nullptr;
    const bool has_resid2 = g.resid2 != nullptr;
    int it = 0;
    for (int w = cta; w < ka.n_work; w += ncta, ++it) {
      WorkItem wi;
      decode_work(ka, w, wi, rank);
      const int buf = ka.nbuf == 2 ? (it & 1) : 0;
      const int n0 = wi.tn * BN;
      mbar_wait(accF_u + 8 * buf, ((uint32_t)(ka.nbuf == 2 ? (it >> 1) : it)) & 1u);
      tc_fence_after();
      if (tr && threadIdx.x == 64 && it < 6) tr[2 + 2 * it] = clock64();
      const int tile_id = wi.tm * ka.tiles_n + wi.tn;
      const size_t pstride = (size_t)ka.mt * TILE_M * BN;
      float* part = (ka.splits > 1 && !ka.fold) ? ka.partial + ((size_t)tile_id * ka.splits + wi.ks) * pstride : nullptr;
      const bool vecpath = ka.vec_ok && !part && !g.transposed;
      const int ntot = wi.mt_eff * nblk;
      const int last_bi = ntot - 1 - ((ntot - 1 - eset) & 1);  // last block of this set (< eset when it has none)
      if (last_bi < eset) {  // nothing to read for this warp: hand the buffer back right away
        tc_fence_before();
        if (lane == 0) { if (PAIR && rank != 0) mbar_arrive_cluster(accE_u + 8 * buf); else mbar_arrive(accE_u + 8 * buf); }
      }
      // NOTE: r[] and the other per-block arrays must only be indexed by compile-time constants (fully unrolled
      // loops): one dynamic index sends the whole array to local memory.
      int m_prev = -1;
      RowInfo ri;
      uint32_t any_ok = 0;
      bool crop_here = false;
      long long base_l0 = 0, flat_l0 = 0;
      int roff = 0, foff = 0;
      int off[8];
      for (int bi = eset; bi < ntot; bi += 2) {
        const int m = bi >= nblk ? 1 : 0;
        const int cb = (bi - m * nblk) * 32;
        if (m != m_prev) {
          m_prev = m;
          ri = row_info(ka, wi, m, lg * 32 + lane);  // this thread's own row
          any_ok = __ballot_sync(0xffffffffu, ri.ok);
          // cropped outputs (transposed-conv margins): only the rows at a clip's edges reach outside [out_lo, out_hi) — a
          // warp whose 32 rows all lie inside takes the same branch-free store passes as an uncropped layer
          crop_here = ka.need_crop && __any_sync(0xffffffffu, ri.ok && (ri.flat0 < g.out_lo || ri.flat0 + (long long)g.N > g.out_hi));
          base_l0 = __shfl_sync(0xffffffffu, ri.base, 0);
          flat_l0 = __shfl_sync(0xffffffffu, ri.flat0, 0);
          roff = (int)(ri.base - base_l0);   // row offsets inside the tile (host checked: fit 32 bits)
          foff = (int)(ri.flat0 - flat_l0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = 4 * i + rsub;
            off[i] = __shfl_sync(0xffffffffu, roff, rr);
            if (!((any_ok >> rr) & 1u)) off[i] = -1;
          }
        }
        uint32_t r[32];
        const bool stamp = tr && threadIdx.x == 64 && it == 1 && bi < 32;   // trace build: stamps of this warp's blocks of item 1
        if (stamp) tr[800 + 5 * (bi >> 1)] = clock64();
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * ka.acc_cols + m * BN + cb);
        const int ncols = min(32, BN - cb);  // BN is a multiple of 16
        if (ncols == 32) tc_ld32(taddr, r); else tc_ld16(taddr, r);
        if (ka.fold) {
          // folded split-K: the splits' accumulators are added in split order — the same f32 additions, in the same
          // order, as the last-arriver reduction of the global-partial path, so both modes give the same bits
          tc_wait_ld();
          for (int sp = 1; sp < ka.splits; ++sp) {
            const uint32_t ta2 = taddr + (uint32_t)(sp * ka.mt * BN);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {   // 16 columns at a time: a 32-register temporary would spill
              if (hf * 16 < ncols) {
                uint32_t r2[16];
                tc_ld16b(ta2 + (uint32_t)(hf * 16), r2);
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) r[hf * 16 + j] = __float_as_uint(__uint_as_float(r[hf * 16 + j]) + __uint_as_float(r2[j]));
              }
            }
          }
        }
        int nb = n0 + cb;
        if (PAIR && ka.n_iss == 2) {   // two issuers in pair mode: column block -> (issuer, CTA half, row) -> output column
          const int NI = BN >> 1, q = NI >> 1;
          const int iu = cb / NI, r = cb - iu * NI, hf = r / q;
          nb = n0 + hf * (BN >> 1) + iu * q + (r - hf * q);
        }
        const int n = nb + c4;
        const bool col_ok = c4 < ncols && n < g.N;
        // bias / residual loads of the vector path are issued under the TMEM load
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 rv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (vecpath && any_ok) {
          if (col_ok && has_bias) bias4 = __ldg(reinterpret_cast<const float4*>(g.bias + n));
          if (has_resid) {
            if (!crop_here) {  // predicated loads, no per-cell branches
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (col_ok && off[i] >= 0) rv[i] = __ldg(reinterpret_cast<const float4*>(g.resid + base_l0 + off[i] + n));
              if (has_resid2) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  if (col_ok && off[i] >= 0) {
                    const float4 r2 = __ldg(reinterpret_cast<const float4*>(g.resid2 + base_l0 + off[i] + n));
                    rv[i].x += r2.x; rv[i].y += r2.y; rv[i].z += r2.z; rv[i].w += r2.w;
                  }
              }
            } else {              // cropped cells (transposed-conv margins) may lie outside the residual tensor
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const long long fl = flat_l0 + __shfl_sync(0xffffffffu, foff, 4 * i + rsub) + n;
                if (col_ok && off[i] >= 0 && fl >= g.out_lo && fl < g.out_hi) {
                  rv[i] = __ldg(reinterpret_cast<const float4*>(g.resid + base_l0 + off[i] + n));
                  if (has_resid2) {
                    const float4 r2 = __ldg(reinterpret_cast<const float4*>(g.resid2 + base_l0 + off[i] + n));
                    rv[i].x += r2.x; rv[i].y += r2.y; rv[i].z += r2.z; rv[i].w += r2.w;
                  }
                }
              }
            }
          }
        }
        tc_wait_ld();
        if (stamp) tr[800 + 5 * (bi >> 1) + 1] = clock64();
        if (bi == last_bi) {  // last TMEM read of this buffer by this warp: hand it back to the MMA warps
          tc_fence_before();
          if (lane == 0) { if (PAIR && rank != 0) mbar_arrive_cluster(accE_u + 8 * buf); else mbar_arrive(accE_u + 8 * buf); }
        }
        if (!any_ok) continue;
        if (g.transposed && !part) {
          // out[b][n][pix]: consecutive rows are consecutive addresses -> already coalesced per column
          if (ri.ok) {
            const float* rb = g.rowbias ? g.rowbias + (long long)ri.b * g.rowbias_stride : nullptr;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols && nb + j < g.N) finish1(g, __uint_as_float(r[j]), ri.base + (long long)(nb + j) * g.out_n_stride, nb + j, rb);
          }
          continue;
        }
        // transpose through shared memory: thread = row  ->  8 lanes per row, 4 rows per instruction
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (j < ncols)
            *reinterpret_cast<float4*>(stg + lane * STAGE_LD + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                               __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        __syncwarp();
        if (stamp) { tr[800 + 5 * (bi >> 1) + 2] = clock64(); tr[800 + 5 * (bi >> 1) + 3] = (unsigned long long)(__float_as_uint(rv[0].x) & 1u) + clock64(); }
        if (part) {
          // raw partial sums, row-major [mt*128][BN]
          if (c4 < ncols) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = 4 * i + rsub;
              if (off[i] >= 0)
                __stcg(reinterpret_cast<float4*>(part + ((size_t)(m * TILE_M + lg * 32 + rr)) * BN + cb + c4),
                       *reinterpret_cast<const float4*>(stg + rr * STAGE_LD + c4));
            }
          }
        } else if (vecpath) {
          // 4 rows x 128 contiguous bytes per store instruction
          if (ka.epi_plain && !crop_here) {
            // plain f32 output (+bias, +residual): branch-free passes, only the store is predicated, so the passes
            // interleave (the general variant below serialises on its per-pass uniform branches)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              float4 a[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(stg + (4 * (4 * half + i) + rsub) * STAGE_LD + c4);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int k = 4 * half + i;
                float4 o;
                o.x = fmaf(a[i].x, g.alpha, bias4.x) + rv[k].x; o.y = fmaf(a[i].y, g.alpha, bias4.y) + rv[k].y;
                o.z = fmaf(a[i].z, g.alpha, bias4.z) + rv[k].z; o.w = fmaf(a[i].w, g.alpha, bias4.w) + rv[k].w;
                if (col_ok && off[k] >= 0) {
                  *reinterpret_cast<float4*>(g.out32 + base_l0 + off[k] + n) = o;
                  if (has_out16) {  // f16 copy for a following tensor-core layer (uniform flag)
                    __half2 h0 = __floats2half2_rn(o.x, o.y), h1 = __floats2half2_rn(o.z, o.w);
                    uint2 pk;
                    pk.x = *reinterpret_cast<unsigned*>(&h0);
                    pk.y = *reinterpret_cast<unsigned*>(&h1);
                    *reinterpret_cast<uint2*>(g.out16 + base_l0 + off[k] + n) = pk;
                  }
                }
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = 4 * i + rsub;
              const int fo = __shfl_sync(0xffffffffu, foff, rr);   // convergent: before any per-lane skip
              const int bb = __shfl_sync(0xffffffffu, ri.b, rr);
              bool ok = col_ok && off[i] >= 0;
              if (crop_here) {
                const long long fl = flat_l0 + fo + n;
                ok = ok && fl >= g.out_lo && fl < g.out_hi;
              }
              if (ok) {
                const float4 a = *reinterpret_cast<const float4*>(stg + rr * STAGE_LD + c4);
                float x[4] = {fmaf(a.x, g.alpha, bias4.x), fmaf(a.y, g.alpha, bias4.y), fmaf(a.z, g.alpha, bias4.z), fmaf(a.w, g.alpha, bias4.w)};
                if (g.rowbias) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.rowbias + (long long)bb * g.rowbias_stride + n));
                  x[0] += b4.x; x[1] += b4.y; x[2] += b4.z; x[3] += b4.w;
                }
                if (g.act) {
#pragma unroll
                  for (int u = 0; u < 4; ++u) x[u] = egr_apply_act(x[u], g.act);
                }
                x[0] = (x[0] + rv[i].x) * g.post; x[1] = (x[1] + rv[i].y) * g.post;
                x[2] = (x[2] + rv[i].z) * g.post; x[3] = (x[3] + rv[i].w) * g.post;
                const long long idx = base_l0 + off[i] + n;
                if (g.out32) *reinterpret_cast<float4*>(g.out32 + idx) = make_float4(x[0], x[1], x[2], x[3]);
                if (g.out16) {
                  __half2 h0 = __floats2half2_rn(x[0], x[1]), h1 = __floats2half2_rn(x[2], x[3]);
                  uint2 pk;
                  pk.x = *reinterpret_cast<unsigned*>(&h0);
                  pk.y = *reinterpret_cast<unsigned*>(&h1);
                  *reinterpret_cast<uint2*>(g.out16 + idx) = pk;
                }
              }
            }
          }
        } else {
          // scalar path (odd alignments): lane = column, one row per pass
#pragma unroll 4
          for (int rr = 0; rr < 32; ++rr) {
            const int o = __shfl_sync(0xffffffffu, roff, rr);
            const int fo = __shfl_sync(0xffffffffu, foff, rr);
            const int bb = __shfl_sync(0xffffffffu, ri.b, rr);
            const int nn = nb + lane;
            const long long fl = flat_l0 + fo + nn;
            if (((any_ok >> rr) & 1u) && lane < ncols && nn < g.N && fl >= g.out_lo && fl < g.out_hi) {
              const float* rb = g.rowbias ? g.rowbias + (long long)bb * g.rowbias_stride : nullptr;
              finish1(g, stg[rr * STAGE_LD + lane], base_l0 + o + nn, nn, rb);
            }
          }
        }
        __syncwarp();
        if (stamp) tr[800 + 5 * (bi >> 1) + 4] = clock64();
      }
      if (part && !ka.defer) {
        // split-K: the last CTA to finish this output tile reduces all partials in split order
        __threadfence();
        epi_bar_sync();
        if (et == 0) {
          const unsigned int old = atomicAdd(ka.counters + tile_id, 1u);
          const unsigned int last = (old == (unsigned int)(ka.splits - 1)) ? 1u : 0u;
          if (last) ka.counters[tile_id] = 0u;  // ready for the next launch
          *last_flag = last;
        }
        epi_bar_sync();
        const bool is_last = *last_flag != 0u;
        epi_bar_sync();  // everyone has read the flag before a later item may overwrite it
        if (is_last) {
          __threadfence();
          // all 128 epilogue threads share the valid elements, whatever TMEM lane group produced them
          const float* pt = ka.partial + (size_t)tile_id * ka.splits * pstride;
          const int c4n = BN >> 2;
          const int total = wi.mt_eff * TILE_M * c4n;
          for (int e = et; e < total; e += 256) {
            const int row = fdiv(e, ka.d_c4n), c = (e - row * c4n) * 4;
            const int m = row >> 7;
            const RowInfo ri = row_info(ka, wi, m, row & 127);
            const int n = n0 + c;
            if (!ri.ok || n >= g.N) continue;
            const float* pe = pt + (size_t)row * BN + c;
            float4 acc = __ldcg(reinterpret_cast<const float4*>(pe));
            int sidx = 1;
            for (; sidx + 4 <= ka.splits; sidx += 4) {  // four independent loads in flight, summed in split order
              const float4 v0 = __ldcg(reinterpret_cast<const float4*>(pe + (size_t)sidx * pstride));
              const float4 v1 = __ldcg(reinterpret_cast<const float4*>(pe + (size_t)(sidx + 1) * pstride));
              const float4 v2 = __ldcg(reinterpret_cast<const float4*>(pe + (size_t)(sidx + 2) * pstride));
              const float4 v3 = __ldcg(reinterpret_cast<const float4*>(pe + (size_t)(sidx + 3) * pstride));
              acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
              acc.x += v1.x; acc.y += v1.y; acc.z += v1.z; acc.w += v1.w;
              acc.x += v2.x; acc.y += v2.y; acc.z += v2.z; acc.w += v2.w;
              acc.x += v3.x; acc.y += v3.y; acc.z += v3.z; acc.w += v3.w;
            }
            for (; sidx < ka.splits; ++sidx) {
              const float4 v = __ldcg(reinterpret_cast<const float4*>(pe + (size_t)sidx * pstride));
              acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            const float* rb = g.rowbias ? g.rowbias + (long long)ri.b * g.rowbias_stride : nullptr;
            const float a4[4] = {acc.x, acc.y, acc.z, acc.w};
            if (g.transposed) {
              for (int u = 0; u < 4 && n + u < g.N; ++u) finish1(g, a4[u], ri.base + (long long)(n + u) * g.out_n_stride, n + u, rb);
            } else if (ka.vec_ok) {
              const long long fl = ri.flat0 + n;
              if (fl >= g.out_lo && fl < g.out_hi) finish4(g, acc, ri.base + n, n, rb);
            } else {
              for (int u = 0; u < 4 && n + u < g.N; ++u) {
                const long long fl = ri.flat0 + n + u;
                if (fl >= g.out_lo && fl < g.out_hi) finish1(g, a4[u], ri.base + n + u, n + u, rb);
              }
            }
          }
        }
      }
      if (tr && threadIdx.x == 64 && it < 6) tr[3 + 2 * it] = clock64();
    }
  }
}
