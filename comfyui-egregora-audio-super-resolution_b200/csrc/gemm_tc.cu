// gemm_tc.cu — the dense contraction of the FlashSR plan on 5th-gen tensor cores (sm_100a):
//
//     out[pix, n] = act(alpha * sum_tap sum_k A[pix + tap, k] * B[tap|batch, n, k] + bias[n] + rowbias[b,n]) + resid
//
// One kernel covers 3x3 / 1x1 / strided 2-D convs, dilated 1-D convs, transposed convs (as 2-tap GEMMs with
// N = stride*Cout), linears and the two attention GEMMs: the A operand is a channels-last f16 activation viewed
// through a rank-5 TMA tensor map, so every tap is just a shifted box load and zero padding is TMA's
// out-of-bounds fill — no im2col buffer ever exists in HBM.
//
// CTA = MT x 128 output pixels (MT = 1 or 2 sub-tiles sharing every B load) x BLOCK_N channels, 192 threads:
//   warp 0      TMA producer  (cp.async.bulk.tensor, SWIZZLE_128B, mbarrier expect_tx)          [UTMALDG]
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=BLOCK_N, K=16 f16)   [UTCHMMA]
//   warps 2..5  epilogue: tcgen05.ld 32x32b -> registers -> bias/act/residual -> vector stores   [LDTM]
// Two shared-memory rings, A and B, each with its own full/empty mbarriers:
//   normal mode : one A stage (MT boxes of 128 rows x 64 ch) and one B stage (BLOCK_N x 64) per (tap, k-chunk)
//   halo mode   : (taps along one axis, 1-D convs) the A stage is a super-tile of MT*128 + tap-span rows loaded
//                 ONCE per k-chunk; every tap multiplies a row-shifted window of it (the UMMA descriptor start
//                 address moves by whole 128-byte rows), so A traffic from L2 drops by the tap count and the
//                 loop streams only the per-tap weight tiles.
// Accumulators: MT x 128 lanes x BLOCK_N f32 columns in TMEM.
#include <cuda.h>
#include <cstdlib>
#include "ops.cuh"

namespace egr {

static constexpr int TILE_M = 128;
static constexpr int KBLK = 64;  // f16 elements per smem row = 128 B = one swizzle span
static constexpr int A_STAGE_BYTES = TILE_M * KBLK * 2;

struct TcPrepared {
  CUtensorMap tmA, tmB;
  GemmArgs g;
  Taps taps;
  int a_rank;
  int mt, halo, n_outer, n_inner, kchunks, nboxA, boxA_bytes, a_stage_bytes, b_stage_bytes, SA, SB, tmin, tiles, tiles_w;
  int tmem_cols, smem_bytes;
  dim3 grid;
  int vec_ok;
  char name[48];
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

}  // namespace egr

using namespace egr;

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

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

// ------------------------------------------------------------------------------------------------ the kernel
struct TcKernelArgs {
  GemmArgs g;
  Taps taps;
  int mt, halo, n_outer, n_inner, kchunks, nboxA, boxA_bytes, a_stage_bytes, b_stage_bytes, SA, SB, tmin, tiles, tiles_w;
  int tmem_cols, vec_ok;
};

static constexpr int HALO_BOX_ROWS = 64;

__device__ __forceinline__ void tile_origin(const GemmArgs& g, const TcKernelArgs& ka, int sub, int& w0, int& h0, int& b0, bool& valid) {
  const int tiles_h = (g.Ho + g.bh - 1) / g.bh;
  if (ka.halo) {  // CTA tile = mt*128 consecutive positions along W inside one (h, b) row
    int mt_ = blockIdx.x;
    w0 = (mt_ % ka.tiles_w) * (128 * ka.mt) + 128 * sub; mt_ /= ka.tiles_w;
    h0 = (mt_ % tiles_h) * g.bh; mt_ /= tiles_h;
    b0 = mt_ * g.bb;
    valid = w0 < g.Wo;
  } else {
    int mt_ = blockIdx.x * ka.mt + sub;
    valid = mt_ < ka.tiles;
    w0 = (mt_ % ka.tiles_w) * g.bw; mt_ /= ka.tiles_w;
    h0 = (mt_ % tiles_h) * g.bh; mt_ /= tiles_h;
    b0 = mt_ * g.bb;
  }
}

__global__ void __launch_bounds__(192, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                          const __grid_constant__ CUtensorMap tmB,
                                                          const __grid_constant__ TcKernelArgs ka) {
  extern __shared__ uint8_t smem_raw[];
  const GemmArgs& g = ka.g;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int BN = g.block_n;
  uint8_t* ringA = smem;
  uint8_t* ringB = smem + (size_t)ka.SA * ka.a_stage_bytes;
  uint64_t* fullA = reinterpret_cast<uint64_t*>(ringB + (size_t)ka.SB * ka.b_stage_bytes);
  uint64_t* emptyA = fullA + ka.SA;
  uint64_t* fullB = emptyA + ka.SA;
  uint64_t* emptyB = fullB + ka.SB;
  uint64_t* accum_bar = emptyB + ka.SB;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < ka.SA; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
    for (int s = 0; s < ka.SB; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)ka.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // valid sub-tiles of this CTA (the second one may fall off the end of the tile list)
  int w0s[2], h0s[2], b0s[2];
  int mt_eff = 0;
  for (int m = 0; m < ka.mt; ++m) {
    bool v;
    tile_origin(g, ka, m, w0s[m], h0s[m], b0s[m], v);
    if (v) mt_eff = m + 1;
  }

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
      const int nA = ka.halo ? ka.nboxA : mt_eff;
      for (int io = 0; io < ka.n_outer; ++io) {
        const int sa = io % ka.SA;
        mbar_wait(&emptyA[sa], (((uint32_t)(io / ka.SA)) & 1u) ^ 1u);
        uint8_t* dstA = ringA + (size_t)sa * ka.a_stage_bytes;
        mbar_expect_tx(&fullA[sa], (uint32_t)(nA * ka.boxA_bytes));
        int tap = 0, kc = io;
        if (!ka.halo) { tap = io / ka.kchunks; kc = io - tap * ka.kchunks; }
        for (int bx = 0; bx < nA; ++bx) {
          int c[5];
          if (ka.halo) {
#pragma unroll
            for (int d = 0; d < 5; ++d) c[d] = 0;
            c[0] = kc * KBLK;
            c[g.dimW] = w0s[0] + ka.tmin + bx * HALO_BOX_ROWS;
            c[g.dimH] += h0s[0]; c[g.dimB] += b0s[0];
          } else {
#pragma unroll
            for (int d = 0; d < 5; ++d) c[d] = ka.taps.t[tap][d];
            c[0] += kc * KBLK;
            c[g.dimW] += w0s[bx]; c[g.dimH] += h0s[bx]; c[g.dimB] += b0s[bx];
          }
          tma_load_5d(dstA + (size_t)bx * ka.boxA_bytes, &tmA, &fullA[sa], c[0], c[1], c[2], c[3], c[4]);
        }
        for (int ii = 0; ii < ka.n_inner; ++ii) {
          const int ib = io * ka.n_inner + ii;
          const int sb = ib % ka.SB;
          mbar_wait(&emptyB[sb], (((uint32_t)(ib / ka.SB)) & 1u) ^ 1u);
          mbar_expect_tx(&fullB[sb], (uint32_t)ka.b_stage_bytes);
          const int z = g.wz_batch ? b0s[0] : (ka.halo ? ii : tap);
          tma_load_3d(ringB + (size_t)sb * ka.b_stage_bytes, &tmB, &fullB[sb], kc * KBLK, n0, z);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=f16, K-major both, N>>3 @17, M>>4 @24
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
      for (int io = 0; io < ka.n_outer; ++io) {
        const int sa = io % ka.SA;
        mbar_wait(&fullA[sa], ((uint32_t)(io / ka.SA)) & 1u);
        tc_fence_after();
        const uint32_t aBase = smem_u32(ringA + (size_t)sa * ka.a_stage_bytes);
        for (int ii = 0; ii < ka.n_inner; ++ii) {
          const int ib = io * ka.n_inner + ii;
          const int sb = ib % ka.SB;
          mbar_wait(&fullB[sb], ((uint32_t)(ib / ka.SB)) & 1u);
          tc_fence_after();
          const uint64_t bdesc = make_smem_desc(smem_u32(ringB + (size_t)sb * ka.b_stage_bytes));
          for (int m = 0; m < mt_eff; ++m) {
            const uint32_t aoff = ka.halo ? (uint32_t)((m * TILE_M + ka.taps.t[ii][g.dimW] - ka.tmin) * 128)
                                          : (uint32_t)(m * A_STAGE_BYTES);
            const uint64_t adesc = make_smem_desc(aBase + aoff);
#pragma unroll
            for (int k = 0; k < KBLK / 16; ++k) {
              // advance 16 f16 = 32 B inside the 128 B swizzle span: +2 in the (addr >> 4) field
              tc_mma_f16(tmem_base + (uint32_t)(m * BN), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                         (io | ii | k) ? 1u : 0u);
            }
          }
          tc_commit(&emptyB[sb]);
        }
        tc_commit(&emptyA[sa]);
      }
      tc_commit(accum_bar);
    }
  } else {
    // epilogue warps 2..5; a warp may only touch TMEM lanes 32*(warp%4) .. +31
    const int lg = warp & 3;
    const int row = lg * 32 + lane;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    for (int m = 0; m < mt_eff; ++m) {
    int w, h, b;
    if (ka.halo) { w = w0s[m] + row; h = h0s[m]; b = b0s[m]; }
    else {
      const int wl = row % g.bw, hl = (row / g.bw) % g.bh, bl = row / (g.bw * g.bh);
      w = w0s[m] + wl; h = h0s[m] + hl; b = b0s[m] + bl;
    }
    const bool row_ok = (w < g.Wo) && (h < g.Ho) && (b < g.Bo);
    const long long pix = (long long)h * g.Wo + w;
    for (int cb = 0; cb < BN; cb += 32) {
      uint32_t r[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(m * BN + cb);
      const int ncols = min(32, BN - cb);  // BN is a multiple of 16
      if (ncols == 32) tc_ld32(taddr, r); else tc_ld16(taddr, r);
      tc_wait_ld();
      const int nb = n0 + cb;
      if (!row_ok) {
        // nothing to store for padded rows
      } else if (g.transposed) {
        for (int j = 0; j < ncols; ++j) {
          const int n = nb + j;
          if (n >= g.N) break;
          float v = __uint_as_float(r[j]) * g.alpha;
          if (g.bias) v += g.bias[n];
          if (g.rowbias) v += g.rowbias[(long long)b * g.rowbias_stride + n];
          v = egr_apply_act(v, g.act);
          const long long idx = (long long)b * g.out_batch_stride + (long long)n * g.out_n_stride + pix + g.out_offset;
          if (g.resid) v += g.resid[idx];
          if (g.out32) g.out32[idx] = v;
          if (g.out16) g.out16[idx] = __float2half_rn(v);
        }
      } else {
      const long long flat0 = pix * g.out_pix_stride + g.out_offset + nb;
      const long long base = (long long)b * g.out_batch_stride + flat0;
      const float* rb = g.rowbias ? g.rowbias + (long long)b * g.rowbias_stride + nb : nullptr;
      if (ka.vec_ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (j >= ncols || nb + j >= g.N) break;
          const long long fl = flat0 + j;
          if (fl < g.out_lo || fl >= g.out_hi) continue;
          float v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = __uint_as_float(r[j + u]) * g.alpha;
          if (g.bias) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(g.bias + nb + j));
            v[0] += bv.x; v[1] += bv.y; v[2] += bv.z; v[3] += bv.w;
          }
          if (rb) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(rb + j));
            v[0] += bv.x; v[1] += bv.y; v[2] += bv.z; v[3] += bv.w;
          }
          if (g.act) {
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = egr_apply_act(v[u], g.act);
          }
          if (g.resid) {
            const float4 rv = __ldg(reinterpret_cast<const float4*>(g.resid + base + j));
            v[0] += rv.x; v[1] += rv.y; v[2] += rv.z; v[3] += rv.w;
          }
          if (g.out32) *reinterpret_cast<float4*>(g.out32 + base + j) = make_float4(v[0], v[1], v[2], v[3]);
          if (g.out16) {
            __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
            uint2 pk;
            pk.x = *reinterpret_cast<unsigned*>(&h0);
            pk.y = *reinterpret_cast<unsigned*>(&h1);
            *reinterpret_cast<uint2*>(g.out16 + base + j) = pk;
          }
        }
      } else {
        for (int j = 0; j < ncols; ++j) {
          const int n = nb + j;
          if (n >= g.N) break;
          const long long fl = flat0 + j;
          if (fl < g.out_lo || fl >= g.out_hi) continue;
          float v = __uint_as_float(r[j]) * g.alpha;
          if (g.bias) v += g.bias[n];
          if (rb) v += rb[j];
          v = egr_apply_act(v, g.act);
          if (g.resid) v += g.resid[base + j];
          if (g.out32) g.out32[base + j] = v;
          if (g.out16) g.out16[base + j] = __float2half_rn(v);
        }
      }
      }
      __syncwarp();  // reconverge before the next .sync.aligned TMEM load
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)ka.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host side
int egr::tc_global_init() {
  if (g_encode) return EGR_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess)
    return fail(EGR_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
  g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  EGR_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  return EGR_OK;
}

static int encode_map(CUtensorMap* tm, void* base, int rank, const long long* dim, const long long* stride_elems,
                      const int* box, const char* what) {
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int d = 0; d < rank; ++d) {
    gdim[d] = (cuuint64_t)dim[d];
    bx[d] = (cuuint32_t)box[d];
    es[d] = 1;
    if (d > 0) gstr[d - 1] = (cuuint64_t)stride_elems[d] * 2;  // f16
    if (box[d] < 1 || box[d] > 256) return fail(EGR_ERR_ARG, "%s: TMA box[%d]=%d out of range", what, d, box[d]);
    if (d > 0 && ((stride_elems[d] * 2) % 16 != 0 || stride_elems[d] <= 0))
      return fail(EGR_ERR_ARG, "%s: TMA stride[%d]=%lld elements is not a positive multiple of 16 bytes", what, d, stride_elems[d]);
  }
  if (reinterpret_cast<uintptr_t>(base) % 16) return fail(EGR_ERR_ARG, "%s: TMA base not 16-byte aligned", what);
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, base, gdim, gstr, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(EGR_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed with CUresult %d", what, (int)r);
  return EGR_OK;
}

int egr::tc_prepare(const Spaces& s, const egr_op& op, TcPrepared** out) {
  int rc = tc_global_init();
  if (rc) return rc;
  TcPrepared* p = new TcPrepared();
  View a;
  rc = gemm_args_from_op(s, op, &p->g, &p->taps, &a);
  if (rc) { delete p; return rc; }
  GemmArgs& g = p->g;
  snprintf(p->name, sizeof(p->name), "%s", op.name);
  auto bail = [&](int code) { delete p; return code; };
  if (a.elem != 1) return bail(fail(EGR_ERR_ARG, "%s: tensor-core path needs an f16 A operand", op.name));
  if (g.block_n < 16 || g.block_n > 256 || g.block_n % 16) return bail(fail(EGR_ERR_ARG, "%s: BLOCK_N=%d must be a multiple of 16 in [16,256]", op.name, g.block_n));
  if (op.i[EGR_I_KBLOCK] && op.i[EGR_I_KBLOCK] != KBLK) return bail(fail(EGR_ERR_UNSUPPORTED, "%s: only K block 64 is built", op.name));
  // halo mode: every tap moves along dimW only and the tile is 128 consecutive W positions
  const int kchunks = ceil_div(g.K, KBLK);
  bool halo = g.ntaps > 1 && !g.wz_batch && g.bw == 128 && g.bh == 1 && g.bb == 1 && getenv("EGR_TC_NO_HALO") == nullptr;
  int tmin = 0, tmax = 0;
  for (int t = 0; t < g.ntaps && halo; ++t) {
    for (int d = 0; d < 5; ++d)
      if (d != g.dimW && p->taps.t[t][d] != 0) halo = false;
    const int o = p->taps.t[t][g.dimW];
    tmin = o < tmin ? o : tmin; tmax = o > tmax ? o : tmax;
  }
  if (halo && tmax - tmin > 128) halo = false;
  const int tiles_w128 = ceil_div(g.Wo, g.bw), tiles_h = ceil_div(g.Ho, g.bh), tiles_b = ceil_div(g.Bo, g.bb);
  const int tiles1 = tiles_w128 * tiles_h * tiles_b;
  const int ntn = ceil_div(g.N, g.block_n);
  const int sms = devinfo().sm_count ? devinfo().sm_count : 148;
  // two sub-tiles per CTA (shared B loads) when that still leaves at least ~a wave of CTAs
  int mt = 1;
  if (!g.wz_batch && g.block_n <= 256 && getenv("EGR_TC_NO_MT2") == nullptr) {
    const long long ctas2 = halo ? (long long)ceil_div(g.Wo, 256) * tiles_h * tiles_b * ntn : (long long)ceil_div(tiles1, 2) * ntn;
    if (ctas2 >= sms) mt = 2;
  }
  p->mt = mt; p->halo = halo ? 1 : 0; p->kchunks = kchunks; p->tmin = tmin;
  p->tiles = tiles1;
  p->tiles_w = halo ? ceil_div(g.Wo, 128 * mt) : tiles_w128;
  if (halo) {
    p->n_outer = kchunks; p->n_inner = g.ntaps;
    p->boxA_bytes = HALO_BOX_ROWS * KBLK * 2;
    p->nboxA = ceil_div(128 * mt + (tmax - tmin), HALO_BOX_ROWS);
  } else {
    p->n_outer = g.ntaps * kchunks; p->n_inner = 1;
    p->boxA_bytes = A_STAGE_BYTES;
    p->nboxA = mt;
  }
  p->a_stage_bytes = p->nboxA * p->boxA_bytes;
  p->b_stage_bytes = g.block_n * KBLK * 2;
  // A map: always rank 5 (missing dims are size 1 with a harmless stride)
  long long dim[5], str[5]; int box[5];
  for (int d = 0; d < 5; ++d) { dim[d] = a.dim[d]; str[d] = a.stride[d]; }
  for (int d = a.rank; d < 5; ++d) { dim[d] = 1; str[d] = dim[d - 1] * str[d - 1]; }
  for (int d = 0; d < 5; ++d) box[d] = 1;
  box[0] = KBLK; box[g.dimW] = halo ? HALO_BOX_ROWS : g.bw; box[g.dimH] = g.bh; box[g.dimB] = g.bb;
  if (g.dimW == g.dimH || g.dimW == g.dimB || g.dimH == g.dimB)
    return bail(fail(EGR_ERR_ARG, "%s: tile dims must be distinct A dims", op.name));
  rc = encode_map(&p->tmA, const_cast<void*>(a.p), 5, dim, str, box, op.name);
  if (rc) return bail(rc);
  // B map: [K, N, Z]
  long long bdim[3] = {g.K, g.N, g.wz_batch ? (long long)g.Bo : (long long)g.ntaps};
  long long bstr[3] = {1, g.wstride_n, g.wstride_z > 0 ? g.wstride_z : g.wstride_n * g.N};
  int bbox[3] = {KBLK, g.block_n, 1};
  rc = encode_map(&p->tmB, const_cast<void*>(g.W), 3, bdim, bstr, bbox, op.name);
  if (rc) return bail(rc);
  if (g.wz_batch && g.bb != 1) return bail(fail(EGR_ERR_ARG, "%s: batch-indexed B operand needs BB == 1", op.name));

  // ring depths inside ~200 KB: B first (it is the stream in halo mode), the rest to A
  const int budget = 200 * 1024;
  int SB, SA;
  if (halo) {
    SB = (96 * 1024) / p->b_stage_bytes; if (SB > 6) SB = 6; if (SB < 2) SB = 2;
    SA = (budget - SB * p->b_stage_bytes) / p->a_stage_bytes; if (SA > 3) SA = 3;
    if (SA < 2) { SA = 2; SB = (budget - 2 * p->a_stage_bytes) / p->b_stage_bytes; }
    if (SB < 2) return bail(fail(EGR_ERR_UNSUPPORTED, "%s: halo tile does not fit shared memory", op.name));
  } else {
    SA = budget / (p->a_stage_bytes + p->b_stage_bytes); if (SA > 8) SA = 8; if (SA < 2) SA = 2;
    SB = SA;
  }
  p->SA = SA; p->SB = SB;
  p->smem_bytes = SA * p->a_stage_bytes + SB * p->b_stage_bytes + 1024 /*align slack*/ + (2 * SA + 2 * SB + 1) * 8 + 16;
  if (p->smem_bytes > 227 * 1024) return bail(fail(EGR_ERR_UNSUPPORTED, "%s: %d B of shared memory needed", op.name, p->smem_bytes));
  int cols = 32;
  while (cols < mt * g.block_n) cols <<= 1;
  p->tmem_cols = cols;
  const int ctas_m = halo ? p->tiles_w * tiles_h * tiles_b : ceil_div(tiles1, mt);
  p->grid = dim3(ctas_m, ntn, 1);
  // vector epilogue needs 16-byte aligned rows of 4
  bool vec = !g.transposed && (g.N % 4 == 0) && (g.out_pix_stride % 4 == 0) && (g.out_batch_stride % 4 == 0) &&
             (g.out_offset % 4 == 0) && (g.out_lo % 4 == 0) && (g.out_hi % 4 == 0) && (g.rowbias_stride % 4 == 0);
  auto al = [](const void* q, int a) { return q == nullptr || reinterpret_cast<uintptr_t>(q) % a == 0; };
  vec = vec && al(g.out32, 16) && al(g.out16, 8) && al(g.resid, 16) && al(g.bias, 16) && al(g.rowbias, 16);
  p->vec_ok = vec ? 1 : 0;
  *out = p;
  return EGR_OK;
}

int egr::tc_launch(const TcPrepared* p, cudaStream_t st) {
  TcKernelArgs ka;
  ka.g = p->g;
  ka.taps = p->taps;
  ka.mt = p->mt; ka.halo = p->halo; ka.n_outer = p->n_outer; ka.n_inner = p->n_inner; ka.kchunks = p->kchunks;
  ka.nboxA = p->nboxA; ka.boxA_bytes = p->boxA_bytes; ka.a_stage_bytes = p->a_stage_bytes; ka.b_stage_bytes = p->b_stage_bytes;
  ka.SA = p->SA; ka.SB = p->SB; ka.tmin = p->tmin; ka.tiles = p->tiles; ka.tiles_w = p->tiles_w;
  ka.tmem_cols = p->tmem_cols;
  ka.vec_ok = p->vec_ok;
  gemm_tc_kernel<<<p->grid, 192, p->smem_bytes, st>>>(p->tmA, p->tmB, ka);
  EGR_CHECK_LAUNCH(p->name);
  return EGR_OK;
}

void egr::tc_free(TcPrepared* p) { delete p; }
