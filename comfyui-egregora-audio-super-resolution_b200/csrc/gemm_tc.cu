// gemm_tc.cu — the dense contraction of the FlashSR plan on 5th-gen tensor cores (sm_100a):
//
//     out[pix, n] = act(alpha * sum_tap sum_k A[pix + tap, k] * B[tap|batch, n, k] + bias[n] + rowbias[b,n]) + resid
//
// One kernel covers 3x3 / 1x1 / strided 2-D convs, dilated 1-D convs, transposed convs (as 2-tap GEMMs with
// N = stride*Cout), linears and the two attention GEMMs: the A operand is a channels-last f16 activation viewed
// through a rank-5 TMA tensor map, so every tap is just a shifted box load and zero padding is TMA's
// out-of-bounds fill — no im2col buffer ever exists in HBM.
//
// PERSISTENT, warp-specialised CTA (one per SM, 384 threads) looping over work items (m-tile, n-tile, k-split):
//   warp 0 / 6  TMA producers of the A / B rings (cp.async.bulk.tensor, SWIZZLE_128B, expect_tx)  [UTMALDG]
//   warp 1 / 7  tcgen05.mma issuers (M=128, K=16 f16), each owning half of the CTA tile; warp 1 allocates TMEM [UTCHMMA]
//   warps 2..5  epilogue: tcgen05.ld 32x32b -> registers -> shared-memory transpose -> bias/act/residual ->
//               row-contiguous (coalesced) vector stores                                         [LDTM]
// The accumulator is double-buffered in TMEM (2 x MT x BLOCK_N f32 columns of the 512), so the epilogue of
// work item i drains buffer i&1 while the tensor pipe already fills the other one with item i+1.
// Two shared-memory rings, A and B, each with its own full/empty mbarriers:
//   normal mode : one A stage (MT boxes of 128 rows x 64 ch) and one B stage (BLOCK_N x 64) per (tap, k-chunk)
//   halo mode   : (taps along one axis, 1-D convs) the A stage is a super-tile of MT*128 + tap-span rows loaded
//                 ONCE per k-chunk; every tap multiplies a row-shifted window of it (the UMMA descriptor start
//                 address moves by whole 128-byte rows), so A traffic from L2 drops by the tap count and the
//                 loop streams only the per-tap weight tiles.
// Split-K: layers with few output tiles (the UNet at 8x4 .. 64x32 latent pixels) would leave most SMs idle while a
// handful of CTAs stream megabytes of weights; their reduction range is cut into `splits` work items, each CTA
// writes its raw f32 partial tile to a library-owned workspace and the LAST one to arrive (one atomic counter per
// output tile) sums the partials in split order — deterministic for a given launch geometry — and applies the
// epilogue.  Splitting is used only when the launch as a whole cannot fill the SMs (small batches).
//
// The single-thread producer / issuer loops are kept free of integer divisions and of dynamically indexed local
// arrays: a clock trace of the previous version showed ~1400 cycles per k-step spent in exactly that scalar code.
#include "gemm_tc.cuh"

namespace egr {
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;
}  // namespace egr

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(384, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                          const __grid_constant__ CUtensorMap tmB,
                                                          const __grid_constant__ TcKernelArgs ka) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  TcSmemView sv;
  sv.ringA = smem;
  sv.ringB = sv.ringA + (size_t)ka.SA * ka.a_stage_bytes;
  sv.stage_all = reinterpret_cast<float*>(sv.ringB + (size_t)ka.SB * ka.b_stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sv.stage_all) + 8 * STAGE_BYTES_PER_WARP);
  sv.fullA = bars;
  sv.emptyA = sv.fullA + ka.SA;
  sv.fullB = sv.emptyA + ka.SA;
  sv.emptyB = sv.fullB + ka.SB;
  sv.acc_full = sv.emptyB + ka.SB;   // [2]
  sv.acc_empty = sv.acc_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sv.acc_empty + 2);
  sv.last_flag = tmem_slot + 1;

  const int warp = threadIdx.x >> 5;
#ifdef EGR_TC_TRACE  // debug build only (make TRACE=1): clock stamps of CTA 0; the optimiser drops all of it otherwise
  unsigned long long* tr = (ka.trace && blockIdx.x == 0) ? ka.trace : nullptr;
  if (tr && threadIdx.x == 0) tr[0] = clock64();
  if (ka.trace && threadIdx.x == 0 && blockIdx.x < 148) {  // per-CTA wall-clock start (ns)
    unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    ka.trace[1100 + 2 * blockIdx.x] = gt;
  }
#else
  unsigned long long* const tr = nullptr;
#endif

  if (threadIdx.x == 64) {  // descriptor fetches overlap the barrier / TMEM set-up instead of delaying the first TMA
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (threadIdx.x == 0) tc_init_barriers(ka, sv);
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tr && threadIdx.x == 0) tr[1] = clock64();
  // programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) touched no global memory and may
  // have run while the previous kernel of the stream was draining; from here on its results are needed
  egr_pdl_sync();

  tc_roles(&tmA, &tmB, ka, sv, tmem_base, (int)blockIdx.x, (int)gridDim.x, tr);

  tc_fence_before();
  __syncthreads();
#ifdef EGR_TC_TRACE
  if (ka.trace && threadIdx.x == 64 && blockIdx.x < 148) {  // per-CTA wall-clock end (ns), after the barrier released
    unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    ka.trace[1100 + 2 * blockIdx.x + 1] = gt;
  }
#endif
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ the CTA-pair kernel
// cta_group::2: a cluster of two CTAs (one TPC) owns a 256-row x BLOCK_N output tile.  Each CTA stages its own 128 rows
// of A and HALF of the weight tile (BLOCK_N/2 rows); the leader's tcgen05.mma reads both halves, so a k-step costs each
// SM 16 KB + BLOCK_N x 64 B of L2 -> shared-memory traffic for 128 x BLOCK_N x 64 MACs — half of what the single-CTA
// tile moves per FLOP.  That traffic, not the tensor pipe, bounds the big 3x3 layers: ncu (profiles/r2_f_ncu_full_summary)
// shows ~8 TB/s of TMA loads at 26-40 % tensor-pipe activity, i.e. ~50 B/clk per SM, the rate B300_MICROARCH.md gives
// for TMA service.  Same warp roles as gemm_tc_kernel (tc_roles<true>): stage barriers live in the leader, commits are
// multicast to both CTAs, both CTAs' epilogues drain their own TMEM lanes.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ TcKernelArgs ka) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  TcSmemView sv;
  sv.ringA = smem;
  sv.ringB = sv.ringA + (size_t)ka.SA * ka.a_stage_bytes;
  sv.stage_all = reinterpret_cast<float*>(sv.ringB + (size_t)ka.SB * ka.b_stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sv.stage_all) + 8 * STAGE_BYTES_PER_WARP);
  sv.fullA = bars;
  sv.emptyA = sv.fullA + ka.SA;
  sv.fullB = sv.emptyA + ka.SA;
  sv.emptyB = sv.fullB + ka.SB;
  sv.acc_full = sv.emptyB + ka.SB;
  sv.acc_empty = sv.acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sv.acc_empty + 2);
  sv.last_flag = tmem_slot + 1;
  const int warp = threadIdx.x >> 5;
  const int rank = (int)cluster_ctarank();
  if (threadIdx.x == 64) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (threadIdx.x == 0) tc_init_barriers(ka, sv);
  if (warp == 1) {   // both CTAs of the pair, same shared-memory slot offset
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / TMA completion can reach them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  egr_pdl_sync();   // launched as a cluster, never as a programmatic dependent itself: only the trigger matters here

  tc_roles<true>(&tmA, &tmB, ka, sv, tmem_base, (int)(blockIdx.x >> 1), (int)(gridDim.x >> 1), nullptr, rank);

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's shared memory and barriers stay alive until the leader's last MMA / commit has landed
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
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
  EGR_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
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

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

int egr::tc_prepare(const Spaces& s, const egr_op& op, TcPrepared** out) {
  int rc = tc_global_init();
  if (rc) return rc;
  TcPrepared* p = new TcPrepared();
  memset(&p->ka, 0, sizeof(p->ka));
  TcKernelArgs& ka = p->ka;
  View a;
  rc = gemm_args_from_op(s, op, &ka.g, &ka.taps, &a);
  if (rc) { delete p; return rc; }
  GemmArgs& g = ka.g;
  snprintf(p->name, sizeof(p->name), "%s", op.name);
  auto bail = [&](int code) { delete p; return code; };
  if (a.elem != 1) return bail(fail(EGR_ERR_ARG, "%s: tensor-core path needs an f16 A operand", op.name));
  if (g.N % 16) return bail(fail(EGR_ERR_ARG, "%s: N=%d must be a multiple of 16", op.name, g.N));
  if (op.i[EGR_I_KBLOCK] && op.i[EGR_I_KBLOCK] != KBLK) return bail(fail(EGR_ERR_UNSUPPORTED, "%s: only K block 64 is built", op.name));
  if (g.dimW == g.dimH || g.dimW == g.dimB || g.dimH == g.dimB)
    return bail(fail(EGR_ERR_ARG, "%s: tile dims must be distinct A dims", op.name));
  if (g.wz_batch && g.bb != 1) return bail(fail(EGR_ERR_ARG, "%s: batch-indexed B operand needs BB == 1", op.name));
  const int sms = devinfo().sm_count ? devinfo().sm_count : 148;
  const int kchunks = ceil_div(g.K, KBLK);

  // ---- halo mode: every tap moves along dimW only and the tile is 128 consecutive W positions
  bool halo = g.ntaps > 1 && !g.wz_batch && g.bw == 128 && g.bh == 1 && g.bb == 1 && env_int("EGR_TC_NO_HALO", 0) == 0;
  int tmin = 0, tmax = 0;
  for (int t = 0; t < g.ntaps; ++t) {
    for (int d = 0; d < 5; ++d)
      if (d != g.dimW && ka.taps.t[t][d] != 0) halo = false;
    const int o = ka.taps.t[t][g.dimW];
    ka.tapw[t] = (short)o;
    tmin = o < tmin ? o : tmin; tmax = o > tmax ? o : tmax;
  }
  if (halo && tmax - tmin > 128) halo = false;
  if (!halo) { tmin = 0; tmax = 0; }
  const int tiles_w128 = ceil_div(g.Wo, g.bw), tiles_h = ceil_div(g.Ho, g.bh), tiles_b = ceil_div(g.Bo, g.bb);
  const int tiles1 = tiles_w128 * tiles_h * tiles_b;
  const int n_outer = halo ? kchunks : g.ntaps * kchunks;
  const int n_inner = halo ? g.ntaps : 1;

  // ---- split-K: a constant of the LAYER, not of the launch.  The split count fixes the f32 summation order of
  // every output element, so it is derived from the geometry of ONE batch item (output tiles of a single item x
  // n-tiles): a chunk-channel then gets bit-identical results whether it runs alone or inside a batch of 8.  (The
  // first version sized the split for the whole launch; alone-vs-batched outputs differed by ~7e-4 RMS through the
  // f16 operand roundings downstream.)  EGR_TC_SPLIT_POLICY=launch restores that rule for A/B timing only.
  static const bool split_by_launch = getenv("EGR_TC_SPLIT_POLICY") && !strcmp(getenv("EGR_TC_SPLIT_POLICY"), "launch");
  const int units = (split_by_launch ? tiles1 : tiles_w128 * tiles_h) * ceil_div(g.N, 128);
  int splits = 1;
  if (!g.wz_batch && env_int("EGR_TC_NO_SPLITK", 0) == 0) {
    splits = sms / (units > 0 ? units : 1);
    if (splits > MAX_SPLITS) splits = MAX_SPLITS;
    // at least sixteen outer steps per split: a shorter reduction is cheaper left whole than split and reduced again (the
    // reduction pass costs several microseconds of latency; swept 8 ... 32 in the plan: batch-8 UNet 6.6 -> 6.1 ms at 16,
    // batch 1 flat)
    static const int split_min = env_int("EGR_TC_SPLIT_MIN", 16) > 0 ? env_int("EGR_TC_SPLIT_MIN", 16) : 16;   // outer steps a split keeps (A/B knob)
    if (splits > n_outer / split_min) splits = n_outer / split_min;
    if (splits < 1) splits = 1;
  }
  if (env_int("EGR_TC_SPLITS", 0) > 0) splits = env_int("EGR_TC_SPLITS", 0) > n_outer ? n_outer : env_int("EGR_TC_SPLITS", 0);
  const int ops = ceil_div(n_outer, splits);
  splits = ceil_div(n_outer, ops);  // no empty split

  // ---- folded split-K.  The split count is a constant of the layer (bit-identical results at any batch size), but a
  // batched launch has enough output tiles to fill the GPU without spreading one tile's reduction over several CTAs.
  // Then ONE work item streams all the splits of its tile back to back, each split into its own TMEM accumulator
  // region, and the epilogue adds the regions in split order: the same f32 additions in the same order as the
  // last-arriver reduction over global partial tiles — without the partial traffic, the atomics and the extra CTAs.
  // Needs splits * BLOCK_N <= 512 TMEM columns and enough (tile, n-block) items to occupy the SMs.
  int fold = 0, fold_bn = 0;
  if (splits > 1 && env_int("EGR_TC_NO_FOLD", 0) == 0) {
    for (int c = 16; c <= 128; c += 16)
      if (g.N % c == 0 && splits * c <= 512) fold_bn = c;
    if (fold_bn && (long long)tiles1 * (g.N / fold_bn) * 2 >= (long long)sms) fold = 1;   // at least half the SMs get an item
  }

  // ---- BLOCK_N and sub-tiles per CTA: cheapest (waves x per-item cycles) under a coarse cost model
  int bn = 0, mt = 1;
  double best = 1e30;
  if (fold) { bn = fold_bn; mt = 1; }
  else {
    const int cands[] = {256, 192, 160, 128, 96, 80, 64, 48, 32, 16};
    for (int c : cands) {
      if (g.N % c) continue;
      for (int m = 1; m <= 2; ++m) {
        if (m * c > 256) continue;
        if (m == 2 && (g.wz_batch || env_int("EGR_TC_NO_MT2", 0))) continue;
        if (splits > 1 && (m == 2 || c > 128)) continue;  // bounded partial-tile workspace
        const long long tiles_m = halo ? (long long)ceil_div(g.Wo, TILE_M * m) * tiles_h * tiles_b : ceil_div(tiles1, m);
        const long long ctas = tiles_m * (g.N / c) * splits;
        const long long waves = (ctas + sms - 1) / sms;
        // per k-step: MMA pace (N < 64 still costs a 64-column pass) vs the L2->SM stream at ~48 B/clk/SM
        const double a_bytes = halo ? (double)(TILE_M * m + (tmax - tmin)) * 128.0 / n_inner : (double)m * A_BOX_BYTES;
        const double step_mma = (double)m * 4.0 * (c > 64 ? c : 64) / 2.0;
        const double step_mem = (a_bytes + c * 128.0) / 48.0;
        const double step = step_mma > step_mem ? step_mma : step_mem;
        const double epi = 600.0 + (double)m * c * 14.0;
        const double item = (double)ops * n_inner * step + 1500.0;
        const double cost = (double)waves * (item > epi ? item : epi) + epi;
        if (cost < best) { best = cost; bn = c; mt = m; }
      }
    }
    if (bn == 0) return bail(fail(EGR_ERR_ARG, "%s: no BLOCK_N fits N=%d", op.name, g.N));
  }
  if (!fold && env_int("EGR_TC_BN", 0) > 0 && g.N % env_int("EGR_TC_BN", 0) == 0) { bn = env_int("EGR_TC_BN", 0); if (mt * bn > 256) mt = 1; }
  // ---- CTA-pair mode (gemm_tc_pair_kernel): big unsplit non-halo layers, same cost model with the pair's traffic (each
  // SM stages 16 KB of A + BLOCK_N x 64 B of weights per k-step) and MMA pace (128 x BLOCK_N per SM)
  int pair = 0;
  if (!fold && !halo && splits == 1 && !g.wz_batch && !(op.flags & EGR_FLAG_MEGA) && tiles1 >= 2 && (sms & 1) == 0 &&
      env_int("EGR_TC_NO_PAIR", 0) == 0) {
    // Measured (round-2 probes, tools/gemm_probe.py, batch 8): a pair k-step costs ~520 + 2.3 * BLOCK_N cycles (a 2-CTA
    // MMA has a ~160-190 cycle floor whatever its N, and wait + commit are not hidden behind the MMAs), so only the
    // 256-wide tile pays: 1090 / 1060 / 1150 TF/s against 935 / 940 / 1010 single-CTA on the 1024-, 256- and 512-channel
    // 3x3 layers, while BLOCK_N = 128 pairs lose (670 against 820).  Rule: pair mode iff a 256-wide tile divides N
    // and the pair items fill at least 80 % of their waves (a function of the launch only: results do not depend on it).
    int pbn = 0;
    double pbest = 1e30;
    for (int c = 256; c >= 32; c -= 32) {
      if (g.N % c) continue;
      if (env_int("EGR_TC_PAIR_BN", 0) > 0 && c != env_int("EGR_TC_PAIR_BN", 0)) continue;
      if (env_int("EGR_TC_PAIR_BN", 0) == 0 && c != 256) continue;
      const long long items = (long long)ceil_div(tiles1, 2) * (g.N / c);
      const long long waves = (items + sms / 2 - 1) / (sms / 2);
      const double fill = (double)items / (double)(waves * (sms / 2));
      if (fill < 0.8 && env_int("EGR_TC_FORCE_PAIR", 0) == 0) continue;
      pbest = 0.0; pbn = c;
      break;
    }
    if (pbn && (pbest < 0.95 * best || env_int("EGR_TC_FORCE_PAIR", 0))) { pair = 1; bn = pbn; mt = 1; }
  }
  g.block_n = bn;

  ka.mt = mt; ka.halo = halo ? 1 : 0; ka.kchunks = kchunks; ka.tmin = tmin;
  ka.pair = pair;
  // two issuers: each owns a sub-tile (MT = 2) or a column half (MT = 1; in pair mode the halves must be whole 32-column
  // blocks per CTA half: BLOCK_N a multiple of 128)
  // (pair mode: ONE issuer — two issuers on column halves double the number of 2-CTA MMAs, each with its ~170-cycle floor:
  //  950 against 1090 TF/s measured; EGR_TC_PAIR_TWO_ISSUERS=1 keeps the variant reachable)
  ka.n_iss = (env_int("EGR_TC_ONE_ISSUER", 0) == 0 &&
              (pair ? (bn % 128 == 0 && env_int("EGR_TC_PAIR_TWO_ISSUERS", 0)) : (mt == 2 || bn % 32 == 0))) ? 2 : 1;
  ka.n_outer = n_outer; ka.n_inner = n_inner;
  ka.tiles1 = tiles1;
  ka.tiles_w = halo ? ceil_div(g.Wo, TILE_M * mt) : tiles_w128;
  ka.tiles_h = tiles_h;
  ka.tiles_m = halo ? ka.tiles_w * tiles_h * tiles_b : (pair ? ceil_div(tiles1, 2) : ceil_div(tiles1, mt));
  ka.tiles_n = g.N / bn;
  ka.splits = splits; ka.outer_per_split = ops;
  ka.fold = fold;
  ka.n_work = ka.tiles_m * ka.tiles_n * (fold ? 1 : splits);
  ka.acc_cols = mt * bn * (fold ? splits : 1);
  ka.nbuf = 2 * ka.acc_cols <= 512 ? 2 : 1;
  ka.d_splits = make_fastdiv(splits); ka.d_tiles_n = make_fastdiv(ka.tiles_n); ka.d_tiles_w = make_fastdiv(ka.tiles_w);
  ka.d_tiles_h = make_fastdiv(tiles_h); ka.d_kchunks = make_fastdiv(kchunks); ka.d_bw = make_fastdiv(g.bw);
  ka.d_bh = make_fastdiv(g.bh); ka.d_c4n = make_fastdiv(bn / 4);
  if (halo) {
    ka.boxA_bytes = HALO_BOX_ROWS * KBLK * 2;
    ka.nboxA = ceil_div(TILE_M * mt + (tmax - tmin), HALO_BOX_ROWS);
  } else {
    ka.boxA_bytes = A_BOX_BYTES;
    ka.nboxA = mt;
  }
  ka.a_stage_bytes = ka.nboxA * ka.boxA_bytes;
  ka.b_stage_bytes = (pair ? bn / 2 : bn) * KBLK * 2;   // pair mode: each CTA stages half of the weight rows
  // A map: always rank 5 (missing dims are size 1 with a harmless stride)
  long long dim[5], str[5]; int box[5];
  for (int d = 0; d < 5; ++d) { dim[d] = a.dim[d]; str[d] = a.stride[d]; }
  for (int d = a.rank; d < 5; ++d) { dim[d] = 1; str[d] = dim[d - 1] * str[d - 1]; }
  for (int d = 0; d < 5; ++d) box[d] = 1;
  box[0] = KBLK; box[g.dimW] = halo ? HALO_BOX_ROWS : g.bw; box[g.dimH] = g.bh; box[g.dimB] = g.bb;
  rc = encode_map(&p->tmA, const_cast<void*>(a.p), 5, dim, str, box, op.name);
  if (rc) return bail(rc);
  // B map: [K, N, Z]
  long long bdim[3] = {g.K, g.N, g.wz_batch ? (long long)g.Bo : (long long)g.ntaps};
  long long bstr[3] = {1, g.wstride_n, g.wstride_z > 0 ? g.wstride_z : g.wstride_n * g.N};
  int bbox[3] = {KBLK, pair ? bn / 2 : bn, 1};
  rc = encode_map(&p->tmB, const_cast<void*>(g.W), 3, bdim, bstr, bbox, op.name);
  if (rc) return bail(rc);

  // ---- ring depths: ~195 KB of the 227 KB for the two rings (the rest: epilogue staging, barriers, alignment slack)
  const int budget = 184 * 1024;
  int SB, SA;
  if (halo) {
    SB = (96 * 1024) / ka.b_stage_bytes; if (SB > 8) SB = 8; if (SB < 2) SB = 2;
    SA = (budget - SB * ka.b_stage_bytes) / ka.a_stage_bytes; if (SA > 3) SA = 3;
    if (SA < 2) { SA = 2; SB = (budget - 2 * ka.a_stage_bytes) / ka.b_stage_bytes; if (SB > 8) SB = 8; }
    if (SB < 2) return bail(fail(EGR_ERR_UNSUPPORTED, "%s: halo tile does not fit shared memory", op.name));
  } else {
    SA = budget / (ka.a_stage_bytes + ka.b_stage_bytes); if (SA > 10) SA = 10; if (SA < 2) SA = 2;
    SB = SA;
  }
  if (env_int("EGR_TC_STAGES", 0) >= 2 && !halo) { SA = SA < env_int("EGR_TC_STAGES", 0) ? SA : env_int("EGR_TC_STAGES", 0); SB = SA; }
  ka.SA = SA; ka.SB = SB;
  ka.dbg = env_int("EGR_TC_DBG_SKIP", 0);
  ka.gmax = halo ? env_int("EGR_TC_GMAX_HALO", 1) : env_int("EGR_TC_GMAX", TC_GMAX);
  ka.hgroup = (halo && !pair) ? env_int("EGR_TC_HGROUP", 1) : 1;
  if (ka.hgroup > TC_GMAX) ka.hgroup = TC_GMAX;
  if (ka.hgroup > SB - 1) ka.hgroup = SB - 1;
  if (ka.hgroup < 1) ka.hgroup = 1;
  if (ka.gmax > TC_GMAX) ka.gmax = TC_GMAX;
  if (ka.gmax > SB) ka.gmax = SB;
  if (ka.gmax < 1) ka.gmax = 1;
  p->smem_bytes = SA * ka.a_stage_bytes + SB * ka.b_stage_bytes + 8 * STAGE_BYTES_PER_WARP + (2 * SA + 2 * SB + 4) * 8 + 16 + 1024;
  if (p->smem_bytes > 227 * 1024) return bail(fail(EGR_ERR_UNSUPPORTED, "%s: %d B of shared memory needed", op.name, p->smem_bytes));
  p->grid = pair ? 2 * (ka.n_work < sms / 2 ? ka.n_work : sms / 2) : (ka.n_work < sms ? ka.n_work : sms);
  p->partial_bytes = (splits > 1 && !fold) ? (size_t)ka.tiles_m * ka.tiles_n * splits * mt * TILE_M * bn * sizeof(float) : 0;
  p->n_counters = (splits > 1 && !fold) ? ka.tiles_m * ka.tiles_n : 0;
  // vector epilogue needs 16-byte aligned groups of 4 columns
  bool vec = !g.transposed && (g.N % 4 == 0) && (g.out_pix_stride % 4 == 0) && (g.out_h_stride % 4 == 0) && (g.out_batch_stride % 4 == 0) &&
             (g.out_offset % 4 == 0) && (g.out_lo % 4 == 0) && (g.out_hi % 4 == 0) && (g.rowbias_stride % 4 == 0);
  auto al = [](const void* q, int a) { return q == nullptr || reinterpret_cast<uintptr_t>(q) % a == 0; };
  vec = vec && al(g.out32, 16) && al(g.out16, 8) && al(g.resid, 16) && al(g.resid2, 16) && al(g.bias, 16) && al(g.rowbias, 16);
  {
    // row offsets inside a tile are exchanged as 32-bit values
    const long long span = (long long)(g.Bo + g.bb) * (g.out_batch_stride > 0 ? g.out_batch_stride : 1) +
                           (long long)(g.Ho + TILE_M) * g.out_h_stride + (long long)(g.Wo + TILE_M) * (g.out_pix_stride > 0 ? g.out_pix_stride : 1) +
                           (long long)g.N * (g.out_n_stride > 0 ? g.out_n_stride : 1);
    if (span >= (1ll << 31)) return bail(fail(EGR_ERR_UNSUPPORTED, "%s: output of %lld elements exceeds the 32-bit tile offsets", op.name, span));
    ka.epi_plain = 0;
    ka.need_crop = (g.out_lo > 0 || g.out_hi < (long long)(g.Ho - 1) * g.out_h_stride + (long long)g.Wo * g.out_pix_stride + g.out_offset + g.N) ? 1 : 0;
  }
  ka.vec_ok = vec ? 1 : 0;
  ka.epi_plain = (vec && g.out32 && g.post == 1.0f && g.act == EGR_ACT_NONE && !g.rowbias) ? 1 : 0;   // cropped edge rows: decided per warp
  *out = p;
  return EGR_OK;
}

size_t egr::tc_partial_bytes(const TcPrepared* p) { return p->partial_bytes; }
int egr::tc_num_counters(const TcPrepared* p) { return p->n_counters; }
void egr::tc_bind_scratch(TcPrepared* p, float* partial, unsigned int* counters) {
  p->ka.partial = partial;
  p->ka.counters = counters;
}

static unsigned long long* g_trace_dev = nullptr;
// debug hook (not part of the public header): first call enables the per-launch clock trace of CTA 0, later calls
// read it back and clear it
extern "C" int egr_debug_tc_trace(unsigned long long* h_out, int n) {
  if (!g_trace_dev) {
    if (cudaMalloc(&g_trace_dev, 1400 * sizeof(unsigned long long)) != cudaSuccess) return -1;
    cudaMemset(g_trace_dev, 0, 1400 * sizeof(unsigned long long));
    return 0;
  }
  cudaDeviceSynchronize();
  if (h_out && n > 0) cudaMemcpy(h_out, g_trace_dev, sizeof(unsigned long long) * (n > 1400 ? 1400 : n), cudaMemcpyDeviceToHost);
  cudaMemset(g_trace_dev, 0, 1400 * sizeof(unsigned long long));
  return 0;
}

int egr::tc_launch(const TcPrepared* p, cudaStream_t st) {
  if (p->ka.splits > 1 && !p->ka.fold && (!p->ka.partial || !p->ka.counters))
    return fail(EGR_ERR_STATE, "%s: split-K scratch not bound", p->name);
  TcKernelArgs ka = p->ka;
  ka.trace = g_trace_dev;
  if (ka.pair) gemm_tc_pair_kernel<<<p->grid, 384, p->smem_bytes, st>>>(p->tmA, p->tmB, ka);
  else if (egr::pdl_enabled()) EGR_CUDA(egr::launch_pdl(gemm_tc_kernel, dim3(p->grid), dim3(384), (size_t)p->smem_bytes, st, p->tmA, p->tmB, ka));
  else gemm_tc_kernel<<<p->grid, 384, p->smem_bytes, st>>>(p->tmA, p->tmB, ka);
  EGR_CHECK_LAUNCH(p->name);
  return EGR_OK;
}

void egr::tc_describe(const TcPrepared* p, int* o) {
  o[0] = p->ka.g.block_n; o[1] = p->ka.mt; o[2] = p->ka.splits; o[3] = p->ka.halo; o[4] = p->ka.n_work; o[5] = p->grid;
  o[6] = p->ka.SA; o[7] = p->ka.SB + 100 * p->ka.fold + 1000 * p->ka.pair;
}

void egr::tc_free(TcPrepared* p) { delete p; }
