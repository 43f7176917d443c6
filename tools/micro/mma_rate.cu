// Microbenchmark: how fast does ONE SM retire tcgen05.mma (kind::f16, f32 accumulate) with operands that sit in shared
// memory the whole time (no TMA, no epilogue)?  Tells the tensor pipe's own pace apart from everything gemm_tc.cuh adds.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
// variants: N (instruction width), issuers (1 or 2 warps on disjoint TMEM columns), stages (operands rotate over S slots
// of shared memory), commit every k-step or only at the end, pair (cta_group::2, M = 256 across a cluster of two CTAs).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}"
      ::"r"(bar), "r"(parity) : "memory");
}
template <int CG>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint32_t bar) {
  if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                    ::"r"(bar), "h"((unsigned short)3) : "memory");
}

struct P { int N, issuers, stages, commit_each, ksteps, kper, pollers, ring, nofence, nowait, wmode, G; };

template <int CG>
__global__ void __launch_bounds__(384, 1) k(P p, unsigned long long* out) {
  long long* stamps = reinterpret_cast<long long*>(out + 148 * 4);
  extern __shared__ uint8_t raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[12];
  __shared__ uint64_t full[8], empty[8], never, never2;
  __shared__ unsigned flag[8];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  // operand slots: A 16 KB, B (N/CG)*128 B each, per stage
  const int a_bytes = 16384, b_bytes = (p.N / CG) * 128;
  for (int i = threadIdx.x; i < p.stages * (a_bytes + b_bytes) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;  // f16 1.0
  if (threadIdx.x < 8) flag[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 12; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(1));
    for (int i = 0; i < 8; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[i])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[i])), "r"(p.issuers));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&never)), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&never2)), "r"(1));
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&never2)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  const int NI = p.N / p.issuers;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(NI >> 3) << 17) | ((uint32_t)((CG * 128) >> 4) << 24);
  unsigned long long t0 = 0, t1 = 0;
  if (warp < p.issuers && rank == 0) {
    const int u = warp;
    const uint32_t sm_u = smem_u32(sm);
    const uint32_t bar_done = smem_u32(&bars[u]);
    const uint32_t bar_step = smem_u32(&bars[4 + u]);
    t0 = clock64();
    int s = 0;
    unsigned long long rounds = 0;
    long long ph[7] = {0, 0, 0, 0, 0, 0, 0}, cend = 0, c0 = clock64(); int elected = 0;
    if (p.ring == 2) {
      int ks = 0;
      while (ks < p.ksteps) {
        mbar_wait(smem_u32(&full[s]), (uint32_t)((ks / p.stages) & 1));
        int nb = 1;
        while (nb < p.G && ks + nb < p.ksteps) {
          const int s2 = (s + nb) % p.stages;
          uint32_t ok;
          asm volatile("{\n\t.reg .pred q;\n\tmbarrier.test_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.b32 %0, 1, 0, q;\n\t}"
                       : "=r"(ok) : "r"(smem_u32(&full[s2])), "r"((uint32_t)(((ks + nb) / p.stages) & 1)) : "memory");
          if (!__all_sync(0xffffffffu, ok)) break;
          ++nb;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int j = 0; j < nb; ++j) {
          const uint32_t aB = sm_u + s * (a_bytes + b_bytes);
          const uint32_t bB = aB + a_bytes + (uint32_t)(u * (NI / CG) * 128);
          const uint64_t ad = make_desc(aB), bd = make_desc(bB);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              mma<CG>(tm + (uint32_t)(u * NI), ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), idesc, (ks | kk) ? 1u : 0u);
            commit<CG>(smem_u32(&empty[s]));
          }
          __syncwarp();
          ++ks;
          if (++s == p.stages) s = 0;
        }
        ++rounds;
      }
    } else
    for (int ks = 0; ks < p.ksteps; ++ks) {
      const uint32_t aB = sm_u + s * (a_bytes + b_bytes);
      const uint32_t bB = aB + a_bytes + (uint32_t)(u * (NI / CG) * 128);
      const uint64_t ad = make_desc(aB), bd = make_desc(bB);
      const long long cw0 = clock64();
      if (p.ring && !p.nowait && (p.wmode != 5 || ks == 0)) {
        const uint32_t par = (uint32_t)((ks / p.stages) & 1);
        if (p.wmode == 0 || p.wmode == 5) mbar_wait(smem_u32(&full[s]), par);
        else if (p.wmode == 6) {
          asm volatile("{\n\t.reg .pred q;\n\tWL6:\n\tmbarrier.try_wait.parity.relaxed.cta.shared::cta.b64 q, [%0], %1;\n\t@q bra DN6;\n\tbra WL6;\n\tDN6:\n\t}" ::"r"(smem_u32(&full[s])), "r"(par) : "memory");
        } else if (p.wmode == 7) {   // plain volatile load poll of the barrier word's phase bit? not portable: poll a flag the producer writes
          volatile unsigned* f = (volatile unsigned*)&flag[s];
          while (*f != (unsigned)(ks / p.stages + 1)) { }
        }
        else if (p.wmode == 1) {
          uint32_t ok = 0;
          while (!ok) asm volatile("{\n\t.reg .pred q;\n\tmbarrier.test_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.b32 %0, 1, 0, q;\n\t}" : "=r"(ok) : "r"(smem_u32(&full[s])), "r"(par) : "memory");
        } else if (p.wmode == 2) {
          mbar_wait(smem_u32(&full[s]), par);
          mbar_wait(smem_u32(&never2), 0);   // a second, already completed barrier: the cost of one more wait per round
        } else if (p.wmode == 3) {
          if ((threadIdx.x & 31) == 0) mbar_wait(smem_u32(&full[s]), par);
          __syncwarp();
        }
      }
      if (p.ring && !p.nofence) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const long long c1 = clock64();
      if (u == 0 && blockIdx.x == 0 && ks >= 200 && ks < 216 && (threadIdx.x & 31) == 0) { stamps[(ks - 200) * 5 + 0] = cw0; stamps[(ks - 200) * 5 + 1] = c1; }
      if (elect_one()) {
        const long long c2 = clock64();
        mma<CG>(tm + (uint32_t)(u * NI), ad, bd, idesc, ks ? 1u : 0u);
        const long long c3 = clock64();
#pragma unroll 4
        for (int kk = 1; kk < p.kper; ++kk)
          mma<CG>(tm + (uint32_t)(u * NI), ad + (uint64_t)(2 * (kk & 3)), bd + (uint64_t)(2 * (kk & 3)), idesc, 1u);
        const long long c4 = clock64();
        if (p.wmode == 5 && ks + 1 < p.ksteps) { const int s1 = (s + 1 == p.stages) ? 0 : s + 1; mbar_wait(smem_u32(&full[s1]), (uint32_t)(((ks + 1) / p.stages) & 1)); }
        if (p.ring) commit<CG>(smem_u32(&empty[s])); else if (p.commit_each) commit<CG>(bar_step);
        const long long c5 = clock64();
        if (u == 0 && blockIdx.x == 0 && ks >= 200 && ks < 216) stamps[(ks - 200) * 5 + 2] = c5;
        ph[1] += c2 - c1; ph[2] += c3 - c2; ph[3] += c4 - c3; ph[4] += c5 - c4; cend = c5; elected = 1;
      }
      __syncwarp();
      { const long long c6 = clock64(); if (elected) ph[5] += c6 - cend; ph[0] += c1 - c0; ph[6] += c6 - c0; c0 = c6; }
      if (++s == p.stages) s = 0;
    }
    if (elect_one()) commit<CG>(bar_done);
    __syncwarp();
    mbar_wait(bar_done, 0);
    t1 = clock64();
    if (threadIdx.x % 32 == 0) { out[(blockIdx.x * 2 + u) * 2] = t0; out[(blockIdx.x * 2 + u) * 2 + 1] = t1; if (blockIdx.x == 0 && u == 0) out[148 * 4 - 1] = rounds; }
    if (elected && blockIdx.x == 0 && u == 0) for (int i = 0; i < 7; ++i) out[148 * 4 - 9 + i] = (unsigned long long)ph[i];
    if (u == 0 && threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&never)) : "memory");
  } else if (warp == 2 && p.ring && rank == 0) {
    // producer stand-in: hands every stage back as soon as the MMAs that read it have completed (no loads)
    int s = 0;
    for (int ks = 0; ks < p.ksteps; ++ks) {
      mbar_wait(smem_u32(&empty[s]), (uint32_t)(((ks / p.stages) & 1) ^ 1));
      const long long pw = clock64();
      if ((threadIdx.x & 31) == 0) { ((volatile unsigned*)flag)[s] = (unsigned)(ks / p.stages + 1); }
      if (elect_one()) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[s])) : "memory");
      if (blockIdx.x == 0 && ks >= 200 && ks < 216 && (threadIdx.x & 31) == 0) { stamps[(ks - 200) * 5 + 3] = pw; stamps[(ks - 200) * 5 + 4] = clock64(); }
      __syncwarp();
      if (++s == p.stages) s = 0;
    }
  } else if (warp >= 3 && warp < 3 + p.pollers && rank == 0) {
    mbar_wait(smem_u32(&never), 0);   // what the epilogue warps do during a main loop
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  if (warp == 0) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
  }
}

template <int CG>
static void run(P p, int grid, const char* tag) {
  unsigned long long* d;
  cudaMalloc(&d, 148 * 4 * 8 + 80 * 8);
  cudaMemset(d, 0, 148 * 4 * 8 + 80 * 8);
  const int smem = p.stages * (16384 + (p.N / CG) * 128) + 2048;
  cudaFuncSetAttribute(k<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    if (CG == 1) {
      k<CG><<<grid, 384, smem>>>(p, d);
    } else {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, k<CG>, p, d);
    }
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("%s: CUDA error %s\n", tag, cudaGetErrorString(err)); exit(1); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  unsigned long long h[148 * 4];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const double cyc = (double)(h[1] - h[0]);
  const double n_mma = (double)p.ksteps * p.kper;   // per issuer
  const double flops = 2.0 * (CG * 128) * p.N * 16 * n_mma * (grid / CG);
  printf("%-34s N=%3d iss=%d stages=%d ring=%d pollers=%d commit_each=%d : %7.1f clk per k-step(%d MMAs/issuer)  %6.1f clk/MMA  floor %5.1f   chip %7.1f TF/s (%.3f ms)\n",
         tag, p.N, p.issuers, p.stages, p.ring, p.pollers, p.commit_each, cyc / p.ksteps, p.kper, cyc / n_mma,
         128.0 * (p.N / p.issuers) / 256.0, flops / best / 1e9, best);
  if (p.ring != 2) printf("      per round: top-of-loop+wait %.0f | ->elected %.0f | 1st MMA %.0f | rest MMAs %.0f | commit %.0f | ->reconverged %.0f | total %.0f\n",
         (double)h[148*4-9] / p.ksteps, (double)h[148*4-8] / p.ksteps, (double)h[148*4-7] / p.ksteps, (double)h[148*4-6] / p.ksteps, (double)h[148*4-5] / p.ksteps, (double)h[148*4-4] / p.ksteps, (double)h[148*4-3] / p.ksteps);
  if (p.ring == 1) { long long st[80]; cudaMemcpy(st, d + 148 * 4, sizeof(st), cudaMemcpyDeviceToHost); long long b = st[0];
    for (int r = 0; r < 16; ++r) printf("      r%02d issuer: at wait %6lld, passed %6lld, committed %6lld | producer for this round: woke %6lld, armed %6lld\n", r, st[r*5]-b, st[r*5+1]-b, st[r*5+2]-b, st[r*5+3]-b, st[r*5+4]-b); }
  if (p.ring == 2) printf("      adaptive: %.2f k-steps per round\n", (double)p.ksteps / (double)h[148 * 4 - 1]);
  cudaFree(d);
}

int main() {
  const int KS = 4000;
  const int grid = 148;
  run<1>({96, 1, 3, 1, KS, 4, 0, 1, 0, 0, 0, 1}, grid, "ring N=96 acquire try_wait");
  run<1>({96, 1, 3, 1, KS, 4, 0, 1, 0, 0, 6, 1}, grid, "ring N=96 relaxed try_wait");
  run<1>({96, 1, 3, 1, KS, 4, 0, 1, 0, 0, 7, 1}, grid, "ring N=96 flag poll (ld.volatile)");
  run<1>({256, 1, 3, 1, KS, 4, 0, 1, 0, 0, 6, 1}, grid, "ring N=256 relaxed try_wait");
  run<1>({256, 1, 3, 1, KS, 4, 0, 1, 0, 0, 7, 1}, grid, "ring N=256 flag poll");
  return 0;
}
