// common.cuh — shared host/device helpers for libegregora_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include "../../include/egregora_b200.h"

#define EGR_OK 0
#define EGR_ERR_ARG -1
#define EGR_ERR_CUDA -2
#define EGR_ERR_UNSUPPORTED -3
#define EGR_ERR_STATE -4

namespace egr {

// thread-local last-error text, exposed through egr_last_error()
char* err_buf();
int fail(int code, const char* fmt, ...);
// kernels launched by this library in this process (every EGR_CHECK_LAUNCH site counts one)
unsigned long long& launch_count();

#define EGR_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return egr::fail(EGR_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),   \
                       __FILE__, __LINE__);                                                     \
  } while (0)

#define EGR_CHECK_LAUNCH(what)                                                                  \
  do {                                                                                          \
    ++egr::launch_count();                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ != cudaSuccess)                                                                     \
      return egr::fail(EGR_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e__)); \
  } while (0)

struct DeviceInfo {
  int device = -1;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  bool inited = false;
};
DeviceInfo& devinfo();

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Programmatic dependent launch (EGR_PDL=1): a kernel launched through launch_pdl() may be scheduled while its predecessor in
// the stream is still draining — its CTAs take the SM slots the predecessor's last CTAs free, run their set-up (barrier
// init, TMEM allocation, descriptor prefetch) and block in egr_pdl_wait() until the predecessor has completed and its
// writes are visible.  Every kernel launched that way calls egr_pdl_wait() before its first global-memory access (reads AND
// writes: workspace buffers are recycled) and egr_pdl_trigger() right after it, so at most one kernel runs ahead.
bool pdl_enabled();
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// resolved pointers of a plan
struct Spaces {
  char* ws = nullptr;
  size_t ws_bytes = 0;
  const char* wt = nullptr;
  size_t wt_bytes = 0;
};

inline void* resolve(const Spaces& s, uint64_t addr) {
  uint64_t space = addr >> 60, off = addr & 0x0FFFFFFFFFFFFFFFull;
  switch (space) {
    case EGR_SPACE_WS: return s.ws + off;
    case EGR_SPACE_WT: return const_cast<char*>(s.wt) + off;
    case EGR_SPACE_ABS: return reinterpret_cast<void*>(off);
    default: return nullptr;
  }
}

}  // namespace egr

// ---------------------------------------------------------------- device helpers
// no-ops when the kernel was not launched as a programmatic dependent (and under the CPU emulator)
__device__ __forceinline__ void egr_pdl_wait() {
#ifdef __CUDACC__
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void egr_pdl_trigger() {
#ifdef __CUDACC__
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void egr_pdl_sync() { egr_pdl_wait(); egr_pdl_trigger(); }
// x * sigmoid(x) as FMUL + MUFU.EX2 + FADD + MUFU.RCP + FMUL (2 ulp): the IEEE division this replaces was ~20 instructions
// per element (Newton steps + a range check with a slow-path call) and made the GroupNorm apply pass issue bound
__device__ __forceinline__ float egr_silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------- packed f32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2)
// A three-register FFMA issues every other cycle per scheduler on this part; the packed forms do two IEEE operations per
// issue slot (each lane rounded exactly like the scalar instruction), so f32 streaming kernels that are bound by the FMA pipe
// run two independent elements (two channels, re/im) per thread.  The plain-C bodies are what the CPU emulator compiles.
__device__ __forceinline__ float2 egr_fma2(float2 a, float2 b, float2 c) {
#ifdef __CUDACC__
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
#else
  return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
__device__ __forceinline__ float2 egr_mul2(float2 a, float2 b) {
#ifdef __CUDACC__
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
#else
  return make_float2(a.x * b.x, a.y * b.y);
#endif
}
__device__ __forceinline__ float2 egr_add2(float2 a, float2 b) {
#ifdef __CUDACC__
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
