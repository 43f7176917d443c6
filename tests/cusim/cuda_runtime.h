// cusim — TEST INFRASTRUCTURE ONLY.  A single-threaded CPU emulation of the slice of the CUDA execution model the SIMT
// kernels of libegregora_b200 use (thread blocks, shared memory, __syncthreads, warp shuffles, atomics), so that the
// kernels' real source can be compiled with g++ and executed in the `-m "not gpu"` tests when no GPU is at hand.
// Every CUDA thread of a block is a ucontext fiber; __syncthreads / shuffles yield to a scheduler that releases a barrier
// when all live participants have arrived.  Blocks run one after another.  Nothing in the product loads this: the
// package's _abi.py binds libegregora_b200.so (nvcc, sm_100a) and raises without a GPU.  This header shadows
// <cuda_runtime.h> only for the emulator build (tests/cusim/build.py puts this directory first on the include path).
// Not emulated: tcgen05 / TMA / mbarrier (gemm_tc.cu is replaced by the plain loops of gemm_tc_ref.cpp, so a whole
// plan can run).  Thread-block clusters exist only in the CUSIM_CLUSTERS build variant: there every CTA of a cluster is an
// OS thread running its own fiber scheduler, `__shared__` is thread_local (so each CTA has its own copy), cluster.sync()
// is a pthread barrier and map_shared_rank() adds the constant distance between two threads' TLS blocks.  thread_local
// access is slow, so the default build has no clusters (the low-pass kernel and GroupNorm on 2-8 CTAs return an error).  Half precision is _Float16.  Timing means nothing; arithmetic differs from
// the GPU only where nvcc contracts a*b+c into FMA and g++ (-ffp-contract=off) does not, and in the last bit of libm
// versus SFU transcendentals.
#pragma once
#include <ucontext.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <type_traits>
#include <vector>

#define CUSIM 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#ifdef CUSIM_CLUSTERS
#define CUSIM_TLS thread_local
#else
#define CUSIM_TLS
#endif
#define __shared__ static CUSIM_TLS
#define __align__(n) alignas(n)

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3 { unsigned x, y, z; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct alignas(8) int2 { int x, y; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline int2 make_int2(int x, int y) { return int2{x, y}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }

// ------------------------------------------------------------------------------------------------ runtime API subset
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
typedef void* cudaStream_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
struct cudaDeviceProp {
  char name[256];
  int major, minor, multiProcessorCount;
  size_t sharedMemPerBlockOptin, totalGlobalMem;
};
inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "cusim error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 148; return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  memset(p, 0, sizeof(*p));
  snprintf(p->name, sizeof(p->name), "cusim (CPU emulation of an sm_100 device)");
  p->major = 10; p->minor = 0; p->multiProcessorCount = 148; p->sharedMemPerBlockOptin = 227 * 1024;
  return cudaSuccess;
}
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n); }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <class F> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 8; return cudaSuccess; }
// CUDA graphs: capture is refused, so egr_plan_run stays on its eager path
typedef void* cudaGraph_t;
typedef void* cudaGraphExec_t;
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal, cudaStreamCaptureModeThreadLocal, cudaStreamCaptureModeRelaxed };
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return cudaErrorInvalidValue; }
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = nullptr; return cudaErrorInvalidValue; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t*, cudaGraph_t, unsigned long long) { return cudaErrorInvalidValue; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorInvalidValue; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }

// ------------------------------------------------------------------------------------------------ the emulator
namespace cusim {

enum State { READY, WAIT_BLOCK, WAIT_WARP, WAIT_CLUSTER, DONE };

struct Fiber {
  ucontext_t uc;
  uint3 tid;
  State st;
  unsigned wmask;  // lanes this fiber waits for at a warp barrier
};

struct Block {
  dim3 bdim, gdim;
  uint3 bid;
  std::vector<Fiber> f;
  ucontext_t sched;
  int cur = -1;
  const std::function<void()>* body = nullptr;
  alignas(16) unsigned char wbuf[64][32][16];  // shuffle exchange slots: [warp][lane][bytes]
};

Block& blk();                  // the block being executed
void* dyn_smem();              // dynamic shared memory of the running block
void yield(State s, unsigned mask);
// cluster > 1 (CUSIM_CLUSTERS build only): consecutive x-blocks form a cluster; returns false when it cannot be done
bool launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body, unsigned cluster = 1);
unsigned cluster_size();
unsigned cluster_rank();
void cluster_sync();
ptrdiff_t cluster_delta(unsigned rank);   // bytes from this CTA's shared memory to the same object in CTA `rank`
unsigned long long launches();

inline int linear_tid() { return blk().cur; }
inline bool lane_alive(int warp, int lane) {
  Block& b = blk();
  const size_t i = (size_t)warp * 32 + lane;
  return i < b.f.size() && b.f[i].st != DONE;
}

template <class T, class PICK>
inline T shuffle(unsigned mask, T v, PICK pick) {
  static_assert(sizeof(T) <= 16, "shuffle payload too large");
  Block& b = blk();
  const int lt = linear_tid(), warp = lt >> 5, lane = lt & 31;
  memcpy(b.wbuf[warp][lane], &v, sizeof(T));
  yield(WAIT_WARP, mask);
  const int src = pick(lane);
  T r = v;
  if (src >= 0 && src < 32 && ((mask >> src) & 1u) && lane_alive(warp, src)) memcpy(&r, b.wbuf[warp][src], sizeof(T));
  yield(WAIT_WARP, mask);
  return r;
}

}  // namespace cusim

// extended launch (cluster dimension along x only)
enum cudaLaunchAttributeID { cudaLaunchAttributeClusterDimension = 4 };
struct cudaLaunchAttributeValue { struct { unsigned x, y, z; } clusterDim; };
struct cudaLaunchAttribute { cudaLaunchAttributeID id; cudaLaunchAttributeValue val; };
struct cudaLaunchConfig_t {
  dim3 gridDim, blockDim;
  size_t dynamicSmemBytes = 0;
  cudaStream_t stream = nullptr;
  cudaLaunchAttribute* attrs = nullptr;
  unsigned numAttrs = 0;
};
template <class K, class... A>
inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, K kernel, A... args) {
  unsigned ncl = 1;
  for (unsigned i = 0; i < cfg->numAttrs; ++i)
    if (cfg->attrs[i].id == cudaLaunchAttributeClusterDimension && (cfg->attrs[i].val.clusterDim.y != 1 || cfg->attrs[i].val.clusterDim.z != 1))
      return cudaErrorInvalidValue;
    else if (cfg->attrs[i].id == cudaLaunchAttributeClusterDimension)
      ncl = cfg->attrs[i].val.clusterDim.x;
  return cusim::launch(cfg->gridDim, cfg->blockDim, cfg->dynamicSmemBytes, [&]() { kernel(args...); }, ncl) ? cudaSuccess : cudaErrorInvalidValue;
}
#define __cluster_dims__(...)

// built-in variables: plain globals the scheduler sets before it resumes a fiber (one OS thread runs everything)
extern CUSIM_TLS uint3 threadIdx, blockIdx;
extern CUSIM_TLS dim3 blockDim, gridDim;

inline void __syncthreads() { cusim::yield(cusim::WAIT_BLOCK, 0); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { cusim::yield(cusim::WAIT_WARP, mask); }
template <class T> inline T __shfl_xor_sync(unsigned m, T v, int x, int = 32) { return cusim::shuffle(m, v, [x](int l) { return l ^ x; }); }
template <class T> inline T __shfl_up_sync(unsigned m, T v, unsigned d, int = 32) { return cusim::shuffle(m, v, [d](int l) { return l - (int)d; }); }
template <class T> inline T __shfl_down_sync(unsigned m, T v, unsigned d, int = 32) { return cusim::shuffle(m, v, [d](int l) { return l + (int)d; }); }
template <class T> inline T __shfl_sync(unsigned m, T v, int s, int = 32) { return cusim::shuffle(m, v, [s](int) { return s; }); }

// one OS thread runs everything, so plain read-modify-write is atomic
template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }

// ------------------------------------------------------------------------------------------------ device intrinsics
template <class T> inline T __ldg(const T* p) { return *p; }
inline float __fmul_rn(float a, float b) { return a * b; }      // build with -ffp-contract=off: separately rounded
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
inline float __expf(float a) { return expf(a); }
inline float __sinf(float a) { return sinf(a); }
inline float __fdividef(float a, float b) { return a / b; }
inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline float __cosf(float a) { return cosf(a); }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline int __float2int_rn(float f) { return (int)nearbyintf(f); }
inline void sincospi(double x, double* s, double* c) {
  double r = fmod(x, 2.0);              // exact; sin/cos of pi*r with r in (-2, 2)
  if (r > 1.0) r -= 2.0; else if (r < -1.0) r += 2.0;
  if (r == 0.5) { *s = 1.0; *c = 0.0; } else if (r == -0.5) { *s = -1.0; *c = 0.0; }
  else if (r == 1.0 || r == -1.0) { *s = 0.0; *c = -1.0; }
  else { *s = sin(M_PI * r); *c = cos(M_PI * r); }
}
inline double cospi(double x) { double s, c; sincospi(x, &s, &c); return c; }
inline double sinpi(double x) { double s, c; sincospi(x, &s, &c); return s; }
inline void sincospif(float x, float* s, float* c) { double sd, cd; sincospi((double)x, &sd, &cd); *s = (float)sd; *c = (float)cd; }

template <class A, class B> inline typename std::common_type<A, B>::type min(A a, B b) { return b < a ? b : a; }
template <class A, class B> inline typename std::common_type<A, B>::type max(A a, B b) { return a < b ? b : a; }
