// core.cu — lifecycle, error reporting, small utility kernels (absmax, PCM-16 wire format).
#include <cstdlib>
#include "common.cuh"

namespace egr {

char* err_buf() {
  static thread_local char buf[1024] = "";
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 1024, fmt, ap);
  va_end(ap);
  return code;
}

bool pdl_enabled() {
#ifdef __CUDACC__
  static const int on = [] { const char* e = getenv("EGR_PDL"); return e ? atoi(e) : 0; }();
  return on != 0;
#else
  return false;
#endif
}

unsigned long long& launch_count() {
  static unsigned long long n = 0;
  return n;
}

DeviceInfo& devinfo() {
  static DeviceInfo d;
  return d;
}

}  // namespace egr

using namespace egr;

extern "C" int egr_abi_version(void) { return EGR_ABI_VERSION; }
extern "C" const char* egr_last_error(void) { return err_buf(); }
extern "C" int egr_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(egr_tensor);
    case 1: return (int)sizeof(egr_op);
    default: return -1;
  }
}

extern "C" int egr_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(EGR_ERR_CUDA, "no CUDA device visible (%s); libegregora_b200 has no CPU fallback",
                e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(EGR_ERR_ARG, "device %d out of range [0,%d)", device, n);
  EGR_CUDA(cudaSetDevice(device));
  cudaDeviceProp p;
  EGR_CUDA(cudaGetDeviceProperties(&p, device));
  if (p.major != 10)
    return fail(EGR_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                p.major, p.minor);
  DeviceInfo& d = devinfo();
  d.device = device;
  d.sm_count = p.multiProcessorCount;
  d.cc_major = p.major;
  d.cc_minor = p.minor;
  d.inited = true;
  return EGR_OK;
}

extern "C" int egr_sm_count(void) { return devinfo().sm_count; }
extern "C" int64_t egr_launch_count(void) { return (int64_t)launch_count(); }

// ------------------------------------------------------------------------------------------------
// absmax: grid-stride float4 loads, warp-shuffle + one atomicMax (as uint bits; values are >= 0).
// ------------------------------------------------------------------------------------------------
__global__ void absmax_kernel(const float* __restrict__ x, int64_t n, unsigned int* __restrict__ out) {
  float m = 0.f;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (int64_t j = i; j < n4; j += stride) {
    float4 v = __ldg(x4 + j);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  for (int64_t j = (n4 << 2) + i; j < n; j += stride) m = fmaxf(m, fabsf(x[j]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

extern "C" int egr_absmax(const float* d_in, int64_t n, float* d_out, void* stream) {
  if (!d_in || !d_out || n < 0) return fail(EGR_ERR_ARG, "egr_absmax: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  EGR_CUDA(cudaMemsetAsync(d_out, 0, sizeof(float), st));
  if (n == 0) return EGR_OK;
  if ((reinterpret_cast<uintptr_t>(d_in) & 15) != 0) return fail(EGR_ERR_ARG, "egr_absmax: input must be 16-byte aligned");
  int sms = devinfo().sm_count ? devinfo().sm_count : 148;
  int blocks = (int)((n / 4 + 255) / 256);
  if (blocks > sms * 8) blocks = sms * 8;
  if (blocks < 1) blocks = 1;
  absmax_kernel<<<blocks, 256, 0, st>>>(d_in, n, reinterpret_cast<unsigned int*>(d_out));
  EGR_CHECK_LAUNCH("absmax_kernel");
  return EGR_OK;
}

// ------------------------------------------------------------------------------------------------
// PCM-16 wire format.  libsndfile float->short with clipping on (soundfile's default for float input to a
// PCM_16 file): scaled = x * 32768.0 (normalised float mode), clip to [-32768, 32767], lrintf (round
// half to even).  Read back: sample / 32768.0.
// ------------------------------------------------------------------------------------------------
__global__ void pcm16_quant_kernel(const float* __restrict__ x, int16_t* __restrict__ y, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float s = x[i] * 32768.0f;
    int v;
    if (s >= 32767.0f) v = 32767;
    else if (s <= -32768.0f) v = -32768;
    else v = __float2int_rn(s);
    y[i] = (int16_t)v;
  }
}

__global__ void pcm16_deq_kernel(const int16_t* __restrict__ x, float* __restrict__ y, int64_t n, float scale) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) y[i] = (float)x[i] * scale;
}

static int grid_for(int64_t n, int per_thread = 4) {
  int sms = devinfo().sm_count ? devinfo().sm_count : 148;
  int64_t b = (n + 256LL * per_thread - 1) / (256LL * per_thread);
  if (b > (int64_t)sms * 16) b = (int64_t)sms * 16;
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" int egr_pcm16_quantize(const float* d_in, int16_t* d_out, int64_t n, void* stream) {
  if (!d_in || !d_out || n < 0) return fail(EGR_ERR_ARG, "egr_pcm16_quantize: bad arguments");
  if (n == 0) return EGR_OK;
  pcm16_quant_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(d_in, d_out, n);
  EGR_CHECK_LAUNCH("pcm16_quant_kernel");
  return EGR_OK;
}

extern "C" int egr_pcm16_to_float(const int16_t* d_in, float* d_out, int64_t n, float scale, void* stream) {
  if (!d_in || !d_out || n < 0) return fail(EGR_ERR_ARG, "egr_pcm16_to_float: bad arguments");
  if (n == 0) return EGR_OK;
  pcm16_deq_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(d_in, d_out, n, scale);
  EGR_CHECK_LAUNCH("pcm16_deq_kernel");
  return EGR_OK;
}

// ------------------------------------------------------------------------------------------------
// Diffusion start noise x_T.  Counter-based (Philox4x32-10, Salmon et al. 2011) so that the value of element e of
// chunk-channel row r depends on (seed, r, e) only: a rank that owns rows [lo, hi) of a clip generates exactly the
// numbers a single-GPU run uses for those rows, whatever the world size or sub-batch split.  Four uniforms per
// counter -> two Box-Muller pairs.  u = ((bits >> 8) + 1) * 2^-24 in (0, 1].
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void noise_fill_kernel(uint32_t k0, uint32_t k1, int64_t row0, int64_t n_rows, int64_t row_elems,
                                  float* __restrict__ out) {
  const int64_t quads = (row_elems + 3) >> 2;
  const int64_t total = n_rows * quads;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / quads, q = i - r * quads;
    const uint64_t row = (uint64_t)(row0 + r);
    uint32_t x[4];
    philox4x32_10((uint32_t)q, (uint32_t)((uint64_t)q >> 32), (uint32_t)row, (uint32_t)(row >> 32), k0, k1, x);
    float z[4];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const float u1 = (float)((x[2 * p] >> 8) + 1u) * 5.9604644775390625e-8f;      // 2^-24
      const float u2 = (float)((x[2 * p + 1] >> 8) + 1u) * 5.9604644775390625e-8f;
      const float rad = sqrtf(-2.0f * logf(u1));
      float s, c;
      sincospif(2.0f * u2, &s, &c);
      z[2 * p] = rad * c; z[2 * p + 1] = rad * s;
    }
    float* dst = out + r * row_elems + 4 * q;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (4 * q + u < row_elems) dst[u] = z[u];
  }
}

extern "C" int egr_noise_fill(uint64_t seed, int64_t row0, int64_t n_rows, int64_t row_elems, float* d_out, void* stream) {
  if (!d_out || n_rows < 0 || row_elems <= 0 || row0 < 0) return fail(EGR_ERR_ARG, "egr_noise_fill: bad arguments");
  if (n_rows == 0) return EGR_OK;
  const int64_t total = n_rows * ((row_elems + 3) >> 2);
  noise_fill_kernel<<<grid_for(total, 1), 256, 0, (cudaStream_t)stream>>>((uint32_t)seed, (uint32_t)(seed >> 32), row0, n_rows,
                                                                         row_elems, d_out);
  EGR_CHECK_LAUNCH("noise_fill_kernel");
  return EGR_OK;
}

// x *= scale iff *d_ref > threshold — the "rescale integer-scaled data by 2^(8*sw-1) when its peak exceeds 1" rule of
// the reference's patched write_audio (egregora_fat_llama_gpu.py:195-200), decided on the device from egr_absmax's
// result so the node body never reads the peak back to the host.
__global__ void scale_if_above_kernel(float* __restrict__ x, int64_t n, const float* __restrict__ ref, float threshold, float scale) {
  if (!(__ldg(ref) > threshold)) return;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) x[i] *= scale;
}

extern "C" int egr_scale_if_above(float* d_x, int64_t n, const float* d_ref, float threshold, float scale, void* stream) {
  if (!d_x || !d_ref || n < 0) return fail(EGR_ERR_ARG, "egr_scale_if_above: bad arguments");
  if (n == 0) return EGR_OK;
  scale_if_above_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(d_x, n, d_ref, threshold, scale);
  EGR_CHECK_LAUNCH("scale_if_above_kernel");
  return EGR_OK;
}
