// fft.cu — path B transform engine: plan cache, the generic two-level complex FFT behind egr_fft_exec
// (natural order in and out, any length; Bluestein for lengths that are not 2..13-smooth), no cuFFT.
// The Fat-Llama fast path (fatllama.cu) uses the same plans but its own fused kernels.
#include <cmath>
#include <functional>
#include <map>
#include <mutex>
#include <cstdlib>
#include "fft_plan.cuh"

using namespace egr;
using namespace egrfft;

// ------------------------------------------------------------------------------------------------ planning
static const int kPrimes[] = {2, 3, 5, 7, 11, 13};
static const int kMaxLen = 8192;  // longest shared-memory transform (64 KB of float2)

static bool smooth(int64_t n) {
  if (n < 1) return false;
  for (int p : kPrimes)
    while (n % p == 0) n /= p;
  return n == 1;
}

static int max_radix() {  // tuning override: cap on the in-register radix (default 16)
  if (const char* e = getenv("EGR_FFT_MAXR")) {
    const int r = atoi(e);
    if (r >= 7 && r <= 16) return r;
  }
  return 16;
}

// fewest stages with radices in [2,16]; larger radices first
static bool plan_radices(int L, Radices* out) {
  std::vector<int> best;
  std::vector<int> cur;
  std::map<int, int> memo;  // L -> min stages
  std::function<int(int)> cost = [&](int l) -> int {
    if (l == 1) return 0;
    auto it = memo.find(l);
    if (it != memo.end()) return it->second;
    int b = 1 << 20;
    for (int r = max_radix(); r >= 2; --r)
      if (l % r == 0) {
        int c = cost(l / r);
        if (c + 1 < b) b = c + 1;
      }
    memo[l] = b;
    return b;
  };
  if (cost(L) >= (1 << 20)) return false;
  int l = L;
  out->n = 0;
  while (l > 1) {
    int pick = 0;
    for (int r = max_radix(); r >= 2; --r)
      if (l % r == 0 && cost(l / r) + 1 == cost(l)) { pick = r; break; }
    if (!pick || out->n >= EGR_FFT_MAX_STAGES) return false;
    out->r[out->n++] = pick;
    l /= pick;
  }
  return true;
}

static void split(int64_t M, int* R1, int* R2) {
  *R1 = 0; *R2 = 0;
  if (M <= kMaxLen) { *R1 = 1; *R2 = (int)M; return; }
  if (const char* e = getenv("EGR_FFT_R1")) {  // tuning override: column length (must divide M, both factors <= kMaxLen)
    const int64_t a = atoll(e);
    if (a >= 2 && M % a == 0 && a <= kMaxLen && M / a <= kMaxLen) { *R1 = (int)a; *R2 = (int)(M / a); return; }
  }
  for (int64_t d = (int64_t)std::floor(std::sqrt((double)M)) + 1; d >= 2; --d) {
    if (M % d) continue;
    int64_t a = d, b = M / d;
    if (a > b) std::swap(a, b);
    if (b <= kMaxLen) { *R1 = (int)a; *R2 = (int)b; return; }
  }
}

bool egr::fft2_plannable(int64_t M) {
  if (M < 1 || !smooth(M)) return false;
  int a, b;
  split(M, &a, &b);
  return a > 0;
}

int64_t egr::fft2_next_plannable(int64_t n) {
  // 2,3,5,7-smooth candidates; the split constraint only bites above 8192^2
  for (int64_t m = n < 1 ? 1 : n;; ++m) {
    int64_t t = m;
    for (int p : {2, 3, 5, 7})
      while (t % p == 0) t /= p;
    if (t == 1 && fft2_plannable(m)) return m;
    if (m > n + (1ll << 26)) return -1;
  }
}

static void digit_perm(int L, const Radices& rd, int s, std::vector<int>& out) {
  // out[pos] = frequency index held at position pos after the DIF stages s.. of a length-L transform
  if (s == rd.n) { out.assign(1, 0); return; }
  const int r = rd.r[s], Ls = L / r;
  std::vector<int> sub;
  digit_perm(Ls, rd, s + 1, sub);
  out.resize(L);
  for (int p = 0; p < r; ++p)
    for (int i = 0; i < Ls; ++i) out[p * Ls + i] = r * sub[i] + p;
}

static std::mutex g_mu;
static std::map<std::pair<int, int64_t>, Fft2Plan*> g_plans;

template <class T>
static T* carve(char*& p, size_t n) {
  T* r = reinterpret_cast<T*>(p);
  p += (n * sizeof(T) + 255) / 256 * 256;
  return r;
}

const Fft2Plan* egr::fft2_get_plan(int64_t M) {
  std::lock_guard<std::mutex> lk(g_mu);
  int dev = 0;
  cudaGetDevice(&dev);
  auto key = std::make_pair(dev, M);
  auto it = g_plans.find(key);
  if (it != g_plans.end()) return it->second;
  if (!fft2_plannable(M)) {
    fail(EGR_ERR_UNSUPPORTED, "fft: length %lld has no two-level plan (needs 2..13-smooth, factors <= %d)", (long long)M, kMaxLen);
    return nullptr;
  }
  Fft2Plan* p = new Fft2Plan();
  p->M = M;
  split(M, &p->R1, &p->R2);
  if (!plan_radices(p->R1, &p->rd1) || !plan_radices(p->R2, &p->rd2)) {
    delete p;
    fail(EGR_ERR_UNSUPPORTED, "fft: could not factor %lld into radix-2..16 stages", (long long)M);
    return nullptr;
  }
  const int R1 = p->R1, R2 = p->R2;
  const int64_t nhi = (M >> 10) + 2;
  std::vector<float2> tw1(R1), tw2(R2), mlo(1024), mhi(nhi), nlo(1024), nhi_(nhi), twh(R2);
  auto unit = [](double num, double den) {
    const double a = -2.0 * M_PI * (num / den);
    return make_float2((float)std::cos(a), (float)std::sin(a));
  };
  for (int t = 0; t < R1; ++t) tw1[t] = unit(t, R1);
  for (int t = 0; t < R2; ++t) { tw2[t] = unit(t, R2); twh[t] = unit(t, 2.0 * (double)R2); }
  for (int l = 0; l < 1024; ++l) { mlo[l] = unit(l, (double)M); nlo[l] = unit(l, 2.0 * (double)M); }
  for (int64_t h = 0; h < nhi; ++h) { mhi[h] = unit((double)(h << 10), (double)M); nhi_[h] = unit((double)(h << 10), 2.0 * (double)M); }
  std::vector<int> perm1, perm2, pos1(R1), pos2(R2);
  digit_perm(R1, p->rd1, 0, perm1);
  digit_perm(R2, p->rd2, 0, perm2);
  for (int i = 0; i < R1; ++i) pos1[perm1[i]] = i;
  for (int i = 0; i < R2; ++i) pos2[perm2[i]] = i;
  std::vector<int> pairT(R2), pairZ(R2);
  std::vector<float2> twhp(R2);
  for (int i = 0; i < R2; ++i) {
    const int k2 = perm2[i];
    pairT[i] = pos2[R2 - 1 - k2];
    pairZ[i] = pos2[(R2 - k2) % R2];
    twhp[i] = twh[k2];
  }
  size_t bytes = 256 * 16 + sizeof(float2) * (R1 + 3 * R2 + 2048 + 2 * nhi) + sizeof(int) * 2 * (R1 + 2 * R2);
  if (cudaMalloc(&p->d_block, bytes) != cudaSuccess) {
    delete p;
    fail(EGR_ERR_CUDA, "fft: cudaMalloc of %zu table bytes failed", bytes);
    return nullptr;
  }
  char* q = (char*)p->d_block;
  p->tw1 = carve<float2>(q, R1); p->tw2 = carve<float2>(q, R2);
  p->perm1 = carve<int>(q, R1); p->pos1 = carve<int>(q, R1);
  p->perm2 = carve<int>(q, R2); p->pos2 = carve<int>(q, R2);
  p->twM_lo = carve<float2>(q, 1024); p->twM_hi = carve<float2>(q, nhi);
  p->twN_lo = carve<float2>(q, 1024); p->twN_hi = carve<float2>(q, nhi);
  p->twH = carve<float2>(q, R2); p->twHp = carve<float2>(q, R2);
  p->pairT = carve<int>(q, R2); p->pairZ = carve<int>(q, R2);
  auto up = [](void* d, const void* h, size_t n) { return cudaMemcpy(d, h, n, cudaMemcpyHostToDevice) == cudaSuccess; };
  bool ok = up(p->tw1, tw1.data(), sizeof(float2) * R1) && up(p->tw2, tw2.data(), sizeof(float2) * R2) &&
            up(p->perm1, perm1.data(), sizeof(int) * R1) && up(p->pos1, pos1.data(), sizeof(int) * R1) &&
            up(p->perm2, perm2.data(), sizeof(int) * R2) && up(p->pos2, pos2.data(), sizeof(int) * R2) &&
            up(p->twM_lo, mlo.data(), sizeof(float2) * 1024) && up(p->twM_hi, mhi.data(), sizeof(float2) * nhi) &&
            up(p->twN_lo, nlo.data(), sizeof(float2) * 1024) && up(p->twN_hi, nhi_.data(), sizeof(float2) * nhi) &&
            up(p->twH, twh.data(), sizeof(float2) * R2) && up(p->twHp, twhp.data(), sizeof(float2) * R2) &&
            up(p->pairT, pairT.data(), sizeof(int) * R2) && up(p->pairZ, pairZ.data(), sizeof(int) * R2);
  if (!ok) {
    cudaFree(p->d_block);
    delete p;
    fail(EGR_ERR_CUDA, "fft: table upload failed");
    return nullptr;
  }
  // column tile width: keep tile + staged twiddles under ~56 KB so FOUR CTAs share an SM (measured on c4: 2 CTAs/SM
  // at cw=4 330 us/iteration, 4 CTAs/SM at cw=2 281 us; the loop is latency-bound between its block-wide barriers)
  if (getenv("EGR_FFT_CW")) {
    p->cw = 1;
    while (p->cw * 2 <= atoi(getenv("EGR_FFT_CW")) && p->cw < 256) p->cw <<= 1;  // power of two (kernels use shifts)
    while (p->cw > 2 && (size_t)R1 * p->cw * sizeof(float2) > 96 * 1024) p->cw >>= 1;
  } else {
    p->cw = 8;
    while (p->cw > 2 && (size_t)R1 * (p->cw + 1) * sizeof(float2) > 56 * 1024) p->cw >>= 1;
  }
  if (R1 == 1) p->cw = 256;
  if (const char* e = getenv("EGR_FL_MINB")) { const int t = atoi(e); if (t >= 2 && t <= 4) p->min_blocks = t; }
  const int tmax = p->min_blocks == 4 ? 256 : 320;
  if (const char* e = getenv("EGR_FL_ROW_T")) { const int t = atoi(e); if (t >= 64 && t <= tmax && t % 32 == 0) p->row_threads = t; }
  if (const char* e = getenv("EGR_FL_COL_T")) { const int t = atoi(e); if (t >= 64 && t <= tmax && t % 32 == 0) p->col_threads = t; }
  g_plans[key] = p;
  return p;
}

// ------------------------------------------------------------------------------------------------ generic kernels
struct Fft2Dev {
  int R1, R2, cw;
  long long M;
  Radices rd1, rd2;
  const float2 *tw1, *tw2, *twM_lo, *twM_hi;
  const int *perm1, *pos1, *perm2, *pos2;
};
static Fft2Dev dev_of(const Fft2Plan* p) {
  Fft2Dev d;
  d.R1 = p->R1; d.R2 = p->R2; d.cw = p->cw; d.M = p->M; d.rd1 = p->rd1; d.rd2 = p->rd2;
  d.tw1 = p->tw1; d.tw2 = p->tw2; d.twM_lo = p->twM_lo; d.twM_hi = p->twM_hi;
  d.perm1 = p->perm1; d.pos1 = p->pos1; d.perm2 = p->perm2; d.pos2 = p->pos2;
  return d;
}

__device__ __forceinline__ float2 twiddle_M(const Fft2Dev& d, long long idx) {
  return cmulf(__ldg(d.twM_hi + (idx >> 10)), __ldg(d.twM_lo + (idx & 1023)));
}

// column pass: forward  natural in[n1*R2+n2]  -> positions out[p1*R2+n2]
//              inverse  positions in          -> natural out
__global__ void __launch_bounds__(256) fft2_col_kernel(Fft2Dev d, const float2* __restrict__ in, float2* __restrict__ out, int inverse) {
  extern __shared__ float2 sm[];
  const int cw = d.cw, R1 = d.R1, R2 = d.R2;
  const int c0 = blockIdx.x * cw;
  const float2* src = in + (long long)blockIdx.y * d.M;
  float2* dst = out + (long long)blockIdx.y * d.M;
  for (int i = threadIdx.x; i < R1 * cw; i += blockDim.x) {
    const int r = i / cw, c = i - r * cw;
    sm[i] = (c0 + c < R2) ? src[(long long)r * R2 + c0 + c] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  const Tile g{R1, cw, cw, 1};
  if (!inverse) fft_forward<true>(sm, g, d.rd1, d.tw1);
  else fft_inverse<true>(sm, g, d.rd1, d.tw1);
  for (int i = threadIdx.x; i < R1 * cw; i += blockDim.x) {
    const int r = i / cw, c = i - r * cw;
    if (c0 + c < R2) dst[(long long)r * R2 + c0 + c] = sm[i];
  }
}

// row pass, in place on positions: forward = twiddle, DIF ; inverse = adjoint DIF, conj twiddle, scale
__global__ void __launch_bounds__(256) fft2_row_kernel(Fft2Dev d, float2* __restrict__ data, int inverse, float scale) {
  extern __shared__ float2 sm[];
  const int R2 = d.R2;
  const int p1 = blockIdx.x;
  const int k1 = d.perm1[p1];
  float2* row = data + (long long)blockIdx.y * d.M + (long long)p1 * R2;
  const Tile g{R2, 1, 1, R2};
  if (!inverse) {
    for (int i = threadIdx.x; i < R2; i += blockDim.x) sm[i] = cmulf(row[i], twiddle_M(d, (long long)i * k1));
    __syncthreads();
    fft_forward<false>(sm, g, d.rd2, d.tw2);
    for (int i = threadIdx.x; i < R2; i += blockDim.x) row[i] = sm[i];
  } else {
    for (int i = threadIdx.x; i < R2; i += blockDim.x) sm[i] = row[i];
    __syncthreads();
    fft_inverse<false>(sm, g, d.rd2, d.tw2);
    for (int i = threadIdx.x; i < R2; i += blockDim.x) {
      const float2 v = cmulc(sm[i], twiddle_M(d, (long long)i * k1));
      row[i] = make_float2(v.x * scale, v.y * scale);
    }
  }
}

// positions <-> natural frequency order
__global__ void fft2_perm_kernel(Fft2Dev d, const float2* __restrict__ in, float2* __restrict__ out, int to_natural) {
  const long long b = blockIdx.y;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < d.M; k += (long long)gridDim.x * blockDim.x) {
    const int k1 = (int)(k % d.R1), k2 = (int)(k / d.R1);
    const long long pidx = (long long)d.pos1[k1] * d.R2 + d.pos2[k2];
    if (to_natural) out[b * d.M + k] = in[b * d.M + pidx];
    else out[b * d.M + pidx] = in[b * d.M + k];
  }
}

static int set_smem(const void* fn, size_t bytes) {
  if (bytes > 48 * 1024) EGR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return EGR_OK;
}

namespace egr {
// natural -> positions (forward) / positions -> natural (inverse, unnormalised * scale); in != out for the column pass
int fft2_forward_to_positions(const Fft2Plan* p, const float2* d_in, float2* d_out, int batch, cudaStream_t st) {
  const Fft2Dev d = dev_of(p);
  const size_t smc = (size_t)p->R1 * p->cw * sizeof(float2), smr = (size_t)p->R2 * sizeof(float2);
  int rc = set_smem((const void*)fft2_col_kernel, smc); if (rc) return rc;
  rc = set_smem((const void*)fft2_row_kernel, smr); if (rc) return rc;
  fft2_col_kernel<<<dim3(ceil_div(p->R2, p->cw), batch), 256, smc, st>>>(d, d_in, d_out, 0);
  EGR_CHECK_LAUNCH("fft2_col_kernel");
  fft2_row_kernel<<<dim3(p->R1, batch), 256, smr, st>>>(d, d_out, 0, 1.0f);
  EGR_CHECK_LAUNCH("fft2_row_kernel");
  return EGR_OK;
}
int fft2_inverse_from_positions(const Fft2Plan* p, float2* d_pos, float2* d_out, int batch, float scale, cudaStream_t st) {
  const Fft2Dev d = dev_of(p);
  const size_t smc = (size_t)p->R1 * p->cw * sizeof(float2), smr = (size_t)p->R2 * sizeof(float2);
  int rc = set_smem((const void*)fft2_col_kernel, smc); if (rc) return rc;
  rc = set_smem((const void*)fft2_row_kernel, smr); if (rc) return rc;
  fft2_row_kernel<<<dim3(p->R1, batch), 256, smr, st>>>(d, d_pos, 1, scale);
  EGR_CHECK_LAUNCH("fft2_row_kernel");
  fft2_col_kernel<<<dim3(ceil_div(p->R2, p->cw), batch), 256, smc, st>>>(d, d_pos, d_out, 1);
  EGR_CHECK_LAUNCH("fft2_col_kernel");
  return EGR_OK;
}
}  // namespace egr

// ------------------------------------------------------------------------------------------------ Bluestein
// X[k] = w[k] * sum_j (x[j] w[j]) conj(w[k-j]),  w[j] = exp(-i pi j^2 / n): a circular convolution of length
// P >= 2n-1 done with the two-level FFT in position order (the filter spectrum is stored in the same order).
__global__ void chirp_kernel(long long n, float2* __restrict__ w) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
    const long long q = (j * j) % (2 * n);
    double s, c;
    sincospi((double)q / (double)n, &s, &c);
    w[j] = make_float2((float)c, (float)-s);
  }
}
__global__ void bluestein_filter_kernel(long long n, long long P, const float2* __restrict__ w, float2* __restrict__ b) {
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < P; m += (long long)gridDim.x * blockDim.x) {
    float2 v = make_float2(0.f, 0.f);
    if (m < n) v = make_float2(w[m].x, -w[m].y);
    else if (P - m < n) v = make_float2(w[P - m].x, -w[P - m].y);
    b[m] = v;
  }
}
// a[j] = (conj_in ? conj(x[j]) : x[j]) * w[j], zero padded to P
__global__ void bluestein_pre_kernel(long long n, long long P, const float2* __restrict__ x, const float2* __restrict__ w,
                                     float2* __restrict__ a, int conj_in) {
  const long long b = blockIdx.y;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (long long)gridDim.x * blockDim.x) {
    float2 v = make_float2(0.f, 0.f);
    if (j < n) {
      float2 xv = x[b * n + j];
      if (conj_in) xv.y = -xv.y;
      v = cmulf(xv, w[j]);
    }
    a[b * P + j] = v;
  }
}
__global__ void bluestein_mul_kernel(long long P, float2* __restrict__ a, const float2* __restrict__ bhat) {
  const long long b = blockIdx.y;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (long long)gridDim.x * blockDim.x)
    a[b * P + j] = cmulf(a[b * P + j], bhat[j]);
}
__global__ void bluestein_post_kernel(long long n, long long P, const float2* __restrict__ c, const float2* __restrict__ w,
                                      float2* __restrict__ x, int conj_out, float scale) {
  const long long b = blockIdx.y;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
    float2 v = cmulf(c[b * P + k], w[k]);
    if (conj_out) v.y = -v.y;
    x[b * n + k] = make_float2(v.x * scale, v.y * scale);
  }
}

// ------------------------------------------------------------------------------------------------ C ABI
struct egr_fft_plan {
  int64_t n = 0;
  int batch = 1;
  const Fft2Plan* direct = nullptr;  // n itself is plannable
  const Fft2Plan* conv = nullptr;    // Bluestein: plan of length P
  int64_t P = 0;
  float2* d_chirp = nullptr;         // w[n]
  float2* d_bhat = nullptr;          // FFT_P(filter) in position order
};

static unsigned grid1(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)(devinfo().sm_count ? devinfo().sm_count : 148) * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" int egr_fft_plan_create(int64_t n, int batch, egr_fft_plan** out) {
  if (!out || n < 1 || batch < 1 || batch > 65535) return fail(EGR_ERR_ARG, "egr_fft_plan_create: bad arguments");
  if (!devinfo().inited) return fail(EGR_ERR_STATE, "egr_fft_plan_create: call egr_init first");
  egr_fft_plan* p = new egr_fft_plan();
  p->n = n; p->batch = batch;
  if (fft2_plannable(n)) {
    p->direct = fft2_get_plan(n);
    if (!p->direct) { delete p; return EGR_ERR_UNSUPPORTED; }
  } else {
    p->P = fft2_next_plannable(2 * n - 1);
    if (p->P < 0) { delete p; return fail(EGR_ERR_UNSUPPORTED, "fft: length %lld too large for the Bluestein path", (long long)n); }
    p->conv = fft2_get_plan(p->P);
    if (!p->conv) { delete p; return EGR_ERR_UNSUPPORTED; }
    float2* tmp = nullptr;
    if (cudaMalloc(&p->d_chirp, sizeof(float2) * n) != cudaSuccess || cudaMalloc(&p->d_bhat, sizeof(float2) * p->P) != cudaSuccess ||
        cudaMalloc(&tmp, sizeof(float2) * p->P) != cudaSuccess) {
      cudaFree(p->d_chirp); cudaFree(p->d_bhat); cudaFree(tmp);
      delete p;
      return fail(EGR_ERR_CUDA, "fft: cudaMalloc for the Bluestein tables failed");
    }
    chirp_kernel<<<grid1(n), 256>>>(n, p->d_chirp);
    bluestein_filter_kernel<<<grid1(p->P), 256>>>(n, p->P, p->d_chirp, tmp);
    int rc = fft2_forward_to_positions(p->conv, tmp, p->d_bhat, 1, 0);
    cudaError_t e = cudaDeviceSynchronize();
    cudaFree(tmp);
    if (rc || e != cudaSuccess) {
      cudaFree(p->d_chirp); cudaFree(p->d_bhat);
      delete p;
      return rc ? rc : fail(EGR_ERR_CUDA, "fft: Bluestein table build failed: %s", cudaGetErrorString(e));
    }
  }
  *out = p;
  return EGR_OK;
}

extern "C" size_t egr_fft_plan_workspace_bytes(const egr_fft_plan* p) {
  if (!p) return 0;
  if (p->direct) return sizeof(float2) * (size_t)p->n * p->batch;
  return 2 * sizeof(float2) * (size_t)p->P * p->batch;
}

extern "C" int egr_fft_plan_passes(const egr_fft_plan* p) {
  if (!p) return 0;
  const Fft2Plan* q = p->direct ? p->direct : p->conv;
  const int per = (q->R1 > 1 ? 2 : 1) + 1;  // column + row (+ reorder)
  return p->direct ? per : 2 * per + 3;
}

extern "C" int egr_fft_exec(egr_fft_plan* p, float* d_data, float* d_work, int inverse, int scale_inverse, void* stream) {
  if (!p || !d_data || !d_work) return fail(EGR_ERR_ARG, "egr_fft_exec: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  float2* x = reinterpret_cast<float2*>(d_data);
  float2* wk = reinterpret_cast<float2*>(d_work);
  const float scale = (inverse && scale_inverse) ? (float)(1.0 / (double)p->n) : 1.0f;
  if (p->direct) {
    const Fft2Dev d = dev_of(p->direct);
    if (!inverse) {
      int rc = fft2_forward_to_positions(p->direct, x, wk, p->batch, st);
      if (rc) return rc;
      fft2_perm_kernel<<<dim3(grid1(p->n), p->batch), 256, 0, st>>>(d, wk, x, 1);
      EGR_CHECK_LAUNCH("fft2_perm_kernel");
    } else {
      fft2_perm_kernel<<<dim3(grid1(p->n), p->batch), 256, 0, st>>>(d, x, wk, 0);
      EGR_CHECK_LAUNCH("fft2_perm_kernel");
      int rc = fft2_inverse_from_positions(p->direct, wk, x, p->batch, scale, st);
      if (rc) return rc;
    }
    return EGR_OK;
  }
  // Bluestein; the inverse uses ifft(X) = conj(fft(conj(X)))
  float2* a = wk;
  float2* a2 = wk + (size_t)p->P * p->batch;
  bluestein_pre_kernel<<<dim3(grid1(p->P), p->batch), 256, 0, st>>>(p->n, p->P, x, p->d_chirp, a, inverse);
  EGR_CHECK_LAUNCH("bluestein_pre_kernel");
  int rc = fft2_forward_to_positions(p->conv, a, a2, p->batch, st);
  if (rc) return rc;
  bluestein_mul_kernel<<<dim3(grid1(p->P), p->batch), 256, 0, st>>>(p->P, a2, p->d_bhat);
  EGR_CHECK_LAUNCH("bluestein_mul_kernel");
  rc = fft2_inverse_from_positions(p->conv, a2, a, p->batch, (float)(1.0 / (double)p->P), st);
  if (rc) return rc;
  bluestein_post_kernel<<<dim3(grid1(p->n), p->batch), 256, 0, st>>>(p->n, p->P, a, p->d_chirp, x, inverse, scale);
  EGR_CHECK_LAUNCH("bluestein_post_kernel");
  return EGR_OK;
}

extern "C" void egr_fft_plan_destroy(egr_fft_plan* p) {
  if (!p) return;
  cudaFree(p->d_chirp);
  cudaFree(p->d_bhat);
  delete p;
}
