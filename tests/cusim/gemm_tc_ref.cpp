// cusim stand-in for csrc/gemm_tc.cu — TEST INFRASTRUCTURE ONLY.  tcgen05 / TMA cannot be emulated, so under the
// emulator an EGR_OP_GEMM_TC op is evaluated by plain host loops with the op's exact semantics (f16 operands, f32
// accumulation, taps as shifted reads of the A view with zero fill, batch-indexed B, bias / row bias / activation /
// residuals / post scale, dense, strided-phase and transposed output layouts, crop window).  This says nothing about
// the real kernel — that one is only ever checked on the GPU — it lets the REST of a plan (the SIMT kernels, the
// executor) run on the CPU around it.
#include <vector>
#include "ops.cuh"
#include "mega.cuh"

namespace egr {

struct TcPrepared {
  GemmArgs g;
  Taps taps;
  View a;
  char name[48];
};

int tc_global_init() { return EGR_OK; }

int tc_prepare(const Spaces& s, const egr_op& op, TcPrepared** out) {
  TcPrepared* p = new TcPrepared();
  int rc = gemm_args_from_op(s, op, &p->g, &p->taps, &p->a);
  if (rc) { delete p; return rc; }
  snprintf(p->name, sizeof(p->name), "%s", op.name);
  if (p->a.elem != 1) { delete p; return fail(EGR_ERR_ARG, "%s: tensor-core path needs an f16 A operand", op.name); }
  if (p->g.N % 16) { delete p; return fail(EGR_ERR_ARG, "%s: N=%d must be a multiple of 16", op.name, p->g.N); }
  *out = p;
  return EGR_OK;
}

int tc_launch(const TcPrepared* p, cudaStream_t) {
  const GemmArgs& g = p->g;
  const View& a = p->a;
  const __half* A = reinterpret_cast<const __half*>(a.p);
  const __half* W = reinterpret_cast<const __half*>(g.W);
  const long long wz = g.wstride_z > 0 ? g.wstride_z : g.wstride_n * g.N;
  std::vector<float> arow((size_t)g.ntaps * g.K);
  for (int b = 0; b < g.Bo; ++b)
    for (int h = 0; h < g.Ho; ++h)
      for (int w = 0; w < g.Wo; ++w) {
        for (int t = 0; t < g.ntaps; ++t) {   // gather this pixel's taps x K operand row once
          long long c[5];
          for (int d = 0; d < 5; ++d) c[d] = p->taps.t[t][d];
          c[g.dimW] += w; c[g.dimH] += h; c[g.dimB] += b;
          bool inb = true;
          long long off = 0;
          for (int d = 1; d < 5; ++d) { inb = inb && c[d] >= 0 && c[d] < a.dim[d]; off += c[d] * a.stride[d]; }
          for (int k = 0; k < g.K; ++k) {
            const long long ch = c[0] + k;
            arow[(size_t)t * g.K + k] = (inb && ch >= 0 && ch < a.dim[0]) ? __half2float(A[off + ch * a.stride[0]]) : 0.f;
          }
        }
        const long long pix = (long long)h * g.Wo + w;
        for (int n = 0; n < g.N; ++n) {
          float acc = 0.f;
          for (int t = 0; t < g.ntaps; ++t) {
            const __half* wr = W + (long long)(g.wz_batch ? b : t) * wz + (long long)n * g.wstride_n;
            const float* ar = arow.data() + (size_t)t * g.K;
            for (int k = 0; k < g.K; ++k) acc += ar[k] * __half2float(wr[k]);
          }
          float v = acc * g.alpha;
          if (g.bias) v += g.bias[n];
          if (g.rowbias) v += g.rowbias[(long long)b * g.rowbias_stride + n];
          v = egr_apply_act(v, g.act);
          long long idx;
          if (g.transposed) {
            idx = (long long)b * g.out_batch_stride + (long long)n * g.out_n_stride + pix + g.out_offset;
          } else {
            const long long flat = (long long)h * g.out_h_stride + (long long)w * g.out_pix_stride + g.out_offset + n;
            if (flat < g.out_lo || flat >= g.out_hi) continue;
            idx = (long long)b * g.out_batch_stride + flat;
          }
          if (g.resid) v += g.resid[idx];
          if (g.resid2) v += g.resid2[idx];
          v *= g.post;
          if (g.out32) g.out32[idx] = v;
          if (g.out16) g.out16[idx] = __float2half_rn(v);
        }
      }
  ++launch_count();
  return EGR_OK;
}

void tc_free(TcPrepared* p) { delete p; }
size_t tc_partial_bytes(const TcPrepared*) { return 0; }
int tc_num_counters(const TcPrepared*) { return 0; }
void tc_bind_scratch(TcPrepared*, float*, unsigned int*) {}
void tc_describe(const TcPrepared*, int* o) { for (int i = 0; i < 8; ++i) o[i] = 0; }

// The persistent UNet kernel (csrc/mega.cu) is built on the tcgen05 warp roles: not emulated.  Flagged ops run one by one.
struct MegaRun { int unused; };
bool mega_supports(const egr_op&) { return false; }
int mega_build(const Spaces&, const egr_op*, TcPrepared* const*, int, int, MegaRun** out) { *out = nullptr; return EGR_OK; }
int mega_launch(const MegaRun*, cudaStream_t) { return fail(EGR_ERR_UNSUPPORTED, "cusim: no persistent kernel"); }
void mega_describe(const MegaRun*, int* o) { for (int i = 0; i < 8; ++i) o[i] = 0; }
int mega_aborted(const MegaRun*) { return 0; }
int mega_trace(const MegaRun*, unsigned long long*, int*, int) { return 0; }
void mega_free(MegaRun*) {}

}  // namespace egr
