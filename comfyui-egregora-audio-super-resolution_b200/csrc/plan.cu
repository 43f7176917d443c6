// plan.cu — the FlashSR plan executor: a straight-line list of egr_op over one workspace and one weight
// blob.  Creation validates every op and encodes the TMA tensor maps of the tensor-core GEMMs once; run just
// enqueues kernels on the caller's stream (no allocation, no synchronisation), so a whole forward pass can be
// captured into a CUDA graph by the host.
#include <cstdlib>
#include <vector>
#include "ops.cuh"
#include "mega.cuh"

using namespace egr;

struct egr_plan {
  std::vector<egr_op> ops;
  std::vector<TcPrepared*> tc;  // per op, nullptr unless GEMM_TC
  Spaces sp;
  void* scratch = nullptr;      // split-K partial tiles followed by the per-tile arrival counters (library-owned)
  std::vector<MegaRun*> runs;   // persistent-kernel runs (mega.cu): consecutive ops flagged EGR_FLAG_MEGA
  std::vector<int> run_at;      // per op: index into `runs` of the run that STARTS here, else -1
  // whole-plan CUDA graph: the op list is static (fixed addresses, shapes, launch geometry), so after one eager pass
  // the ~900 launches are captured once and replayed with a single cudaGraphLaunch
  cudaGraphExec_t graph_exec = nullptr;
  int full_runs = 0;
  unsigned long long launches_per_run = 0;
};

static int run_op(const egr_plan* p, int i, cudaStream_t st) {
  const egr_op& op = p->ops[i];
  switch (op.code) {
    case EGR_OP_GEMM_TC: return tc_launch(p->tc[i], st);
    case EGR_OP_GEMM_SIMT: return launch_gemm_simt(p->sp, op, st);
    case EGR_OP_GN_STATS: return launch_gn_stats(p->sp, op, st);
    case EGR_OP_GN_APPLY: return launch_gn_apply(p->sp, op, st);
    case EGR_OP_LAYERNORM: return launch_layernorm(p->sp, op, st);
    case EGR_OP_SOFTMAX: return launch_softmax(p->sp, op, st);
    case EGR_OP_ATTN_SMALL: return launch_attn_small(p->sp, op, st);
    case EGR_OP_GEGLU: return launch_geglu(p->sp, op, st);
    case EGR_OP_ELTWISE: return launch_eltwise(p->sp, op, st);
    case EGR_OP_SNAKE_AA: return launch_snake_aa(p->sp, op, st);
    case EGR_OP_STFT_MEL: return launch_stft_mel(p->sp, op, st);
    case EGR_OP_LOWPASS: return launch_lowpass(p->sp, op, st);
    case EGR_OP_TIME_EMBED: return launch_time_embed(p->sp, op, st);
    case EGR_OP_ZERO: return launch_zero(p->sp, op, st);
    default: return fail(EGR_ERR_ARG, "op %d (%s): unknown code %d", i, op.name, op.code);
  }
}

extern "C" int egr_plan_create(const egr_op* h_ops, int n_ops, void* d_workspace, size_t ws_bytes,
                               const void* d_weights, size_t wt_bytes, egr_plan** out) {
  if (!h_ops || n_ops <= 0 || !out) return fail(EGR_ERR_ARG, "egr_plan_create: bad arguments");
  if (!devinfo().inited) return fail(EGR_ERR_STATE, "egr_plan_create: call egr_init first");
  egr_plan* p = new egr_plan();
  p->sp.ws = (char*)d_workspace; p->sp.ws_bytes = ws_bytes;
  p->sp.wt = (const char*)d_weights; p->sp.wt_bytes = wt_bytes;
  p->ops.assign(h_ops, h_ops + n_ops);
  p->tc.assign(n_ops, nullptr);
  for (int i = 0; i < n_ops; ++i) {
    egr_op& op = p->ops[i];
    op.name[sizeof(op.name) - 1] = 0;
    // bounds-check every address against its space
    auto chk = [&](uint64_t a) -> bool {
      uint64_t space = a >> 60, off = a & 0x0FFFFFFFFFFFFFFFull;
      if (space == EGR_SPACE_WS) return off < ws_bytes;
      if (space == EGR_SPACE_WT) return off < wt_bytes;
      return true;
    };
    bool ok = chk(op.x0.addr) && chk(op.x1.addr);
    for (int k = 0; k < 10; ++k) ok = ok && chk(op.ptr[k]);
    if (!ok) {
      int rc = fail(EGR_ERR_ARG, "op %d (%s): address outside its workspace/weights space", i, op.name);
      egr_plan_destroy(p);
      return rc;
    }
    if (op.code == EGR_OP_GEMM_TC) {
      int rc = tc_prepare(p->sp, op, &p->tc[i]);
      if (rc) { egr_plan_destroy(p); return rc; }
    }
  }
  // split-K scratch: ops run in stream order, so one buffer sized for the largest op serves them all
  size_t part = 0; int ctr = 0;
  for (TcPrepared* t : p->tc)
    if (t) { part = tc_partial_bytes(t) > part ? tc_partial_bytes(t) : part; ctr = tc_num_counters(t) > ctr ? tc_num_counters(t) : ctr; }
  if (part > 0) {
    part = (part + 255) / 256 * 256;
    cudaError_t e = cudaMalloc(&p->scratch, part + (size_t)ctr * sizeof(unsigned int));
    if (e != cudaSuccess) {
      int rc = fail(EGR_ERR_CUDA, "egr_plan_create: cudaMalloc of %zu B split-K scratch failed: %s", part, cudaGetErrorString(e));
      egr_plan_destroy(p);
      return rc;
    }
    unsigned int* counters = reinterpret_cast<unsigned int*>(static_cast<char*>(p->scratch) + part);
    cudaMemset(counters, 0, (size_t)ctr * sizeof(unsigned int));
    for (TcPrepared* t : p->tc)
      if (t && tc_partial_bytes(t) > 0) tc_bind_scratch(t, static_cast<float*>(p->scratch), counters);
  }
  // persistent-kernel runs: maximal stretches of flagged ops the megakernel supports (the UNet of every diffusion step)
  p->run_at.assign(n_ops, -1);
  if (getenv("EGR_NO_MEGA") == nullptr) {
    int i = 0;
    while (i < n_ops) {
      int j = i;
      while (j < n_ops && (p->ops[j].flags & EGR_FLAG_MEGA) && mega_supports(p->ops[j])) ++j;
      if (j - i >= 4) {
        MegaRun* r = nullptr;
        int rc = mega_build(p->sp, p->ops.data(), p->tc.data(), i, j, &r);
        if (rc) { egr_plan_destroy(p); return rc; }
        if (r) { p->run_at[i] = (int)p->runs.size(); p->runs.push_back(r); }
      }
      i = j > i ? j : i + 1;
    }
  }
  *out = p;
  return EGR_OK;
}

// ops [first, last) in stream order; a persistent-kernel run that lies wholly inside the range is ONE launch
static int run_range(const egr_plan* plan, int first, int last, cudaStream_t st) {
  for (int i = first; i < last;) {
    const int r = plan->run_at[i];
    if (r >= 0) {
      int d[8];
      mega_describe(plan->runs[r], d);
      if (d[1] <= last) {
        int rc = mega_launch(plan->runs[r], st);
        if (rc) return rc;
        i = d[1];
        continue;
      }
    }
    int rc = run_op(plan, i, st);
    if (rc) return rc;
    ++i;
  }
  return EGR_OK;
}

// debug hook (not in the public header): persistent-kernel runs of a plan -> out[0] = count, then 7 ints per run
// (first, last, ops, gemm ops, grid barriers, smem bytes, grid), at most `max_runs` of them
extern "C" int egr_debug_mega_info(const egr_plan* plan, int* out, int max_runs) {
  if (!plan || !out) return -1;
  out[0] = (int)plan->runs.size();
  for (int r = 0; r < (int)plan->runs.size() && r < max_runs; ++r) {
    int d[8];
    mega_describe(plan->runs[r], d);
    for (int k = 0; k < 7; ++k) out[1 + 7 * r + k] = d[k];
  }
  return 0;
}

// debug hook (EGR_MEGA_TRACE=1): clock trace of run `run` -> number of ops copied
extern "C" int egr_debug_mega_trace(const egr_plan* plan, int run, unsigned long long* h_stamps, int* h_codes, int max_ops) {
  if (!plan || run < 0 || run >= (int)plan->runs.size()) return -1;
  return mega_trace(plan->runs[run], h_stamps, h_codes, max_ops);
}

// debug hook: number of persistent-kernel launches of this plan that a barrier watchdog abandoned (must be 0)
extern "C" int egr_debug_mega_aborted(const egr_plan* plan) {
  if (!plan) return -1;
  int n = 0;
  for (MegaRun* r : plan->runs) n += mega_aborted(r) > 0;
  return n;
}

// debug hook (not in the public header): tile configuration the library chose for op `op_index`
extern "C" int egr_debug_tc_config(const egr_plan* plan, int op_index, int* out8) {
  if (!plan || op_index < 0 || op_index >= (int)plan->ops.size() || !plan->tc[op_index]) return -1;
  tc_describe(plan->tc[op_index], out8);
  return 0;
}

extern "C" int egr_plan_run(egr_plan* plan, int first, int last, void* stream) {
  if (!plan) return fail(EGR_ERR_ARG, "egr_plan_run: null plan");
  const int n = (int)plan->ops.size();
  if (first < 0) first = 0;
  if (last < 0 || last > n) last = n;
  cudaStream_t st = (cudaStream_t)stream;
  const bool full = (first == 0 && last == n);
  static const bool use_graph = getenv("EGR_NO_GRAPH") == nullptr;
  if (full && use_graph && plan->graph_exec) {
    EGR_CUDA(cudaGraphLaunch(plan->graph_exec, st));
    launch_count() += plan->launches_per_run;
    return EGR_OK;
  }
  if (full && use_graph && plan->full_runs >= 1 && !plan->graph_exec) {
    // second full pass: every lazy one-time setting (function attributes, occupancy queries) has been made
    const unsigned long long before = launch_count();
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      int rc = run_range(plan, first, last, st);
      cudaError_t e = cudaStreamEndCapture(st, &graph);
      if (rc == EGR_OK && e == cudaSuccess && graph && cudaGraphInstantiate(&plan->graph_exec, graph, 0) == cudaSuccess) {
        plan->launches_per_run = launch_count() - before;
        cudaGraphDestroy(graph);
        launch_count() = before;
        EGR_CUDA(cudaGraphLaunch(plan->graph_exec, st));
        launch_count() += plan->launches_per_run;
        ++plan->full_runs;
        return EGR_OK;
      }
      if (graph) cudaGraphDestroy(graph);
      plan->graph_exec = nullptr;
      launch_count() = before;
      cudaGetLastError();
      if (rc) return rc;
      // capture failed: fall through to the eager path (and do not try again)
      plan->full_runs = -1000000;
    } else {
      cudaGetLastError();  // e.g. the legacy default stream cannot be captured: stay eager on this plan
      plan->full_runs = -1000000;
    }
  }
  {
    int rc = run_range(plan, first, last, st);
    if (rc) return rc;
  }
  if (full) ++plan->full_runs;
  return EGR_OK;
}

extern "C" int egr_plan_num_launches(const egr_plan* plan, int first, int last) {
  if (!plan) return 0;
  const int n = (int)plan->ops.size();
  if (first < 0) first = 0;
  if (last < 0 || last > n) last = n;
  return last > first ? last - first : 0;
}

// ops that execute inside a persistent-kernel run are not launches of their own: the measurement helpers skip them
static bool in_mega_run(const egr_plan* plan, int i) {
  for (const MegaRun* r : plan->runs) {
    int d[8];
    mega_describe(r, d);
    if (i >= d[0] && i < d[1]) return true;
  }
  return false;
}

extern "C" int egr_plan_run_code(egr_plan* plan, int code, void* stream) {
  if (!plan) return fail(EGR_ERR_ARG, "egr_plan_run_code: null plan");
  for (int i = 0; i < (int)plan->ops.size(); ++i) {
    if (plan->ops[i].code != code || in_mega_run(plan, i)) continue;
    int rc = run_op(plan, i, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return EGR_OK;
}

extern "C" int egr_plan_count_code(const egr_plan* plan, int code) {
  if (!plan) return 0;
  int n = 0;
  for (int i = 0; i < (int)plan->ops.size(); ++i) n += (plan->ops[i].code == code && !in_mega_run(plan, i));
  return n;
}

extern "C" void egr_plan_destroy(egr_plan* plan) {
  if (!plan) return;
  for (TcPrepared* t : plan->tc)
    if (t) tc_free(t);
  for (MegaRun* r : plan->runs) mega_free(r);
  if (plan->graph_exec) cudaGraphExecDestroy(plan->graph_exec);
  if (plan->scratch) cudaFree(plan->scratch);
  delete plan;
}

// debug hook (not in the public header): 1 when the whole-plan CUDA graph has been instantiated, 0 when the plan runs
// eagerly (before its second full pass, or because capture failed)
extern "C" int egr_debug_plan_graphed(const egr_plan* plan) { return plan && plan->graph_exec ? 1 : 0; }
