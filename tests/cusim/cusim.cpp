// cusim scheduler — see cuda_runtime.h in this directory.  TEST INFRASTRUCTURE ONLY.
#include "cuda_runtime.h"
#ifdef CUSIM_CLUSTERS
#include <pthread.h>
#include <thread>
#endif

CUSIM_TLS uint3 threadIdx, blockIdx;
CUSIM_TLS dim3 blockDim, gridDim;

namespace cusim {

struct ClusterCtx {
  unsigned n = 1;
  char* anchor[16] = {};
#ifdef CUSIM_CLUSTERS
  pthread_barrier_t bar;
#endif
};

static CUSIM_TLS Block g_blk;
static unsigned long long g_launches = 0;
alignas(256) static CUSIM_TLS unsigned char g_dyn[232448];       // 227 KB
static CUSIM_TLS unsigned char* g_stacks = nullptr;
static CUSIM_TLS size_t g_stack_threads = 0;
static CUSIM_TLS ClusterCtx* g_cluster = nullptr;
static CUSIM_TLS unsigned g_rank = 0;
static CUSIM_TLS char g_anchor;   // any thread_local object: the distance between two threads' copies is the TLS-block distance
static const size_t STACK = 64 * 1024;

Block& blk() { return g_blk; }
void* dyn_smem() { return g_dyn; }
unsigned long long launches() { return g_launches; }
unsigned cluster_size() { return g_cluster ? g_cluster->n : 1; }
unsigned cluster_rank() { return g_rank; }
ptrdiff_t cluster_delta(unsigned r) { return g_cluster ? g_cluster->anchor[r] - &g_anchor : 0; }

void yield(State s, unsigned mask) {
  Fiber& f = g_blk.f[g_blk.cur];
  f.st = s;
  f.wmask = mask;
  swapcontext(&f.uc, &g_blk.sched);
}

void cluster_sync() { yield(g_cluster ? WAIT_CLUSTER : WAIT_BLOCK, 0); }

static void trampoline() {
  (*g_blk.body)();
  Fiber& f = g_blk.f[g_blk.cur];
  f.st = DONE;
  swapcontext(&f.uc, &g_blk.sched);
}

static void run_block() {
  Block& b = g_blk;
  const size_t n = b.f.size();
  for (size_t i = 0; i < n; ++i) {
    Fiber& f = b.f[i];
    getcontext(&f.uc);
    f.uc.uc_stack.ss_sp = g_stacks + i * STACK;
    f.uc.uc_stack.ss_size = STACK;
    f.uc.uc_link = nullptr;
    makecontext(&f.uc, trampoline, 0);
    f.st = READY;
  }
  // CUSIM_ORDER=reverse resumes the fibers of a block in descending order: code that misses a barrier tends to give
  // different results under the two orders (a poor man's racecheck; tests/test_cusim.py runs the new kernels both ways)
  static const bool reverse = getenv("CUSIM_ORDER") && !strcmp(getenv("CUSIM_ORDER"), "reverse");
  for (;;) {
    bool ran = false;
    for (size_t j = 0; j < n; ++j) {
      const size_t i = reverse ? n - 1 - j : j;
      if (b.f[i].st != READY) continue;
      b.cur = (int)i;
      threadIdx = b.f[i].tid;
      swapcontext(&b.sched, &b.f[i].uc);
      ran = true;
    }
    b.cur = -1;
    size_t done = 0, at_block = 0, at_cluster = 0;
    for (size_t i = 0; i < n; ++i) { done += b.f[i].st == DONE; at_block += b.f[i].st == WAIT_BLOCK; at_cluster += b.f[i].st == WAIT_CLUSTER; }
    if (done == n) return;
    bool released = false;
    // warp barriers: a waiting lane is released when every live lane of its mask waits at a warp barrier too
    for (size_t w0 = 0; w0 < n; w0 += 32) {
      const size_t w1 = std::min(n, w0 + 32);
      bool any = false, ok = true;
      unsigned need = 0;
      for (size_t i = w0; i < w1; ++i) if (b.f[i].st == WAIT_WARP) { any = true; need |= b.f[i].wmask; }
      if (!any) continue;
      for (size_t i = w0; i < w1; ++i)
        if (((need >> (i - w0)) & 1u) && b.f[i].st != WAIT_WARP && b.f[i].st != DONE) ok = false;
      if (ok) { for (size_t i = w0; i < w1; ++i) if (b.f[i].st == WAIT_WARP) b.f[i].st = READY; released = true; }
    }
    if (!released && at_block > 0 && at_block + done == n) {
      for (size_t i = 0; i < n; ++i) if (b.f[i].st == WAIT_BLOCK) b.f[i].st = READY;
      released = true;
    }
#ifdef CUSIM_CLUSTERS
    if (!released && at_cluster > 0 && at_cluster + done == n) {   // the whole CTA has arrived: meet the other CTAs
      pthread_barrier_wait(&g_cluster->bar);
      for (size_t i = 0; i < n; ++i) if (b.f[i].st == WAIT_CLUSTER) b.f[i].st = READY;
      released = true;
    }
#endif
    if (!released) {
      bool any_ready = false;
      for (size_t i = 0; i < n; ++i) any_ready |= b.f[i].st == READY;
      if (!any_ready) {
        fprintf(stderr, "cusim: deadlock / divergent barrier in block (%u,%u,%u): %zu done, %zu at __syncthreads, %zu at cluster.sync of %zu threads\n",
                b.bid.x, b.bid.y, b.bid.z, done, at_block, at_cluster, n);
        abort();
      }
    }
    (void)ran;
  }
}

static void setup_block(dim3 grid, dim3 block, const std::function<void()>& body) {
  const size_t n = (size_t)block.x * block.y * block.z;
  if (g_stack_threads < n) { free(g_stacks); g_stacks = (unsigned char*)aligned_alloc(4096, n * STACK); g_stack_threads = n; }
  Block& b = g_blk;
  b.bdim = block;
  b.gdim = grid;
  blockDim = block;
  gridDim = grid;
  b.body = &body;
  b.f.resize(n);
  for (size_t i = 0; i < n; ++i) b.f[i].tid = uint3{(unsigned)(i % block.x), (unsigned)((i / block.x) % block.y), (unsigned)(i / ((size_t)block.x * block.y))};
}

bool launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body, unsigned cluster) {
  const size_t n = (size_t)block.x * block.y * block.z;
  if (n == 0 || n > 1024 || smem > sizeof(g_dyn) || (size_t)grid.x * grid.y * grid.z == 0) {
    fprintf(stderr, "cusim: invalid launch configuration (%zu threads, %zu bytes of dynamic shared memory)\n", n, smem);
    abort();
  }
  if (cluster == 0 || cluster > 16 || grid.x % cluster) return false;
  ++g_launches;
  if (cluster == 1) {
    setup_block(grid, block, body);
    for (unsigned z = 0; z < grid.z; ++z)
      for (unsigned y = 0; y < grid.y; ++y)
        for (unsigned x = 0; x < grid.x; ++x) {
          g_blk.bid = uint3{x, y, z};
          blockIdx = g_blk.bid;
          run_block();
        }
    g_blk.body = nullptr;
    return true;
  }
#ifdef CUSIM_CLUSTERS
  for (unsigned z = 0; z < grid.z; ++z)
    for (unsigned y = 0; y < grid.y; ++y)
      for (unsigned x0 = 0; x0 < grid.x; x0 += cluster) {
        ClusterCtx ctx;
        ctx.n = cluster;
        pthread_barrier_init(&ctx.bar, nullptr, cluster);
        std::vector<std::thread> cta;
        for (unsigned r = 0; r < cluster; ++r)
          cta.emplace_back([&, r]() {
            g_cluster = &ctx;
            g_rank = r;
            ctx.anchor[r] = &g_anchor;
            setup_block(grid, block, body);
            g_blk.bid = uint3{x0 + r, y, z};
            blockIdx = g_blk.bid;
            pthread_barrier_wait(&ctx.bar);   // every CTA's anchor is published before any kernel code runs
            run_block();
            pthread_barrier_wait(&ctx.bar);   // shared memory of a CTA stays alive until the whole cluster is done
            free(g_stacks);
            g_stacks = nullptr;
            g_stack_threads = 0;
          });
        for (auto& t : cta) t.join();
        pthread_barrier_destroy(&ctx.bar);
      }
  return true;
#else
  return false;
#endif
}

}  // namespace cusim
