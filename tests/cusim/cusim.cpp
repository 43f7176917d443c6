// cusim scheduler — see cuda_runtime.h in this directory.  TEST INFRASTRUCTURE ONLY.
#include "cuda_runtime.h"

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace cusim {

static Block g_blk;
static unsigned long long g_launches = 0;
alignas(256) static unsigned char g_dyn[232448];       // 227 KB
static unsigned char* g_stacks = nullptr;
static size_t g_stack_threads = 0;
static const size_t STACK = 64 * 1024;

Block& blk() { return g_blk; }
void* dyn_smem() { return g_dyn; }
unsigned long long launches() { return g_launches; }

void yield(State s, unsigned mask) {
  Fiber& f = g_blk.f[g_blk.cur];
  f.st = s;
  f.wmask = mask;
  swapcontext(&f.uc, &g_blk.sched);
}

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
    size_t done = 0, at_block = 0;
    for (size_t i = 0; i < n; ++i) { done += b.f[i].st == DONE; at_block += b.f[i].st == WAIT_BLOCK; }
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
    if (!released && !ran) {
      fprintf(stderr, "cusim: deadlock in block (%u,%u,%u): %zu done, %zu at __syncthreads of %zu threads\n", b.bid.x, b.bid.y, b.bid.z, done, at_block, n);
      abort();
    }
    if (!released) {
      bool any_ready = false;
      for (size_t i = 0; i < n; ++i) any_ready |= b.f[i].st == READY;
      if (!any_ready) {
        fprintf(stderr, "cusim: divergent barrier in block (%u,%u,%u): %zu done, %zu at __syncthreads of %zu threads\n", b.bid.x, b.bid.y, b.bid.z, done, at_block, n);
        abort();
      }
    }
  }
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const size_t n = (size_t)block.x * block.y * block.z;
  if (n == 0 || n > 1024 || smem > sizeof(g_dyn) || (size_t)grid.x * grid.y * grid.z == 0) {
    fprintf(stderr, "cusim: invalid launch configuration (%zu threads, %zu bytes of dynamic shared memory)\n", n, smem);
    abort();
  }
  ++g_launches;
  if (g_stack_threads < n) { free(g_stacks); g_stacks = (unsigned char*)aligned_alloc(4096, n * STACK); g_stack_threads = n; }
  Block& b = g_blk;
  b.bdim = block;
  b.gdim = grid;
  blockDim = block;
  gridDim = grid;
  b.body = &body;
  b.f.resize(n);
  for (size_t i = 0; i < n; ++i) b.f[i].tid = uint3{(unsigned)(i % block.x), (unsigned)((i / block.x) % block.y), (unsigned)(i / ((size_t)block.x * block.y))};
  for (unsigned z = 0; z < grid.z; ++z)
    for (unsigned y = 0; y < grid.y; ++y)
      for (unsigned x = 0; x < grid.x; ++x) {
        b.bid = uint3{x, y, z};
        blockIdx = b.bid;
        run_block();
      }
  b.body = nullptr;
}

}  // namespace cusim
