// fft_plan.cuh — host-side plan of the two-level ("four-step") complex FFT used by path B.
//
//   M = R1 * R2,  n = R2*n1 + n2,  k = k1 + R1*k2
//   X[k1 + R1 k2] = sum_n2 [ w_M^(n2 k1) * sum_n1 z[R2 n1 + n2] w_R1^(n1 k1) ] w_R2^(n2 k2)
//
// Column pass: length-R1 transforms over n1 (stride R2) for a tile of adjacent columns, in shared memory.
// Row pass   : length-R2 transforms over n2 on contiguous rows, with the w_M twiddle folded into the load/store.
// Both passes are in-place decimation-in-frequency, so the spectrum is kept in "position" order:
//   work[p1*R2 + p2] = X[perm1[p1] + R1*perm2[p2]]
// and the inverse retraces the same steps backwards — no transposition or reordering pass touches HBM.
#pragma once
#include <vector>
#include "common.cuh"
#include "fft_device.cuh"

namespace egr {

struct Fft2Plan {
  int64_t M = 0;
  int R1 = 1, R2 = 1;
  egrfft::Radices rd1{}, rd2{};
  // device tables (one cudaMalloc block)
  void* d_block = nullptr;
  float2 *tw1 = nullptr, *tw2 = nullptr;       // exp(-2 pi i t / R1), exp(-2 pi i t / R2)
  int *perm1 = nullptr, *pos1 = nullptr;       // position -> frequency digit, and its inverse
  int *perm2 = nullptr, *pos2 = nullptr;
  float2 *twM_lo = nullptr, *twM_hi = nullptr; // exp(-2 pi i l / M), l < 1024 ; exp(-2 pi i 1024 h / M)
  float2 *twN_lo = nullptr, *twN_hi = nullptr; // same for N = 2M (real-transform split)
  float2 *twHp = nullptr;                      // twH in position order: twHp[p2] = twH[perm2[p2]]
  int *pairT = nullptr, *pairZ = nullptr;      // pairT[p2] = pos2[R2-1-perm2[p2]], pairZ[p2] = pos2[(R2-perm2[p2]) % R2]
  float2 *twH = nullptr;                       // exp(-2 pi i k2 / (2 R2)) = w_N^(R1 k2), k2 < R2
  int cw = 4;                                  // columns per column-pass CTA
  int col_threads = 256, row_threads = 256;
  int min_blocks = 4;                          // __launch_bounds__ min CTAs/SM variant of the fused Fat-Llama kernels
};

// true when M is {2,3,5,7,11,13}-smooth and splits into R1*R2 with both factors <= 8192
bool fft2_plannable(int64_t M);
// cached per device; returns nullptr (and sets the error text) when M is not plannable
const Fft2Plan* fft2_get_plan(int64_t M);
// smallest plannable length >= n (2,3,5,7-smooth)
int64_t fft2_next_plannable(int64_t n);

}  // namespace egr
