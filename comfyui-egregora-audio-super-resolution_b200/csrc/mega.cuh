// mega.cuh — host interface of the persistent UNet kernel (mega.cu): a run of consecutive plan ops flagged EGR_FLAG_MEGA
// is executed by ONE cooperative launch that walks them with grid barriers in between.
#pragma once
#include "ops.cuh"

namespace egr {

struct MegaRun;
bool mega_supports(const egr_op& op);
// Build the device-side op table of ops [first, last) (GEMM ops must already be prepared: tc[i] != nullptr).
int  mega_build(const Spaces& sp, const egr_op* ops, TcPrepared* const* tc, int first, int last, MegaRun** out);
int  mega_launch(const MegaRun* r, cudaStream_t st);
void mega_describe(const MegaRun* r, int* out8);   // first, last, ops, gemm ops, grid barriers, smem bytes, grid
int  mega_trace(const MegaRun* r, unsigned long long* h_stamps, int* h_codes, int max_ops);   // debug (EGR_MEGA_TRACE=1)
int  mega_aborted(const MegaRun* r);   // debug: 1 when a grid barrier's watchdog abandoned a launch (synchronises)
void mega_free(MegaRun* r);

}  // namespace egr
