// temporary: path-B entry points land in fft.cu / fatllama.cu
#include "common.cuh"
using namespace egr;
extern "C" int egr_fft_plan_create(int64_t, int, egr_fft_plan**) { return fail(EGR_ERR_UNSUPPORTED, "fft: not built yet"); }
extern "C" size_t egr_fft_plan_workspace_bytes(const egr_fft_plan*) { return 0; }
extern "C" int egr_fft_plan_passes(const egr_fft_plan*) { return 0; }
extern "C" int egr_fft_exec(egr_fft_plan*, float*, float*, int, int, void*) { return fail(EGR_ERR_UNSUPPORTED, "fft: not built yet"); }
extern "C" void egr_fft_plan_destroy(egr_fft_plan*) {}
extern "C" size_t egr_fatllama_workspace_bytes(int, int64_t, int) { return 0; }
extern "C" int egr_fatllama_run(const float*, float*, int, int64_t, int, int, float, uint32_t, void*, size_t, void*) { return fail(EGR_ERR_UNSUPPORTED, "fatllama: not built yet"); }
