// Error plumbing and version of the grf_b200 C ABI (include/grf_b200.h).
#include <stdarg.h>

#include "grf_common.cuh"

namespace grf {

static thread_local char g_error[512] = "";

char *error_buffer() { return g_error; }

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

int check_cuda(cudaError_t err, const char *what) {
    if (err == cudaSuccess) return GRF_OK;
    return fail(GRF_ERR_CUDA, "%s: %s", what, cudaGetErrorString(err));
}

}  // namespace grf

extern "C" int grf_abi_version(void) { return GRF_B200_ABI_VERSION; }
extern "C" const char *grf_last_error(void) { return grf::error_buffer(); }
