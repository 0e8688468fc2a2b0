// common.cu -- error reporting for the C ABI (include/dfb.h)
#include "common.h"

namespace dfb {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return DFB_OK;
    set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
    return DFB_ERR_CUDA;
}
}  // namespace dfb

extern "C" int dfb_version(void) { return 100; }
extern "C" const char* dfb_last_error(void) { return dfb::g_err; }
