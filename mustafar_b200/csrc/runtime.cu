// runtime.cu — ABI version and thread-local error reporting for libmustafar_b200.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace mfb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error in %s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return MFB200_ECUDA;
}

int current_device_sm_count(int* out) {
    static int cached[kMaxDevices] = {0};
    int dev = 0;
    MFB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices || cached[dev] == 0) {
        int n = 0;
        MFB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        if (dev >= 0 && dev < kMaxDevices) cached[dev] = n;
        *out = n;
    } else {
        *out = cached[dev];
    }
    return MFB200_OK;
}

}  // namespace mfb

extern "C" int mfb200_abi_version(void) { return MFB200_ABI_VERSION; }
extern "C" const char* mfb200_last_error(void) { return mfb::g_err; }
