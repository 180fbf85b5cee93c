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

// ---- peer-to-peer plumbing (head-sharded decode, include/mustafar_b200.h) -----------------------------------
namespace mfb {
__global__ void peer_wait_kernel(const uint32_t* flags, int n, uint32_t epoch, int32_t* timed_out) {
    // Launched with programmatic stream serialization: it may start polling while the attention launch in front of it is
    // still finishing, and it lets the NEXT launch (the following layer's attention, which waits for this kernel's
    // completion before it touches q / window / workspace) run its prologue and first fetches meanwhile.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (timed_out && *reinterpret_cast<volatile int32_t*>(timed_out)) return;  // an earlier wait gave up: fail fast from here on
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        uint32_t spins = 0;
        while (true) {
            uint32_t v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
            if (static_cast<int32_t>(v - epoch) >= 0) break;
            if (++spins > (1u << 22)) {  // ~2 s of 500 ns naps: a peer is gone, do not hang the stream
                if (timed_out) *timed_out = 1;
                break;
            }
            __nanosleep(500);
        }
    }
}
}  // namespace mfb

extern "C" int mfb200_peer_wait(const uint32_t* flags, int n, uint32_t epoch, int32_t* timed_out, mfb200_stream_t stream) {
    MFB_REQUIRE(flags != nullptr && n > 0, "peer_wait: null flags / empty range");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(n < 256 ? ((n + 31) / 32) * 32 : 256);
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MFB_CUDA(cudaLaunchKernelEx(&cfg, mfb::peer_wait_kernel, flags, n, epoch, timed_out));
    return mfb::launch_status("peer_wait_kernel");
}
extern "C" int mfb200_peer_alloc(size_t bytes, void** ptr) {
    MFB_REQUIRE(ptr != nullptr && bytes > 0, "peer_alloc: null pointer / zero size");
    MFB_CUDA(cudaMalloc(ptr, bytes));
    MFB_CUDA(cudaMemset(*ptr, 0, bytes));
    return MFB200_OK;
}
extern "C" int mfb200_peer_free(void* ptr) {
    MFB_CUDA(cudaFree(ptr));
    return MFB200_OK;
}
extern "C" int mfb200_ipc_export(void* ptr, unsigned char handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    MFB_REQUIRE(ptr != nullptr && handle != nullptr, "ipc_export: null pointer");
    cudaIpcMemHandle_t h;
    MFB_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle, &h, 64);
    return MFB200_OK;
}
extern "C" int mfb200_ipc_open(const unsigned char handle[64], void** ptr) {
    MFB_REQUIRE(ptr != nullptr && handle != nullptr, "ipc_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    MFB_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return MFB200_OK;
}
extern "C" int mfb200_ipc_close(void* ptr) {
    MFB_CUDA(cudaIpcCloseMemHandle(ptr));
    return MFB200_OK;
}

extern "C" int mfb200_abi_version(void) { return MFB200_ABI_VERSION; }
extern "C" const char* mfb200_last_error(void) { return mfb::g_err; }
