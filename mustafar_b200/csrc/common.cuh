// common.cuh — shared device/host helpers for libmustafar_b200 (sm_100a only).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mustafar_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libmustafar_b200 is written for sm_100a only"
#endif

namespace mfb {

constexpr int kHeadDim = 128;
constexpr int kTile = 64;           // elements per bitmap tile
constexpr int kBlockTokens = 64;    // tokens per compressed block (= 128 tiles of K or of V)
constexpr int kTilesPerBlock = 128;

// ---- host-side error plumbing -------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define MFB_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            mfb::set_error(__VA_ARGS__);  \
            return MFB200_EINVAL;         \
        }                                 \
    } while (0)

#define MFB_CUDA(expr)                                              \
    do {                                                            \
        cudaError_t _e = (expr);                                    \
        if (_e != cudaSuccess) return mfb::cuda_fail(_e, #expr);    \
    } while (0)

// Per-device state.  cudaFuncSetAttribute and the SM count are properties of a DEVICE, not of the process: a host that
// drives several GPUs from one process (e.g. HF device_map sharding, pred_long_bench.py:165) must configure every
// kernel on every device it launches on.
constexpr int kMaxDevices = 64;
int current_device_sm_count(int* out);  // cached per device ordinal
// Raises a kernel's dynamic shared-memory limit on the CURRENT device if this device has not seen `bytes` yet.
// `cache` is one zero-initialised array per kernel instantiation; races only repeat an idempotent call.
template <class Kernel>
inline int ensure_dynamic_smem(Kernel kernel, size_t (&cache)[kMaxDevices], size_t bytes) {
    int dev = 0;
    MFB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices || cache[dev] < bytes) {
        MFB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
        if (dev >= 0 && dev < kMaxDevices) cache[dev] = bytes;
    }
    return MFB200_OK;
}

inline int launch_status(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, what);
    return MFB200_OK;
}

// ---- device helpers -------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    // volatile + memory clobber: the load must stay BEFORE the mbarrier arrive that hands the ring slot
    // back to the producer.  (A plain asm is a pure function of `addr` to the compiler, which may sink it
    // — together with the FMA chain it feeds — below the arrive; the slot is then refilled by the next
    // bulk copy before the load executes.  Seen as sporadic NaNs in one query head of the G=4 build.)
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float h2f_bits(uint32_t bits16) {
    return __half2float(__ushort_as_half(static_cast<unsigned short>(bits16)));
}

// streaming 16-byte global load (read once: keep it out of L1)
__device__ __forceinline__ uint4 ldg_stream_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream_v2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

// the same without .nc: for memory the kernel itself writes later (the read-only path requires the data to stay
// unmodified for the whole kernel)
__device__ __forceinline__ uint2 ld_coherent_v2(const void* p) {
    uint2 r;
    asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
    return r;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, no tensor map needed) ------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// Required before the memory of a used mbarrier is initialised again (re-initialising a valid mbarrier is
// undefined behaviour; it showed up as a hang when one CTA processed two work segments).
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"  // suspend-time hint: sleep, don't spin
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// Same, for warps that are off the critical path (producer, softmax): back off with nanosleep between polls so
// that their polling does not take issue slots from the tile warps.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(256);
    }
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#endif  // __CUDACC__

}  // namespace mfb
