// ubench_ablate.cu — which instruction of the tile-decode loop sets its throughput?  The G=1 loop with one piece removed
// at a time (results are wrong on purpose; only the timing matters).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I mustafar_b200/csrc -o tools/ubench_ablate tools/ubench_ablate.cu
#include <stdio.h>
#include "sparse_tile.cuh"
using namespace mfb;
namespace mfb { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*) { return -1; } }
constexpr int kIters = 2000;

// MODE bits: 1 = no POPC, 2 = one value load, 4 = record from registers, 8 = no FMA, 16 = no predicates (plain FMAs), 32 = no value loads
template <int MODE>
__global__ void __launch_bounds__(256, 2) loop_kernel(float* out, int iters) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bmp = reinterpret_cast<uint64_t*>(smem);
    uint2* recs = reinterpret_cast<uint2*>(smem + 2048);
    uint8_t* oper = smem + 2048 + 4096;
    uint8_t* nz = oper + 2048;
    const int warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();
    uint32_t st = 1234567u + threadIdx.x * 7919u;
    uint64_t bm = 0;
    for (int b = 0; b < 64; ++b) {
        st = st * 1664525u + 1013904223u;
        if ((st >> 8) % 100 < 50) bm |= 1ull << b;
    }
    bmp[warp * 32 + lane] = bm;
    for (int i = threadIdx.x; i < 2048 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(oper)[i] = 0x3c003c00u;
    for (int i = threadIdx.x; i < 8 * 4096 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(nz)[i] = 0x38003800u;
    __syncthreads();
    const LaneConst lc = make_lane_const();
    uint2* rec = recs + warp * 64;
    build_records(bmp + warp * 32, smem_u32(nz + warp * 4096), rec);
    __syncwarp();
    const uint2* my_rec = rec + lc.half;
    float a0 = 0.f, a1 = 0.f;
    const uint32_t* w32 = reinterpret_cast<const uint32_t*>(oper);
    const uint32_t above1 = lc.above | lc.bit0;
    uint2 rreg = my_rec[0];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            uint2 r = (MODE & 4) ? rreg : my_rec[2 * j];
            if (MODE & 4) rreg.y ^= 2;  // keep it loop-variant
            uint32_t rank = (MODE & 1) ? (r.x & above1 & 31u) : __popc(r.x & above1);
            uint32_t addr1;
            asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(addr1) : "r"(rank), "r"(r.y));
            const bool b0 = (r.x & lc.bit0) != 0, b1 = (r.x & lc.bit1) != 0;
            uint32_t x = (MODE & 32) ? addr1 : lds_u16(addr1 - 2);
            uint32_t y = (MODE & (2 | 32)) ? x ^ 0x100u : lds_u16(addr1);
            const uint16_t w = static_cast<uint16_t>(w32[j >> 1] >> (16 * (j & 1)));
            if (MODE & 8) {
                a0 += __uint_as_float(x << 13);
                a1 += __uint_as_float(y << 13);
            } else if (MODE & 16) {
                a0 = fhfma(static_cast<uint16_t>(x), w, a0);
                a1 = fhfma(static_cast<uint16_t>(y), w, a1);
            } else {
                if (b0) a0 = fhfma(static_cast<uint16_t>(x), w, a0);
                if (b1) a1 = fhfma(static_cast<uint16_t>(y), w, a1);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1;
}

template <int MODE>
void run(const char* name) {
    int sms, clk;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out;
    const int threads = 256, blocks = sms * 2, smem = 2048 + 4096 + 2048 + 8 * 4096;
    cudaMalloc(&out, sizeof(float) * threads * blocks);
    cudaFuncSetAttribute(loop_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    loop_kernel<MODE><<<blocks, threads, smem>>>(out, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    loop_kernel<MODE><<<blocks, threads, smem>>>(out, kIters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-46s %5.2f clk per tile per SM [%s]\n", name, ms * 1e-3 * clk * 1e3 / (2.0 * 8 * 32 * kIters), cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    run<0>("full G=1 loop");
    run<1>("no POPC");
    run<2>("one value load instead of two");
    run<4>("record from registers (no LDS.64)");
    run<8>("no FHFMA (plain FADDs)");
    run<16>("unpredicated FHFMA");
    run<32>("no value loads");
    run<1 | 4>("no POPC, no record load");
    run<2 | 4>("one value load, no record load");
    run<32 | 4>("no loads at all");
    run<32 | 4 | 1>("no loads, no POPC");
    return 0;
}
