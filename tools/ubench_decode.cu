// ubench_decode.cu — stand-alone throughput of the tile-decode inner loops (no TMA, no barriers, no softmax): the same
// device functions as the kernel, fed from a static shared-memory block, 8 tile warps per CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I mustafar_b200/csrc -o tools/ubench_decode tools/ubench_decode.cu
#include <stdio.h>
#include "gqa_mma.cuh"

using namespace mfb;
namespace mfb { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*) { return -1; } }

constexpr int kIters = 2000;

// G = 4: HMMA block-diagonal path; G = 1: FHFMA path (decode_pair + 2 predicated FMAs)
template <int G, int DENS_PCT>
__global__ void __launch_bounds__(256, 3) loop_kernel(float* out, int iters) {
    extern __shared__ __align__(128) uint8_t smem[];
    // layout: [8 warps][32 bitmaps 8 B] | [8 warps][64 recs 8 B] | operand blocks 2 KB | nz 8 warps x 4 KB
    uint64_t* bmp = reinterpret_cast<uint64_t*>(smem);
    uint2* recs = reinterpret_cast<uint2*>(smem + 2048);
    uint8_t* oper = smem + 2048 + 4096;
    uint8_t* nz = oper + 2048;
    const int warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();
    // pseudo-random bitmaps of the requested density
    uint32_t st = 1234567u + threadIdx.x * 7919u;
    uint64_t bm = 0;
    for (int b = 0; b < 64; ++b) {
        st = st * 1664525u + 1013904223u;
        if ((st >> 8) % 100 < DENS_PCT) bm |= 1ull << b;
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
    float res = 0.f;
    if constexpr (G >= 4) {
        const GqaLane<G> gl = make_gqa_lane<G>();
        uint32_t op[G / 2];
        for (int m = 0; m < G / 2; ++m) op[m] = smem_u32(oper) + 8u * (gl.live_m == (uint32_t)m ? gl.g_live : (uint32_t)G);
        float acc[G / 2][4] = {};
        for (int it = 0; it < iters; ++it) tiles32_mma<G, true>(my_rec, lc, nullptr, op, acc);
        for (int m = 0; m < G / 2; ++m) res += acc[m][0] + acc[m][1] + acc[m][2] + acc[m][3];
    } else {
        float a0 = 0.f, a1 = 0.f;
        const uint32_t* w32 = reinterpret_cast<const uint32_t*>(oper);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const DecodedPair d = decode_pair<true>(my_rec + 2 * j, lc, nullptr);
                const uint16_t w = static_cast<uint16_t>(w32[j >> 1] >> (16 * (j & 1)));
                if (d.b0) a0 = fhfma(d.x, w, a0);
                if (d.b1) a1 = fhfma(d.y, w, a1);
            }
        }
        res = a0 + a1;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = res;
}

template <int G, int D>
void run(const char* name, int ctas_per_sm = 2) {
    int sms, clk;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out;
    const int threads = 256, blocks = sms * ctas_per_sm;
    const int smem = 2048 + 4096 + 2048 + 8 * 4096;
    cudaMalloc(&out, sizeof(float) * threads * blocks);
    cudaFuncSetAttribute(loop_kernel<G, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    loop_kernel<G, D><<<blocks, threads, smem>>>(out, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    loop_kernel<G, D><<<blocks, threads, smem>>>(out, kIters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tiles_per_sm = 1.0 * ctas_per_sm * 8 * 32 * kIters;
    printf("%-34s %8.3f ms  %5.2f clk per tile per SM (%d tile warps/SM) [%s]\n", name, ms, ms * 1e-3 * clk * 1e3 / tiles_per_sm,
           8 * ctas_per_sm, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    run<4, 30>("G=4 HMMA block-diag, 30 % dense");
    run<4, 50>("G=4 HMMA block-diag, 50 % dense");
    run<1, 30>("G=1 FHFMA, 30 % dense");
    run<1, 50>("G=1 FHFMA, 50 % dense");
    for (int c = 1; c <= 4; ++c) {
        run<4, 50>("G=4 HMMA block-diag, 50 % dense", c);
        run<1, 50>("G=1 FHFMA, 50 % dense", c);
    }
    return 0;
}
