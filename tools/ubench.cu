// ubench.cu — instruction-throughput probe for the design of the tile decoder (sm_100a).
// Reports warp-instructions per clock per SM for the handful of ops the decode loop is made of.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 4096

template <int OP>
__global__ void k(uint32_t* out, uint32_t seed) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 8 + i;
    __shared__ uint32_t sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i * 2654435761u;
    __syncthreads();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) a[i] = __popc(a[i]) + seed;                     // POPC (+IADD)
            if (OP == 1) a[i] = __clz(a[i]) + seed;                      // FLO
            if (OP == 2) a[i] = __brev(a[i]) + seed;                     // BREV
            if (OP == 3) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1);   // SHFL
            if (OP == 4) a[i] = sm[(a[i] >> 7) & 2047] ;                 // LDS.32 random
            if (OP == 5) a[i] = (a[i] & seed) + it;                      // LOP+IADD
            if (OP == 6) a[i] = a[i] * seed + it;                        // IMAD
            if (OP == 7) { float f = __uint_as_float(a[i]); f = fmaf(f, 1.0001f, 0.5f); a[i] = __float_as_uint(f); }  // FFMA
            if (OP == 8) a[i] = ((const uint16_t*)sm)[(a[i] >> 7) & 4095] + seed;  // LDS.U16 random
            if (OP == 9) a[i] = __ballot_sync(0xffffffffu, a[i] & 1) + a[i];  // VOTE
            if (OP == 10) a[i] = __reduce_add_sync(0xffffffffu, a[i]);   // REDUX
            if (OP == 11) a[i] = (a[i] & 1) ? a[i] >> 1 : seed;           // SEL-ish
            // pipe-sharing probes: 2 LDS + 1 POPC per iteration (the decode loop's mix), and the same with the POPC replaced by a LOP3
            if (OP == 12) { uint32_t t = sm[(a[i] >> 7) & 2047]; t += ((const uint16_t*)sm)[(a[i] >> 9) & 4095]; a[i] = __popc(a[i] ^ t) + t; }
            if (OP == 13) { uint32_t t = sm[(a[i] >> 7) & 2047]; t += ((const uint16_t*)sm)[(a[i] >> 9) & 4095]; a[i] = ((a[i] ^ t) & seed) + t; }
            if (OP == 14) { uint32_t t = sm[(a[i] >> 7) & 2047]; a[i] = __shfl_xor_sync(0xffffffffu, a[i] + t, 1); }   // 1 LDS + 1 SHFL
            if (OP == 15) { a[i] = __popc(a[i]) + __shfl_xor_sync(0xffffffffu, a[i], 1); }                              // 1 POPC + 1 SHFL
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int OP>
void run(const char* name, int extra_ops) {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    uint32_t* out;
    const int threads = 512, blocks = sms * 4;
    cudaMalloc(&out, sizeof(uint32_t) * threads * blocks);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, 12345);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, 12345);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double warp_instr = (double)blocks * (threads / 32) * ITERS * 8;
    double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-22s %8.3f ms  %6.2f warp-iters/clk/SM at nominal %d MHz (each iter = op + %d helper)\n", name, ms,
           warp_instr / cycles / sms, clk / 1000, extra_ops);
    cudaFree(out);
}

// ---- shared-memory load shapes: how many wavefronts does a broadcast-ish wide load cost? ----
template <int MODE>
__global__ void lds_shape(uint32_t* out, uint32_t seed) {
    __shared__ __align__(16) uint32_t sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 2654435761u + seed;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    uint32_t acc = 0, off = (threadIdx.x >> 5) * 64;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t base = (off + i * 8 + (it & 7) * 4) & 2047;  // word index, multiple of 4
            if (MODE == 0) acc += sm[base];                                                   // LDS.32, 1 address
            if (MODE == 1) acc += sm[base + (lane >> 4)];                                     // LDS.32, 2 addresses
            if (MODE == 2) { uint2 v = *reinterpret_cast<uint2*>(&sm[base]); acc += v.x ^ v.y; }                 // LDS.64, 1 address
            if (MODE == 3) { uint2 v = *reinterpret_cast<uint2*>(&sm[base + 2 * (lane >> 4)]); acc += v.x ^ v.y; }  // LDS.64, 2 addresses
            if (MODE == 4) { uint4 v = *reinterpret_cast<uint4*>(&sm[base]); acc += v.x ^ v.y ^ v.z ^ v.w; }     // LDS.128, 1 address
            if (MODE == 5) { uint4 v = *reinterpret_cast<uint4*>(&sm[base + 4 * (lane >> 4)]); acc += v.x ^ v.y ^ v.z ^ v.w; }  // LDS.128, 2 addr
            if (MODE == 6) acc += reinterpret_cast<uint16_t*>(sm)[2 * base + lane + (lane >> 2)];  // LDS.U16, ~40 consecutive halves
            if (MODE == 7) acc += sm[base + lane];                                                // LDS.32, 32 consecutive words
            // B-fragment shapes of the GQA HMMA path: 16 live lanes read one of 4 rows (g = gid % 4), 16 dead lanes a zero block
            if (MODE >= 8 && MODE <= 13) {
                const uint32_t gid = lane >> 2, tig = lane & 3;
                const bool live = (tig & 1) == (gid >> 2);
                uint32_t w;  // word index
                if (MODE == 8 || MODE == 10 || MODE == 12) w = live ? base + 36 * (gid & 3) : (base + 1024 + 6) & 2047;   // rows 144 B apart, zeros elsewhere
                else w = (base & ~15u) + (live ? 2 * (gid & 3) : 8);                                            // 4 rows + zeros inside one 64-byte block
                if (MODE <= 9) { uint2 v = *reinterpret_cast<uint2*>(&sm[w & ~1u]); acc += v.x ^ v.y; }
                else if (MODE <= 11) acc += sm[w];
                else if (tig < 2 ? live : false) { uint2 v = *reinterpret_cast<uint2*>(&sm[w & ~1u]); acc += v.x ^ v.y; }  // predicated: 8 lanes
            }
            if (MODE == 14) { uint4 v = make_uint4(acc, lane, it, i); *reinterpret_cast<uint4*>(&sm[(off + 4 * lane) & 2047]) = v; }       // STS.128 contiguous
            if (MODE == 15) { uint2 v = make_uint2(acc, lane); *reinterpret_cast<uint2*>(&sm[(off + 4 * lane) & 2047]) = v;
                              *reinterpret_cast<uint2*>(&sm[(off + 4 * lane + 2) & 2047]) = v; }                                           // 2 x STS.64, 16-byte lane stride
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run_lds(const char* name) {
    int sms, clk;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    uint32_t* out;
    const int threads = 512, blocks = sms * 4;
    cudaMalloc(&out, sizeof(uint32_t) * threads * blocks);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    lds_shape<MODE><<<blocks, threads>>>(out, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    lds_shape<MODE><<<blocks, threads>>>(out, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double n = (double)blocks * (threads / 32) * ITERS * 8;
    printf("%-28s %8.3f ms  %6.2f clk per warp-load per SM\n", name, ms, ms * 1e-3 * clk * 1e3 * sms / n);
    cudaFree(out);
}

int main() {
    run_lds<0>("LDS.32 1 addr");
    run_lds<1>("LDS.32 2 addr");
    run_lds<2>("LDS.64 1 addr");
    run_lds<3>("LDS.64 2 addr (half-warps)");
    run_lds<4>("LDS.128 1 addr");
    run_lds<5>("LDS.128 2 addr (half-warps)");
    run_lds<6>("LDS.U16 ~40 consecutive");
    run_lds<7>("LDS.32 32 consecutive");
    run_lds<8>("LDS.64 B-frag rows 144B apart");
    run_lds<9>("LDS.64 B-frag in one 64B block");
    run_lds<10>("LDS.32 B-frag rows 144B apart");
    run_lds<11>("LDS.32 B-frag in one 64B block");
    run_lds<12>("@P LDS.64 8 lanes rows apart");
    run_lds<13>("@P LDS.64 8 lanes one block");
    run_lds<14>("STS.128 contiguous");
    run_lds<15>("2 x STS.64 16B lane stride");
    run<0>("popc+iadd", 1);
    run<1>("flo(clz)+iadd", 1);
    run<2>("brev+iadd", 1);
    run<3>("shfl", 0);
    run<4>("lds.32 random", 2);
    run<5>("lop+iadd", 1);
    run<6>("imad", 0);
    run<7>("ffma", 0);
    run<8>("lds.u16 random+iadd", 3);
    run<9>("vote+lop+iadd", 2);
    run<10>("redux.add", 0);
    run<11>("sel-ish", 2);
    run<12>("2 lds + popc (+3 alu)", 3);
    run<13>("2 lds + lop (+3 alu)", 3);
    run<14>("lds + shfl (+2 alu)", 2);
    run<15>("popc + shfl (+1 alu)", 1);
    return 0;
}
