// umma_probe.cu — standalone check of the tcgen05 (UMMA) operand layouts the GQA decode kernel relies on.
//
// The GQA kernel decompresses a 64-token block straight into an MN-major, un-swizzled UMMA operand tile with a
// padded core-matrix pitch (so that the decode's STS.32 are bank-conflict free) and contracts it with tcgen05.mma.
// This probe pins, on the real hardware, every descriptor assumption that design makes:
//   variant 0: scores   D[64 tokens x 8]   = A(MN-major, M=64,  K=128 channels) * B(K-major, N=8)
//   variant 1: same, A's LBO/SBO swapped (must FAIL if variant 0 is right)
//   variant 2: output   D[128 chan  x 16]  = A(MN-major, M=128, K=64 tokens)    * B(K-major, N=16, 16 real rows)
//   variant 3: same, B's second 8-row group aliased onto the first (SBO = 0)
// and reports how D's rows map to TMEM lanes.  Integers are used so that fp32 results are exact.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                   \
        }                                                                              \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

struct Cfg {
    int M, N, K;           // MMA tile: D[M x N] += A[M x K] * B[N x K]^T, K in steps of 16
    int a_sbo, a_lbo;      // bytes: A core-matrix strides as WE lay the data out: mn-group stride, k-group stride
    int b_sbo, b_lbo;      // bytes: B (K-major): n-group stride, k-group stride (data layout)
    int desc_a_sbo, desc_a_lbo, desc_b_sbo, desc_b_lbo;  // what goes into the descriptors
    int b_rows;            // rows of B actually stored (8 when the second group is aliased)
};

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr >> 4) & 0x3fff);
    d |= static_cast<uint64_t>((lbo >> 4) & 0x3fff) << 16;
    d |= static_cast<uint64_t>((sbo >> 4) & 0x3fff) << 32;
    d |= static_cast<uint64_t>(1) << 46;  // version = 1 (Blackwell)
    return d;                              // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

__global__ void __launch_bounds__(128, 1) probe(Cfg c, const __half* A, const __half* B, float* D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* sa = smem;
    uint8_t* sb = smem + 48 * 1024;
    for (int i = tid; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x7e007e00u;  // NaN poison
    __syncthreads();
    // A[m][k] (row-major in global) -> MN-major core matrices: (m&7)*2 + (k&7)*16 + (m>>3)*a_sbo + (k>>3)*a_lbo
    for (int i = tid; i < c.M * c.K; i += 128) {
        const int m = i / c.K, k = i % c.K;
        *reinterpret_cast<__half*>(sa + (m & 7) * 2 + (k & 7) * 16 + (m >> 3) * c.a_sbo + (k >> 3) * c.a_lbo) = A[i];
    }
    // B[n][k] -> K-major core matrices: (k&7)*2 + (n&7)*16 + (n>>3)*b_sbo + (k>>3)*b_lbo
    for (int i = tid; i < c.b_rows * c.K; i += 128) {
        const int n = i / c.K, k = i % c.K;
        *reinterpret_cast<__half*>(sb + (k & 7) * 2 + (n & 7) * 16 + (n >> 3) * c.b_sbo + (k >> 3) * c.b_lbo) = B[i];
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> async proxy (UMMA)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        // instruction descriptor: D fp32, A/B fp16, A MN-major, B K-major
        const uint32_t idesc = (1u << 4) | (1u << 15) | (static_cast<uint32_t>(c.N >> 3) << 17) | (static_cast<uint32_t>(c.M >> 4) << 24);
        for (int ks = 0; ks < c.K / 16; ++ks) {
            const uint64_t da = make_desc(smem_u32(sa) + 2 * ks * c.a_lbo, c.desc_a_lbo, c.desc_a_sbo);
            const uint64_t db = make_desc(smem_u32(sb) + 2 * ks * c.b_lbo, c.desc_b_lbo, c.desc_b_sbo);
            const uint32_t acc = ks > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {   // wait for the MMAs
        uint32_t done = 0;
        for (int spin = 0; !done && spin < (1 << 22); ++spin) {  // bounded: a faulted MMA must not hang the box
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[16];
    const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * 16 + j] = __uint_as_float(r[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    Cfg c;
    memset(&c, 0, sizeof(c));
    if (variant <= 1) {
        c.M = 64; c.N = 8; c.K = 128;
        c.a_sbo = 144; c.a_lbo = 8 * 144;
        c.b_sbo = 128; c.b_lbo = 128; c.b_rows = 8;
    } else {
        c.M = 128; c.N = 16; c.K = 64;
        c.a_sbo = 144; c.a_lbo = 16 * 144;
        c.b_rows = variant == 3 ? 8 : 16;
        c.b_sbo = variant == 3 ? 0 : 128;
        c.b_lbo = variant == 3 ? 128 : 256;
    }
    c.desc_a_sbo = c.a_sbo; c.desc_a_lbo = c.a_lbo; c.desc_b_sbo = c.b_sbo; c.desc_b_lbo = c.b_lbo;
    if (variant == 1) { c.desc_a_sbo = c.a_lbo; c.desc_a_lbo = c.a_sbo; }
    const int M = c.M, N = c.N, K = c.K;
    __half* hA = (__half*)malloc(M * K * 2);
    __half* hB = (__half*)malloc(16 * K * 2);
    srand(7);
    for (int i = 0; i < M * K; ++i) hA[i] = __float2half((float)(rand() % 9 - 4));
    for (int i = 0; i < 16 * K; ++i) hB[i] = __float2half((float)(rand() % 7 - 3));
    float* ref = (float*)calloc(M * 16, 4);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            const int nb = (variant == 3) ? (n & 7) : n;  // aliased second group repeats rows 0..7
            float s = 0;
            for (int k = 0; k < K; ++k) s += __half2float(hA[m * K + k]) * __half2float(hB[nb * K + k]);
            ref[m * 16 + n] = s;
        }
    __half *dA, *dB;
    float* dD;
    CK(cudaMalloc(&dA, M * K * 2));
    CK(cudaMalloc(&dB, 16 * K * 2));
    CK(cudaMalloc(&dD, 128 * 16 * 4));
    CK(cudaMemcpy(dA, hA, M * K * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB, 16 * K * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, 128 * 16 * 4));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    probe<<<1, 128, 64 * 1024>>>(c, dA, dB, dD);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    float* hD = (float*)malloc(128 * 16 * 4);
    CK(cudaMemcpy(hD, dD, 128 * 16 * 4, cudaMemcpyDeviceToHost));
    // hypotheses for row -> lane
    int bad_identity = 0, bad_m64 = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            const float e = ref[m * 16 + n];
            if (hD[m * 16 + n] != e) ++bad_identity;
            const int l64 = (m & 15) + 32 * (m >> 4);
            if (M == 64 && hD[l64 * 16 + n] != e) ++bad_m64;
        }
    printf("variant %d: M=%d N=%d K=%d  mismatches: lane=row %d / %d", variant, M, N, K, bad_identity, M * N);
    if (M == 64) printf(", lane=(row%%16)+32*(row/16) %d / %d", bad_m64, M * N);
    printf("  -> %s\n", (bad_identity == 0 || (M == 64 && bad_m64 == 0)) ? "OK" : "MISMATCH");
    if (bad_identity && (M != 64 || bad_m64)) {
        printf("first rows of D (lane: cols 0..7) vs ref row 0: ");
        for (int n = 0; n < 8; ++n) printf("%g ", ref[n]);
        printf("\n");
        for (int l = 0; l < 4; ++l) {
            printf("lane %d: ", l);
            for (int n = 0; n < 8; ++n) printf("%g ", hD[l * 16 + n]);
            printf("\n");
        }
    }
    return 0;
}
