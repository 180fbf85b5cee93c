// spmv_compat.cu — the two batched SpMV operators with the reference's own semantics.
//
//   mfb200_key_formulation   == Key_SplitK_API   (kernel/csrc/SpMM_API.cu:86-139  -> Key_Kernel,   SpMM_Kernel.cuh:156-419)
//   mfb200_value_formulation == Value_SplitK_API (kernel/csrc/SpMM_API.cu:193-254 -> Value_Kernel, SpMM_Kernel.cuh:421-676)
//
// Same buffers, same indexing rules (SpMM_Kernel.cuh:174-185: head = batch / groups, NZ base in
// uint4 units, idx has tiles+1 entries per head), same [Batch, 8, M] fp16 output with fp32
// accumulation.  They exist so that models/llama_mustafar_kernel.py:274 and :314 run unmodified; the
// fused kernel in decode_attn.cu is the fast path.  Differences on purpose: no dense tile is rebuilt
// in shared memory, no read past the end of NZ (SpMM_Kernel.cuh:71), the value op is split along the
// sequence and merged in-kernel instead of one CTA walking the whole sequence, rows 1..7 of B (the
// F.pad zeros) are detected per block and skipped, and launch errors are returned.
#include "sparse_tile.cuh"

namespace mfb {

constexpr int kSpmvThreads = 128;
constexpr int kSpmvWarps = 4;

struct SpmvSmem {
    __align__(16) uint64_t bmp[kTilesPerBlock];            // first: the decode may read the 2 bytes in front of nz
    __align__(16) uint8_t nz[kTilesPerBlock * kTile * 2];  // worst case 16 KB
    uint2 rec[kSpmvWarps][64];
    float vec[8 * kHeadDim];            // key: B rows as [c][n]; value: P block as [t][n] (64*8)
    float part[kSpmvWarps][8][64];      // key: per-warp partial scores; value: final cross-warp reduce
    uint32_t seg[5];
};

// cooperative load of one block (bitmaps + nonzeros + 32-tile segment offsets) into shared memory
__device__ __forceinline__ void load_block_sync(SpmvSmem& s, const uint64_t* bmp_h, const uint32_t* idx_h,
                                                const uint8_t* nz_h, int blk) {
    const int tid = threadIdx.x;
    s.bmp[tid] = bmp_h[static_cast<int64_t>(blk) * 128 + tid];
    if (tid < 5) s.seg[tid] = idx_h[static_cast<int64_t>(blk) * 128 + tid * 32];
    const uint32_t off0 = idx_h[static_cast<int64_t>(blk) * 128];
    const uint32_t off1 = idx_h[static_cast<int64_t>(blk) * 128 + 128];
    const uint32_t n16 = (off1 - off0) >> 2;  // uint4 count
    const uint4* src = reinterpret_cast<const uint4*>(nz_h + static_cast<uint64_t>(off0) * 4u);
    uint4* dst = reinterpret_cast<uint4*>(s.nz);
    for (uint32_t i = tid; i < n16; i += kSpmvThreads) dst[i] = ldg_stream_v4(src + i);
}

// C[bq, n, 64*blk + t] for one block.  grid = (M/64, Batch).
__global__ void __launch_bounds__(kSpmvThreads)
key_formulation_kernel(const uint64_t* __restrict__ bmp, const uint8_t* __restrict__ nz, const uint32_t* __restrict__ idx,
                       const uint32_t* __restrict__ nz_offset, const __half* __restrict__ B, __half* __restrict__ C,
                       int M, int groups) {
    __shared__ SpmvSmem s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t lane = lane_id();
    const int bq = blockIdx.y, blk = blockIdx.x;
    const int h = bq / groups;
    const int64_t tiles = static_cast<int64_t>(M) * 2;
    load_block_sync(s, bmp + h * tiles, idx + h * (tiles + 1), nz + static_cast<uint64_t>(nz_offset[h]) * 16u, blk);
    // B rows: [bq][n][c] -> vec[c*8 + n]; note whether any of the pad rows 1..7 is nonzero
    int pad_nonzero = 0;
    {
        const __half* bp = B + static_cast<int64_t>(bq) * 8 * kHeadDim;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const __half v = bp[n * kHeadDim + tid];
            s.vec[tid * 8 + n] = __half2float(v);
            if (n > 0 && (__half_as_ushort(v) & 0x7fffu) != 0) pad_nonzero = 1;
        }
    }
    const int general = __syncthreads_or(pad_nonzero);
    const LaneConst lc = make_lane_const();
    build_records(s.bmp + warp * 32, smem_u32(s.nz) + (s.seg[warp] - s.seg[0]) * 4u, s.rec[warp]);
    __syncwarp();
    const uint2* rec_base = s.rec[warp] + lc.half;
    float sc[8][2];
#pragma unroll
    for (int n = 0; n < 8; ++n) sc[n][0] = sc[n][1] = 0.f;
    if (!general) {
#pragma unroll 4
        for (int j = 0; j < 32; ++j) {
            float v0, v1;
            decode_pair<true>(rec_base + 2 * j, lc, nullptr, v0, v1);
            const float q = s.vec[(warp * 32 + j) * 8];
            sc[0][0] = fmaf(q, v0, sc[0][0]);
            sc[0][1] = fmaf(q, v1, sc[0][1]);
        }
    } else {
#pragma unroll 2
        for (int j = 0; j < 32; ++j) {
            float v0, v1;
            decode_pair<true>(rec_base + 2 * j, lc, nullptr, v0, v1);
            const float* q = s.vec + (warp * 32 + j) * 8;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                sc[n][0] = fmaf(q[n], v0, sc[n][0]);
                sc[n][1] = fmaf(q[n], v1, sc[n][1]);
            }
        }
    }
#pragma unroll
    for (int n = 0; n < 8; ++n) *reinterpret_cast<float2*>(&s.part[warp][n][2 * lane]) = make_float2(sc[n][0], sc[n][1]);
    __syncthreads();
    // 8 rows x 64 tokens = 512 outputs, 4 per thread; row-contiguous fp16 stores
    __half* cp = C + static_cast<int64_t>(bq) * 8 * M + static_cast<int64_t>(blk) * 64;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int o = r * kSpmvThreads + tid, n = o >> 6, t = o & 63;
        const float v = s.part[0][n][t] + s.part[1][n][t] + s.part[2][n][t] + s.part[3][n][t];
        cp[static_cast<int64_t>(n) * M + t] = __float2half_rn(v);
    }
}

// C[bq, n, c] = sum_t V[t, c] * B[bq, n, t].  grid = (n_split, Batch); partials merged by the last CTA.
__global__ void __launch_bounds__(kSpmvThreads)
value_formulation_kernel(const uint64_t* __restrict__ bmp, const uint8_t* __restrict__ nz, const uint32_t* __restrict__ idx,
                         const uint32_t* __restrict__ nz_offset, const __half* __restrict__ B, __half* __restrict__ C,
                         int L, int groups, int n_split, int* __restrict__ counters, float* __restrict__ parts) {
    __shared__ SpmvSmem s;
    __shared__ int s_last;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t lane = lane_id();
    const int bq = blockIdx.y, split = blockIdx.x;
    const int h = bq / groups;
    const int nblk = L / 64;
    const int blk0 = static_cast<int>(static_cast<int64_t>(split) * nblk / n_split);
    const int blk1 = static_cast<int>(static_cast<int64_t>(split + 1) * nblk / n_split);
    const int64_t tiles = static_cast<int64_t>(L) * 2;
    const uint64_t* bmp_h = bmp + h * tiles;
    const uint32_t* idx_h = idx + h * (tiles + 1);
    const uint8_t* nz_h = nz + static_cast<uint64_t>(nz_offset[h]) * 16u;
    const __half* bp = B + static_cast<int64_t>(bq) * 8 * L;
    const LaneConst lc = make_lane_const();
    const uint2* rec_base = s.rec[warp] + lc.half;
    float o[8][2];
#pragma unroll
    for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = 0.f;
    for (int blk = blk0; blk < blk1; ++blk) {
        __syncthreads();  // previous block fully consumed
        load_block_sync(s, bmp_h, idx_h, nz_h, blk);
        int pad_nonzero = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {  // P block [8][64] -> vec[t*8 + n]
            const int e = r * kSpmvThreads + tid, n = e >> 6, t = e & 63;
            const __half v = bp[static_cast<int64_t>(n) * L + static_cast<int64_t>(blk) * 64 + t];
            s.vec[t * 8 + n] = __half2float(v);
            if (n > 0 && (__half_as_ushort(v) & 0x7fffu) != 0) pad_nonzero = 1;
        }
        const int general = __syncthreads_or(pad_nonzero);
        build_records(s.bmp + warp * 32, smem_u32(s.nz) + (s.seg[warp] - s.seg[0]) * 4u, s.rec[warp]);
        __syncwarp();
        const float* pw = s.vec + (32 * (warp & 1)) * 8;  // this warp's 32 tokens
        if (!general) {
#pragma unroll 4
            for (int j = 0; j < 32; ++j) {
                float v0, v1;
                decode_pair<true>(rec_base + 2 * j, lc, nullptr, v0, v1);
                const float pj = pw[j * 8];
                o[0][0] = fmaf(pj, v0, o[0][0]);
                o[0][1] = fmaf(pj, v1, o[0][1]);
            }
        } else {
#pragma unroll 2
            for (int j = 0; j < 32; ++j) {
                float v0, v1;
                decode_pair<true>(rec_base + 2 * j, lc, nullptr, v0, v1);
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const float pj = pw[j * 8 + n];
                    o[n][0] = fmaf(pj, v0, o[n][0]);
                    o[n][1] = fmaf(pj, v1, o[n][1]);
                }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int n = 0; n < 8; ++n) *reinterpret_cast<float2*>(&s.part[warp][n][2 * lane]) = make_float2(o[n][0], o[n][1]);
    __syncthreads();
    // warps (0,1) hold channels 0..63, warps (2,3) channels 64..127
    float* mine = parts + (static_cast<int64_t>(bq) * n_split + split) * 8 * kHeadDim;
    {
        const int hf = tid >> 6, e = tid & 63;
#pragma unroll
        for (int n = 0; n < 8; ++n) mine[n * kHeadDim + tid] = s.part[2 * hf][n][e] + s.part[2 * hf + 1][n][e];
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&counters[bq], 1) == n_split - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float* up = parts + static_cast<int64_t>(bq) * n_split * 8 * kHeadDim;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        float acc = 0.f;
        for (int sp = 0; sp < n_split; ++sp) acc += __ldcg(up + (sp * 8 + n) * kHeadDim + tid);
        C[(static_cast<int64_t>(bq) * 8 + n) * kHeadDim + tid] = __float2half_rn(acc);
    }
    if (tid == 0) counters[bq] = 0;
}

static int value_n_split(int L, int batch) {
    const int nblk = L / 64;
    int sms = 0;
    if (current_device_sm_count(&sms) != MFB200_OK || sms <= 0) sms = 148;  // no device (host-only sizing call): B200
    int ns = (sms * 4 + batch - 1) / batch;
    if (ns > 32) ns = 32;
    if (ns > nblk) ns = nblk;
    if (ns < 1) ns = 1;
    return ns;
}

static int check_common(const char* who, const void* bmp, const void* NZ, const void* idx, const void* off, const void* B,
                        const void* C, int N_Global, int Batch_Size, int groups) {
    MFB_REQUIRE(bmp && NZ && idx && off && B && C, "%s: null pointer", who);
    // the reference silently launches nothing for N_Global != 8 (SpMM_API.cu:119-127); refuse instead
    MFB_REQUIRE(N_Global == 8, "%s: N_Global=%d, only 8 is supported (reference: SpMM_API.cu:121)", who, N_Global);
    MFB_REQUIRE(Batch_Size >= 0 && Batch_Size <= 65535, "%s: Batch_Size=%d out of range", who, Batch_Size);
    MFB_REQUIRE(groups >= 1, "%s: num_key_value_groups must be >= 1", who);
    MFB_REQUIRE((reinterpret_cast<uintptr_t>(NZ) & 15) == 0, "%s: NZ must be 16-byte aligned", who);
    return MFB200_OK;
}

}  // namespace mfb

using namespace mfb;

extern "C" size_t mfb200_value_workspace_bytes(int K_Global, int Batch_Size) {
    if (K_Global <= 0 || Batch_Size <= 0) return 256;
    const size_t cbytes = (static_cast<size_t>(Batch_Size) * 4 + 255) & ~static_cast<size_t>(255);
    return cbytes + static_cast<size_t>(Batch_Size) * value_n_split(K_Global, Batch_Size) * 8 * kHeadDim * 4;
}

extern "C" int mfb200_key_formulation(mfb200_stream_t stream, const void* A, const uint64_t* bmp, const void* NZ,
                                      const uint32_t* idx, const uint32_t* NZ_offset, const void* B, void* C,
                                      int M_Global, int N_Global, int K_Global, void* Reduction_Workspace, int Split_K,
                                      int Batch_Size, int num_key_value_groups) {
    (void)A;
    (void)Reduction_Workspace;
    (void)Split_K;
    int rc = check_common("key_formulation", bmp, NZ, idx, NZ_offset, B, C, N_Global, Batch_Size, num_key_value_groups);
    if (rc) return rc;
    MFB_REQUIRE(K_Global == kHeadDim, "key_formulation: K_Global=%d, only head_dim 128 is supported", K_Global);
    MFB_REQUIRE(M_Global >= 0 && M_Global % 64 == 0, "key_formulation: M_Global=%d must be a multiple of 64", M_Global);
    if (Batch_Size == 0 || M_Global == 0) return MFB200_OK;
    dim3 grid(M_Global / 64, Batch_Size);
    key_formulation_kernel<<<grid, kSpmvThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        bmp, static_cast<const uint8_t*>(NZ), idx, NZ_offset, static_cast<const __half*>(B), static_cast<__half*>(C), M_Global,
        num_key_value_groups);
    return launch_status("key_formulation_kernel");
}

extern "C" int mfb200_value_formulation(mfb200_stream_t stream, const void* A, const uint64_t* bmp, const void* NZ,
                                        const uint32_t* idx, const uint32_t* NZ_offset, const void* B, void* C,
                                        int M_Global, int N_Global, int K_Global, void* workspace, int Split_K,
                                        int Batch_Size, int num_key_value_groups) {
    (void)A;
    (void)Split_K;
    int rc = check_common("value_formulation", bmp, NZ, idx, NZ_offset, B, C, N_Global, Batch_Size, num_key_value_groups);
    if (rc) return rc;
    MFB_REQUIRE(M_Global == kHeadDim, "value_formulation: M_Global=%d, only head_dim 128 is supported", M_Global);
    MFB_REQUIRE(K_Global > 0 && K_Global % 64 == 0, "value_formulation: K_Global=%d must be a positive multiple of 64", K_Global);
    MFB_REQUIRE(workspace != nullptr, "value_formulation: workspace must not be null");
    if (Batch_Size == 0) return MFB200_OK;
    const int ns = value_n_split(K_Global, Batch_Size);
    const size_t cbytes = (static_cast<size_t>(Batch_Size) * 4 + 255) & ~static_cast<size_t>(255);
    dim3 grid(ns, Batch_Size);
    value_formulation_kernel<<<grid, kSpmvThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        bmp, static_cast<const uint8_t*>(NZ), idx, NZ_offset, static_cast<const __half*>(B), static_cast<__half*>(C), K_Global,
        num_key_value_groups, ns, static_cast<int*>(workspace), reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + cbytes));
    return launch_status("value_formulation_kernel");
}
