// decode_attn.cu — fused sparse-KV decode attention for sm_100a.
//
// One launch replaces the whole reference decode step of one layer
// (models/llama_mustafar_kernel.py:268-320): F.pad(q) -> torch.cat(NZ) -> Key_Kernel -> window matmul ->
// cat -> /sqrt(d) -> +mask -> fp32 softmax -> F.pad(P) -> torch.cat(NZ) -> Value_Kernel -> window matmul ->
// add, and the (never launched) SplitK_Reduction (kernel/csrc/Reduction_Kernel.cuh:26-48).
//
// Work decomposition: unit = (sequence, KV head); the compressed length is cut into `n_csplit`
// contiguous ranges of 64-token blocks, the dense window into ranges of <= 256 tokens; grid =
// (n_split, units).  Each CTA keeps flash-decoding state (m, l, o) in registers, writes one fp32
// partial and the last CTA of a unit (atomic ticket) merges the partials -> single launch, no
// second kernel, graph-capturable.  All G query heads of a KV head are served by the same CTA, so the
// compressed bytes of a KV head cross HBM once (the reference re-reads them G times,
// kernel/csrc/SpMM_Kernel.cuh:175).
//
// Data path of a compressed split: per 64-token block two pipeline items (K then V), each =
// 1 KB of bitmaps + the block's contiguous nonzero range.  A single elected thread moves them with
// cp.async.bulk (TMA engine) into a 4-slot shared-memory ring, completion through mbarrier
// expect-tx.  Consumers never build a dense tile: see sparse_tile.cuh.
#include "sparse_tile.cuh"

namespace mfb {

constexpr int kAttnThreads = 128;
constexpr int kAttnWarps = 4;
constexpr int kSlots = 4;
constexpr int kMaxBlocksPerSplit = 128;
constexpr int kWinTokensPerSplit = 256;
constexpr int kPartStride = 132;  // floats per (split, head): o[128], m, l, pad
constexpr float kLog2e = 1.4426950408889634f;

struct DecodeArgs {
    mfb200_decode_params p;
    int n_csplit;
    int n_wsplit;
    int slot_nz_bytes;  // capacity of a slot's nonzero area (multiple of 16)
};

struct SmemMap {
    uint32_t slots, bars, rec, qs, spart, ps, segk, segv, total;
};

__host__ __device__ inline SmemMap smem_map(int G, int slot_nz_bytes) {
    SmemMap m;
    uint32_t o = 0;
    m.slots = o;
    o += kSlots * (1024 + slot_nz_bytes);
    m.bars = o;
    o += 64;
    m.rec = o;
    o += kAttnWarps * 64 * 8;  // uint2[64] per warp
    m.qs = o;
    o += kHeadDim * G * 4;
    m.spart = o;  // float [2][warps][G][64]; also reused for the final cross-warp reduction
    o += 2 * kAttnWarps * G * 64 * 4;
    m.ps = o;  // float [warps][64][G]
    o += kAttnWarps * 64 * G * 4;
    m.segk = o;
    o += (kMaxBlocksPerSplit * 4 + 4) * 4;
    m.segv = o;
    o += (kMaxBlocksPerSplit * 4 + 4) * 4;
    m.total = o;
    return m;
}

__device__ __forceinline__ float ref_round_score(float dot, float div, bool ref_rounding) {
    if (ref_rounding) {
        // fp32 accumulate -> fp16 store (SpMM_Kernel.cuh:418), then `/ sqrt(d)` evaluated in fp16
        // (llama_mustafar_kernel.py:284: fp16 tensor / python float -> fp32 divide, fp16 round)
        const float s16 = __half2float(__float2half_rn(dot));
        return __half2float(__float2half_rn(s16 / div));
    }
    return dot / div;
}

template <int G>
__device__ __forceinline__ void write_partial_and_merge(const DecodeArgs& a, int unit, int split, const float* ored,
                                                        /* ored: [G][128] in smem, already reduced */
                                                        const float (&m)[G], const float (&l)[G]) {
    const mfb200_decode_params& p = a.p;
    const int tid = threadIdx.x;
    const int n_split = a.n_csplit + a.n_wsplit;
    const int units = p.batch * p.kv_heads;
    int* counters = static_cast<int*>(p.workspace);
    float* parts = reinterpret_cast<float*>(static_cast<uint8_t*>(p.workspace) + ((units * 4 + 255) & ~255));
    float* mine = parts + (static_cast<int64_t>(unit) * n_split + split) * G * kPartStride;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        mine[g * kPartStride + tid] = ored[g * 128 + tid];
        if (tid == 0) {
            mine[g * kPartStride + 128] = m[g];
            mine[g * kPartStride + 129] = l[g];
        }
    }
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int ticket = atomicAdd(&counters[unit], 1);
        s_last = (ticket == n_split - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float* up = parts + static_cast<int64_t>(unit) * n_split * G * kPartStride;
    const int b = unit / p.kv_heads, h = unit % p.kv_heads;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        float mx = -INFINITY;
        for (int s = 0; s < n_split; ++s) mx = fmaxf(mx, __ldcg(up + (s * G + g) * kPartStride + 128));
        float den = 0.f, num = 0.f;
        for (int s = 0; s < n_split; ++s) {
            const float* ps = up + (s * G + g) * kPartStride;
            const float ms = __ldcg(ps + 128);
            const float w = (ms == -INFINITY) ? 0.f : exp2f((ms - mx) * kLog2e);
            den += __ldcg(ps + 129) * w;
            num += __ldcg(ps + tid) * w;
        }
        const int64_t qh = (static_cast<int64_t>(b) * p.kv_heads + h) * G + g;
        static_cast<__half*>(p.out)[qh * kHeadDim + tid] = __float2half_rn(num / den);
    }
    if (tid == 0) counters[unit] = 0;  // ready for the next launch / graph replay
}

// ------------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ void compressed_split(const DecodeArgs& a, uint8_t* smem, int unit, int split) {
    const mfb200_decode_params& p = a.p;
    const SmemMap sm = smem_map(G, a.slot_nz_bytes);
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t lane = lane_id();
    const int nblk_total = p.comp_len / kBlockTokens;
    const int blk0 = static_cast<int>(static_cast<int64_t>(split) * nblk_total / a.n_csplit);
    const int blk1 = static_cast<int>(static_cast<int64_t>(split + 1) * nblk_total / a.n_csplit);
    const int nb = blk1 - blk0;
    const int b = unit / p.kv_heads, h = unit % p.kv_heads;
    const bool ref_round = (p.flags & MFB200_F_REF_SCORE_ROUNDING) != 0;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sm.bars);
    uint2* rec = reinterpret_cast<uint2*>(smem + sm.rec) + warp * 64;
    float* qs = reinterpret_cast<float*>(smem + sm.qs);        // [c][G]
    float* spart = reinterpret_cast<float*>(smem + sm.spart);  // [2][warps][G][64]
    float* ps = reinterpret_cast<float*>(smem + sm.ps) + warp * 64 * G;  // [64][G]
    uint32_t* segk = reinterpret_cast<uint32_t*>(smem + sm.segk);
    uint32_t* segv = reinterpret_cast<uint32_t*>(smem + sm.segv);
    const uint32_t slot_bytes = 1024u + a.slot_nz_bytes;

    // ---- prologue: q -> fp32 smem, 32-tile segment offsets of this split, barriers ---------------
    {
        const __half* q = static_cast<const __half*>(p.q) + (static_cast<int64_t>(b) * p.kv_heads + h) * G * kHeadDim;
#pragma unroll
        for (int g = 0; g < G; ++g) qs[tid * G + g] = __half2float(q[g * kHeadDim + tid]);
        const uint32_t* ki = p.k_idx + static_cast<int64_t>(unit) * p.idx_stride + static_cast<int64_t>(blk0) * 128;
        const uint32_t* vi = p.v_idx + static_cast<int64_t>(unit) * p.idx_stride + static_cast<int64_t>(blk0) * 128;
        for (int i = tid; i <= nb * 4; i += kAttnThreads) {
            segk[i] = __ldg(ki + i * 32);
            segv[i] = __ldg(vi + i * 32);
        }
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < kSlots; ++s) mbar_init(&bars[s], 1);
            fence_mbar_init();
        }
    }
    __syncthreads();

    const uint8_t* k_nz = static_cast<const uint8_t*>(p.k_nz) + p.k_nz_off[unit] * 16;
    const uint8_t* v_nz = static_cast<const uint8_t*>(p.v_nz) + p.v_nz_off[unit] * 16;
    const uint64_t* k_bmp = p.k_bmp + static_cast<int64_t>(unit) * p.bmp_stride + static_cast<int64_t>(blk0) * 128;
    const uint64_t* v_bmp = p.v_bmp + static_cast<int64_t>(unit) * p.bmp_stride + static_cast<int64_t>(blk0) * 128;

    const int n_items = 2 * nb;
    int issued = 0;
    auto issue_until = [&](int limit) {  // thread 0 only
        for (; issued < limit && issued < n_items; ++issued) {
            const int blk = issued >> 1, is_v = issued & 1, slot = issued & (kSlots - 1);
            const uint32_t* seg = is_v ? segv : segk;
            const uint32_t off0 = seg[blk * 4], off1 = seg[blk * 4 + 4];
            const uint32_t bytes = (off1 - off0) * 4u;
            const bool fits = bytes <= static_cast<uint32_t>(a.slot_nz_bytes);
            uint8_t* dst = smem + sm.slots + slot * slot_bytes;
            mbar_expect_tx(&bars[slot], 1024u + ((fits && bytes) ? bytes : 0u));
            bulk_g2s(dst, (is_v ? v_bmp : k_bmp) + blk * 128, 1024u, &bars[slot]);
            if (fits && bytes) bulk_g2s(dst + 1024, (is_v ? v_nz : k_nz) + static_cast<uint64_t>(off0) * 4u, bytes, &bars[slot]);
        }
    };
    if (tid == 0) issue_until(kSlots);

    const LaneConst lc = make_lane_const();
    const uint32_t rec_base = smem_u32(rec) + lc.half * 8u;

    float m_run[G], l_run[G], o_acc[G][2];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        m_run[g] = -INFINITY;
        l_run[g] = 0.f;
        o_acc[g][0] = o_acc[g][1] = 0.f;
    }
    const __half* mask = p.mask ? static_cast<const __half*>(p.mask) + static_cast<int64_t>(b) * p.mask_stride : nullptr;

    for (int n = 0; n < nb; ++n) {
        // =============================== K item ===============================
        {
            const int item = 2 * n, slot = item & (kSlots - 1);
            mbar_wait(&bars[slot], (item / kSlots) & 1);
            const uint8_t* sl = smem + sm.slots + slot * slot_bytes;
            const uint32_t seg0 = segk[n * 4], segw = segk[n * 4 + warp];
            const bool fits = (segk[n * 4 + 4] - seg0) * 4u <= static_cast<uint32_t>(a.slot_nz_bytes);
            const uint8_t* gblk = k_nz + static_cast<uint64_t>(seg0) * 4u;
            const uint32_t nz_addr = (fits ? smem_u32(sl + 1024) : 0u) + (segw - seg0) * 4u;
            __syncwarp();
            build_records(reinterpret_cast<const uint64_t*>(sl) + warp * 32, nz_addr, rec);
            __syncwarp();
            float sc[G][2];
#pragma unroll
            for (int g = 0; g < G; ++g) sc[g][0] = sc[g][1] = 0.f;
            if (fits) {
#pragma unroll 4
                for (int j = 0; j < 32; ++j) {
                    float v0, v1;
                    decode_pair<true>(rec_base + j * 16, lc, nullptr, v0, v1);
                    const float* qc = qs + (warp * 32 + j) * G;
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        sc[g][0] = fmaf(qc[g], v0, sc[g][0]);
                        sc[g][1] = fmaf(qc[g], v1, sc[g][1]);
                    }
                }
            } else {
#pragma unroll 2
                for (int j = 0; j < 32; ++j) {
                    float v0, v1;
                    decode_pair<false>(rec_base + j * 16, lc, gblk, v0, v1);
                    const float* qc = qs + (warp * 32 + j) * G;
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        sc[g][0] = fmaf(qc[g], v0, sc[g][0]);
                        sc[g][1] = fmaf(qc[g], v1, sc[g][1]);
                    }
                }
            }
            float* sp = spart + (((n & 1) * kAttnWarps + warp) * G) * 64;
#pragma unroll
            for (int g = 0; g < G; ++g) *reinterpret_cast<float2*>(sp + g * 64 + 2 * lane) = make_float2(sc[g][0], sc[g][1]);
        }
        __syncthreads();  // (A) all warps are done with K(n) and V(n-1)
        if (tid == 0) issue_until(2 * n + 5);

        // =============================== softmax update ===============================
        {
            const float* sp = spart + ((n & 1) * kAttnWarps * G) * 64;
            const int tok = (blk0 + n) * kBlockTokens + 2 * lane;
            float mk0 = 0.f, mk1 = 0.f;
            if (mask) {
                mk0 = __half2float(mask[tok]);
                mk1 = __half2float(mask[tok + 1]);
            }
#pragma unroll
            for (int g = 0; g < G; ++g) {
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int w = 0; w < kAttnWarps; ++w) {
                    const float2 t = *reinterpret_cast<const float2*>(sp + (w * G + g) * 64 + 2 * lane);
                    s0 += t.x;
                    s1 += t.y;
                }
                s0 = ref_round_score(s0, p.score_div, ref_round);
                s1 = ref_round_score(s1, p.score_div, ref_round);
                if (mask) {
                    s0 = fmaxf(s0 + mk0, -65504.f);
                    s1 = fmaxf(s1 + mk1, -65504.f);
                }
                const float bm = warp_max(fmaxf(s0, s1));
                const float m_new = fmaxf(m_run[g], bm);
                const float corr = exp2f((m_run[g] - m_new) * kLog2e);  // exp2(-inf) = 0 on the first block
                const float p0 = exp2f((s0 - m_new) * kLog2e), p1 = exp2f((s1 - m_new) * kLog2e);
                l_run[g] = l_run[g] * corr + warp_sum(p0 + p1);
                o_acc[g][0] *= corr;
                o_acc[g][1] *= corr;
                m_run[g] = m_new;
                ps[(2 * lane) * G + g] = p0;
                ps[(2 * lane + 1) * G + g] = p1;
            }
        }
        // =============================== V item ===============================
        {
            const int item = 2 * n + 1, slot = item & (kSlots - 1);
            mbar_wait(&bars[slot], (item / kSlots) & 1);
            const uint8_t* sl = smem + sm.slots + slot * slot_bytes;
            const uint32_t seg0 = segv[n * 4], segw = segv[n * 4 + warp];
            const bool fits = (segv[n * 4 + 4] - seg0) * 4u <= static_cast<uint32_t>(a.slot_nz_bytes);
            const uint8_t* gblk = v_nz + static_cast<uint64_t>(seg0) * 4u;
            const uint32_t nz_addr = (fits ? smem_u32(sl + 1024) : 0u) + (segw - seg0) * 4u;
            __syncwarp();  // ps written, K records no longer read
            build_records(reinterpret_cast<const uint64_t*>(sl) + warp * 32, nz_addr, rec);
            __syncwarp();
            // this warp's tiles: channel half (warp>>1), tokens 32*(warp&1) .. +31
            const float* pw = ps + (32 * (warp & 1)) * G;
            if (fits) {
#pragma unroll 4
                for (int j = 0; j < 32; ++j) {
                    float v0, v1;
                    decode_pair<true>(rec_base + j * 16, lc, nullptr, v0, v1);
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const float pj = pw[j * G + g];
                        o_acc[g][0] = fmaf(pj, v0, o_acc[g][0]);
                        o_acc[g][1] = fmaf(pj, v1, o_acc[g][1]);
                    }
                }
            } else {
#pragma unroll 2
                for (int j = 0; j < 32; ++j) {
                    float v0, v1;
                    decode_pair<false>(rec_base + j * 16, lc, gblk, v0, v1);
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const float pj = pw[j * G + g];
                        o_acc[g][0] = fmaf(pj, v0, o_acc[g][0]);
                        o_acc[g][1] = fmaf(pj, v1, o_acc[g][1]);
                    }
                }
            }
        }
    }
    // ---- cross-warp reduction of o: warps (0,1) hold channel half 0, warps (2,3) half 1 ------------
    __syncthreads();
    float* red = spart;                      // [warps][G][64]
    float* ored = spart + kAttnWarps * G * 64;  // [G][128]
#pragma unroll
    for (int g = 0; g < G; ++g) *reinterpret_cast<float2*>(red + (warp * G + g) * 64 + 2 * lane) = make_float2(o_acc[g][0], o_acc[g][1]);
    __syncthreads();
    {
        const int hf = tid >> 6, e = tid & 63;
#pragma unroll
        for (int g = 0; g < G; ++g) ored[g * 128 + tid] = red[((2 * hf) * G + g) * 64 + e] + red[((2 * hf + 1) * G + g) * 64 + e];
    }
    __syncthreads();
    write_partial_and_merge<G>(a, unit, split, ored, m_run, l_run);
}

// ------------------------------------------------------------------------------------------------
// Dense window split: <= 256 tokens of the fp16 residual window (models/llama_mustafar_kernel.py:278, :316).
template <int G>
__device__ __forceinline__ void window_split(const DecodeArgs& a, uint8_t* smem, int unit, int split) {
    const mfb200_decode_params& p = a.p;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t lane = lane_id();
    const int wsplit = split - a.n_csplit;
    const int t0 = wsplit * kWinTokensPerSplit;
    const int nt = min(kWinTokensPerSplit, p.win_len - t0);
    const int b = unit / p.kv_heads, h = unit % p.kv_heads;
    const bool ref_round = (p.flags & MFB200_F_REF_SCORE_ROUNDING) != 0;

    float* qs = reinterpret_cast<float*>(smem);               // [c][G]
    float* sw = qs + kHeadDim * G;                            // [G][256] scores, then probabilities
    float* red = sw + G * kWinTokensPerSplit;                 // [warps][G][128]
    float* ored = red + kAttnWarps * G * 128;                 // [G][128]
    float* stat = ored + G * 128;                             // [warps][G] scratch for block max / sum

    {
        const __half* q = static_cast<const __half*>(p.q) + (static_cast<int64_t>(b) * p.kv_heads + h) * G * kHeadDim;
#pragma unroll
        for (int g = 0; g < G; ++g) qs[tid * G + g] = __half2float(q[g * kHeadDim + tid]);
    }
    __syncthreads();
    const __half* kw = static_cast<const __half*>(p.k_win) + static_cast<int64_t>(unit) * p.win_stride + static_cast<int64_t>(t0) * kHeadDim;
    const __half* vw = static_cast<const __half*>(p.v_win) + static_cast<int64_t>(unit) * p.win_stride + static_cast<int64_t>(t0) * kHeadDim;

    // ---- scores: 8 lanes per token, lane reads two 16-byte chunks (channels 8s..8s+7, 64+8s..64+8s+7)
    {
        const int tsub = lane >> 3, seg = lane & 7;
        float qr[G][16];
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                qr[g][i] = qs[(8 * seg + i) * G + g];
                qr[g][8 + i] = qs[(64 + 8 * seg + i) * G + g];
            }
        constexpr int kUnroll = 4;
        for (int it = 0; it * 16 * kUnroll < nt; ++it) {
            uint4 ka[kUnroll], kb[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int t = (it * kUnroll + u) * 16 + warp * 4 + tsub;
                if (t < nt) {
                    ka[u] = ldg_stream_v4(kw + static_cast<int64_t>(t) * kHeadDim + 8 * seg);
                    kb[u] = ldg_stream_v4(kw + static_cast<int64_t>(t) * kHeadDim + 64 + 8 * seg);
                } else {
                    ka[u] = make_uint4(0, 0, 0, 0);
                    kb[u] = make_uint4(0, 0, 0, 0);
                }
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int t = (it * kUnroll + u) * 16 + warp * 4 + tsub;
                const uint32_t w[8] = {ka[u].x, ka[u].y, ka[u].z, ka[u].w, kb[u].x, kb[u].y, kb[u].z, kb[u].w};
                float acc[G];
#pragma unroll
                for (int g = 0; g < G; ++g) acc[g] = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
#pragma unroll
                    for (int g = 0; g < G; ++g) acc[g] = fmaf(qr[g][2 * i + 1], f.y, fmaf(qr[g][2 * i], f.x, acc[g]));
                }
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    float s = acc[g];
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    s += __shfl_xor_sync(0xffffffffu, s, 2);
                    s += __shfl_xor_sync(0xffffffffu, s, 4);
                    if (seg == 0 && t < nt) sw[g * kWinTokensPerSplit + t] = s;
                }
            }
        }
    }
    __syncthreads();
    // ---- softmax over the nt window tokens (thread owns tokens tid and tid+128) -----------------------
    float m_loc[G], l_loc[G];
    {
        const __half* mask = p.mask ? static_cast<const __half*>(p.mask) + static_cast<int64_t>(b) * p.mask_stride + p.comp_len + t0 : nullptr;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            float s[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int t = tid + r * 128;
                s[r] = -INFINITY;
                if (t < nt) {
                    s[r] = ref_round_score(sw[g * kWinTokensPerSplit + t], p.score_div, ref_round);
                    if (mask) s[r] = fmaxf(s[r] + __half2float(mask[t]), -65504.f);
                }
            }
            const float wm = warp_max(fmaxf(s[0], s[1]));
            if (lane == 0) stat[warp * G + g] = wm;
            __syncthreads();
            float mx = stat[g];
#pragma unroll
            for (int w = 1; w < kAttnWarps; ++w) mx = fmaxf(mx, stat[w * G + g]);
            __syncthreads();
            float ls = 0.f;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int t = tid + r * 128;
                if (t < nt) {
                    const float e = exp2f((s[r] - mx) * kLog2e);
                    sw[g * kWinTokensPerSplit + t] = e;
                    ls += e;
                }
            }
            ls = warp_sum(ls);
            if (lane == 0) stat[warp * G + g] = ls;
            __syncthreads();
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < kAttnWarps; ++w) tot += stat[w * G + g];
            __syncthreads();
            m_loc[g] = mx;
            l_loc[g] = tot;
        }
    }
    // ---- P.V: warp takes tokens == warp (mod 4); lane owns channels 4*lane .. 4*lane+3 -----------------
    {
        float o[G][4];
#pragma unroll
        for (int g = 0; g < G; ++g) o[g][0] = o[g][1] = o[g][2] = o[g][3] = 0.f;
        constexpr int kUnroll = 8;
        for (int it = 0; it * 4 * kUnroll < nt; ++it) {
            uint2 vv[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int t = (it * kUnroll + u) * 4 + warp;
                vv[u] = (t < nt) ? ldg_stream_v2(vw + static_cast<int64_t>(t) * kHeadDim + 4 * lane) : make_uint2(0, 0);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int t = (it * kUnroll + u) * 4 + warp;
                if (t < nt) {
                    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&vv[u].x));
                    const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&vv[u].y));
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const float pt = sw[g * kWinTokensPerSplit + t];
                        o[g][0] = fmaf(pt, f0.x, o[g][0]);
                        o[g][1] = fmaf(pt, f0.y, o[g][1]);
                        o[g][2] = fmaf(pt, f1.x, o[g][2]);
                        o[g][3] = fmaf(pt, f1.y, o[g][3]);
                    }
                }
            }
        }
#pragma unroll
        for (int g = 0; g < G; ++g) *reinterpret_cast<float4*>(red + (warp * G + g) * 128 + 4 * lane) = make_float4(o[g][0], o[g][1], o[g][2], o[g][3]);
    }
    __syncthreads();
#pragma unroll
    for (int g = 0; g < G; ++g) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kAttnWarps; ++w) s += red[(w * G + g) * 128 + tid];
        ored[g * 128 + tid] = s;
    }
    __syncthreads();
    write_partial_and_merge<G>(a, unit, split, ored, m_loc, l_loc);
}

template <int G>
__global__ void __launch_bounds__(kAttnThreads) sparse_decode_attn_kernel(const __grid_constant__ DecodeArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int unit = blockIdx.y, split = blockIdx.x;
    if (split < a.n_csplit) compressed_split<G>(a, smem, unit, split);
    else window_split<G>(a, smem, unit, split);
}

static size_t window_smem_bytes(int G) {
    return static_cast<size_t>(kHeadDim * G + G * kWinTokensPerSplit + kAttnWarps * G * 128 + G * 128 + kAttnWarps * G) * 4;
}

static int pick_slot_nz_bytes(const mfb200_decode_params* p) {
    // capacity of one ring slot's nonzero area; without a hint the worst case (every element kept).
    int kb = p->slot_kb;
    if (kb <= 0 || kb > 16) kb = 16;
    return kb * 1024;
}

template <int G>
static int launch_decode(const DecodeArgs& a, cudaStream_t s) {
    const SmemMap sm = smem_map(G, a.slot_nz_bytes);
    size_t smem = a.n_csplit > 0 ? sm.total : 0;
    if (a.n_wsplit > 0) smem = smem > window_smem_bytes(G) ? smem : window_smem_bytes(G);
    static size_t configured = 0;
    if (smem > configured) {
        MFB_CUDA(cudaFuncSetAttribute(sparse_decode_attn_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
        configured = smem;
    }
    dim3 grid(a.n_csplit + a.n_wsplit, a.p.batch * a.p.kv_heads);
    sparse_decode_attn_kernel<G><<<grid, kAttnThreads, smem, s>>>(a);
    return launch_status("sparse_decode_attn_kernel");
}

static void split_counts(int comp_len, int win_len, int n_split, int* n_csplit, int* n_wsplit) {
    const int nw = (win_len + kWinTokensPerSplit - 1) / kWinTokensPerSplit;
    *n_wsplit = nw;
    *n_csplit = n_split - nw;
}

}  // namespace mfb

using namespace mfb;

extern "C" int mfb200_decode_plan(int batch, int kv_heads, int groups, int comp_len, int win_len, int sm_count,
                                  size_t* workspace_bytes, size_t* counter_bytes) {
    MFB_REQUIRE(batch > 0 && kv_heads > 0, "decode_plan: batch/kv_heads must be positive");
    MFB_REQUIRE(groups == 1 || groups == 2 || groups == 4 || groups == 8, "decode_plan: groups=%d not in {1,2,4,8}", groups);
    MFB_REQUIRE(comp_len >= 0 && comp_len % 64 == 0, "decode_plan: comp_len=%d must be a multiple of 64", comp_len);
    MFB_REQUIRE(win_len >= 0 && comp_len + win_len >= 1, "decode_plan: empty context");
    if (sm_count <= 0) {
        int dev = 0;
        MFB_CUDA(cudaGetDevice(&dev));
        MFB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    const int64_t units = static_cast<int64_t>(batch) * kv_heads;
    MFB_REQUIRE(units <= 65535, "decode_plan: batch*kv_heads=%lld exceeds 65535", static_cast<long long>(units));
    const int nblk = comp_len / kBlockTokens;
    const int nw = (win_len + kWinTokensPerSplit - 1) / kWinTokensPerSplit;
    int nc = 0;
    if (nblk > 0) {
        const int64_t target_ctas = static_cast<int64_t>(sm_count) * 3;  // ~3 resident CTAs per SM
        int64_t per_unit = (target_ctas + units - 1) / units - nw;
        const int min_c = (nblk + kMaxBlocksPerSplit - 1) / kMaxBlocksPerSplit;
        if (per_unit < min_c) per_unit = min_c;
        if (per_unit > nblk) per_unit = nblk;
        nc = static_cast<int>(per_unit);
    }
    const int n_split = nc + nw;
    const size_t cbytes = (static_cast<size_t>(units) * 4 + 255) & ~static_cast<size_t>(255);
    if (counter_bytes) *counter_bytes = cbytes;
    if (workspace_bytes) *workspace_bytes = cbytes + static_cast<size_t>(units) * n_split * groups * kPartStride * 4;
    return n_split;
}

extern "C" int mfb200_sparse_decode_attention(const mfb200_decode_params* p, mfb200_stream_t stream) {
    MFB_REQUIRE(p != nullptr, "decode: null params");
    MFB_REQUIRE(p->batch > 0 && p->kv_heads > 0, "decode: batch/kv_heads must be positive");
    MFB_REQUIRE(static_cast<int64_t>(p->batch) * p->kv_heads <= 65535, "decode: too many units");
    MFB_REQUIRE(p->groups == 1 || p->groups == 2 || p->groups == 4 || p->groups == 8, "decode: groups=%d not in {1,2,4,8}", p->groups);
    MFB_REQUIRE(p->comp_len >= 0 && p->comp_len % 64 == 0, "decode: comp_len=%d must be a multiple of 64", p->comp_len);
    MFB_REQUIRE(p->win_len >= 0 && p->comp_len + p->win_len >= 1, "decode: empty context");
    MFB_REQUIRE(p->q && p->out && p->workspace, "decode: q/out/workspace must not be null");
    MFB_REQUIRE(p->score_div > 0.f, "decode: score_div must be positive");
    if (p->comp_len > 0) {
        MFB_REQUIRE(p->k_bmp && p->k_idx && p->k_nz && p->k_nz_off && p->v_bmp && p->v_idx && p->v_nz && p->v_nz_off,
                    "decode: compressed cache pointers must not be null when comp_len > 0");
        MFB_REQUIRE(p->bmp_stride >= p->comp_len * 2 && p->idx_stride >= p->comp_len * 2 + 1, "decode: bmp/idx stride too small");
        MFB_REQUIRE(p->bmp_stride % 2 == 0, "decode: bmp_stride must keep 16-byte alignment");
        MFB_REQUIRE(((reinterpret_cast<uintptr_t>(p->k_nz) | reinterpret_cast<uintptr_t>(p->v_nz) |
                      reinterpret_cast<uintptr_t>(p->k_bmp) | reinterpret_cast<uintptr_t>(p->v_bmp)) & 15) == 0,
                    "decode: bitmap / nonzero buffers must be 16-byte aligned");
    }
    if (p->win_len > 0) {
        MFB_REQUIRE(p->k_win && p->v_win, "decode: window pointers must not be null when win_len > 0");
        MFB_REQUIRE(p->win_stride >= static_cast<int64_t>(p->win_len) * kHeadDim && p->win_stride % 8 == 0, "decode: bad win_stride");
        MFB_REQUIRE(((reinterpret_cast<uintptr_t>(p->k_win) | reinterpret_cast<uintptr_t>(p->v_win)) & 15) == 0,
                    "decode: window buffers must be 16-byte aligned");
    }
    if (p->mask) MFB_REQUIRE(p->mask_stride >= p->comp_len + p->win_len, "decode: mask_stride too small");
    DecodeArgs a;
    a.p = *p;
    split_counts(p->comp_len, p->win_len, p->n_split, &a.n_csplit, &a.n_wsplit);
    const int nblk = p->comp_len / kBlockTokens;
    MFB_REQUIRE(p->n_split >= 1 && a.n_csplit >= (nblk + kMaxBlocksPerSplit - 1) / kMaxBlocksPerSplit && a.n_csplit <= nblk &&
                    (nblk == 0 || a.n_csplit >= 1),
                "decode: n_split=%d inconsistent with comp_len=%d win_len=%d (use mfb200_decode_plan)", p->n_split, p->comp_len, p->win_len);
    a.slot_nz_bytes = pick_slot_nz_bytes(p);
    auto s = static_cast<cudaStream_t>(stream);
    switch (p->groups) {
        case 1: return launch_decode<1>(a, s);
        case 2: return launch_decode<2>(a, s);
        case 4: return launch_decode<4>(a, s);
        default: return launch_decode<8>(a, s);
    }
}
