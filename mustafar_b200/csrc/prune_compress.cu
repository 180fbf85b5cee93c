// prune_compress.cu — per-token magnitude pruning and bitmap/packed-nonzero compression.
//
// Replaces (paths relative to /root/reference):
//   models/llama_mustafar_kernel.py:77-153   dh_prune_key / dh_prune_value   (torch.kthvalue + mask)
//   kernel/compression.py:9-115              calculate_bitmap_{key,value}_batched (Triton)
//   kernel/compression.py:118-247            compress_{key,value}_batched         (Triton)
//   kernel/compression.py:294-304            cumsum/cat glue
//
// Design (B200): one CTA owns one 64-token block of one head.  The block (64 x 128 fp16 = 16 KB) is
// read once with 8-byte coalesced loads (a warp reads one 256-byte token row), optionally pruned in
// registers by a warp-level radix select on the 15-bit magnitudes, and parked in shared memory with
// a 65-word row pitch so that both the row-wise (V) and the column-wise (K) tile reads are free of
// bank conflicts.  Bitmap words come from __ballot_sync over lanes that hold the elements in reverse order (MSB = element 0), ranks from
// __popc of the ballot below the lane.  Packed tiles are assembled in a per-warp staging line and
// leave as one contiguous (<=128 B) store per tile.
#include "common.cuh"

#ifndef MFB_LOOKBACK_SLEEP
#define MFB_LOOKBACK_SLEEP 40
#endif
#ifndef MFB_SELECT_EARLY_EXIT
#define MFB_SELECT_EARLY_EXIT 1
#endif

namespace mfb {

constexpr int kPitch = kHeadDim + 2;  // halves; 65 words -> conflict-free column reads
constexpr int kCompressThreads = 256;

// k-th smallest of the 128 magnitudes of a row, held 4 per lane as two registers of two packed 15-bit keys
// (kx = v.x & 0x7fff7fff, ky = v.y & 0x7fff7fff), k in [1,128].
// Greedy MSB-first construction of the largest T with count(key < T) < k  ==  the k-th smallest.
// One round costs ~9 instructions: with c = cand - 1 replicated into both 16-bit fields, (0x8000 + c) - key stays
// inside its field and has bit 15 set iff key < cand, so ONE subtraction compares two keys; a sign-replicating PRMT
// turns the four flag bits into 0x00 / 0xFF bytes, IDP4A sums them (255 per hit), REDUX.SUM adds the lanes.
// `u` carries threshold + this round's bias so that the bias costs no instruction of its own.
// (A compare/select/add chain per key was 17 instructions per round and made the prune ALU-bound at 1.9 TB/s.)
template <int START>
__device__ __forceinline__ uint32_t kth_search(uint32_t u, uint32_t kx, uint32_t ky, uint32_t k255) {
    constexpr uint32_t kRep = 0x10001u;
    // Round `bit` tests cand = t + 2^bit.  u = 0x8000 + cand - 1 in both fields; accepting the bit moves the next
    // round's u up by 2^(bit-1), rejecting it moves it down by 2^(bit-1) (fields never carry or borrow).
#if MFB_SELECT_EARLY_EXIT
    const uint32_t km1_255 = k255 - 255u;
#endif
    bool accept = false;
#pragma unroll
    for (int bit = START; bit >= 0; --bit) {
        uint32_t flags;
        asm("prmt.b32 %0, %1, %2, 0xFDB9;" : "=r"(flags) : "r"(u - kx), "r"(u - ky));  // sign bytes of the 4 fields
        const uint32_t c255 = __reduce_add_sync(0xffffffffu, __dp4a(flags, 0x01010101u, 0u));
#if MFB_SELECT_EARLY_EXIT
        // Exactly k-1 keys lie below cand: the k-th smallest is the smallest key >= cand - one warp minimum instead of the
        // remaining rounds.  A bisection over VALUES keeps halving an interval that soon holds a single key: on Gaussian rows
        // this exit is taken after 10.3 rounds on average instead of 15 (warp-uniform branch).
        if (c255 == km1_255) {
            uint32_t m01, m23;  // 0xffff in the fields whose key is below cand (flag bytes 0xFF), else 0
            asm("prmt.b32 %0, %1, %1, 0x1100;" : "=r"(m01) : "r"(flags));
            asm("prmt.b32 %0, %1, %1, 0x3322;" : "=r"(m23) : "r"(flags));
            const uint32_t cx = kx | m01, cy = ky | m23;  // keys below cand become 0xffff (keys are <= 0x7fff)
            const uint32_t a = min(cx & 0xffffu, cx >> 16), b = min(cy & 0xffffu, cy >> 16);
            return __reduce_min_sync(0xffffffffu, min(a, b));
        }
#endif
        accept = c255 < k255;  // fewer than k keys below cand: the k-th smallest is >= cand
        if (bit > 0) {
            const uint32_t delta = (1u << (bit - 1)) * kRep;
            u += accept ? delta : 0u - delta;
        }
    }
    return (u & 0x7fffu) + (accept ? 1u : 0u);  // round 0: u = 0x8000 + t
}

__device__ __forceinline__ uint32_t warp_kth_smallest(uint32_t kx, uint32_t ky, int k) {
    return kth_search<14>(0x80008000u + ((1u << 14) - 1u) * 0x10001u, kx, ky, static_cast<uint32_t>(k) * 255u);
}

// The same k-th smallest, started from a HINT (the threshold of the previous token row of the same head: neighbouring
// rows have neighbouring thresholds).  One double count checks that the answer lies in the 1024-wide window around the
// hint - count(key < lo) < k <= count(key < lo + 1024), both counts in ONE REDUX (16-bit fields) - and the bisection then
// only walks the window's 10 bits; any row whose threshold is elsewhere takes the full search.  Exact either way.
#ifndef MFB_SELECT_HINT
#define MFB_SELECT_HINT 1
#endif
#ifndef MFB_SELECT_NOINLINE
#define MFB_SELECT_NOINLINE 1
#endif
constexpr uint32_t kNoHint = 0xffffffffu;
#if MFB_SELECT_NOINLINE
__device__ __noinline__
#else
__device__ __forceinline__
#endif
uint32_t warp_kth_smallest_near(uint32_t kx, uint32_t ky, int k, uint32_t hint) {
    constexpr uint32_t kRep = 0x10001u;
    const uint32_t k255 = static_cast<uint32_t>(k) * 255u;
#if MFB_SELECT_HINT
    if (hint <= 0x7fffu) {
        const uint32_t lo = min(hint > 512u ? hint - 512u : 0u, 0x8000u - 1024u);
        const uint32_t ua = (0x7fffu + lo) * kRep, ub = ua + 1024u * kRep;  // u of cand = lo and of cand = lo + 1024
        uint32_t fa, fb;
        asm("prmt.b32 %0, %1, %2, 0xFDB9;" : "=r"(fa) : "r"(ua - kx), "r"(ua - ky));
        asm("prmt.b32 %0, %1, %2, 0xFDB9;" : "=r"(fb) : "r"(ub - kx), "r"(ub - ky));
        const uint32_t c = __reduce_add_sync(0xffffffffu, __dp4a(fa, 0x01010101u, 0u) + (__dp4a(fb, 0x01010101u, 0u) << 16));
        if ((c & 0xffffu) < k255 && (c >> 16) >= k255) return kth_search<9>(ua + 512u * kRep, kx, ky, k255);
    }
#endif
    return kth_search<14>(0x80008000u + ((1u << 14) - 1u) * kRep, kx, ky, k255);
}

// Applies  x * (|x| >= thr)  to 4 halves packed in a uint2; dropped entries keep their sign bit
// (fp16 x * 0 = +-0), exactly what `key_states_flat * mask` produces.
__device__ __forceinline__ uint2 prune4(uint2 v, int k, uint32_t* hint = nullptr) {
    const uint32_t kx = v.x & 0x7fff7fffu, ky = v.y & 0x7fff7fffu;
    uint32_t thr;
    if (hint != nullptr) {  // rows of one head in sequence: start from the previous row's threshold
        thr = warp_kth_smallest_near(kx, ky, k, *hint);
        *hint = thr;
    } else {
        thr = warp_kth_smallest(kx, ky, k);
    }
    // per field: key >= thr  <=>  bit 15 of (0x8000 + key - thr); widen the flag to a keep mask, the sign bit always stays
    const uint32_t tb = thr * 0x10001u;
    const uint32_t gx = ((kx | 0x80008000u) - tb) & 0x80008000u, gy = ((ky | 0x80008000u) - tb) & 0x80008000u;
    const uint32_t keep_lo = (gx - (gx >> 15)) | 0x80008000u;  // 0x8000 -> 0x7fff per field, 0 -> 0; then | sign bits
    const uint32_t keep_hi = (gy - (gy >> 15)) | 0x80008000u;
    return make_uint2(v.x & keep_lo, v.y & keep_hi);
}

// One warp prunes kPruneRowsPerWarp consecutive token rows ([256, 32512, 128]: 1 row 1.52 ms, 2 rows 1.42, 4 rows 1.25, 8 rows 1.21): all loads are issued first, and each select starts from the
// previous row's threshold (warp_kth_smallest_near).
#ifndef MFB_PRUNE_ROWS_PER_WARP
#define MFB_PRUNE_ROWS_PER_WARP 8
#endif
constexpr int kPruneRowsPerWarp = MFB_PRUNE_ROWS_PER_WARP;
__global__ void __launch_bounds__(256) prune_rows_kernel(const uint2* __restrict__ x, uint2* __restrict__ y,
                                                         int64_t rows, int k) {
    const int64_t row0 = (static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5)) * kPruneRowsPerWarp;
    if (row0 >= rows) return;
    const uint32_t lane = lane_id();
    const int n = static_cast<int>(min(static_cast<int64_t>(kPruneRowsPerWarp), rows - row0));
    uint2 v[kPruneRowsPerWarp];
#pragma unroll
    for (int i = 0; i < kPruneRowsPerWarp; ++i)
        if (i < n) v[i] = x[(row0 + i) * 32 + lane];
    uint32_t hint = kNoHint;
#pragma unroll
    for (int i = 0; i < kPruneRowsPerWarp; ++i)
        if (i < n) y[(row0 + i) * 32 + lane] = prune4(v[i], k, &hint);
}

// ---- other pruning policies feeding the same format ------------------------------------------
// Output-aware key pruning (models/llama_mustafar_Kt_Opa_Vt_Mag.py:98-106, :131-156): the score of element c of a row is
// |x[c] * w[c]| (an fp16 product, as the reference's fp16 tensor ops compute it), w = the unit's folded |q|; the n_keep
// highest scores survive.  Threshold form: keep score >= the k-th smallest score, k = 128 - n_keep + 1 - the reference's
// sort + scatter keeps the same set whenever the n_keep-th and (n_keep+1)-th highest scores differ; when they tie it keeps an
// arbitrary subset of the tied elements (torch.sort is not stable), here every tied element survives.
__global__ void __launch_bounds__(256) prune_rows_scored_kernel(const uint2* __restrict__ x, const uint2* __restrict__ w,
                                                                uint2* __restrict__ y, int64_t rows, int64_t rows_per_unit, int k) {
    // like prune_rows_kernel: a warp takes kPruneRowsPerWarp consecutive rows, loads first, hinted selects after
    const int64_t row0 = (static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5)) * kPruneRowsPerWarp;
    if (row0 >= rows) return;
    const uint32_t lane = lane_id();
    const int n = static_cast<int>(min(static_cast<int64_t>(kPruneRowsPerWarp), rows - row0));
    uint2 v[kPruneRowsPerWarp], sc[kPruneRowsPerWarp];
#pragma unroll
    for (int i = 0; i < kPruneRowsPerWarp; ++i)
        if (i < n) {
            v[i] = x[(row0 + i) * 32 + lane];
            // rows_per_unit > 0: w = the unit's weights, score = |x * w|; else w holds the caller's own score rows (the
            // decode-time form: an accumulated score, `:131-145`)
            sc[i] = w[(rows_per_unit > 0 ? (row0 + i) / rows_per_unit : row0 + i) * 32 + lane];
        }
    uint32_t hint = kNoHint;
#pragma unroll
    for (int i = 0; i < kPruneRowsPerWarp; ++i)
        if (i < n) {
            uint32_t kx, ky;
            if (rows_per_unit > 0) {
                const __half2 s0 = __hmul2_rn(*reinterpret_cast<const __half2*>(&v[i].x), *reinterpret_cast<const __half2*>(&sc[i].x));
                const __half2 s1 = __hmul2_rn(*reinterpret_cast<const __half2*>(&v[i].y), *reinterpret_cast<const __half2*>(&sc[i].y));
                kx = *reinterpret_cast<const uint32_t*>(&s0) & 0x7fff7fffu, ky = *reinterpret_cast<const uint32_t*>(&s1) & 0x7fff7fffu;
            } else {
                kx = sc[i].x & 0x7fff7fffu, ky = sc[i].y & 0x7fff7fffu;
            }
            hint = warp_kth_smallest_near(kx, ky, k, hint);
            const uint32_t tb = hint * 0x10001u;
            const uint32_t gx = ((kx | 0x80008000u) - tb) & 0x80008000u, gy = ((ky | 0x80008000u) - tb) & 0x80008000u;
            y[(row0 + i) * 32 + lane] = make_uint2(v[i].x & ((gx - (gx >> 15)) | 0x80008000u), v[i].y & ((gy - (gy >> 15)) | 0x80008000u));
        }
}

// Channel-wise value pruning (models/llama_mustafar_Kt_Mag_Vc_Mag.py:107-170): inside every group of `group` consecutive
// tokens, each CHANNEL keeps the entries with |x| >= the k-th smallest magnitude of that channel's `group` values.
// One CTA of 128 threads per (unit, group); thread c owns channel c: bitwise bisection of the 15-bit magnitude over the
// column held in shared memory (a warp reads 64 contiguous bytes per token: conflict-free).
constexpr int kMaxGroupTokens = 128;  // 32 KB of static shared memory
__global__ void __launch_bounds__(128) prune_token_groups_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int64_t tokens,
                                                                 int group, int k) {
    __shared__ __align__(16) uint16_t tile[kMaxGroupTokens * kHeadDim];
    const int64_t base = (static_cast<int64_t>(blockIdx.y) * tokens + static_cast<int64_t>(blockIdx.x) * group) * (kHeadDim / 8);
    const int n16 = group * (kHeadDim / 8);
    for (int i = threadIdx.x; i < n16; i += 128) reinterpret_cast<uint4*>(tile)[i] = x[base + i];
    __syncthreads();
    const int c = threadIdx.x;
    uint32_t t = 0;
    for (int bit = 14; bit >= 0; --bit) {  // largest t with count(key < t) < k  ==  the k-th smallest
        const uint32_t cand = t | (1u << bit);
        int cnt = 0;
        for (int r = 0; r < group; ++r) cnt += (tile[r * kHeadDim + c] & 0x7fffu) < cand ? 1 : 0;
        if (cnt < k) t = cand;
    }
    for (int r = 0; r < group; ++r) {
        const uint16_t v = tile[r * kHeadDim + c];
        tile[r * kHeadDim + c] = (v & 0x7fffu) >= t ? v : (v & 0x8000u);  // x * False = +-0
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n16; i += 128) y[base + i] = reinterpret_cast<const uint4*>(tile)[i];
}

// Loads one 64-token block of head h into smem (pitch kPitch), pruning on the fly if prune_k > 0.
__device__ __forceinline__ void load_block(const __half* __restrict__ x, int64_t tokens, int64_t h, int tb,
                                           int prune_k, uint16_t* tile) {
    const uint32_t lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const uint2* src = reinterpret_cast<const uint2*>(x + (h * tokens + static_cast<int64_t>(tb) * 64) * kHeadDim);
    uint2 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = ldg_stream_v2(src + (warp * 8 + i) * 32 + lane);
    uint32_t hint = kNoHint;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (prune_k > 0) v[i] = prune4(v[i], prune_k, &hint);
        uint32_t* dst = reinterpret_cast<uint32_t*>(tile + (warp * 8 + i) * kPitch + 4 * lane);
        dst[0] = v[i].x;
        dst[1] = v[i].y;
    }
}

// Elements (tile-local indices 31 - lane and 63 - lane) of tile `tl` (0..127 inside the block).  Lanes take the elements in
// REVERSE order so that a ballot over the lanes IS the bitmap word (bit 31 = element 0 = the format's MSB-first order)
// without a BREV; the elements before a lane's are then the bits ABOVE its own (lane_above / lane_bit).
template <int LAYOUT>
__device__ __forceinline__ void tile_elems(const uint16_t* tile, int tl, uint32_t lane, uint32_t& e0, uint32_t& e1) {
    const uint32_t el = 31u - lane;
    if (LAYOUT == MFB200_LAYOUT_KEY) {  // tile = channel tl, elements = tokens
        e0 = tile[el * kPitch + tl];
        e1 = tile[(el + 32) * kPitch + tl];
    } else {  // tile = (half = tl/64, token = tl%64), elements = channels of that half
        const int r = tl & 63, hf = tl >> 6;
        e0 = tile[r * kPitch + hf * 64 + el];
        e1 = tile[r * kPitch + hf * 64 + 32 + el];
    }
}
__device__ __forceinline__ uint32_t lane_above(uint32_t lane) { return lane == 31 ? 0u : (0xfffffffeu << lane); }  // bits of the elements before mine
__device__ __forceinline__ uint32_t lane_bit(uint32_t lane) { return 1u << lane; }

template <int LAYOUT>
__global__ void __launch_bounds__(kCompressThreads)
compress_count_kernel(const __half* __restrict__ x, int64_t tokens, int prune_k, int64_t* __restrict__ bitmaps,
                      int32_t* __restrict__ counts) {
    __shared__ __align__(16) uint16_t tile[64 * kPitch];
    const int64_t h = blockIdx.y;
    const int tb = blockIdx.x;
    load_block(x, tokens, h, tb, prune_k, tile);
    __syncthreads();
    const uint32_t lane = lane_id();
    const int warp = threadIdx.x >> 5;
    uint32_t my_hi = 0, my_lo = 0;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        uint32_t e0, e1;
        tile_elems<LAYOUT>(tile, warp * 16 + i, lane, e0, e1);
        const uint32_t hi = __ballot_sync(0xffffffffu, (e0 & 0x7fffu) != 0);
        const uint32_t lo = __ballot_sync(0xffffffffu, (e1 & 0x7fffu) != 0);
        if (lane == static_cast<uint32_t>(i)) {
            my_hi = hi;
            my_lo = lo;
        }
    }
    if (lane < 16) {
        const int64_t t = h * (tokens * 2) + static_cast<int64_t>(tb) * 128 + warp * 16 + lane;
        bitmaps[t] = static_cast<int64_t>((static_cast<uint64_t>(my_hi) << 32) | my_lo);
        counts[t] = ((__popc(my_hi) + __popc(my_lo) + 7) & ~7) >> 1;
    }
}

__global__ void __launch_bounds__(1024)
compress_scan_kernel(const int32_t* __restrict__ counts, int64_t tiles, int32_t* __restrict__ accum,
                     int64_t accum_stride, int64_t tile_offset, int32_t* __restrict__ head_total) {
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry_s;
    const int64_t h = blockIdx.x;
    const int32_t* c = counts + h * tiles;
    int32_t* a = accum + h * accum_stride + tile_offset;
    const uint32_t lane = lane_id();
    const int warp = threadIdx.x >> 5;
    int32_t carry = 0;
    if (tile_offset > 0) carry = a[0];
    else if (threadIdx.x == 0) a[0] = 0;
    for (int64_t base = 0; base < tiles; base += 1024) {
        const int64_t i = base + threadIdx.x;
        int32_t v = (i < tiles) ? c[i] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int32_t n = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= static_cast<uint32_t>(o)) v += n;
        }
        if (lane == 31) warp_tot[warp] = v;
        __syncthreads();
        if (warp == 0) {
            int32_t w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int32_t n = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= static_cast<uint32_t>(o)) w += n;
            }
            warp_tot[lane] = w;  // inclusive over warps
            if (lane == 31) carry_s = w;
        }
        __syncthreads();
        const int32_t prev = (warp > 0) ? warp_tot[warp - 1] : 0;
        if (i < tiles) a[i + 1] = carry + prev + v;
        carry += carry_s;
        __syncthreads();
    }
    if (head_total != nullptr && threadIdx.x == 0) head_total[h] = carry;
}

template <int LAYOUT>
__global__ void __launch_bounds__(kCompressThreads)
compress_pack_kernel(const __half* __restrict__ x, int64_t tokens, const int64_t* __restrict__ bitmaps,
                     const int32_t* __restrict__ accum, int64_t accum_stride, int64_t tile_offset,
                     const int64_t* __restrict__ head_base, __half* __restrict__ packed, int64_t head_capacity,
                     int32_t* __restrict__ overflow) {
    __shared__ __align__(16) uint16_t tile[64 * kPitch];
    __shared__ __align__(16) uint16_t stage[8][64];
    const int64_t h = blockIdx.y;
    const int tb = blockIdx.x;
    load_block(x, tokens, h, tb, /*prune_k=*/0, tile);
    __syncthreads();
    const uint32_t lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const uint32_t above = lane_above(lane), mybit = lane_bit(lane);
    const int64_t tile0 = static_cast<int64_t>(tb) * 128 + warp * 16;
    // one coalesced read of this warp's 16 bitmaps and 16 offsets
    uint64_t bm_l = 0;
    int32_t off_l = 0;
    if (lane < 16) {
        bm_l = static_cast<uint64_t>(bitmaps[h * (tokens * 2) + tile0 + lane]);
        off_l = accum[h * accum_stride + tile_offset + tile0 + lane];
    }
    uint16_t* base = reinterpret_cast<uint16_t*>(packed) + head_base[h];
    uint32_t* st32 = reinterpret_cast<uint32_t*>(stage[warp]);
#pragma unroll 2
    for (int i = 0; i < 16; ++i) {
        const uint64_t bm = __shfl_sync(0xffffffffu, bm_l, i);
        const int32_t off = __shfl_sync(0xffffffffu, off_l, i);
        const uint32_t hi = static_cast<uint32_t>(bm >> 32), lo = static_cast<uint32_t>(bm);
        uint32_t e0, e1;
        tile_elems<LAYOUT>(tile, warp * 16 + i, lane, e0, e1);
        st32[lane] = 0;
        __syncwarp();
        const uint32_t pc_hi = __popc(hi);
        if (hi & mybit) stage[warp][__popc(hi & above)] = static_cast<uint16_t>(e0);
        if (lo & mybit) stage[warp][pc_hi + __popc(lo & above)] = static_cast<uint16_t>(e1);
        __syncwarp();
        const uint32_t n_pad = (pc_hi + __popc(lo) + 7u) & ~7u;
        if (head_capacity > 0 && 2 * static_cast<int64_t>(off) + n_pad > head_capacity) {
            if (lane == 0 && overflow != nullptr) atomicExch(overflow, 1);
        } else if (2 * lane < n_pad) {
            uint32_t* dst = reinterpret_cast<uint32_t*>(base + 2 * static_cast<int64_t>(off));
            dst[lane] = st32[lane];
        }
        __syncwarp();
    }
}

// ---- decode-time append: prune + compress the first 256 window rows of every unit, K and V, ONE launch -------
// (models/llama_mustafar_kernel.py:324-398: dh_prune_* + convert_*_batched + the python index/offset surgery
// + dropping the 256 rows from the window).  One CTA of 1024 threads owns one (unit, K|V): the 64 KB chunk
// lives in shared memory, 32 warps prune 8 rows each, bitmap / count / in-CTA scan / pack follow without
// leaving the CTA, results go straight into the cache slabs at the unit's current tile offset, and the
// CTA finally moves the remaining window rows to the front.  No host sync, no temporaries, no second read.
constexpr int kChunkTokens = 256;
constexpr int kChunkTiles = kChunkTokens * 2;  // 512
constexpr int kChunkThreads = 1024;

struct ChunkArgs {
    __half* win[2];          // k_win, v_win: [units, win_stride]
    int64_t win_stride;      // halves
    int win_len;             // rows currently in the window (>= 256)
    int prune_k[2];
    int64_t* bmp[2];         // [units, bmp_stride]
    int32_t* idx[2];         // [units, idx_stride]
    __half* nz[2];
    const int64_t* head_base[2];  // halves
    int64_t bmp_stride, idx_stride, tile_offset, head_capacity;
    int32_t* overflow;
};

template <int LAYOUT>
__device__ __forceinline__ void compress_chunk_body(const ChunkArgs& a, int which, uint16_t* tile, uint32_t* bm_hi,
                                                    uint32_t* bm_lo, int32_t* off, uint16_t (*stage)[64]) {
    const int64_t h = blockIdx.x;
    const uint32_t lane = lane_id();
    const int warp = threadIdx.x >> 5;
    __half* win = a.win[which] + h * a.win_stride;
    // 1. load + prune: warp w owns rows 8w .. 8w+7
    {
        // coherent loads (not ld.global.nc): the same kernel overwrites these window rows in step 5
        const uint2* src = reinterpret_cast<const uint2*>(win);
        uint2 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = ld_coherent_v2(src + (warp * 8 + i) * 32 + lane);
        uint32_t hint = kNoHint;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[i] = prune4(v[i], a.prune_k[which], &hint);
            uint32_t* dst = reinterpret_cast<uint32_t*>(tile + (warp * 8 + i) * kPitch + 4 * lane);
            dst[0] = v[i].x;
            dst[1] = v[i].y;
        }
    }
    __syncthreads();
    // 2. bitmaps + counts: warp w owns tiles 16w .. 16w+15 (token block w/8)
    const int t0 = warp * 16;
    const uint16_t* blk = tile + (t0 >> 7) * 64 * kPitch;
    uint32_t my_hi = 0, my_lo = 0;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        uint32_t e0, e1;
        tile_elems<LAYOUT>(blk, (t0 + i) & 127, lane, e0, e1);
        const uint32_t hi = __ballot_sync(0xffffffffu, (e0 & 0x7fffu) != 0);
        const uint32_t lo = __ballot_sync(0xffffffffu, (e1 & 0x7fffu) != 0);
        if (lane == static_cast<uint32_t>(i)) {
            my_hi = hi;
            my_lo = lo;
        }
    }
    if (lane < 16) {
        bm_hi[t0 + lane] = my_hi;
        bm_lo[t0 + lane] = my_lo;
        a.bmp[which][h * a.bmp_stride + a.tile_offset + t0 + lane] =
            static_cast<int64_t>((static_cast<uint64_t>(my_hi) << 32) | my_lo);
    }
    __syncthreads();
    // 3. exclusive scan of the 512 padded counts (warps 0..15, one tile per lane) on top of the unit's running offset
    __shared__ int32_t warp_tot[16];
    int32_t incl = 0, cnt = 0;
    if (warp < 16) {
        const int t = warp * 32 + lane;
        cnt = ((__popc(bm_hi[t]) + __popc(bm_lo[t]) + 7) & ~7) >> 1;
        incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= static_cast<uint32_t>(o)) incl += n;
        }
        if (lane == 31) warp_tot[warp] = incl;
    }
    __syncthreads();
    if (warp < 16) {
        int32_t* idx = a.idx[which] + h * a.idx_stride + a.tile_offset;
        int32_t base = a.tile_offset > 0 ? idx[0] : 0;
        for (int w = 0; w < warp; ++w) base += warp_tot[w];
        const int t = warp * 32 + lane;
        off[t] = base + incl - cnt;
        idx[t + 1] = base + incl;
        if (t == 0 && a.tile_offset == 0) idx[0] = 0;
    }
    __syncthreads();
    // 4. pack: one contiguous (<= 128 B) store per tile
    const uint32_t above = lane_above(lane), mybit = lane_bit(lane);
    uint16_t* out = reinterpret_cast<uint16_t*>(a.nz[which]) + a.head_base[which][h];
    uint32_t* st32 = reinterpret_cast<uint32_t*>(stage[warp]);
#pragma unroll 2
    for (int i = 0; i < 16; ++i) {
        const uint32_t hi = bm_hi[t0 + i], lo = bm_lo[t0 + i];
        const int32_t o2 = off[t0 + i];
        uint32_t e0, e1;
        tile_elems<LAYOUT>(blk, (t0 + i) & 127, lane, e0, e1);
        st32[lane] = 0;
        __syncwarp();
        const uint32_t pc_hi = __popc(hi);
        if (hi & mybit) stage[warp][__popc(hi & above)] = static_cast<uint16_t>(e0);
        if (lo & mybit) stage[warp][pc_hi + __popc(lo & above)] = static_cast<uint16_t>(e1);
        __syncwarp();
        const uint32_t n_pad = (pc_hi + __popc(lo) + 7u) & ~7u;
        if (a.head_capacity > 0 && 2 * static_cast<int64_t>(o2) + n_pad > a.head_capacity) {
            if (lane == 0 && a.overflow != nullptr) atomicExch(a.overflow, 1);
        } else if (2 * lane < n_pad) {
            reinterpret_cast<uint32_t*>(out + 2 * static_cast<int64_t>(o2))[lane] = st32[lane];
        }
        __syncwarp();
    }
    // 5. drop the 256 compressed rows: move rows [256, win_len) to the front (all reads of rows < 256 are done)
    const int rest16 = (a.win_len - kChunkTokens) * 16;  // uint4 per remaining rows
    uint4* w4 = reinterpret_cast<uint4*>(win);
    for (int i = threadIdx.x; i < rest16; i += kChunkThreads) {
        const uint4 v = w4[kChunkTokens * 16 + i];
        w4[i] = v;  // rest <= 256 rows: source and destination ranges never overlap
    }
}

__global__ void __launch_bounds__(kChunkThreads) compress_append_chunk_kernel(const ChunkArgs a) {
    extern __shared__ __align__(16) uint8_t csm[];
    uint16_t* tile = reinterpret_cast<uint16_t*>(csm);                                  // [256][kPitch]
    uint32_t* bm_hi = reinterpret_cast<uint32_t*>(csm + kChunkTokens * kPitch * 2);     // [512]
    uint32_t* bm_lo = bm_hi + kChunkTiles;
    int32_t* off = reinterpret_cast<int32_t*>(bm_lo + kChunkTiles);                     // [512]
    uint16_t(*stage)[64] = reinterpret_cast<uint16_t(*)[64]>(off + kChunkTiles);         // [32][64]
    if (blockIdx.y == 0) compress_chunk_body<MFB200_LAYOUT_KEY>(a, 0, tile, bm_hi, bm_lo, off, stage);
    else compress_chunk_body<MFB200_LAYOUT_VALUE>(a, 1, tile, bm_hi, bm_lo, off, stage);
}
constexpr size_t kChunkSmem = kChunkTokens * kPitch * 2 + 3 * kChunkTiles * 4 + 32 * 64 * 2;

// ---- prefill: prune + compress a whole prompt, K and V, in ONE single-pass launch ----------------------------
// (models/llama_mustafar_kernel.py:416-442: dh_prune_* + convert_*_batched on the prompt.)  CTA = one 64-token
// block of one (unit, K|V): the block is read once (strided source, no .contiguous() copy), pruned in registers,
// turned into bitmaps + padded counts, and packed straight into the cache slabs.  The packed offset of a block
// is the sum of the padded counts of all earlier blocks of the unit; instead of a count pass + scan pass + a
// second read of the input, blocks exchange that sum with a decoupled look-back (Merrill & Garland): each CTA
// publishes its aggregate as soon as it is known, then warp 0 walks back over the predecessors' status words
// (32 per probe) until it meets an inclusive prefix.  A CTA's logical block index is a TICKET drawn from a
// per-(unit, K|V) atomic counter, not blockIdx.x: whoever holds ticket t knows that tickets < t were drawn by CTAs
// that are already running or finished, so the wait cannot deadlock whatever order the hardware dispatches CTAs in.
// Algorithmic bytes per token-head and stream: 256 read + 16 bitmap + 8 idx + the padded nonzeros written.
struct PrefillArgs {
    const __half* x[2];                   // key_states, value_states: [B, Hkv, tokens, 128], innermost stride 1
    int64_t stride_b[2], stride_h[2], stride_t[2];  // halves
    int kv_heads;
    int prune_k[2];
    int64_t* bmp[2];
    int32_t* idx[2];
    __half* nz[2];
    const int64_t* head_base[2];
    int64_t bmp_stride, idx_stride, tile_offset, head_capacity;
    int32_t* overflow;
    unsigned long long* status;           // [2][units][blocks], zeroed by the host: (state << 32) | value
    unsigned int* ticket;                 // [2][units], zeroed by the host: next logical block of the chain
};
constexpr unsigned long long kStAggregate = 1ull << 32, kStPrefix = 2ull << 32;

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Warp-collective.  Returns the exclusive prefix of block `tb` (= carry + totals of blocks 0..tb-1) and publishes
// this block's status.  The status word carries its own payload, so relaxed (L2-coherent) accesses are enough.
__device__ __forceinline__ int32_t lookback_exclusive(unsigned long long* st, int tb, int32_t total, int32_t carry) {
    const uint32_t lane = lane_id();
    if (lane == 0) st_relaxed_u64(&st[tb], kStAggregate | static_cast<uint32_t>(total));
    int32_t excl = 0;
    for (int look = tb - 1;; look -= 32) {
        const int j = look - static_cast<int>(lane);
        unsigned long long s;
        while (true) {  // block -1 is a virtual predecessor whose inclusive prefix is the unit's carry-in
            s = j >= 0 ? ld_relaxed_u64(&st[j]) : (kStPrefix | static_cast<uint32_t>(j == -1 ? carry : 0));
            if (!__any_sync(0xffffffffu, (s >> 32) == 0)) break;
            __nanosleep(MFB_LOOKBACK_SLEEP);
        }
        const int32_t val = static_cast<int32_t>(static_cast<uint32_t>(s));
        const uint32_t pm = __ballot_sync(0xffffffffu, (s >> 32) == 2);
        if (pm) {  // nearest inclusive prefix: add the aggregates in front of it and stop
            const uint32_t first = __ffs(pm) - 1;
            excl += __reduce_add_sync(0xffffffffu, lane <= first ? val : 0);
            break;
        }
        excl += __reduce_add_sync(0xffffffffu, val);
    }
    if (lane == 0) st_relaxed_u64(&st[tb], kStPrefix | static_cast<uint32_t>(excl + total));
    return excl;
}

template <int LAYOUT>
__device__ __forceinline__ void prefill_block_body(const PrefillArgs& a, int which, uint16_t* tile,
                                                   uint16_t (*stage)[16][64]) {
    __shared__ int32_t warp_tot[8];
    __shared__ int32_t s_excl;
    __shared__ int s_tb;
    // Units vary fastest in dispatch order: the predecessor of a block in its unit's look-back chain (same unit, previous
    // ticket) was dispatched `units` CTAs earlier and has normally published its inclusive prefix by the time this CTA asks
    // for it.  (Blocks of one unit dispatched back to back all reached the look-back together: warp 0 of every CTA spun on
    // its neighbours - 29 % of the kernel's executed instructions were that loop.)
    const int64_t u = blockIdx.x, units = gridDim.x, nblk = gridDim.y;
    if (threadIdx.x == 0) s_tb = static_cast<int>(atomicAdd(&a.ticket[static_cast<int64_t>(which) * units + u], 1u));
    __syncthreads();
    const int tb = s_tb;
    const uint32_t lane = lane_id();
    const int warp = threadIdx.x >> 5;
    // 1. load + prune: warp w owns token rows 8w .. 8w+7 of the block
    {
        const __half* src = a.x[which] + (u / a.kv_heads) * a.stride_b[which] + (u % a.kv_heads) * a.stride_h[which] +
                            static_cast<int64_t>(tb) * 64 * a.stride_t[which];
        uint2 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = ldg_stream_v2(reinterpret_cast<const uint2*>(src + (warp * 8 + i) * a.stride_t[which]) + lane);
        uint32_t hint = kNoHint;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (a.prune_k[which] > 0) v[i] = prune4(v[i], a.prune_k[which], &hint);
            uint32_t* dst = reinterpret_cast<uint32_t*>(tile + (warp * 8 + i) * kPitch + 4 * lane);
            dst[0] = v[i].x;
            dst[1] = v[i].y;
        }
    }
    // the warp's packing area: 16 tiles x 64 halves, zero = the padding behind each tile's nonzeros
    {
        uint32_t* z = reinterpret_cast<uint32_t*>(stage[warp]);
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i * 32 + lane] = 0;
    }
    __syncthreads();
    // 2. bitmaps + padded counts + packing in ONE pass over the tile's elements: warp w owns tiles 16w .. 16w+15,
    //    lane i < 16 keeps the bitmap of tile 16w+i; every lane drops its two elements at their rank in the packing area.
    //    The 16 tiles of a warp are neighbours in the packed stream too, so they are packed back to back (`run` = the
    //    warp-uniform fill level in halves, padded per tile) and leave as ONE contiguous run.
    const int t0 = warp * 16;
    const uint32_t above = lane_above(lane);  // bits of elements before mine
    uint16_t* pack = &stage[warp][0][0];
    uint32_t my_hi = 0, my_lo = 0, run = 0, my_end = 0;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        uint32_t e0, e1;
        tile_elems<LAYOUT>(tile, t0 + i, lane, e0, e1);
        const bool nz0 = (e0 & 0x7fffu) != 0, nz1 = (e1 & 0x7fffu) != 0;
        const uint32_t hi = __ballot_sync(0xffffffffu, nz0);
        const uint32_t lo = __ballot_sync(0xffffffffu, nz1);
        const uint32_t pc_hi = __popc(hi);
        if (nz0) pack[run + __popc(hi & above)] = static_cast<uint16_t>(e0);
        if (nz1) pack[run + pc_hi + __popc(lo & above)] = static_cast<uint16_t>(e1);
        run += (pc_hi + __popc(lo) + 7u) & ~7u;
        if (lane == static_cast<uint32_t>(i)) {
            my_hi = hi;
            my_lo = lo;
            my_end = run;
        }
    }
    const int64_t tile_g = a.tile_offset + static_cast<int64_t>(tb) * 128 + t0 + lane;  // lanes < 16
    if (lane < 16) a.bmp[which][u * a.bmp_stride + tile_g] = static_cast<int64_t>((static_cast<uint64_t>(my_hi) << 32) | my_lo);
    if (lane == 0) warp_tot[warp] = static_cast<int32_t>(run >> 1);  // counts are kept in units of 2 halves
    __syncthreads();
    // 3. this block's offset inside the unit's packed stream
    int32_t* idx = a.idx[which] + u * a.idx_stride;
    if (warp == 0) {
        int32_t total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) total += warp_tot[w];
        const int32_t carry = a.tile_offset > 0 ? idx[a.tile_offset] : 0;
        unsigned long long* st = a.status + (static_cast<int64_t>(which) * units + u) * nblk;
        const int32_t excl = lookback_exclusive(st, tb, total, carry);
        if (lane == 0) {
            s_excl = excl;
            if (tb == 0 && a.tile_offset == 0) idx[0] = 0;
        }
    }
    __syncthreads();
    int32_t base = s_excl;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    if (lane < 16) idx[tile_g + 1] = base + static_cast<int32_t>(my_end >> 1);
    // 4. copy out the warp's run: 16-byte chunks (every tile start is 16-byte aligned relative to the unit's base)
    uint16_t* out = reinterpret_cast<uint16_t*>(a.nz[which]) + a.head_base[which][u] + 2 * static_cast<int64_t>(base);
    uint32_t chunks = run >> 3;
    if (a.head_capacity > 0 && 2 * static_cast<int64_t>(base) + run > a.head_capacity) {
        if (lane == 0 && a.overflow != nullptr) atomicExch(a.overflow, 1);
        const int64_t room = a.head_capacity - 2 * static_cast<int64_t>(base);  // halves that still fit (never write past the slab)
        chunks = room > 0 ? static_cast<uint32_t>(room >> 3) : 0u;
    }
    __syncwarp();
    for (uint32_t c = lane; c < chunks; c += 32) reinterpret_cast<uint4*>(out)[c] = reinterpret_cast<const uint4*>(pack)[c];
}

__global__ void __launch_bounds__(kCompressThreads) compress_prefill_kernel(const PrefillArgs a) {
    __shared__ __align__(16) uint16_t tile[64 * kPitch];
    __shared__ __align__(16) uint16_t stage[8][16][64];  // per warp: its 16 tiles, packed and zero-padded
    if (blockIdx.z == 0) prefill_block_body<MFB200_LAYOUT_KEY>(a, 0, tile, stage);
    else prefill_block_body<MFB200_LAYOUT_VALUE>(a, 1, tile, stage);
}

// ---- window append: win[u, pos, :] = row[u, :] (K and V in one launch) -------------------------
__global__ void __launch_bounds__(256)
window_append_kernel(uint4* __restrict__ k_win, uint4* __restrict__ v_win, int64_t win_stride_v4,
                     const uint4* __restrict__ k_row, const uint4* __restrict__ v_row, int64_t units, int64_t pos) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // 16 uint4 per row
    if (i >= units * 16) return;
    const int64_t u = i >> 4, j = i & 15;
    k_win[u * win_stride_v4 + pos * 16 + j] = k_row[i];
    v_win[u * win_stride_v4 + pos * 16 + j] = v_row[i];
}

}  // namespace mfb

using namespace mfb;

extern "C" int mfb200_prune_rows(const void* x, void* y, int64_t rows, int k, mfb200_stream_t stream) {
    MFB_REQUIRE(x && y, "prune_rows: null pointer");
    MFB_REQUIRE(k >= 1 && k <= kHeadDim, "prune_rows: k=%d out of [1,128]", k);
    MFB_REQUIRE(rows >= 0, "prune_rows: rows < 0");
    MFB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 7) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0,
                "prune_rows: pointers must be 8-byte aligned");
    if (rows == 0) return MFB200_OK;
    const int64_t grid = (rows + 8 * kPruneRowsPerWarp - 1) / (8 * kPruneRowsPerWarp);
    prune_rows_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint2*>(x), static_cast<uint2*>(y), rows, k);
    return launch_status("prune_rows_kernel");
}

extern "C" int mfb200_prune_rows_scored(const void* x, const void* w, void* y, int64_t rows, int64_t rows_per_unit, int k,
                                        mfb200_stream_t stream) {
    MFB_REQUIRE(x && w && y, "prune_rows_scored: null pointer");
    MFB_REQUIRE(k >= 1 && k <= kHeadDim, "prune_rows_scored: k=%d out of [1,128]", k);
    MFB_REQUIRE(rows >= 0 && rows_per_unit >= 0, "prune_rows_scored: rows < 0 or rows_per_unit < 0");
    MFB_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(y)) & 7) == 0,
                "prune_rows_scored: pointers must be 8-byte aligned");
    if (rows == 0) return MFB200_OK;
    prune_rows_scored_kernel<<<static_cast<unsigned>((rows + 8 * kPruneRowsPerWarp - 1) / (8 * kPruneRowsPerWarp)), 256, 0,
                               static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint2*>(x), static_cast<const uint2*>(w), static_cast<uint2*>(y), rows, rows_per_unit, k);
    return launch_status("prune_rows_scored_kernel");
}

extern "C" int mfb200_prune_token_groups(const void* x, void* y, int64_t units, int64_t tokens, int group, int k, mfb200_stream_t stream) {
    MFB_REQUIRE(x && y, "prune_token_groups: null pointer");
    MFB_REQUIRE(group >= 1 && group <= kMaxGroupTokens, "prune_token_groups: group=%d out of [1,%d]", group, kMaxGroupTokens);
    MFB_REQUIRE(k >= 1 && k <= group, "prune_token_groups: k=%d out of [1,group]", k);
    MFB_REQUIRE(units >= 0 && units <= 65535 && tokens >= 0 && tokens % group == 0,
                "prune_token_groups: tokens=%lld must be a multiple of group=%d (units <= 65535)", static_cast<long long>(tokens), group);
    MFB_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, "prune_token_groups: pointers must be 16-byte aligned");
    if (units == 0 || tokens == 0) return MFB200_OK;
    const dim3 grid(static_cast<unsigned>(tokens / group), static_cast<unsigned>(units));
    prune_token_groups_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(x), static_cast<uint4*>(y),
                                                                                     tokens, group, k);
    return launch_status("prune_token_groups_kernel");
}

extern "C" int mfb200_compress_count(const void* x, int64_t heads, int64_t tokens, int layout, int prune_k,
                                     int64_t* bitmaps, int32_t* counts, mfb200_stream_t stream) {
    MFB_REQUIRE(x && bitmaps && counts, "compress_count: null pointer");
    MFB_REQUIRE(tokens % 64 == 0 && tokens >= 0, "compress_count: tokens=%lld must be a multiple of 64",
                static_cast<long long>(tokens));
    MFB_REQUIRE(layout == MFB200_LAYOUT_KEY || layout == MFB200_LAYOUT_VALUE, "compress_count: bad layout %d", layout);
    MFB_REQUIRE(prune_k >= 0 && prune_k <= kHeadDim, "compress_count: prune_k=%d out of [0,128]", prune_k);
    MFB_REQUIRE(heads >= 0 && heads <= 65535, "compress_count: heads=%lld out of range", static_cast<long long>(heads));
    MFB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 7) == 0, "compress_count: x must be 8-byte aligned");
    if (heads == 0 || tokens == 0) return MFB200_OK;
    dim3 grid(static_cast<unsigned>(tokens / 64), static_cast<unsigned>(heads));
    auto s = static_cast<cudaStream_t>(stream);
    if (layout == MFB200_LAYOUT_KEY)
        compress_count_kernel<MFB200_LAYOUT_KEY><<<grid, kCompressThreads, 0, s>>>(static_cast<const __half*>(x), tokens,
                                                                                     prune_k, bitmaps, counts);
    else
        compress_count_kernel<MFB200_LAYOUT_VALUE><<<grid, kCompressThreads, 0, s>>>(static_cast<const __half*>(x),
                                                                                       tokens, prune_k, bitmaps, counts);
    return launch_status("compress_count_kernel");
}

extern "C" int mfb200_compress_scan(const int32_t* counts, int64_t heads, int64_t tiles, int32_t* accum,
                                    int64_t accum_stride, int64_t tile_offset, int32_t* head_total,
                                    mfb200_stream_t stream) {
    MFB_REQUIRE(counts && accum, "compress_scan: null pointer");
    MFB_REQUIRE(heads >= 0 && tiles >= 0 && tile_offset >= 0, "compress_scan: negative size");
    MFB_REQUIRE(accum_stride >= tile_offset + tiles + 1, "compress_scan: accum_stride too small");
    if (heads == 0) return MFB200_OK;
    compress_scan_kernel<<<static_cast<unsigned>(heads), 1024, 0, static_cast<cudaStream_t>(stream)>>>(
        counts, tiles, accum, accum_stride, tile_offset, head_total);
    return launch_status("compress_scan_kernel");
}

extern "C" int mfb200_compress_pack(const void* x, int64_t heads, int64_t tokens, int layout, const int64_t* bitmaps,
                                    const int32_t* accum, int64_t accum_stride, int64_t tile_offset,
                                    const int64_t* head_base, void* packed, int64_t head_capacity, int32_t* overflow,
                                    mfb200_stream_t stream) {
    MFB_REQUIRE(x && bitmaps && accum && head_base && packed, "compress_pack: null pointer");
    MFB_REQUIRE(tokens % 64 == 0 && tokens >= 0, "compress_pack: tokens must be a multiple of 64");
    MFB_REQUIRE(layout == MFB200_LAYOUT_KEY || layout == MFB200_LAYOUT_VALUE, "compress_pack: bad layout %d", layout);
    MFB_REQUIRE(heads >= 0 && heads <= 65535, "compress_pack: heads out of range");
    MFB_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15) == 0, "compress_pack: packed must be 16-byte aligned");
    if (heads == 0 || tokens == 0) return MFB200_OK;
    dim3 grid(static_cast<unsigned>(tokens / 64), static_cast<unsigned>(heads));
    auto s = static_cast<cudaStream_t>(stream);
    if (layout == MFB200_LAYOUT_KEY)
        compress_pack_kernel<MFB200_LAYOUT_KEY><<<grid, kCompressThreads, 0, s>>>(
            static_cast<const __half*>(x), tokens, bitmaps, accum, accum_stride, tile_offset, head_base,
            static_cast<__half*>(packed), head_capacity, overflow);
    else
        compress_pack_kernel<MFB200_LAYOUT_VALUE><<<grid, kCompressThreads, 0, s>>>(
            static_cast<const __half*>(x), tokens, bitmaps, accum, accum_stride, tile_offset, head_base,
            static_cast<__half*>(packed), head_capacity, overflow);
    return launch_status("compress_pack_kernel");
}

extern "C" int mfb200_window_append(void* k_win, void* v_win, int64_t win_stride, const void* k_row,
                                    const void* v_row, int64_t units, int64_t pos, mfb200_stream_t stream) {
    MFB_REQUIRE(k_win && v_win && k_row && v_row, "window_append: null pointer");
    MFB_REQUIRE(win_stride % 8 == 0 && pos >= 0 && (pos + 1) * kHeadDim <= win_stride,
                "window_append: pos=%lld outside the window capacity", static_cast<long long>(pos));
    if (units == 0) return MFB200_OK;
    const int64_t n = units * 16;
    window_append_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<uint4*>(k_win), static_cast<uint4*>(v_win), win_stride / 8, static_cast<const uint4*>(k_row),
        static_cast<const uint4*>(v_row), units, pos);
    return launch_status("window_append_kernel");
}

extern "C" int mfb200_compress_append_chunk(void* k_win, void* v_win, int64_t win_stride, int64_t units, int win_len,
                                            int prune_k_key, int prune_k_value, int64_t* k_bmp, int32_t* k_idx, void* k_nz,
                                            const int64_t* k_head_base, int64_t* v_bmp, int32_t* v_idx, void* v_nz,
                                            const int64_t* v_head_base, int64_t bmp_stride, int64_t idx_stride,
                                            int64_t tile_offset, int64_t head_capacity, int32_t* overflow,
                                            mfb200_stream_t stream) {
    MFB_REQUIRE(k_win && v_win && k_bmp && k_idx && k_nz && k_head_base && v_bmp && v_idx && v_nz && v_head_base,
                "compress_append_chunk: null pointer");
    MFB_REQUIRE(win_len >= kChunkTokens && win_len <= 2 * kChunkTokens, "compress_append_chunk: win_len=%d outside [256, 512]", win_len);
    MFB_REQUIRE(win_stride >= static_cast<int64_t>(win_len) * kHeadDim && win_stride % 8 == 0, "compress_append_chunk: bad win_stride");
    MFB_REQUIRE(prune_k_key >= 0 && prune_k_key <= kHeadDim && prune_k_value >= 0 && prune_k_value <= kHeadDim,
                "compress_append_chunk: prune_k out of [0,128]");
    MFB_REQUIRE(tile_offset >= 0 && tile_offset % 128 == 0 && bmp_stride >= tile_offset + kChunkTiles && idx_stride >= tile_offset + kChunkTiles + 1,
                "compress_append_chunk: cache slab too small for tile_offset=%lld", static_cast<long long>(tile_offset));
    MFB_REQUIRE(units >= 0 && units <= (1 << 20), "compress_append_chunk: units out of range");
    if (units == 0) return MFB200_OK;
    static size_t configured[kMaxDevices] = {0};
    {
        const int rc = ensure_dynamic_smem(compress_append_chunk_kernel, configured, kChunkSmem);
        if (rc) return rc;
    }
    ChunkArgs a;
    a.win[0] = static_cast<__half*>(k_win);
    a.win[1] = static_cast<__half*>(v_win);
    a.win_stride = win_stride;
    a.win_len = win_len;
    a.prune_k[0] = prune_k_key;
    a.prune_k[1] = prune_k_value;
    a.bmp[0] = k_bmp;
    a.bmp[1] = v_bmp;
    a.idx[0] = k_idx;
    a.idx[1] = v_idx;
    a.nz[0] = static_cast<__half*>(k_nz);
    a.nz[1] = static_cast<__half*>(v_nz);
    a.head_base[0] = k_head_base;
    a.head_base[1] = v_head_base;
    a.bmp_stride = bmp_stride;
    a.idx_stride = idx_stride;
    a.tile_offset = tile_offset;
    a.head_capacity = head_capacity;
    a.overflow = overflow;
    dim3 grid(static_cast<unsigned>(units), 2);
    compress_append_chunk_kernel<<<grid, kChunkThreads, kChunkSmem, static_cast<cudaStream_t>(stream)>>>(a);
    return launch_status("compress_append_chunk_kernel");
}

extern "C" int mfb200_compress_prefill(const void* k, const void* v, const int64_t* k_strides, const int64_t* v_strides,
                                       int batch, int kv_heads, int64_t tokens, int prune_k_key, int prune_k_value,
                                       int64_t* k_bmp, int32_t* k_idx, void* k_nz, const int64_t* k_head_base,
                                       int64_t* v_bmp, int32_t* v_idx, void* v_nz, const int64_t* v_head_base,
                                       int64_t bmp_stride, int64_t idx_stride, int64_t tile_offset, int64_t head_capacity,
                                       int32_t* overflow, void* status_ws, mfb200_stream_t stream) {
    MFB_REQUIRE(k && v && k_strides && v_strides && k_bmp && k_idx && k_nz && k_head_base && v_bmp && v_idx && v_nz &&
                    v_head_base && status_ws, "compress_prefill: null pointer");
    MFB_REQUIRE(tokens >= 0 && tokens % 64 == 0, "compress_prefill: tokens=%lld must be a multiple of 64", static_cast<long long>(tokens));
    MFB_REQUIRE(batch >= 0 && kv_heads >= 1 && static_cast<int64_t>(batch) * kv_heads <= 65535, "compress_prefill: batch*kv_heads out of range");
    MFB_REQUIRE(prune_k_key >= 0 && prune_k_key <= kHeadDim && prune_k_value >= 0 && prune_k_value <= kHeadDim,
                "compress_prefill: prune_k out of [0,128]");
    MFB_REQUIRE(tile_offset >= 0 && tile_offset % 128 == 0 && bmp_stride >= tile_offset + tokens * 2 && idx_stride >= tile_offset + tokens * 2 + 1,
                "compress_prefill: cache slab too small for tile_offset=%lld", static_cast<long long>(tile_offset));
    MFB_REQUIRE(tokens / 64 <= 0x7fffffff, "compress_prefill: too many blocks");
    MFB_REQUIRE(((reinterpret_cast<uintptr_t>(k_nz) | reinterpret_cast<uintptr_t>(v_nz)) & 15) == 0,
                "compress_prefill: nonzero slabs must be 16-byte aligned (and every head base a multiple of 8 halves)");
    for (int w = 0; w < 2; ++w) {
        const int64_t* st = w ? v_strides : k_strides;
        MFB_REQUIRE((reinterpret_cast<uintptr_t>(w ? v : k) & 7) == 0 && st[0] % 4 == 0 && st[1] % 4 == 0 && st[2] % 4 == 0 && st[2] >= kHeadDim,
                    "compress_prefill: %s must be 8-byte aligned with strides that are multiples of 4 halves", w ? "v" : "k");
    }
    const int64_t units = static_cast<int64_t>(batch) * kv_heads, nblk = tokens / 64;
    if (units == 0 || nblk == 0) return MFB200_OK;
    auto s = static_cast<cudaStream_t>(stream);
    MFB_CUDA(cudaMemsetAsync(status_ws, 0, static_cast<size_t>(2 * units * nblk + 2 * units) * sizeof(unsigned long long), s));
    PrefillArgs a;
    a.x[0] = static_cast<const __half*>(k);
    a.x[1] = static_cast<const __half*>(v);
    for (int w = 0; w < 2; ++w) {
        const int64_t* st = w ? v_strides : k_strides;
        a.stride_b[w] = st[0];
        a.stride_h[w] = st[1];
        a.stride_t[w] = st[2];
    }
    a.kv_heads = kv_heads;
    a.prune_k[0] = prune_k_key;
    a.prune_k[1] = prune_k_value;
    a.bmp[0] = k_bmp;
    a.bmp[1] = v_bmp;
    a.idx[0] = k_idx;
    a.idx[1] = v_idx;
    a.nz[0] = static_cast<__half*>(k_nz);
    a.nz[1] = static_cast<__half*>(v_nz);
    a.head_base[0] = k_head_base;
    a.head_base[1] = v_head_base;
    a.bmp_stride = bmp_stride;
    a.idx_stride = idx_stride;
    a.tile_offset = tile_offset;
    a.head_capacity = head_capacity;
    a.overflow = overflow;
    a.status = static_cast<unsigned long long*>(status_ws);
    a.ticket = reinterpret_cast<unsigned int*>(a.status + 2 * units * nblk);
    MFB_REQUIRE(nblk <= 65535, "compress_prefill: more than 65535 blocks per unit");
    dim3 grid(static_cast<unsigned>(units), static_cast<unsigned>(nblk), 2);
    compress_prefill_kernel<<<grid, kCompressThreads, 0, s>>>(a);
    return launch_status("compress_prefill_kernel");
}
