// sparse_tile.cuh — warp-level decoding of bitmap+packed-nonzero tiles straight out of shared memory.
//
// The reference kernels (kernel/csrc/SpMM_Kernel.cuh:109-151) rebuild a dense 256x64 tile in shared
// memory with a per-thread serial __clzll walk and then run mma.sync over it.  Here nothing is
// materialised: a warp owns 32 consecutive tiles of a 64-token block; inside a tile every lane owns
// the two ADJACENT positions (2*lane, 2*lane+1), so the rank of its first position needs ONE popc of
// the 32-bit half-word that holds it and the second rank is rank+bit0.  Lanes 0-15 work on the high
// word (positions 0-31), lanes 16-31 on the low word; a per-warp "record" prepared once per tile
// group stores {word, smem byte address of the word's first nonzero} so the inner loop is
//     LDS.64 record -> LOP -> POPC -> IMAD -> 2x LDS.U16 -> 2 selects -> FMA(s)
// per 64 positions.  K tiles (one channel x 64 tokens) give the lane two token scores; V tiles
// (one token x 64 channels) give it two output channels.
#pragma once
#include "common.cuh"

namespace mfb {

struct LaneConst {
    uint32_t above;  // mask of the bits (positions) before this lane's pair inside its 32-bit word
    uint32_t shift;  // left shift that brings the lane's pair to bits 31,30
    uint32_t half;   // 0: high word (positions 0..31), 1: low word
};

__device__ __forceinline__ LaneConst make_lane_const() {
    LaneConst lc;
    const uint32_t l = lane_id();
    const uint32_t sub = l & 15;
    lc.shift = 2 * sub;
    lc.above = sub == 0 ? 0u : (0xffffffffu << (32 - 2 * sub));
    lc.half = l >> 4;
    return lc;
}

// Per-warp record table: rec[tile j in 0..31][half] = {bitmap word, byte address of its first value}.
// `bmp` points at the 32 bitmaps of this warp's tile group in shared memory, `nz_addr` is the
// (shared-space byte address | byte offset from a global base) of the group's first nonzero.
__device__ __forceinline__ void build_records(const uint64_t* bmp, uint32_t nz_addr, uint2* rec) {
    const uint32_t l = lane_id();
    const uint64_t bm = bmp[l];
    const uint32_t hi = static_cast<uint32_t>(bm >> 32), lo = static_cast<uint32_t>(bm);
    const uint32_t pc_hi = __popc(hi);
    const uint32_t padded = (pc_hi + __popc(lo) + 7u) & ~7u;  // halves
    uint32_t incl = padded;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
        if (l >= static_cast<uint32_t>(o)) incl += n;
    }
    const uint32_t start = nz_addr + 2u * (incl - padded);
    rec[2 * l] = make_uint2(hi, start);
    rec[2 * l + 1] = make_uint2(lo, start + 2u * pc_hi);
}

// Values (as fp16 bit patterns, 0 when the bit is clear) of this lane's two positions in tile j.
template <bool NZ_SHARED>
__device__ __forceinline__ void decode_pair(uint32_t rec_addr, const LaneConst& lc, const uint8_t* gbase,
                                            float& v0, float& v1) {
    const uint2 r = lds_v2(rec_addr);
    const uint32_t w = r.x;
    const uint32_t addr = r.y + 2u * __popc(w & lc.above);
    const uint32_t sh = w << lc.shift;
    const bool b0 = (sh & 0x80000000u) != 0, b1 = (sh & 0x40000000u) != 0;
    uint32_t x, y;
    if (NZ_SHARED) {
        // unconditional: a clear bit reads the next value / padding / stale bytes inside our own
        // shared allocation and is discarded by the selects below.
        x = lds_u16(addr);
        y = lds_u16(addr + 2);
    } else {
        // overflow path (block larger than the staging slot): values come straight from global,
        // predicated so that nothing is read past the end of the buffer.
        const uint16_t* g = reinterpret_cast<const uint16_t*>(gbase + addr);
        x = (b0 || b1) ? g[0] : 0;
        y = (b0 && b1) ? g[1] : 0;
    }
    v0 = b0 ? h2f_bits(x) : 0.f;
    v1 = b1 ? h2f_bits(b0 ? y : x) : 0.f;
}

}  // namespace mfb
