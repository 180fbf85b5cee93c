// sparse_tile.cuh — warp-level decoding of bitmap+packed-nonzero tiles straight out of shared memory.
//
// The reference kernels (kernel/csrc/SpMM_Kernel.cuh:109-151) rebuild a dense 256x64 tile in shared
// memory with a per-thread serial __clzll walk and then run mma.sync over it.  Here nothing is
// materialised: inside a tile every lane owns the two ADJACENT positions (2*lane, 2*lane+1), so the
// rank of its first position needs ONE popc of the 32-bit half-word that holds it and the second
// rank is rank+bit0.  Lanes 0-15 work on the high word (positions 0-31), lanes 16-31 on the low
// word; a "record" prepared once per tile stores {word, smem byte address of the word's first
// nonzero} so the inner loop is, per 64 positions,
//     LDS.64 record, LOP3 (mask), POPC, IMAD (address), 2x LDS.U16, 2x LOP3->predicate, 2x @p FHFMA
// FHFMA is Blackwell's mixed-precision FMA (PTX fma.rn.f32.f16: fp16 x fp16 + fp32 -> fp32, one
// rounding): the fp16 nonzero and the fp16 operand (q or p) are multiplied without any conversion
// instruction and accumulated in fp32.
//
// Measured on B200 (tools/ubench.cu): POPC/FLO/BREV issue at 0.5 warp-instr/clk/SM, integer ALU
// (LOP3/IADD3/SHF/SEL) at 2, IMAD ~2, FFMA 4, LDS ~1 — shared-memory wavefronts (3 per tile here) and
// issue slots are what bound the loop, so both are kept minimal.
#pragma once
#include "common.cuh"

namespace mfb {

struct LaneConst {
    uint32_t above;  // mask of the bits (positions) before this lane's pair inside its 32-bit word
    uint32_t bit0;   // mask of the lane's first position
    uint32_t bit1;   // mask of the lane's second position
    uint32_t half;   // 0: high word (positions 0..31), 1: low word
};

__device__ __forceinline__ LaneConst make_lane_const() {
    LaneConst lc;
    const uint32_t l = lane_id();
    const uint32_t sub = l & 15;
    lc.bit0 = 0x80000000u >> (2 * sub);
    lc.bit1 = 0x40000000u >> (2 * sub);
    lc.above = sub == 0 ? 0u : (0xffffffffu << (32 - 2 * sub));
    lc.half = l >> 4;
    return lc;
}

// Record table of one 32-tile group: rec[2*j + half] = {bitmap word, byte address of its first value}.
// Executed by one full warp (lane j <-> tile j).  `bmp` points at the group's 32 bitmaps in shared
// memory, `nz_addr` is the (shared-space byte address | byte offset from a global base) of the
// group's first nonzero.
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
    // one 16-byte store per lane (two 8-byte stores with a 16-byte lane stride cost 4 wavefronts each instead of 2)
    reinterpret_cast<uint4*>(rec)[l] = make_uint4(hi, start, lo, start + 2u * pc_hi);
}

// acc + a*b with a, b fp16 (bit patterns) and fp32 accumulation: SASS FHFMA.
__device__ __forceinline__ float fhfma(uint16_t a, uint16_t b, float acc) {
    asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc) : "h"(a), "h"(b));
    return acc;
}

// fp16 bit patterns of this lane's two positions in one tile: x is valid iff b0, y iff b1 (otherwise
// they hold whatever the load returned and must not be used -> callers predicate their FMAs).
struct DecodedPair {
    uint16_t x, y;
    bool b0, b1;
};

#ifndef MFB_INCL_RANK
#define MFB_INCL_RANK 1
#endif
template <bool NZ_SHARED>
__device__ __forceinline__ DecodedPair decode_pair(const uint2* rec, const LaneConst& lc, const uint8_t* gbase) {
    const uint2 r = *rec;
    const uint32_t w = r.x;
    DecodedPair d;
    d.b0 = (w & lc.bit0) != 0;
    d.b1 = (w & lc.bit1) != 0;
#if MFB_INCL_RANK
    // Rank taken INCLUSIVE of the lane's first position: the second value then sits at that rank and the first one slot
    // before it - neither load depends on a bit test and no select is needed afterwards.  (Rank 0 with the first bit
    // clear reads the 2 bytes in front of the group - the bitmap area of the slot -, never used.)
    uint32_t addr1;  // r.y + 2*rank as one IMAD: keeps the half-rate integer-ALU pipe free
    asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(addr1) : "r"(__popc(w & (lc.above | lc.bit0))), "r"(r.y));
    uint32_t x, y;
    if (NZ_SHARED) {
        x = lds_u16(addr1 - 2);
        y = lds_u16(addr1);
    } else {
        // overflow path (block larger than the staging slot): values come straight from global,
        // predicated so that nothing is read outside the buffer.
        const uint16_t* g = reinterpret_cast<const uint16_t*>(gbase + addr1);
        x = d.b0 ? g[-1] : 0;
        y = d.b1 ? g[0] : 0;
    }
    d.x = static_cast<uint16_t>(x);
    d.y = static_cast<uint16_t>(y);
#else
    uint32_t addr;
    asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(addr) : "r"(__popc(w & lc.above)), "r"(r.y));
    uint32_t x, y;
    if (NZ_SHARED) {
        x = lds_u16(addr);
        y = lds_u16(addr + 2);
    } else {
        const uint16_t* g = reinterpret_cast<const uint16_t*>(gbase + addr);
        x = (d.b0 || d.b1) ? g[0] : 0;
        y = (d.b0 && d.b1) ? g[1] : 0;
    }
    d.x = static_cast<uint16_t>(x);
    d.y = static_cast<uint16_t>(d.b0 ? y : x);  // the pair's second value sits one slot further iff b0
#endif
    return d;
}

// Same, converted to fp32 with cleared positions forced to 0 (reference-compatible SpMV kernels).
template <bool NZ_SHARED>
__device__ __forceinline__ void decode_pair(const uint2* rec, const LaneConst& lc, const uint8_t* gbase, float& v0,
                                            float& v1) {
    const DecodedPair d = decode_pair<NZ_SHARED>(rec, lc, gbase);
    v0 = d.b0 ? h2f_bits(d.x) : 0.f;
    v1 = d.b1 ? h2f_bits(d.y) : 0.f;
}

}  // namespace mfb
