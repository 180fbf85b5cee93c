// gqa_tc.cuh — tcgen05 (5th-generation tensor core) contraction for the GQA variant, G >= 4 query heads per KV head.
//
// north_star: "Tensor cores are used only for the GQA case, where the path becomes a small dense contraction after
// in-smem decompression."  A 64-token block of one KV head serves all G query heads from ONE decode:
//   * the tile warps decompress the block straight into un-swizzled, MN-major UMMA operand tiles in shared memory
//     (one conflict-free STS.32 per 64-position tile: the lane's two adjacent positions); no ldmatrix, no register
//     fragments, no per-warp MMA issue;
//   * scores  S[64 tokens x 8]   = Kd(A: M = tokens,   K = 128 channels) . q^T(B, N = 8)      8 x tcgen05.mma M64 N8  K16
//     output  O[128 chan x 16]   = Vd(A: M = channels, K = 64 tokens)    . p^T(B, N = 16)     4 x tcgen05.mma M128 N16 K16
//     issued by ONE thread each, accumulators in TMEM, completion through tcgen05.commit -> mbarrier;
//   * four epilogue warps read S from TMEM (tcgen05.ld), run the online softmax (the reference's fp16 rounding
//     points included), write p as the next B operand, and fold every block's O into fp32 registers with the
//     softmax rescale — nothing is ever rescaled inside TMEM.
// Operand layout (verified on the hardware by tools/umma_probe.cu, profiles/r2_umma_probe.log): core matrix = 8 MN
// elements (16 B) x 8 K rows, K rows 16 B apart; core matrices 144 B apart along MN (SBO; 128 B + 16 B of padding so
// that the 8 lanes groups of a decode store fall into different banks) and kLboK / kLboV apart along K (LBO).
// The B operands are K-major: [k/8][n (8 rows x 16 B)][k%8], k-groups 128 B apart; the V-side N = 16 aliases its second
// 8-row group onto the first (SBO = 0), columns 8..15 of O are never read.
// M = 64 accumulators live in lanes (row % 16) + 32 * (row / 16), M = 128 in lane = row.
#pragma once
#include "sparse_tile.cuh"

namespace mfb {

constexpr int kTcSbo = 144;                 // bytes between core matrices along MN
constexpr int kLboK = 8 * kTcSbo;           // K operand: 64 tokens = 8 MN groups per 8-channel K group
constexpr int kLboV = 16 * kTcSbo;          // V operand: 128 channels = 16 MN groups per 8-token K group
constexpr int kDenseBytes = 16 * kLboK;     // = 8 * kLboV = 18432: one dense 64-token block of K or of V
constexpr int kQbBytes = 16 * 128;          // q as B operand: 16 channel groups x (8 rows x 16 B)
constexpr int kPbBytes = 8 * 128;           // p as B operand: 8 token groups x (8 rows x 16 B)
constexpr uint32_t kTmemCols = 64;          // S[2] at columns 0 / 8, O[2] at columns 16 / 32
constexpr uint32_t kIdescS = (1u << 4) | (1u << 15) | (1u << 17) | (4u << 24);   // f32 acc, f16 x f16, A MN-major, N=8,  M=64
constexpr uint32_t kIdescO = (1u << 4) | (1u << 15) | (2u << 17) | (8u << 24);   //                                  N=16, M=128

// Shared-memory operand descriptor (64 bit): [0,14) address >> 4, [16,30) LBO >> 4, [32,46) SBO >> 4, [46,48) version = 1,
// layout type (bits 61-63) 0 = no swizzle.  Split into its two words: only the low one changes from MMA to MMA.
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t addr, uint32_t lbo) { return ((addr >> 4) & 0x3fffu) | (((lbo >> 4) & 0x3fffu) << 16); }
__device__ __forceinline__ constexpr uint32_t umma_desc_hi(uint32_t sbo) { return ((sbo >> 4) & 0x3fffu) | (1u << 14); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {  // arrives on `bar` when all MMAs issued so far by this thread are done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base) {  // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(kTmemCols) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {  // this thread's lane, 8 consecutive columns
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void sts_b32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// Decompress the warp's 32 tiles (records at `rec`) into a UMMA operand tile: tile j lands at `dst + (j & 7) * 16 +
// (j >> 3) * lbo` (8 K rows of one core-matrix row group, then the next K group), `dst` already carries the lane's MN
// offset ((lane >> 2) * 144 + (lane & 3) * 4) and the warp's first K group.  One instantiation serves K and V tiles
// (lbo is a run-time value) and only 8 tiles are unrolled: the kernel's warp roles all run different code, so its
// instruction footprint matters (the fully unrolled first version stalled on instruction fetch).
// Per tile: LDS.64 record, LOP3, POPC, IMAD, 2 x LDS.U16, 2 x LOP3->P, 2 x SEL, PRMT, STS.32.  The rank is taken
// INCLUSIVE of the lane's first position (one popc over `above | bit0`): the second value then sits at that rank and
// the first one slot before it, so neither load depends on a bit test.
template <bool NZ_SHARED>
__device__ __forceinline__ void decode_to_umma32(const uint2* rec, const LaneConst& lc, const uint8_t* gbase, uint32_t dst, uint32_t lbo) {
    const uint32_t above1 = lc.above | lc.bit0;
#pragma unroll 1
    for (int j0 = 0; j0 < 32; j0 += 8, rec += 16, dst += lbo) {
        uint32_t packed[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint2 r = rec[2 * i];
            uint32_t addr1;  // address of the value of the lane's SECOND position (if set); the first one's is addr1 - 2
            asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(addr1) : "r"(__popc(r.x & above1)), "r"(r.y));
            const bool b0 = (r.x & lc.bit0) != 0, b1 = (r.x & lc.bit1) != 0;
            uint32_t a, b;
            if (NZ_SHARED) {
                a = lds_u16(addr1 - 2);  // rank 0 with b0 clear reads the 2 bytes in front of the group (bitmap area): unused
                b = lds_u16(addr1);
            } else {
                const uint16_t* g = reinterpret_cast<const uint16_t*>(gbase + addr1);
                a = b0 ? g[-1] : 0;
                b = b1 ? g[0] : 0;
            }
            packed[i] = __byte_perm(b0 ? a : 0u, b1 ? b : 0u, 0x5410);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) sts_b32(dst + i * 16, packed[i]);
    }
}

}  // namespace mfb
