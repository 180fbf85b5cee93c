// gqa_mma.cuh — tensor-core contraction for the GQA variant (G >= 4 query heads per KV head).
//
// north_star: "Tensor cores are used only for the GQA case, where the path becomes a small dense contraction
// after in-smem decompression."  With G heads sharing one decoded tile the CUDA-core path needs 2*G FHFMA per
// 64-position tile; here a warp instead decompresses its 32 tiles of a block into a private dense fp16 buffer in
// shared memory (one conflict-free STS.32 per tile: the lane's two adjacent positions) and contracts the whole
// 32 x 64 buffer against the G query rows (K side) or probability rows (V side) with 16 mma.sync.m16n8k16
// (SASS HMMA.16816.F32): rows of the buffer are the contraction index (channels for K tiles, tokens for V
// tiles), so the B fragments come from ldmatrix.x4.trans; A holds the G <= 8 live rows (rows 8..15 are zero).
// A 144-byte row pitch keeps both the STS.32 rows and the ldmatrix 8x8 blocks free of bank conflicts.
// (tcgen05/TMEM would need the decompressed operand as a canonical UMMA shared-memory tile plus a TMEM round
// trip for a 16 x 64 result; for this M = G <= 8 contraction the register-fragment MMA is the better fit and
// the kernel stays bound by the decode, not by the tensor pipe.)
#pragma once
#include "sparse_tile.cuh"

namespace mfb {

constexpr int kDensePitch = 144;                  // bytes per dense row (64 halves + 16 B pad)
constexpr int kDenseWarpBytes = 32 * kDensePitch;  // one warp's 32 x 64 buffer
constexpr int kTcRowPitch = 72;  // elements per [g] row of the score / probability buffers in the tensor-core variant
                                 // (64 + 8: the 4..8 live fragment rows then fall into different banks)

__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr)
                 : "memory");
}
// d += A(16x16, rows 8..15 zero) * B(16x8):  a_lo = A[gid][2tig..2tig+1], a_hi = A[gid][8+2tig..9+2tig]
__device__ __forceinline__ void mma16816_toprows(float (&d)[4], uint32_t a_lo, uint32_t a_hi, uint32_t b0, uint32_t b1) {
    const uint32_t z = 0u;
    asm(  // pure function of its operands: left schedulable
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a_lo), "r"(z), "r"(a_hi), "r"(z), "r"(b0), "r"(b1));
}

// Decompress the warp's 32 tiles (records at `rec`) into its dense buffer: row j = tile j, column = position.
// Eight tiles are decoded into registers before their eight stores are issued, so the shared-memory loads of
// a group are in flight together instead of queueing behind the previous tile's store.
template <bool NZ_SHARED>
__device__ __forceinline__ void decode_to_dense32(const uint2* rec, const LaneConst& lc, const uint8_t* gbase,
                                                  uint32_t dense_addr) {
    const uint32_t col = dense_addr + 4u * lane_id();  // positions (2*lane, 2*lane+1)
#pragma unroll
    for (int j0 = 0; j0 < 32; j0 += 8) {
        uint32_t packed[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const DecodedPair d = decode_pair<NZ_SHARED>(rec + 2 * (j0 + i), lc, gbase);
            const uint32_t v0 = d.b0 ? d.x : 0u, v1 = d.b1 ? d.y : 0u;
            packed[i] = __byte_perm(v0, v1, 0x5410);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) sts_u32(col + (j0 + i) * kDensePitch, packed[i]);
    }
}

// acc[nt] (16 x 8 tiles over the 64 columns) += A(16 x 32) * dense(32 x 64); afrag[ks] = {a_lo, a_hi} of k-step ks.
// The four ldmatrix of a k-step are issued together, then their eight MMAs (no load->MMA bubble per pair).
__device__ __forceinline__ void mma_dense32(uint32_t dense_addr, const uint32_t (&afrag)[2][2], float (&acc)[8][4]) {
    const uint32_t lane = lane_id();
    const uint32_t mi = lane >> 3, ri = lane & 7;
    // ldmatrix.x4: matrices {rows +0..7, rows +8..15} x {column block 2*ntp, 2*ntp+1}
    const uint32_t base = dense_addr + (8u * (mi & 1) + ri) * kDensePitch + 16u * (mi >> 1);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        uint32_t bfrag[4][4];
#pragma unroll
        for (int ntp = 0; ntp < 4; ++ntp)
            ldmatrix_x4_trans(base + ks * 16 * kDensePitch + ntp * 32, bfrag[ntp][0], bfrag[ntp][1], bfrag[ntp][2], bfrag[ntp][3]);
#pragma unroll
        for (int ntp = 0; ntp < 4; ++ntp) {
            mma16816_toprows(acc[2 * ntp], afrag[ks][0], afrag[ks][1], bfrag[ntp][0], bfrag[ntp][1]);
            mma16816_toprows(acc[2 * ntp + 1], afrag[ks][0], afrag[ks][1], bfrag[ntp][2], bfrag[ntp][3]);
        }
    }
}

}  // namespace mfb
