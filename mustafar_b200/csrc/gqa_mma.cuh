// gqa_mma.cuh — tensor-core contraction for the GQA variant (G >= 4 query heads per KV head), register fragments only.
//
// north_star: "Tensor cores are used only for the GQA case, where the path becomes a small dense contraction after
// decompression."  All G query heads of a KV head are served from ONE decode of a tile.  The decoded values never
// touch shared memory again: a lane's two adjacent positions of four consecutive tiles ARE an A fragment of
// mma.sync.m16n8k16 (SASS HMMA.16816.F32), and a block-diagonal B operand keeps the 32 lanes' positions apart:
//
//   A (16 x 16, row-major fragment):  thread (gid = lane/4, tig = lane%4) supplies rows gid and gid+8 at the k-slots
//       K_tig = {2tig, 2tig+1, 2tig+8, 2tig+9}.  It fills row gid   with its EVEN position (2*lane)   of tiles j..j+3
//                                                     and row gid+8 with its ODD  position (2*lane+1) of tiles j..j+3.
//       A row is shared by the four threads of a quad, but each owns its own k-slots.
//   B (16 x 8 per MMA, column n = jq*G + g):  B[k][jq*G + g] = operand[g][tile of slot k]  if k in K_jq, else 0
//       (operand = q for K tiles, p for V tiles).  The zero blocks route the k-slots of quad member jq to its own
//       columns, so D[gid][jq*G + g] = sum over the 4 tiles of (thread (gid, jq)'s even position) * operand[g][tile]
//       and D[gid+8][...] the same for its odd position: 4 tiles x 64 positions x G heads per G/2 MMAs, no shuffles.
//   Thread (gid, tig) holds B[k in K_tig][n = 8m + gid] of MMA m: non-zero only if tig == jq(n), i.e. a lane is "live"
//       in at most ONE of the G/2 MMAs, where its fragment is 8 contiguous bytes operand[g][4 tiles]; in the others it
//       reads the zero row of the same operand block (one LDS.64 per MMA either way, one wavefront).
//
// Per 64-position tile this costs LDS.64 record, LOP3, POPC, IMAD, 2 x LOP3->P, 2 x predicated LDS.U16 (into zeroed
// registers), PRMT, plus per 4 tiles G/2 x (LDS.64 + HMMA): ~10.5 instructions and ~3.5 shared-memory
// wavefronts.  The two earlier GQA variants both staged a dense operand in shared memory - mma.sync + ldmatrix
// (round 1) and tcgen05 with TMEM accumulators (gqa_tc.cuh) - and paid one store wavefront plus one operand-read
// wavefront per tile on the resource that bounds this kernel (ncu: shared-memory data pipe 71 % busy, 6.5 wavefronts
// per tile; the dense operand is 2 of them), plus a fence.proxy.async per block and three more warp roles.
// Non-finite cache values: 0 x Inf/NaN = NaN inside an MMA taints the 4-lane quad instead of one position; the
// reference's softmax spreads a non-finite score over the whole row anyway.
#pragma once
#include "sparse_tile.cuh"

namespace mfb {

#ifndef MFB_SCORE_PITCH
#define MFB_SCORE_PITCH 74
#endif
constexpr int kTcRowPitch = MFB_SCORE_PITCH;  // floats per [g] row of the partial-score buffers of the G >= 4 variants: 64 + 10 keeps the K warps'
                                              // fragment stores (rows g0 / g0+2 within a half-warp) and the softmax warp's 8-byte reads
                                              // (two heads per half-warp) free of bank conflicts (72: 2-way conflicts on the stores; -1 %)
// Operand blocks: the B fragments of one 4-tile group sit in ONE aligned block, row g = operand[g][4 tiles] (8 bytes), row G
// = zeros (what the lanes that are not live in an MMA read).  An LDS.64 whose 5..9 distinct addresses share a 128-byte
// line costs one wavefront; the same rows 144 bytes apart cost 2.4 (tools/ubench.cu, profiles/r2_ubench_lds.txt).
#ifndef MFB_GQA_UNROLL
#define MFB_GQA_UNROLL 8  // fully unrolled: the next group's record loads overlap this group's value loads (2 -> 8: -5..7 %)
#endif
constexpr int kGqaUnroll = MFB_GQA_UNROLL;  // 4-tile groups decoded per loop body
__host__ __device__ constexpr int oper_block_bytes(int G) { return G <= 4 ? 64 : 128; }

__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint2 lds_u64(uint32_t addr) {  // volatile: must stay before the barrier that frees the buffer
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
// d += A(16x16) * B(16x8), fp16 inputs, fp32 accumulation
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint2 b) {
    asm(  // pure function of its operands: left schedulable
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

// Per-lane constants of the block-diagonal scheme for G heads (G = 4 or 8): NM = G/2 MMAs per 4 tiles.
template <int G>
struct GqaLane {
    static constexpr int NM = G / 2;
    uint32_t live_m;   // the MMA this lane's B fragment belongs to, or NM if the lane is never live
    uint32_t g_live;   // the head whose operand row it loads there
    uint32_t g0;       // D columns of this lane: heads g0, g0+1 ...
    uint32_t pos[NM];  // ... of positions pos[m] (c0, c1) and pos[m]+1 (c2, c3) in MMA m
};
template <int G>
__device__ __forceinline__ GqaLane<G> make_gqa_lane() {
    static_assert(G == 4 || G == 8, "block-diagonal HMMA scheme: G in {4, 8}");
    GqaLane<G> gl;
    const uint32_t lane = lane_id(), gid = lane >> 2, tig = lane & 3;
    gl.live_m = GqaLane<G>::NM;
#pragma unroll
    for (int m = 0; m < GqaLane<G>::NM; ++m) {
        const uint32_t n = 8 * m + gid;           // B column held by this lane in MMA m
        if (n / G == tig) gl.live_m = m;          // column (jq, g) with jq == tig: this lane's own k-slots
        const uint32_t nd = 8 * m + 2 * tig;      // first D column held by this lane in MMA m
        gl.pos[m] = 8 * gid + 2 * (nd / G);       // owner thread (gid, jq = nd / G) -> its even position
    }
    gl.g_live = (8 * (gl.live_m < GqaLane<G>::NM ? gl.live_m : 0) + gid) % G;
    gl.g0 = (2 * tig) % G;
    return gl;
}

// Decode 4 consecutive tiles (records rec[0], rec[2], rec[4], rec[6] of this lane's half) straight into an A fragment.
// The rank is taken INCLUSIVE of the lane's first position: the second value then sits at that rank and the first one
// slot before it, so neither load depends on the other bit; a cleared position keeps the zero its register was given.
template <bool NZ_SHARED>
__device__ __forceinline__ void decode_group4(const uint2* rec, const LaneConst& lc, const uint8_t* gbase, uint32_t (&a)[4]) {
    const uint32_t above1 = lc.above | lc.bit0;
    uint32_t x[4], y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint2 r = rec[2 * i];
        uint32_t addr1;
        asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(addr1) : "r"(__popc(r.x & above1)), "r"(r.y));
        x[i] = 0u;
        y[i] = 0u;
        if (NZ_SHARED) {
            if (r.x & lc.bit0) x[i] = lds_u16(addr1 - 2);
            if (r.x & lc.bit1) y[i] = lds_u16(addr1);
        } else {  // overflow path (block larger than the staging slot): straight from global
            const uint16_t* g = reinterpret_cast<const uint16_t*>(gbase + addr1);
            if (r.x & lc.bit0) x[i] = g[-1];
            if (r.x & lc.bit1) y[i] = g[0];
        }
    }
    a[0] = __byte_perm(x[0], x[1], 0x5410);
    a[1] = __byte_perm(y[0], y[1], 0x5410);
    a[2] = __byte_perm(x[2], x[3], 0x5410);
    a[3] = __byte_perm(y[2], y[3], 0x5410);
}

// acc[m] += (the warp's 32 tiles) x operand.  oper[m] = shared address this lane reads its B fragment of MMA m from, in
// the operand block of the warp's first 4-tile group: row g_live in the MMA it is live in, the zero row in the others
// (a plain load per MMA and group - no predicates, no register that would have to survive as "never written").
// (Holding the K side's q fragments in 32 registers for the whole split instead measured 2 % slower.)
template <int G, bool NZ_SHARED>
__device__ __forceinline__ void tiles32_mma(const uint2* rec, const LaneConst& lc, const uint8_t* gbase,
                                            const uint32_t (&oper)[G / 2], float (&acc)[G / 2][4]) {
    constexpr int NM = G / 2;
#pragma unroll kGqaUnroll
    for (int grp = 0; grp < 8; ++grp) {
        uint32_t a[4];
        decode_group4<NZ_SHARED>(rec + 8 * grp, lc, gbase, a);
        uint2 bfr[NM];
#pragma unroll
        for (int m = 0; m < NM; ++m) bfr[m] = lds_u64(oper[m] + oper_block_bytes(G) * grp);
#pragma unroll
        for (int m = 0; m < NM; ++m) mma16816(acc[m], a, bfr[m]);
    }
}

}  // namespace mfb
