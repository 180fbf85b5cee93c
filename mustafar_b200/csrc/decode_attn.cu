// decode_attn.cu — fused sparse-KV decode attention for sm_100a.
//
// One launch replaces the whole reference decode step of one layer
// (models/llama_mustafar_kernel.py:268-320): F.pad(q) -> torch.cat(NZ) -> Key_Kernel -> window matmul ->
// cat -> /sqrt(d) -> +mask -> fp32 softmax -> F.pad(P) -> torch.cat(NZ) -> Value_Kernel -> window matmul ->
// add, and the (never launched) SplitK_Reduction (kernel/csrc/Reduction_Kernel.cuh:26-48).
//
// Work decomposition: unit = (sequence, KV head).  The compressed length is cut into contiguous ranges of 64-token
// blocks - per unit (uniform plan) or across all units (flat plan, large MHA launches) -, the dense window into
// ranges of <= 64 tokens; 1-D grid, compressed CTAs first.  Each CTA keeps flash-decoding state (m, l, o) and
// contributes one fp32 partial per unit it touches; the partials of a unit are merged inside the same launch
// (ticket or flagged protocol, see publish_or_merge) -> single launch, no second kernel, graph-capturable.
// All G query heads of a KV head are served by the same CTA, so the compressed bytes of a KV head cross HBM once
// (the reference re-reads them G times, SpMM_Kernel.cuh:175).
//
// A compressed-split CTA is warp-specialised and has NO CTA-wide barrier in its steady state:
//   warp 9      producer : one lane issues cp.async.bulk (TMA engine) copies of each block's K item and
//                          V item (1 KB bitmaps + the block's contiguous nonzero range) into two
//                          shared-memory rings; completion through mbarrier expect-tx.
//   warps 0-3   K role   : 32 tiles (channels) each -> partial scores of the block's 64 tokens.
//   warp 8      softmax  : sums the 4 partials, online-softmax update, publishes p[64] and the
//                          rescale factor.
//   warps 4-7   V role   : 32 tiles each (one channel half x 32 tokens) -> o accumulators.
// Hand-offs are mbarriers (TMA completion, one -> many signals) or hardware named barriers (many -> one signals,
// see struct Bars), so the K warps run ahead of the V warps by up to two blocks and nobody waits for the slowest
// warp.  Nothing dense is ever rebuilt in shared memory for G <= 2 (sparse_tile.cuh); G >= 4 contracts a per-warp
// dense buffer on the tensor cores (gqa_mma.cuh).
#include <stdlib.h>

#include <type_traits>

#include "gqa_mma.cuh"
#include "gqa_tc.cuh"
#include "sparse_tile.cuh"

#ifndef MFB_GQA_TC
#define MFB_GQA_TC 0  // 0: G >= 4 contracts register fragments with mma.sync (gqa_mma.cuh); 1: the tcgen05 / TMEM variant (gqa_tc.cuh), kept for A/B runs
#endif

namespace mfb {

#ifndef MFB_G1_CTAS
#define MFB_G1_CTAS 3  // resident CTAs per SM the G=1 kernel is compiled for (64 registers, no spills; +3..6 % at batch 1)
#endif
constexpr int kTileWarps = 4;                 // per role
constexpr int kWarpV0 = 4, kWarpSoftmax = 8, kWarpProducer = 9;  // warps 0-3: K role
constexpr int kAttnWarps = 10;
constexpr int kAttnThreads = kAttnWarps * 32;  // 320: CUDA-core variant (G <= 2)
// tcgen05 variant (G >= 4): warps 0-3 K decode, 4-7 V decode, 8-11 softmax / epilogue (TMEM lane quadrant = warp % 4),
// 12 producer (+ TMEM alloc), 13 issues the score MMAs, 14 the output MMAs
constexpr int kTcWarps = 15;
constexpr int kTcThreads = kTcWarps * 32;  // 480
constexpr int kWarpEpi0 = 8, kWarpProducerTc = 12, kWarpMmaS = 13, kWarpMmaO = 14;
constexpr bool use_tc(int G) { return MFB_GQA_TC != 0 && G >= 4; }
constexpr int cta_threads(int G) { return use_tc(G) ? kTcThreads : kAttnThreads; }
#ifndef MFB_G4_CTAS
#define MFB_G4_CTAS 2  // 3 (64 registers, a few spills) measured 5-12 % slower on the config 3 / 5 shapes
#endif
// resident CTAs per SM each instantiation is compiled for (register cap) and planned with
constexpr int ctas_per_sm(int G) { return G <= 1 ? MFB_G1_CTAS : (G == 2 ? 2 : (G == 4 ? (use_tc(G) ? 2 : MFB_G4_CTAS) : 1)); }
constexpr int kWinWarps = 8;                   // warps used by the dense-window path
constexpr int kWinThreads = kWinWarps * 32;
constexpr int kChunk = 16;                     // tiles per operand-register chunk
constexpr int kMaxDepth = 4;
#ifndef MFB_MAX_BLOCKS
#define MFB_MAX_BLOCKS 128
#endif
constexpr int kMaxBlocksPerSplit = MFB_MAX_BLOCKS;  // blocks of one segment (sizes the segment-offset tables in shared memory)
constexpr int kMaxBlocksPerSplitTc = 64;  // tcgen05 variant: smaller segment tables (shared memory goes to the dense operand buffers)
constexpr int kWinTokensPerSplit = 64;
constexpr int kPartStride = 132;  // {fp32 value, tag} entries per (split, head): o[128], m, l, pad
constexpr float kLog2e = 1.4426950408889634f;

struct DecodeArgs {
    mfb200_decode_params p;
    int n_csplit;   // uniform mode: compressed splits per unit ...
    int n_extra;    //   ... plus one for the units u < n_extra (so that the CTA count can match the resident slots)
    int n_wsplit;   // window chunks per unit
    int flat_ctas;  // flat mode (> 0): number of compressed CTAs n; they partition the B = units*nblk blocks evenly:
    int flat_q;     //   B / n   -> CTA c owns global blocks [c*q + min(c, r), ...) : q + 1 blocks if c < r, else q
    int flat_r;     //   B % n
    int max_split;  // partial slots per unit in the workspace (>= every unit's segment count + n_wsplit)
    int flagged;    // split-merge protocol: 1 = flagged partials + designated merger, 0 = ticket (see publish_or_merge)
    int slot_nz_bytes;  // capacity of a slot's nonzero area (multiple of 1024)
    int depth;          // ring depth per stream (K and V each), 2..4
    int nvd;            // tcgen05 variant: dense V buffers (1 or 2; K always has 2)
    mfb200_peer_out peer;  // head-sharded decode: copy of *p.peer (n_peers = 0: off)
};

// barrier indices inside the bars[] array
// mbarriers (shared memory) carry the hand-offs whose waiter normally finds them already complete, and the TMA
// transaction counts; the many -> one hand-offs whose single waiter sleeps most of the time (K warps -> softmax,
// V warps -> softmax, tile warps -> producer) use hardware named barriers instead: a warp parked in bar.sync issues
// nothing, whereas an mbarrier waiter is woken by every event on the CTA's mbarriers and re-polls (measured: 17 %
// of all issued instructions of a large MHA launch were such polls).
struct Bars {
    enum : int {
        kFullK = 0,                  // [depth] producer (TMA tx) -> K warps
        kFullV = kMaxDepth,          // [depth] producer (TMA tx) -> V warps
        kScEmpty = 2 * kMaxDepth,    // [2] softmax -> K warps           (count 1)
        kPFull = kScEmpty + 2,       // [2] softmax -> V warps           (count 1)
        kCount = kPFull + 2
    };
};
// named barrier ids (0 is __syncthreads); every one joins 4 arriving tile warps and 1 waiting warp
struct NamedBars {
    enum : int {
        kScFull = 1,   // [2] K warps arrive, softmax warp syncs
        kPEmpty = 3,   // [2] V warps arrive, softmax warp syncs
        kEmptyK = 5,   // [depth <= 4] K warps arrive, producer warp syncs
        kEmptyV = 9,   // [depth <= 4] V warps arrive, producer warp syncs
    };
};
constexpr int kHandoffThreads = (kTileWarps + 1) * 32;

struct SmemMap {
    uint32_t slots_k, slots_v, bars, rec, qs, spart, ps, corr, stat, segk, segv, total;
};

__host__ __device__ inline SmemMap smem_map(int G, int slot_nz_bytes, int depth) {
    SmemMap m;
    uint32_t o = 0;
    m.slots_k = o;
    o += depth * (1024 + slot_nz_bytes);
    m.slots_v = o;
    o += depth * (1024 + slot_nz_bytes);
    m.bars = o;
    o += ((Bars::kCount * 8 + 15) / 16) * 16;
    m.rec = o;  // uint2 [8 consumer warps][32 tiles][2]
    o += 2 * kTileWarps * 64 * 8;
    m.qs = o;  // half [128][G]  (G >= 4: 32 operand blocks of 4 channels, gqa_mma.cuh)
    o += G >= 4 ? 32 * oper_block_bytes(G) : kHeadDim * G * 2;
    m.spart = o;  // float [2][4 warps][G][sp_pitch]; reused for the final cross-warp reduction of o
    o += 2 * kTileWarps * G * (G >= 4 ? kTcRowPitch : 64) * 4;
    m.ps = o;  // half [2][64][G]  (G >= 4: [2][16 operand blocks of 4 tokens])
    o += G >= 4 ? 2 * 16 * oper_block_bytes(G) : 2 * 64 * G * 2;
    m.corr = o;  // float [2][8]
    o += 2 * 8 * 4;
    m.stat = o;  // float m[8], l[8]; uint32 tag
    o += 2 * 8 * 4 + 16;
    m.segk = o;
    o += (kMaxBlocksPerSplit * 4 + 4) * 4;
    m.segv = o;
    o += (kMaxBlocksPerSplit * 4 + 4) * 4;
    m.total = o;
    return m;
}

// Programmatic dependent launch (PDL): let the next kernel in the stream start launching, and wait for
// everything the previous kernels wrote.  Both are no-ops when the kernel was launched normally.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_prior_grids() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Timeline instrumentation (`make trace`, tools/trace_attn.py): per-CTA phase timestamps from %globaltimer.
#ifdef MFB_TRACE
constexpr int kTraceSlots = 32, kTraceMaxCtas = 8192;  // 0-15 timestamps / values, 16-31 per-role cycle accumulators
__device__ unsigned long long g_trace[2 * kTraceMaxCtas * kTraceSlots];  // two launches: flag bit 0x100 picks the half
__shared__ int s_trace_half;  // set by thread 0 at kernel entry (barriers follow before any other thread traces)
__device__ __forceinline__ void trace_val(int k, unsigned long long v) {
    const int s_half = s_trace_half;
    if (blockIdx.x < kTraceMaxCtas) g_trace[(static_cast<size_t>(s_half) * kTraceMaxCtas + blockIdx.x) * kTraceSlots + k] = v;
}
__device__ __forceinline__ void trace_at(int k) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trace_val(k, t);
}
#define MFB_TRACE_AT(k) trace_at(k)
#define MFB_TRACE_VAL(k, v) trace_val(k, v)
// cycle accounting of one warp role: MFB_TACC_INIT(n) declares n accumulators, MFB_TACC(i) adds the SM clocks since the
// previous mark to accumulator i, MFB_TACC_STORE(slot0, n, cond) writes them to trace slots slot0.. (one lane)
#define MFB_TACC_INIT(n) long long tacc_[n] = {}; long long tacc_t_ = clock64()
#define MFB_TACC(i) do { const long long now_ = clock64(); tacc_[i] += now_ - tacc_t_; tacc_t_ = now_; } while (0)
#define MFB_TACC_STORE(slot0, n, cond) do { if (cond) for (int i_ = 0; i_ < (n); ++i_) trace_val((slot0) + i_, static_cast<unsigned long long>(tacc_[i_])); } while (0)
#else
#define MFB_TRACE_AT(k) ((void)0)
#define MFB_TRACE_VAL(k, v) ((void)0)
#define MFB_TACC_INIT(n) ((void)0)
#define MFB_TACC(i) ((void)0)
#define MFB_TACC_STORE(slot0, n, cond) ((void)0)
#endif

__device__ __forceinline__ void bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {  // whole warp; orders the warp's earlier accesses
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ex2.approx: what exp2f() uses minus its denormal-range fix-up (results below 2^-126 flush to 0; p is rounded to fp16 anyway)
__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// x / d as q0 = x * r, q = q0 + (x - q0 * d) * r with r = 1/d (one Newton correction on the quotient): the fast path of the
// IEEE division routine without its call, range checks and slow path (about 30 instructions per score).  Correctly rounded
// for the operand ranges here (fp16-valued x, d = sqrt(head_dim)) up to rare last-bit cases that vanish in the fp16
// rounding that follows.
__device__ __forceinline__ float div_fast(float x, float d, float r) {
    const float q0 = x * r;
    return fmaf(fmaf(-q0, d, x), r, q0);
}
// same as ref_round_score below, with the division done by div_fast (rdiv = 1 / div)
__device__ __forceinline__ float ref_round_score_fast(float dot, float div, float rdiv, bool ref_rounding) {
    if (ref_rounding) {
        const float s16 = __half2float(__float2half_rn(dot));
        return __half2float(__float2half_rn(div_fast(s16, div, rdiv)));
    }
    return dot * rdiv;
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

__device__ __forceinline__ uint32_t ld_relaxed_u32_fwd(const void* p) {
    uint32_t r;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
    return r;
}
// Flat mode: the compressed CTAs partition the launch's units*nblk blocks (unit-major global block index, < 2^31)
// evenly; a CTA that crosses a unit boundary contributes one segment (= one partial) to each side.
__host__ __device__ inline uint32_t flat_start(const DecodeArgs& a, uint32_t c) {  // first global block of CTA c
    return c * a.flat_q + (c < static_cast<uint32_t>(a.flat_r) ? c : a.flat_r);
}
__host__ __device__ inline uint32_t flat_owner(const DecodeArgs& a, uint32_t x) {  // CTA that owns global block x
    const uint32_t big = static_cast<uint32_t>(a.flat_r) * (a.flat_q + 1);         // blocks held by the (q+1)-sized CTAs
    return x < big ? x / (a.flat_q + 1) : a.flat_r + (x - big) / a.flat_q;
}
// number of compressed partials of `unit`
__host__ __device__ inline int unit_csplits(const DecodeArgs& a, int unit) {
    if (a.flat_ctas <= 0) return a.n_csplit + (unit < a.n_extra ? 1 : 0);
    const uint32_t nblk = a.p.comp_len / kBlockTokens;
    return static_cast<int>(flat_owner(a, (unit + 1) * nblk - 1) - flat_owner(a, unit * nblk)) + 1;
}

// Window length of this launch: from device memory when the caller runs static (graph-replayed) steps, else the host's value.
// Only to be called after the PDL wait (the counter is advanced by a launch earlier in the stream).
__device__ __forceinline__ int cur_win_len(const DecodeArgs& a) {
    return a.p.win_len_dev != nullptr ? static_cast<int>(ld_relaxed_u32_fwd(a.p.win_len_dev)) : a.p.win_len;
}
// Optional fused rotary embedding (mfb200_decode_params::rope_cos): element c of a 128-half q / new-K row becomes
// x[c]*cos[c] + rotate_half(x)[c]*sin[c], rotate_half(x) = (-x[64..127], x[0..63]).  Products and sum are each rounded to fp16
// (the _rn intrinsics also forbid FMA contraction): bit-identical to transformers' apply_rotary_pos_emb on fp16 tensors, whose
// elementwise kernels compute in fp32 and round once per op (fp32 holds the product of two halves exactly).
struct RopeRows {
    const __half* cs;  // nullptr = off
    const __half* sn;
};
__device__ __forceinline__ RopeRows rope_rows(const mfb200_decode_params& p, int unit) {
    if (p.rope_cos == nullptr) return RopeRows{nullptr, nullptr};
    const int64_t off = static_cast<int64_t>(unit / p.kv_heads) * p.rope_stride;
    return RopeRows{static_cast<const __half*>(p.rope_cos) + off, static_cast<const __half*>(p.rope_sin) + off};
}
__device__ __forceinline__ __half rope_value(__half x, __half partner, __half cs, __half sn, bool low_half) {
    return __hadd_rn(__hmul_rn(x, cs), __hmul_rn(low_half ? __hneg(partner) : partner, sn));
}
__device__ __forceinline__ __half rope_elem(const RopeRows& r, const __half* row, int c) {
    const __half x = row[c];
    return r.cs == nullptr ? x : rope_value(x, row[c ^ 64], r.cs[c], r.sn[c], c < 64);
}

__device__ __forceinline__ int live_wchunks(const DecodeArgs& a) { return (cur_win_len(a) + kWinTokensPerSplit - 1) / kWinTokensPerSplit; }

// ---- split merge ---------------------------------------------------------------------------------------------
// Every contributor of a unit writes its fp32 partial (o[G][128], m, l) into its slot as 8-byte entries
// {value, tag}; tag = the unit's launch epoch + 1 (read from the workspace after the PDL wait).  Two protocols:
//  * ticket (long launches): stores -> CTA barrier -> acq_rel atomic ticket; the LAST ARRIVAL merges.  Merges are
//    spread over the launch and nobody ever waits, at the price of two dependent L2 round trips (release, ticket)
//    before the merge can start.
//  * flagged (short launches, where that tail is ~8 % of the launch): contributors store and leave at once - no
//    fence, no ticket, no barrier.  The owner of the unit's LAST slot (the last window chunk, or the last compressed
//    split when there is no window) merges: its own partial stays in shared memory, the others are folded straight
//    from L2, an entry counting only when its tag matches; a pass that met a stale entry is repeated after a short
//    sleep.  The merger has the highest CTA index of its unit, so everything it waits for was dispatched before
//    it: the wait cannot deadlock.
__device__ __forceinline__ uint4 ld_relaxed_v4(const void* p) {
    uint4 r;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const void* p) {
    uint32_t r;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
    return r;
}
// workspace: [epochs: units u32 | 256-byte rounded][tickets: units s32 | 256-byte rounded][partials]
__host__ __device__ inline size_t ws_counter_bytes(size_t units) { return (units * 4 + 255) & ~static_cast<size_t>(255); }
__device__ __forceinline__ uint32_t* epoch_ptr(const mfb200_decode_params& p, int unit) {
    return static_cast<uint32_t*>(p.workspace) + unit;
}
__device__ __forceinline__ int* ticket_ptr(const mfb200_decode_params& p, int unit) {
    return reinterpret_cast<int*>(static_cast<uint8_t*>(p.workspace) + ws_counter_bytes(p.batch * p.kv_heads)) + unit;
}
__device__ __forceinline__ uint2* partial_ptr(const DecodeArgs& a, int G, int unit, int slot) {
    uint2* parts = reinterpret_cast<uint2*>(static_cast<uint8_t*>(a.p.workspace) + 2 * ws_counter_bytes(a.p.batch * a.p.kv_heads));
    return parts + (static_cast<int64_t>(unit) * a.max_split + slot) * G * kPartStride;
}

// Folds the unit's n_split partials and writes the output rows.  kOwn: slot `split` is taken from shared memory
// (ored, m, l) and the others are accepted only with a matching tag (flagged protocol); otherwise every slot is read
// from the workspace and trusted (ticket protocol: the acquire already ordered them).
// One global round trip per pass: each warp folds a strided subset of one head's partials with all its loads in
// flight (128-bit, L2), the per-warp results are combined through `scratch`.
template <int G, bool kOwn>
__device__ __forceinline__ void merge_unit(const DecodeArgs& a, int unit, int split, int n_split, uint32_t tag,
                                           const float* ored, const float* m, const float* l, float* scratch) {
    const mfb200_decode_params& p = a.p;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int kWpg = (G <= 2 ? kAttnWarps : 8) / G;  // warps per query head: 10, 5, 2, 1
    float (*s_red)[kPartStride] = reinterpret_cast<float (*)[kPartStride]>(scratch);  // per warp: o[128], m, den
    const uint2* up = partial_ptr(a, G, unit, 0);
    for (uint32_t backoff = 128;; backoff = backoff < 1024 ? 2 * backoff : 1024) {
        bool ok = true;
        if (warp < kWpg * G) {
            const int g = warp % G;
            float m_run = -INFINITY, den = 0.f;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int s = warp / G; s < n_split; s += kWpg) {
                float ms, ls;
                float4 o;
                if (kOwn && s == split) {
                    ms = m[g];
                    ls = l[g];
                    o = reinterpret_cast<const float4*>(ored + g * 128)[lane];
                } else {
                    const uint2* ps = up + (s * G + g) * kPartStride;
                    const uint4 e0 = ld_relaxed_v4(ps + 4 * lane), e1 = ld_relaxed_v4(ps + 4 * lane + 2);
                    const uint4 ml = ld_relaxed_v4(ps + 128);  // {m, tag, l, tag}
                    if (kOwn) ok = ok && e0.y == tag && e0.w == tag && e1.y == tag && e1.w == tag && ml.y == tag && ml.w == tag;
                    ms = __uint_as_float(ml.x);
                    ls = __uint_as_float(ml.z);
                    o = make_float4(__uint_as_float(e0.x), __uint_as_float(e0.z), __uint_as_float(e1.x), __uint_as_float(e1.z));
                }
                const float mn = fmaxf(m_run, ms);
                // a partial with no visible token has m = -inf, l = 0, o = 0: weight it 0 (and avoid inf - inf)
                const float so = (m_run == -INFINITY) ? 0.f : exp2f((m_run - mn) * kLog2e);
                const float w = (ms == -INFINITY) ? 0.f : exp2f((ms - mn) * kLog2e);
                den = den * so + ls * w;
                acc.x = acc.x * so + o.x * w;
                acc.y = acc.y * so + o.y * w;
                acc.z = acc.z * so + o.z * w;
                acc.w = acc.w * so + o.w * w;
                m_run = mn;
            }
            reinterpret_cast<float4*>(s_red[warp])[lane] = acc;
            if (lane == 0) {
                s_red[warp][128] = m_run;
                s_red[warp][129] = den;
            }
        }
        if (__syncthreads_and(ok)) break;  // also publishes s_red
        __nanosleep(backoff);              // some contributor has not landed yet (stale tags): look again
    }
    for (int i = tid; i < G * 128; i += blockDim.x) {
        const int g = i >> 7, c = i & 127;
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < kWpg; ++j) mx = fmaxf(mx, s_red[j * G + g][128]);
        float den = 0.f, num = 0.f;
#pragma unroll
        for (int j = 0; j < kWpg; ++j) {
            const float mj = s_red[j * G + g][128];
            const float w = (mj == -INFINITY) ? 0.f : exp2f((mj - mx) * kLog2e);
            den += s_red[j * G + g][129] * w;
            num += s_red[j * G + g][c] * w;
        }
        const int64_t qh = static_cast<int64_t>(unit) * G + g;  // = b*Hq + h*G + g
        const __half hv = __float2half_rn(num / den);
        static_cast<__half*>(p.out)[qh * kHeadDim + c] = hv;
        if (a.peer.n_peers > 1) {
            // head-sharded decode: the same row goes straight into every rank's gathered output (P2P stores over NVLink)
            const int64_t row = static_cast<int64_t>(unit / p.kv_heads) * a.peer.rows_total + a.peer.row0 + (unit % p.kv_heads) * G + g;
            for (int r = 0; r < a.peer.n_peers; ++r) static_cast<__half*>(a.peer.out[r])[row * kHeadDim + c] = hv;
        }
    }
    if (a.peer.n_peers > 1) {
        // arrival flag of (this rank, this unit) on every rank: the CTA's row stores happen-before the barrier, thread 0's
        // system-scope fence makes them visible to whoever acquires the flag (mfb200_peer_wait on the consumer side)
        __syncthreads();
        if (tid == 0) {
            __threadfence_system();
            const int64_t slot = static_cast<int64_t>(a.peer.rank) * p.batch * p.kv_heads + unit;
            for (int r = 0; r < a.peer.n_peers; ++r)
                asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(a.peer.flags[r] + slot), "r"(a.peer.epoch) : "memory");
        }
    }
    if (tid == 0) *epoch_ptr(p, unit) = tag;  // the next launch (stream-ordered / after its PDL wait) tags with tag + 1
}

// Publishes this split's partial (slot `split` of the unit) and merges the unit if it is this CTA's turn.
// ored [G][128], m [G], l [G] live in shared memory.  All threads of the CTA call it.
// `scratch`: kAttnWarps * kPartStride floats of dead dynamic shared memory, must not overlap ored / m / l.
template <int G, bool FLAGGED>
__device__ __forceinline__ void publish_or_merge(const DecodeArgs& a, int unit, int split, int n_split, uint32_t tag,
                                                 const float* ored, const float* m, const float* l, float* scratch) {
    const int tid = threadIdx.x;
    // callers count the window chunks the launch was PLANNED for; with a device-side window length fewer may be live
    if (a.p.win_len_dev != nullptr) n_split += live_wchunks(a) - a.n_wsplit;
    if (FLAGGED && split == n_split - 1) {
        if (tid == 0) {
            MFB_TRACE_AT(9);
            MFB_TRACE_VAL(13, 1ull);
        }
        merge_unit<G, true>(a, unit, split, n_split, tag, ored, m, l, scratch);
        return;
    }
    uint2* mine = partial_ptr(a, G, unit, split);
    for (int i = tid; i < G * 128; i += blockDim.x)
        mine[(i >> 7) * kPartStride + (i & 127)] = make_uint2(__float_as_uint(ored[i]), tag);
    if (tid < G) {
        mine[tid * kPartStride + 128] = make_uint2(__float_as_uint(m[tid]), tag);
        mine[tid * kPartStride + 129] = make_uint2(__float_as_uint(l[tid]), tag);
    }
    int last = 0;
    if constexpr (!FLAGGED) {
        // Release/acquire through the ticket: the CTA's partial stores happen-before thread 0's release (bar.sync),
        // and thread 0's acquire happens-before the merge loads of the whole CTA (bar.sync) - no full fences needed.
        __shared__ int s_last;
        __syncthreads();
        if (tid == 0) {
            int ticket;
            asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(ticket) : "l"(ticket_ptr(a.p, unit)) : "memory");
            s_last = (ticket == n_split - 1);
        }
        __syncthreads();
        last = s_last;
    }
    if (tid == 0) {
        MFB_TRACE_AT(9);
        MFB_TRACE_VAL(13, static_cast<unsigned long long>(last));
    }
    if constexpr (!FLAGGED) {
        if (!last) return;
        merge_unit<G, false>(a, unit, split, n_split, tag, ored, m, l, scratch);
        if (tid == 0) *ticket_ptr(a.p, unit) = 0;  // ready for the next launch / graph replay
    }
}

// One chunk of 16 tiles.  `oper` = the per-tile fp16 FMA operands [16][G] in shared memory (q for K
// tiles, p for V tiles; 16-byte aligned): pulled into registers with 128-bit loads for G <= 4, read
// with one broadcast 128-bit load per tile for G = 8.
template <int G, bool NZ_SHARED>
__device__ __forceinline__ void tiles_chunk(const uint2* rec, const LaneConst& lc, const uint8_t* gbase,
                                            const __half* oper, float (&acc)[G][2]) {
    if constexpr (G <= 4) {
        constexpr int kVec = kChunk * G / 8;  // uint4 loads
        uint4 wv[kVec];
#pragma unroll
        for (int i = 0; i < kVec; ++i) wv[i] = reinterpret_cast<const uint4*>(oper)[i];
        const uint32_t* w32 = reinterpret_cast<const uint32_t*>(wv);
#pragma unroll
        for (int j = 0; j < kChunk; ++j) {
            const DecodedPair d = decode_pair<NZ_SHARED>(rec + 2 * j, lc, gbase);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int e = j * G + g;
                const uint16_t w = static_cast<uint16_t>((e & 1) ? (w32[e >> 1] >> 16) : (w32[e >> 1] & 0xffffu));
                if (d.b0) acc[g][0] = fhfma(d.x, w, acc[g][0]);
                if (d.b1) acc[g][1] = fhfma(d.y, w, acc[g][1]);
            }
        }
    } else {
#pragma unroll 4
        for (int j = 0; j < kChunk; ++j) {
            const DecodedPair d = decode_pair<NZ_SHARED>(rec + 2 * j, lc, gbase);
            const uint4 wv = reinterpret_cast<const uint4*>(oper)[j];
            const uint32_t w32[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const uint16_t w = static_cast<uint16_t>((g & 1) ? (w32[g >> 1] >> 16) : (w32[g >> 1] & 0xffffu));
                if (d.b0) acc[g][0] = fhfma(d.x, w, acc[g][0]);
                if (d.b1) acc[g][1] = fhfma(d.y, w, acc[g][1]);
            }
        }
    }
}

template <int G>
__device__ __forceinline__ void tiles32(bool nz_shared, const uint2* rec, const LaneConst& lc, const uint8_t* gbase,
                                        const __half* oper, float (&acc)[G][2]) {
    if (nz_shared) {
        tiles_chunk<G, true>(rec, lc, gbase, oper, acc);
        tiles_chunk<G, true>(rec + 2 * kChunk, lc, gbase, oper + kChunk * G, acc);
    } else {
        tiles_chunk<G, false>(rec, lc, gbase, oper, acc);
        tiles_chunk<G, false>(rec + 2 * kChunk, lc, gbase, oper + kChunk * G, acc);
    }
}

// ------------------------------------------------------------------------------------------------
template <int G, bool FLAGGED>
__device__ __forceinline__ void compressed_split(const DecodeArgs& a, uint8_t* smem, int unit, int split, int n_split,
                                                 int blk0, int blk1, bool again) {
    const mfb200_decode_params& p = a.p;
    const int D = a.depth;
    const SmemMap sm = smem_map(G, a.slot_nz_bytes, D);
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t lane = lane_id();
    const int nb = blk1 - blk0;
    const int b = unit / p.kv_heads;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sm.bars);
    __half* qs = reinterpret_cast<__half*>(smem + sm.qs);      // [c][G]
    float* spart = reinterpret_cast<float*>(smem + sm.spart);  // [2][4][G][64]
    __half* ps = reinterpret_cast<__half*>(smem + sm.ps);      // [2][64][G]
    float* corr = reinterpret_cast<float*>(smem + sm.corr);    // [2][8]
    float* stat = reinterpret_cast<float*>(smem + sm.stat);    // m[8], l[8]
    uint32_t* segk = reinterpret_cast<uint32_t*>(smem + sm.segk);
    uint32_t* segv = reinterpret_cast<uint32_t*>(smem + sm.segv);
    const uint32_t slot_bytes = 1024u + a.slot_nz_bytes;

    // ---- prologue ------------------------------------------------------------------------------------
    // With MFB200_F_PDL_EARLY_KV the compressed cache (idx, bitmaps, nonzeros) is known not to be written
    // by the kernel that precedes this one in the stream, so the segment offsets and the first ring-full of
    // blocks are requested BEFORE waiting for that kernel: this launch's fetch latency overlaps the
    // previous launch's tail.  q / k_new / window / workspace are only touched after the wait.
    const bool early_kv = (p.flags & MFB200_F_PDL_EARLY_KV) != 0;
    pdl_launch_dependents();
    if (!early_kv) pdl_wait_prior_grids();
    if (again) __syncthreads();  // flat mode, second segment of this CTA: everyone has left the previous one
    {
        const uint32_t* ki = p.k_idx + static_cast<int64_t>(unit) * p.idx_stride + static_cast<int64_t>(blk0) * 128;
        const uint32_t* vi = p.v_idx + static_cast<int64_t>(unit) * p.idx_stride + static_cast<int64_t>(blk0) * 128;
        for (int i = tid; i <= nb * 4; i += blockDim.x) {
            segk[i] = __ldg(ki + i * 32);
            segv[i] = __ldg(vi + i * 32);
        }
        if (tid == 0) {
            if (again) {  // re-initialising a live mbarrier is undefined: invalidate the previous segment's first
#pragma unroll
                for (int i = 0; i < Bars::kCount; ++i) mbar_inval(&bars[i]);
            }
            for (int s = 0; s < kMaxDepth; ++s) {
                mbar_init(&bars[Bars::kFullK + s], 1);
                mbar_init(&bars[Bars::kFullV + s], 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&bars[Bars::kScEmpty + s], 1);
                mbar_init(&bars[Bars::kPFull + s], 1);
            }
            fence_mbar_init();
        }
    }
    __syncthreads();
    if (tid == 0) MFB_TRACE_AT(1);
    const uint8_t* k_nz = static_cast<const uint8_t*>(p.k_nz) + p.k_nz_off[unit] * 16;
    const uint8_t* v_nz = static_cast<const uint8_t*>(p.v_nz) + p.v_nz_off[unit] * 16;
    const uint64_t* k_bmp = p.k_bmp + static_cast<int64_t>(unit) * p.bmp_stride + static_cast<int64_t>(blk0) * 128;
    const uint64_t* v_bmp = p.v_bmp + static_cast<int64_t>(unit) * p.bmp_stride + static_cast<int64_t>(blk0) * 128;
    // producer state: block pn -> ring slot ps_slot.  The whole producer warp runs this (bar.sync is warp-wide);
    // lane 0 issues the copies.  A slot's first use needs no wait; later uses wait for its 4 consumer warps.
    int pn = 0, ps_slot = 0;
    auto produce_until = [&](int limit) {
        for (; pn < limit; ++pn, ++ps_slot) {
            if (ps_slot == D) ps_slot = 0;
#pragma unroll
            for (int is_v = 0; is_v < 2; ++is_v) {
                if (pn >= D) bar_sync((is_v ? NamedBars::kEmptyV : NamedBars::kEmptyK) + ps_slot, kHandoffThreads);
                if (lane == 0) {
                    uint64_t* full = &bars[(is_v ? Bars::kFullV : Bars::kFullK) + ps_slot];
                    const uint32_t* seg = is_v ? segv : segk;
                    const uint32_t off0 = seg[pn * 4], bytes = (seg[pn * 4 + 4] - off0) * 4u;
                    const bool fits = bytes <= static_cast<uint32_t>(a.slot_nz_bytes);
                    uint8_t* dst = smem + (is_v ? sm.slots_v : sm.slots_k) + ps_slot * slot_bytes;
                    mbar_expect_tx(full, 1024u + ((fits && bytes) ? bytes : 0u));
                    bulk_g2s(dst, (is_v ? v_bmp : k_bmp) + pn * 128, 1024u, full);
                    if (fits && bytes) bulk_g2s(dst + 1024, (is_v ? v_nz : k_nz) + static_cast<uint64_t>(off0) * 4u, bytes, full);
                }
                __syncwarp();
            }
        }
    };
    if (warp == kWarpProducer) produce_until(nb < D ? nb : D);  // first ring-full: never blocks
    if (early_kv) pdl_wait_prior_grids();
    if (tid == 0) {
        MFB_TRACE_AT(2);
        reinterpret_cast<uint32_t*>(stat)[16] = ld_relaxed_u32(epoch_ptr(p, unit)) + 1u;  // this launch's partial tag
    }
    {
        const __half* q = static_cast<const __half*>(p.q) + static_cast<int64_t>(unit) * G * kHeadDim;
        const RopeRows rp = rope_rows(p, unit);
        if constexpr (G >= 4) {
            constexpr int BH = oper_block_bytes(G) / 2;  // halves per operand block
            for (int i = tid; i < G * kHeadDim; i += blockDim.x)
                qs[((i & 127) >> 2) * BH + (i >> 7) * 4 + (i & 3)] = rope_elem(rp, q + (i & ~127), i & 127);
            // the zero rows of the q blocks and of both p buffers (64 + 32 blocks, 8 bytes each)
            for (int i = tid; i < 32 + 2 * 16; i += blockDim.x)
                *reinterpret_cast<uint2*>((i < 32 ? qs + i * BH : ps + (i - 32) * BH) + G * 4) = make_uint2(0u, 0u);
        } else {
            for (int i = tid; i < G * kHeadDim; i += blockDim.x) qs[(i & 127) * G + (i >> 7)] = rope_elem(rp, q + (i & ~127), i & 127);
        }
    }
    __syncthreads();  // q staged (threads of every warp contribute) before the K warps read it
    if (tid == 0) {
        MFB_TRACE_AT(3);
        MFB_TRACE_VAL(12, static_cast<unsigned long long>(nb));
    }
    float o_acc[G][2];
#pragma unroll
    for (int g = 0; g < G; ++g) o_acc[g][0] = o_acc[g][1] = 0.f;
    [[maybe_unused]] float gq_o[G >= 4 ? G / 2 : 1][4];  // G >= 4: the V warps' accumulator fragments (see gqa_mma.cuh for the layout)
#pragma unroll
    for (int m = 0; m < (G >= 4 ? G / 2 : 1); ++m) gq_o[m][0] = gq_o[m][1] = gq_o[m][2] = gq_o[m][3] = 0.f;

    if (warp == kWarpProducer) {
        // =========================== producer ===========================
        produce_until(nb);
        // the arrivals of the last ring-full have no refill that would consume them: drain, so that every named
        // barrier is back in its initial state when the split ends (a flat-mode CTA runs another segment)
        for (int j = nb > D ? nb - D : 0; j < nb; ++j) {
            bar_sync(NamedBars::kEmptyK + j % D, kHandoffThreads);
            bar_sync(NamedBars::kEmptyV + j % D, kHandoffThreads);
        }
    } else if (warp == kWarpSoftmax) {
        // =========================== online softmax ===========================
        // Head-parallel lane mapping: 32/G lanes per query head, each lane owns 2G consecutive tokens of the
        // block, so all G heads reduce at once with log2(32/G) shuffle steps (the result gates the V warps).
        // G >= 4: a lane's 2G tokens are chunks of 4 (chunk c = tokens 4*li + 4*LPH*c ..+3): a chunk is one 16-byte read of
        // the partial scores (conflict-free across the lanes of a head) and one 8-byte operand-block row of p.
        constexpr int LPH = 32 / G, TPL = 2 * G;
        const int hg = lane / LPH, li = lane % LPH, t0 = li * TPL;
        auto tok = [&](int i) { return G >= 4 ? 4 * li + 4 * LPH * (i >> 2) + (i & 3) : t0 + i; };
        const bool ref_round = (p.flags & MFB200_F_REF_SCORE_ROUNDING) != 0;
        const __half* mask = p.mask ? static_cast<const __half*>(p.mask) + static_cast<int64_t>(b) * p.mask_stride : nullptr;
        float m_run = -INFINITY, l_run = 0.f;
        for (int n = 0; n < nb; ++n) {
            const int buf = n & 1;
            float mk[TPL];
#pragma unroll
            for (int i = 0; i < TPL; ++i) mk[i] = mask ? __half2float(mask[(blk0 + n) * kBlockTokens + tok(i)]) : 0.f;
            bar_sync(NamedBars::kScFull + buf, kHandoffThreads);  // the 4 K warps' partial scores of block n
            float sc[TPL];
#pragma unroll
            for (int i = 0; i < TPL; ++i) sc[i] = 0.f;
            constexpr int SPP = G >= 4 ? kTcRowPitch : 64;  // row pitch of the partial-score buffer
            const float* sp = spart + buf * kTileWarps * G * SPP + hg * SPP;
#pragma unroll
            for (int w = 0; w < kTileWarps; ++w) {
                if constexpr (G >= 4) {
#pragma unroll
#if MFB_SCORE_PITCH % 4 == 0
                    for (int i = 0; i < TPL; i += 4) {
                        const float4 t = *reinterpret_cast<const float4*>(sp + w * G * SPP + tok(i));
                        sc[i] += t.x;
                        sc[i + 1] += t.y;
                        sc[i + 2] += t.z;
                        sc[i + 3] += t.w;
                    }
#else
                    for (int i = 0; i < TPL; i += 2) {
                        const float2 t = *reinterpret_cast<const float2*>(sp + w * G * SPP + tok(i));
                        sc[i] += t.x;
                        sc[i + 1] += t.y;
                    }
#endif
                } else {
#pragma unroll
                    for (int i = 0; i < TPL; i += 2) {
                        const float2 t = *reinterpret_cast<const float2*>(sp + w * G * SPP + t0 + i);
                        sc[i] += t.x;
                        sc[i + 1] += t.y;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[Bars::kScEmpty + buf]);
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < TPL; ++i) {
                sc[i] = ref_round_score(sc[i], p.score_div, ref_round);
                if (mask) sc[i] = fmaxf(sc[i] + mk[i], -65504.f);
                mx = fmaxf(mx, sc[i]);
            }
#pragma unroll
            for (int o = LPH / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float m_new = fmaxf(m_run, mx);
            const float cr = exp2f((m_run - m_new) * kLog2e);  // exp2(-inf) = 0 on the first block
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < TPL; ++i) {
                sc[i] = exp2f((sc[i] - m_new) * kLog2e);
                sum += sc[i];
            }
#pragma unroll
            for (int o = LPH / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            l_run = l_run * cr + sum;
            m_run = m_new;
            if (n >= 2) bar_sync(NamedBars::kPEmpty + buf, kHandoffThreads);  // V warps are done with p of block n-2
            if constexpr (G >= 4) {  // p as operand blocks: block = 4 tokens, row hg = this head's 4 probabilities
                constexpr int BH = oper_block_bytes(G) / 2;
#pragma unroll
                for (int i = 0; i < TPL; i += 4) {
                    const __half2 lo = __floats2half2_rn(sc[i], sc[i + 1]), hi = __floats2half2_rn(sc[i + 2], sc[i + 3]);
                    *reinterpret_cast<uint2*>(ps + (buf * 16 + (tok(i) >> 2)) * BH + hg * 4) =
                        make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
                }
            } else {
                __half* pb = ps + buf * 64 * G + t0 * G + hg;
#pragma unroll
                for (int i = 0; i < TPL; ++i) pb[i * G] = __float2half_rn(sc[i]);
            }
            if (li == 0) corr[buf * 8 + hg] = cr;
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[Bars::kPFull + buf]);
        }
        if (li == 0) {
            stat[hg] = m_run;
            stat[8 + hg] = l_run;
        }
        for (int j = nb > 2 ? nb - 2 : 0; j < nb; ++j) bar_sync(NamedBars::kPEmpty + (j & 1), kHandoffThreads);  // drain
    } else {
        // =========================== K / V tile warps ===========================
        const bool is_v = warp >= kWarpV0;
        const int w = warp & 3;
        const LaneConst lc = make_lane_const();
        uint2* rec = reinterpret_cast<uint2*>(smem + sm.rec) + warp * 64;
        const uint2* my_rec = rec + lc.half;
        const uint32_t* seg = is_v ? segv : segk;
        const uint8_t* nz_g = is_v ? v_nz : k_nz;
        const uint32_t slots_off = is_v ? sm.slots_v : sm.slots_k;
        const int full0 = is_v ? Bars::kFullV : Bars::kFullK, empty0 = is_v ? NamedBars::kEmptyV : NamedBars::kEmptyK;
        // G >= 4: per-lane constants of the block-diagonal HMMA scheme (gqa_mma.cuh)
        [[maybe_unused]] GqaLane<(G >= 4 ? G : 4)> gl;
        if constexpr (G >= 4) gl = make_gqa_lane<G>();
        // byte offset of this lane's B fragment inside an operand block, per MMA: its head's row where it is live, else the zero row
        [[maybe_unused]] uint32_t oper_row[G >= 4 ? G / 2 : 1];
        if constexpr (G >= 4) {
#pragma unroll
            for (int m = 0; m < G / 2; ++m) oper_row[m] = 8u * (gl.live_m == static_cast<uint32_t>(m) ? gl.g_live : static_cast<uint32_t>(G));
        }
        int s = 0;
        uint32_t par = 0;
        // (trace build) 0 wait TMA, 1 records, 2 wait softmax (V: p ready; K: score buffer free), 3 tiles + hand-off.
        // Caveat: BAR.SYNC is DEFER_BLOCKING - a clock read right after it issues before the warp blocks, the wait is
        // charged to the next phase; mbarrier waits are charged correctly.
        MFB_TACC_INIT(4);
        for (int n = 0; n < nb; ++n, ++s) {
            if (s == D) {
                s = 0;
                par ^= 1;
            }
            const int buf = n & 1;
            const uint32_t par2 = (n >> 1) & 1;
            const uint8_t* sl = smem + slots_off + s * slot_bytes;
            const uint32_t sg0 = seg[n * 4];
            const bool fits = (seg[n * 4 + 4] - sg0) * 4u <= static_cast<uint32_t>(a.slot_nz_bytes);
            const uint8_t* gblk = nz_g + static_cast<uint64_t>(sg0) * 4u;
            const uint32_t nz_addr = (fits ? smem_u32(sl + 1024) : 0u) + (seg[n * 4 + w] - sg0) * 4u;
            MFB_TACC(3);
            mbar_wait(&bars[full0 + s], par);
            MFB_TACC(0);
            build_records(reinterpret_cast<const uint64_t*>(sl) + w * 32, nz_addr, rec);
            __syncwarp();
            MFB_TACC(1);
            if constexpr (G >= 4) {
                // ---------- G >= 4: register-fragment HMMA, all G heads from one decode (gqa_mma.cuh) ----------
                constexpr int NM = G / 2;
                if (!is_v) {
                    // partial scores over channels 32w..32w+31: operand = q[g][channel]
                    float acc[NM][4];
#pragma unroll
                    for (int m = 0; m < NM; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
                    uint32_t oper[NM];
#pragma unroll
                    for (int m = 0; m < NM; ++m)
                        oper[m] = smem_u32(qs) + 8 * w * oper_block_bytes(G) + oper_row[m];
                    if (fits) tiles32_mma<G, true>(my_rec, lc, gblk, oper, acc);
                    else tiles32_mma<G, false>(my_rec, lc, gblk, oper, acc);
                    __syncwarp();
                    bar_arrive(empty0 + s, kHandoffThreads);  // ring slot no longer needed
                    MFB_TACC(3);
                    mbar_wait(&bars[Bars::kScEmpty + buf], par2 ^ 1);
                    MFB_TACC(2);
                    float* sp = spart + ((buf * kTileWarps + w) * G + gl.g0) * kTcRowPitch;
#pragma unroll
                    for (int m = 0; m < NM; ++m) {
                        *reinterpret_cast<float2*>(sp + gl.pos[m]) = make_float2(acc[m][0], acc[m][2]);
                        *reinterpret_cast<float2*>(sp + kTcRowPitch + gl.pos[m]) = make_float2(acc[m][1], acc[m][3]);
                    }
                    __syncwarp();
                    bar_arrive(NamedBars::kScFull + buf, kHandoffThreads);
                    if (tid == 0 && n == 0) MFB_TRACE_AT(4);
                    if (tid == 0 && n == nb - 1) MFB_TRACE_AT(5);
                } else {
                    // out[g][channel] += p[g][token] * V over tokens 32(w&1)..+31: operand = p[g][token]
                    mbar_wait(&bars[Bars::kPFull + buf], par2);
                    MFB_TACC(2);
                    const float c_lo = corr[buf * 8 + gl.g0], c_hi = corr[buf * 8 + gl.g0 + 1];
#pragma unroll
                    for (int m = 0; m < NM; ++m) {
                        gq_o[m][0] *= c_lo;
                        gq_o[m][1] *= c_hi;
                        gq_o[m][2] *= c_lo;
                        gq_o[m][3] *= c_hi;
                    }
                    uint32_t oper[NM];
#pragma unroll
                    for (int m = 0; m < NM; ++m)
                        oper[m] = smem_u32(ps) + (buf * 16 + 8 * (w & 1)) * oper_block_bytes(G) + oper_row[m];
                    if (fits) tiles32_mma<G, true>(my_rec, lc, gblk, oper, gq_o);
                    else tiles32_mma<G, false>(my_rec, lc, gblk, oper, gq_o);
                    __syncwarp();
                    bar_arrive(empty0 + s, kHandoffThreads);
                    bar_arrive(NamedBars::kPEmpty + buf, kHandoffThreads);
                    if (tid == kWarpV0 * 32 && n == 0) MFB_TRACE_AT(6);
                    if (tid == kWarpV0 * 32 && n == nb - 1) MFB_TRACE_AT(7);
                }
            } else if (!is_v) {
                // K item: tiles = channels 32w .. 32w+31 -> partial scores of tokens (2*lane, 2*lane+1)
                float sc[G][2];
#pragma unroll
                for (int g = 0; g < G; ++g) sc[g][0] = sc[g][1] = 0.f;
                tiles32<G>(fits, my_rec, lc, gblk, qs + (32 * w) * G, sc);
                __syncwarp();
                bar_arrive(empty0 + s, kHandoffThreads);
                MFB_TACC(3);
                mbar_wait(&bars[Bars::kScEmpty + buf], par2 ^ 1);
                MFB_TACC(2);
                float* sp = spart + ((buf * kTileWarps + w) * G) * 64;
#pragma unroll
                for (int g = 0; g < G; ++g) *reinterpret_cast<float2*>(sp + g * 64 + 2 * lane) = make_float2(sc[g][0], sc[g][1]);
                __syncwarp();
                bar_arrive(NamedBars::kScFull + buf, kHandoffThreads);
                if (tid == 0 && n == 0) MFB_TRACE_AT(4);
                if (tid == 0 && n == nb - 1) MFB_TRACE_AT(5);
            } else {
                // V item: tiles 32w .. 32w+31 = channel half (w>>1), tokens 32*(w&1)+j -> channels (2*lane, 2*lane+1)
                mbar_wait(&bars[Bars::kPFull + buf], par2);
                MFB_TACC(2);
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const float c = corr[buf * 8 + g];
                    o_acc[g][0] *= c;
                    o_acc[g][1] *= c;
                }
                tiles32<G>(fits, my_rec, lc, gblk, ps + (buf * 64 + 32 * (w & 1)) * G, o_acc);
                __syncwarp();
                bar_arrive(empty0 + s, kHandoffThreads);
                bar_arrive(NamedBars::kPEmpty + buf, kHandoffThreads);
                if (tid == kWarpV0 * 32 && n == 0) MFB_TRACE_AT(6);
                if (tid == kWarpV0 * 32 && n == nb - 1) MFB_TRACE_AT(7);
            }
        }
        MFB_TACC(3);
        MFB_TACC_STORE(is_v ? 20 : 16, 4, lane == 0 && w == 0);
    }
    // ---- cross-warp reduction of o: V warps (0,1) hold channel half 0, (2,3) half 1 -------------------
    __syncthreads();
    if (tid == 0) MFB_TRACE_AT(8);
    float* red = spart;                                          // [4][G][64]
    float* ored = reinterpret_cast<float*>(smem + sm.slots_k);   // [G][128] (the rings are idle now)
    if (warp >= kWarpV0 && warp < kWarpV0 + kTileWarps) {
        const int w = warp & 3;
        if constexpr (G >= 4) {
            const uint32_t gid = lane >> 2, tig = lane & 3;
#pragma unroll
            for (int m = 0; m < G / 2; ++m) {  // same D-fragment layout as the K side: heads g0, g0+1 of positions pos, pos+1
                const uint32_t g0 = (2 * tig) % G, pos = 8 * gid + 2 * ((8 * m + 2 * tig) / G);
                *reinterpret_cast<float2*>(red + (w * G + g0) * 64 + pos) = make_float2(gq_o[m][0], gq_o[m][2]);
                *reinterpret_cast<float2*>(red + (w * G + g0 + 1) * 64 + pos) = make_float2(gq_o[m][1], gq_o[m][3]);
            }
        } else {
#pragma unroll
            for (int g = 0; g < G; ++g) *reinterpret_cast<float2*>(red + (w * G + g) * 64 + 2 * lane) = make_float2(o_acc[g][0], o_acc[g][1]);
        }
    }
    __syncthreads();
    for (int i = tid; i < G * 128; i += blockDim.x) {
        const int g = i >> 7, c = i & 127, hf = c >> 6, e = c & 63;
        ored[i] = red[((2 * hf) * G + g) * 64 + e] + red[((2 * hf + 1) * G + g) * 64 + e];
    }
    __syncthreads();
    publish_or_merge<G, FLAGGED>(a, unit, split, n_split, reinterpret_cast<const uint32_t*>(stat)[16], ored, stat, stat + 8,
                        reinterpret_cast<float*>(smem + sm.slots_v));
}

// ------------------------------------------------------------------------------------------------
// tcgen05 variant of the compressed split (G >= 4), see gqa_tc.cuh.  Pipeline per 64-token block n:
//   producer --TMA--> ring slot --K decode warps--> Kd[n&1] --MMA-S--> S[n&1] (TMEM) --epilogue: softmax--> p[n&1]
//                               --V decode warps--> Vd      ---------------------------MMA-O(p, Vd)--> O[n&1] (TMEM)
//                                                                                      --epilogue: o = o*corr + O
// The decode warps never wait for the softmax: K runs up to two blocks ahead (two dense K buffers), V one or two.
// Hand-offs: everything a WARP signals goes through hardware named barriers (a warp parked in bar.sync issues nothing;
// mbarrier sleepers are woken by every mbarrier event of the CTA and re-poll - the first version spent 10 of its 29
// instructions per tile polling); what the TENSOR CORE signals (tcgen05.commit) has to be an mbarrier, and of the four
// epilogue warps only warp 0 waits on those, the others follow through the epilogue's own named barrier.
struct TcBars {
    enum : int {
        kFullK = 0, kFullV = kMaxDepth,  // [depth] TMA tx -> decode warps
        kKdEmpty = 2 * kMaxDepth,        // [2] commit(MMA-S) -> K decode warps
        kVdEmpty = kKdEmpty + 2,         // [2] commit(MMA-O) -> V decode warps
        kSFull = kVdEmpty + 2,           // [2] commit(MMA-S) -> epilogue
        kPEmpty = kSFull + 2,            // [2] commit(MMA-O) -> epilogue
        kOFull = kPEmpty + 2,            // [2] commit(MMA-O) -> epilogue
        kSEmpty = kOFull + 2,            // [2] epilogue (1 arrival) -> MMA-S
        kOEmpty = kSEmpty + 2,           // [2] epilogue (1 arrival) -> MMA-O
        kCount = kOEmpty + 2
    };
};
// named barrier ids of the tcgen05 variant (ring depth 2): 0 __syncthreads, 5-6 kEmptyK, 9-10 kEmptyV (NamedBars), and
struct TcNamed {
    enum : int {
        kKdFull = 1,   // [2] K decode warps arrive, MMA-S warp syncs
        kVdFull = 3,   // [2] V decode warps arrive, MMA-O warp syncs
        kPFull = 7,    // [2] epilogue warps arrive, MMA-O warp syncs
        kEpi = 13,     // the four epilogue warps
    };
};
struct TcSmemMap {
    uint32_t slots_k, slots_v, kd, vd, qb, pb, bars, rec, segk, segv, wred, stat, total;
};
__host__ __device__ inline TcSmemMap tc_smem_map(int slot_nz_bytes, int depth, int nvd) {
    TcSmemMap m;
    uint32_t o = 0;
    m.slots_k = o;
    o += depth * (1024 + slot_nz_bytes);
    m.slots_v = o;
    o += depth * (1024 + slot_nz_bytes);
    m.kd = o;
    o += 2 * kDenseBytes;
    m.vd = o;
    o += nvd * kDenseBytes;
    m.qb = o;
    o += kQbBytes;
    m.pb = o;
    o += 2 * kPbBytes;
    m.bars = o;
    o += ((TcBars::kCount * 8 + 15) / 16) * 16;
    m.rec = o;  // uint2 [8 decode warps][32 tiles][2]
    o += 2 * kTileWarps * 64 * 8;
    m.segk = o;
    o += (kMaxBlocksPerSplitTc * 4 + 4) * 4;
    m.segv = o;
    o += (kMaxBlocksPerSplitTc * 4 + 4) * 4;
    m.wred = o;  // float [2][8 heads][4 warps] block maxima + [8][4] row sums
    o += (2 * 8 * 4 + 8 * 4) * 4;
    m.stat = o;  // float m[8], l[8]; uint32 tag
    o += 2 * 8 * 4 + 16;
    m.total = o;
    return m;
}
__device__ __forceinline__ void sts_b16(uint32_t addr, uint16_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}

template <int G, bool FLAGGED>
__device__ __forceinline__ void compressed_split_tc(const DecodeArgs& a, uint8_t* smem, uint32_t tmem, int unit, int split,
                                                    int n_split, int blk0, int blk1, bool again) {
    const mfb200_decode_params& p = a.p;
    const int D = a.depth, NVD = a.nvd;
    const TcSmemMap sm = tc_smem_map(a.slot_nz_bytes, D, NVD);
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t lane = lane_id();
    const int nb = blk1 - blk0;
    const int b = unit / p.kv_heads;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sm.bars);
    float* stat = reinterpret_cast<float*>(smem + sm.stat);
    float* wred = reinterpret_cast<float*>(smem + sm.wred);
    uint32_t* segk = reinterpret_cast<uint32_t*>(smem + sm.segk);
    uint32_t* segv = reinterpret_cast<uint32_t*>(smem + sm.segv);
    const uint32_t slot_bytes = 1024u + a.slot_nz_bytes;
    const uint32_t kd_addr = smem_u32(smem + sm.kd), vd_addr = smem_u32(smem + sm.vd);
    const uint32_t qb_addr = smem_u32(smem + sm.qb), pb_addr = smem_u32(smem + sm.pb);

    // ---- prologue (see compressed_split) ----------------------------------------------------------------
    const bool early_kv = (p.flags & MFB200_F_PDL_EARLY_KV) != 0;
    pdl_launch_dependents();
    if (!early_kv) pdl_wait_prior_grids();
    if (again) __syncthreads();
    {
        const uint32_t* ki = p.k_idx + static_cast<int64_t>(unit) * p.idx_stride + static_cast<int64_t>(blk0) * 128;
        const uint32_t* vi = p.v_idx + static_cast<int64_t>(unit) * p.idx_stride + static_cast<int64_t>(blk0) * 128;
        for (int i = tid; i <= nb * 4; i += blockDim.x) {
            segk[i] = __ldg(ki + i * 32);
            segv[i] = __ldg(vi + i * 32);
        }
        for (int i = tid; i < 2 * kPbBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + sm.pb)[i] = 0u;  // rows >= G stay 0
        if (tid == 0) {
            if (again) {
#pragma unroll
                for (int i = 0; i < TcBars::kCount; ++i) mbar_inval(&bars[i]);
            }
            for (int i = 0; i < TcBars::kCount; ++i) mbar_init(&bars[i], 1);  // every mbarrier has exactly one arrival per phase
            fence_mbar_init();
        }
    }
    __syncthreads();
    if (tid == 0) MFB_TRACE_AT(1);
    const uint8_t* k_nz = static_cast<const uint8_t*>(p.k_nz) + p.k_nz_off[unit] * 16;
    const uint8_t* v_nz = static_cast<const uint8_t*>(p.v_nz) + p.v_nz_off[unit] * 16;
    const uint64_t* k_bmp = p.k_bmp + static_cast<int64_t>(unit) * p.bmp_stride + static_cast<int64_t>(blk0) * 128;
    const uint64_t* v_bmp = p.v_bmp + static_cast<int64_t>(unit) * p.bmp_stride + static_cast<int64_t>(blk0) * 128;
    int pn = 0, ps_slot = 0;
    auto produce_until = [&](int limit) {
        for (; pn < limit; ++pn, ++ps_slot) {
            if (ps_slot == D) ps_slot = 0;
#pragma unroll
            for (int is_v = 0; is_v < 2; ++is_v) {
                if (pn >= D) bar_sync((is_v ? NamedBars::kEmptyV : NamedBars::kEmptyK) + ps_slot, kHandoffThreads);
                if (lane == 0) {
                    uint64_t* full = &bars[(is_v ? TcBars::kFullV : TcBars::kFullK) + ps_slot];
                    const uint32_t* seg = is_v ? segv : segk;
                    const uint32_t off0 = seg[pn * 4], bytes = (seg[pn * 4 + 4] - off0) * 4u;
                    const bool fits = bytes <= static_cast<uint32_t>(a.slot_nz_bytes);
                    uint8_t* dst = smem + (is_v ? sm.slots_v : sm.slots_k) + ps_slot * slot_bytes;
                    mbar_expect_tx(full, 1024u + ((fits && bytes) ? bytes : 0u));
                    bulk_g2s(dst, (is_v ? v_bmp : k_bmp) + pn * 128, 1024u, full);
                    if (fits && bytes) bulk_g2s(dst + 1024, (is_v ? v_nz : k_nz) + static_cast<uint64_t>(off0) * 4u, bytes, full);
                }
                __syncwarp();
            }
        }
    };
    if (warp == kWarpProducerTc) produce_until(nb < D ? nb : D);
    if (early_kv) pdl_wait_prior_grids();
    if (tid == 0) {
        MFB_TRACE_AT(2);
        reinterpret_cast<uint32_t*>(stat)[16] = ld_relaxed_u32(epoch_ptr(p, unit)) + 1u;  // this launch's partial tag
    }
    {   // q as the K-major B operand of the score MMAs: [channel group][row g][8 channels]; rows >= G are zero
        const __half* q = static_cast<const __half*>(p.q) + static_cast<int64_t>(unit) * G * kHeadDim;
        __half* qb = reinterpret_cast<__half*>(smem + sm.qb);
        const RopeRows rp = rope_rows(p, unit);
        for (int i = tid; i < kQbBytes / 2; i += blockDim.x) {
            const int kg = i >> 6, row = (i >> 3) & 7, e = i & 7;
            qb[i] = row < G ? rope_elem(rp, q + row * kHeadDim, kg * 8 + e) : __ushort_as_half(0);
        }
    }
    fence_async_smem();  // qb / the zeroed p buffers were written through the generic proxy, the MMAs read them through the async one
    __syncthreads();
    if (tid == 0) {
        MFB_TRACE_AT(3);
        MFB_TRACE_VAL(12, static_cast<unsigned long long>(nb));
    }
    float* ored = reinterpret_cast<float*>(smem + sm.slots_k);  // [G][128], written when the rings are idle

    if (warp == kWarpProducerTc) {
        // =========================== producer ===========================
        produce_until(nb);
        for (int j = nb > D ? nb - D : 0; j < nb; ++j) {  // drain the last ring-full's arrivals (see compressed_split)
            bar_sync(NamedBars::kEmptyK + j % D, kHandoffThreads);
            bar_sync(NamedBars::kEmptyV + j % D, kHandoffThreads);
        }
    } else if (warp == kWarpMmaS) {
        // =========================== score MMAs: S[n&1] = Kd[n&1] . q^T ===========================
        MFB_TACC_INIT(3);  // 0 wait Kd, 1 wait S free, 2 issue
        for (int n = 0; n < nb; ++n) {
            const int buf = n & 1;
            const uint32_t u = n >> 1;
            MFB_TACC(2);
            bar_sync(TcNamed::kKdFull + buf, kHandoffThreads);  // the 4 K decode warps have stored block n
            MFB_TACC(0);
            if (lane == 0) {
                if (u >= 1) mbar_wait(&bars[TcBars::kSEmpty + buf], (u - 1) & 1);
                MFB_TACC(1);
                tc_fence_after();
                const uint32_t a_lo = umma_desc_lo(kd_addr + buf * kDenseBytes, kLboK), b_lo = umma_desc_lo(qb_addr, 128);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    umma_f16(tmem + 8 * buf, a_lo + ks * (2 * kLboK >> 4), umma_desc_hi(kTcSbo), b_lo + ks * (256 >> 4), umma_desc_hi(128),
                             kIdescS, ks > 0);
                umma_commit(&bars[TcBars::kKdEmpty + buf]);
                umma_commit(&bars[TcBars::kSFull + buf]);
            }
            __syncwarp();
        }
        MFB_TACC(2);
        MFB_TACC_STORE(27, 2, lane == 0);  // slots 27 (wait Kd), 28 (wait S free)
        // every commit of this thread has landed before the split ends (a flat-plan CTA re-initialises the barriers); only
        // the LAST completion of a barrier can be waited for (a parity wait on an older phase of a barrier that has
        // advanced twice since blocks forever)
        if (lane == 0) {
            for (int j = nb > 2 ? nb - 2 : 0; j < nb; ++j) {
                mbar_wait(&bars[TcBars::kKdEmpty + (j & 1)], (j >> 1) & 1);
                mbar_wait(&bars[TcBars::kSFull + (j & 1)], (j >> 1) & 1);
            }
        }
    } else if (warp == kWarpMmaO) {
        // =========================== output MMAs: O[n&1] = Vd . p[n&1]^T ===========================
        MFB_TACC_INIT(4);  // 0 wait Vd, 1 wait p, 2 wait O free, 3 issue
        for (int n = 0; n < nb; ++n) {
            const int buf = n & 1;
            const uint32_t u = n >> 1;
            const int vb = NVD == 2 ? buf : 0;
            MFB_TACC(3);
            bar_sync(TcNamed::kVdFull + vb, kHandoffThreads);  // the 4 V decode warps have stored block n
            MFB_TACC(0);
            bar_sync(TcNamed::kPFull + buf, kHandoffThreads);  // the 4 epilogue warps have stored p of block n
            MFB_TACC(1);
            if (lane == 0) {
                if (u >= 1) mbar_wait(&bars[TcBars::kOEmpty + buf], (u - 1) & 1);
                MFB_TACC(2);
                tc_fence_after();
                const uint32_t a_lo = umma_desc_lo(vd_addr + vb * kDenseBytes, kLboV), b_lo = umma_desc_lo(pb_addr + buf * kPbBytes, 128);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma_f16(tmem + 16 + 16 * buf, a_lo + ks * (2 * kLboV >> 4), umma_desc_hi(kTcSbo), b_lo + ks * (256 >> 4), umma_desc_hi(0),
                             kIdescO, ks > 0);
                umma_commit(&bars[TcBars::kVdEmpty + vb]);
                umma_commit(&bars[TcBars::kPEmpty + buf]);
                umma_commit(&bars[TcBars::kOFull + buf]);
            }
            __syncwarp();
        }
        MFB_TACC(3);
        MFB_TACC_STORE(29, 3, lane == 0);  // slots 29 (wait Vd), 30 (wait p), 31 (wait O free)
        if (lane == 0) {
            for (int j = nb > 2 ? nb - 2 : 0; j < nb; ++j) {
                if (NVD == 2 || j == nb - 1) {
                    const int vb = NVD == 2 ? (j & 1) : 0;
                    const uint32_t uv = NVD == 2 ? (j >> 1) : static_cast<uint32_t>(j);
                    mbar_wait(&bars[TcBars::kVdEmpty + vb], uv & 1);
                }
                mbar_wait(&bars[TcBars::kPEmpty + (j & 1)], (j >> 1) & 1);
                mbar_wait(&bars[TcBars::kOFull + (j & 1)], (j >> 1) & 1);
            }
        }
    } else if (warp >= kWarpEpi0) {
        // =========================== epilogue: online softmax + output accumulation ===========================
        // S rows (tokens) 16e..16e+15 of an M = 64 accumulator live in lanes 0..15 of TMEM lane quadrant e = this warp;
        // O rows (channels) 32e..32e+31 of the M = 128 accumulator in all 32 lanes of the same quadrant.
        // Iteration n: [S(n) and O(n-2) are complete] -> load both -> fold O(n-2) -> softmax(n) -> publish p(n).
        const int ew = warp - kWarpEpi0;
        const bool valid = lane < 16;
        const int t = 16 * ew + static_cast<int>(lane & 15);
        const uint32_t tlane = tmem + (static_cast<uint32_t>(32 * ew) << 16);
        const bool ref_round = (p.flags & MFB200_F_REF_SCORE_ROUNDING) != 0;
        const float rdiv = 1.0f / p.score_div;
        const __half* mask = p.mask ? static_cast<const __half*>(p.mask) + static_cast<int64_t>(b) * p.mask_stride : nullptr;
        float m_run[G], l_part[G], o_acc[G], corr1[G], corr2[G];  // corr1 / corr2: rescale factors of blocks n-1 / n-2
#pragma unroll
        for (int g = 0; g < G; ++g) {
            m_run[g] = -INFINITY;
            l_part[g] = o_acc[g] = 0.f;
            corr1[g] = corr2[g] = 0.f;
        }
        auto fold = [&](int j, const float (&cr)[G]) {  // o = o * corr_j + O_j ; the caller has made sure O_j is complete
            float o[8];
            tmem_ld8(tlane + 16 + 16 * (j & 1), o);
#pragma unroll
            for (int g = 0; g < G; ++g) o_acc[g] = fmaf(o_acc[g], cr[g], o[g]);
        };
        MFB_TACC_INIT(3);  // 0 wait for S / O, 1 load + maxima, 2 second barrier + exp + publish
        for (int n = 0; n < nb; ++n) {
            const int buf = n & 1;
            const uint32_t u = n >> 1;
            const float mk = (mask && valid) ? __half2float(mask[(blk0 + n) * kBlockTokens + t]) : 0.f;
            MFB_TACC(2);
            if (ew == 0) {
                mbar_wait(&bars[TcBars::kSFull + buf], u & 1);
                if (u >= 1) {
                    mbar_wait(&bars[TcBars::kOFull + buf], (u - 1) & 1);   // MMA-O(n-2) done: O[buf] complete ...
                    mbar_wait(&bars[TcBars::kPEmpty + buf], (u - 1) & 1);  // ... and p[buf] free again
                }
            }
            bar_sync(TcNamed::kEpi, 128);
            MFB_TACC(0);
            tc_fence_after();
            float s[8];
            tmem_ld8(tlane + 8 * buf, s);
            if (u >= 1) fold(n - 2, corr2);
            tc_fence_before();
            float x[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                float v = ref_round_score_fast(s[g], p.score_div, rdiv, ref_round);
                if (mask) v = fmaxf(v + mk, -65504.f);
                x[g] = valid ? v * kLog2e : -INFINITY;  // log2 domain from here on (m, the stored partial maxima are converted back)
                float mx = x[g];
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                if (lane == 0) wred[(buf * 8 + g) * 4 + ew] = mx;
            }
            MFB_TACC(1);
            bar_sync(TcNamed::kEpi, 128);  // the four warps' block maxima are in wred; every warp has read S[buf] and O[buf]
            if (ew == 0 && lane == 0) {
                mbar_arrive_cta(&bars[TcBars::kSEmpty + buf]);
                if (u >= 1) mbar_arrive_cta(&bars[TcBars::kOEmpty + buf]);
            }
            uint16_t ph[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float4 w4 = *reinterpret_cast<const float4*>(wred + (buf * 8 + g) * 4);
                const float m_new = fmaxf(m_run[g], fmaxf(fmaxf(w4.x, w4.y), fmaxf(w4.z, w4.w)));
                corr2[g] = corr1[g];
                corr1[g] = ex2_fast(m_run[g] - m_new);  // ex2(-inf) = 0 on the first block
                const float pv = ex2_fast(x[g] - m_new);  // invalid lanes: ex2(-inf) = 0
                l_part[g] = fmaf(l_part[g], corr1[g], pv);
                m_run[g] = m_new;
                ph[g] = __half_as_ushort(__float2half_rn(pv));
            }
            if (valid) {
                const uint32_t pa = pb_addr + buf * kPbBytes + (t >> 3) * 128 + (t & 7) * 2;
#pragma unroll
                for (int g = 0; g < G; ++g) sts_b16(pa + g * 16, ph[g]);
            }
            fence_async_smem();
            __syncwarp();
            bar_arrive(TcNamed::kPFull + buf, kHandoffThreads);
        }
        MFB_TACC(2);
        MFB_TACC_STORE(24, 3, ew == 0 && lane == 0);
        // the last two blocks' outputs: block nb-2 (rescaled by corr of nb-2 ... see below) then nb-1
        for (int j = nb > 2 ? nb - 2 : 0; j < nb; ++j) {
            if (ew == 0) mbar_wait(&bars[TcBars::kOFull + (j & 1)], (j >> 1) & 1);
            bar_sync(TcNamed::kEpi, 128);
            tc_fence_after();
            if (j == nb - 1) fold(j, corr1);
            else fold(j, corr2);
            tc_fence_before();
        }
        // split results: o [G][128], m, l
#pragma unroll
        for (int g = 0; g < G; ++g) {
            ored[g * 128 + 32 * ew + lane] = o_acc[g];
            float v = l_part[g];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) wred[64 + g * 4 + ew] = v;
        }
        bar_sync(TcNamed::kEpi, 128);
        if (ew == 0 && lane == 0) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                stat[g] = m_run[g] * (1.0f / kLog2e);  // back from the log2 domain: partial maxima are merged in natural units
                stat[8 + g] = wred[64 + g * 4] + wred[64 + g * 4 + 1] + wred[64 + g * 4 + 2] + wred[64 + g * 4 + 3];
            }
        }
    } else {
        // =========================== K / V decode warps ===========================
        // (one instantiation per stream: the stream-dependent constants fold away instead of being re-selected per block)
        auto decode_role = [&](auto is_v_tag) {
            constexpr bool is_v = decltype(is_v_tag)::value;
            const int w = warp & 3;
            const LaneConst lc = make_lane_const();
            uint2* rec = reinterpret_cast<uint2*>(smem + sm.rec) + warp * 64;
            const uint2* my_rec = rec + lc.half;
            const uint32_t* seg = is_v ? segv : segk;
            const uint8_t* nz_g = is_v ? v_nz : k_nz;
            const uint32_t slots_off = is_v ? sm.slots_v : sm.slots_k;
            constexpr int full0 = is_v ? TcBars::kFullV : TcBars::kFullK, empty0 = is_v ? NamedBars::kEmptyV : NamedBars::kEmptyK;
            constexpr int dfull0 = is_v ? TcNamed::kVdFull : TcNamed::kKdFull, dempty0 = is_v ? TcBars::kVdEmpty : TcBars::kKdEmpty;
            constexpr uint32_t lbo = is_v ? kLboV : kLboK;
            // destination of tile 0 of this warp for this lane: MN offset of the lane's position pair + the warp's K groups
            //   K: tile = channel 32w + j, positions = tokens     -> K group 4w + (j >> 3), MN = tokens
            //   V: tile = token 32(w&1) + j of channel half w>>1  -> K group 4(w&1) + (j >> 3), MN = channels 64(w>>1) + ...
            const uint32_t lane_off = (lane >> 2) * kTcSbo + (lane & 3) * 4;
            const uint32_t dst0 = is_v ? vd_addr + 8 * (w >> 1) * kTcSbo + 4 * (w & 1) * kLboV + lane_off
                                       : kd_addr + 4 * w * kLboK + lane_off;
            const bool single = is_v && NVD == 1;  // one dense buffer, used by every block
            int s = 0;
            uint32_t par = 0;
            MFB_TACC_INIT(4);  // 0 wait TMA, 1 records, 2 wait dense buffer, 3 decode
            for (int n = 0; n < nb; ++n, ++s) {
                if (s == D) {
                    s = 0;
                    par ^= 1;
                }
                const uint8_t* sl = smem + slots_off + s * slot_bytes;
                const uint32_t sg0 = seg[n * 4];
                const bool fits = (seg[n * 4 + 4] - sg0) * 4u <= static_cast<uint32_t>(a.slot_nz_bytes);
                const uint8_t* gblk = nz_g + static_cast<uint64_t>(sg0) * 4u;
                const uint32_t nz_addr = (fits ? smem_u32(sl + 1024) : 0u) + (seg[n * 4 + w] - sg0) * 4u;
                MFB_TACC(3);
                mbar_wait(&bars[full0 + s], par);
                MFB_TACC(0);
                build_records(reinterpret_cast<const uint64_t*>(sl) + w * 32, nz_addr, rec);
                __syncwarp();
                MFB_TACC(1);
                const int db = single ? 0 : (n & 1);
                const uint32_t ud = single ? static_cast<uint32_t>(n) : static_cast<uint32_t>(n >> 1);
                if (ud >= 1) mbar_wait(&bars[dempty0 + db], (ud - 1) & 1);  // the MMAs that read this dense buffer last are done
                MFB_TACC(2);
                if (fits) decode_to_umma32<true>(my_rec, lc, gblk, dst0 + db * kDenseBytes, lbo);
                else decode_to_umma32<false>(my_rec, lc, gblk, dst0 + db * kDenseBytes, lbo);
                fence_async_smem();  // this lane's dense stores -> visible to the tensor core's (async proxy) reads
                __syncwarp();
                bar_arrive(empty0 + s, kHandoffThreads);   // ring slot free
                bar_arrive(dfull0 + db, kHandoffThreads);  // dense block ready
                if (!is_v && tid == 0 && n == 0) MFB_TRACE_AT(4);
                if (!is_v && tid == 0 && n == nb - 1) MFB_TRACE_AT(5);
                if (is_v && tid == kWarpV0 * 32 && n == 0) MFB_TRACE_AT(6);
                if (is_v && tid == kWarpV0 * 32 && n == nb - 1) MFB_TRACE_AT(7);
            }
            MFB_TACC(3);
            MFB_TACC_STORE(is_v ? 20 : 16, 4, lane == 0 && w == 0);
        };
        if (warp >= kWarpV0) decode_role(std::true_type{});
        else decode_role(std::false_type{});
    }
    __syncthreads();
    if (tid == 0) MFB_TRACE_AT(8);
    publish_or_merge<G, FLAGGED>(a, unit, split, n_split, reinterpret_cast<const uint32_t*>(stat)[16], ored, stat, stat + 8,
                                 reinterpret_cast<float*>(smem + sm.slots_v));
}

// ------------------------------------------------------------------------------------------------
// Dense window split: <= 64 tokens of the fp16 residual window (models/llama_mustafar_kernel.py:278, :316).
// The chunk's K rows and V rows are contiguous (64 x 256 B each): one elected thread fetches both with
// two bulk copies at kernel entry, so the whole split costs a single DRAM round trip; scores, softmax
// and P.V then run out of shared memory.  Uses the first 8 warps of the CTA.
// q in the window path: a lane reads the 8-channel chunks `seg` and `8 + seg` of every head, the 8 lanes of a token
// differ in seg -> a chunk pitch of 9 floats spreads them over distinct banks (a [channel][G] layout made every
// such load an 8-way conflict: 7 % of all shared-memory wavefronts of a GQA launch).
constexpr int kWinQPitch = 16 * 9;
__device__ __forceinline__ int win_q_index(int g, int c) { return g * kWinQPitch + (c >> 3) * 9 + (c & 7); }
struct WinSmem {
    uint32_t kw, vw, bar, qs, sw, red, ored, ml, total;
};
__host__ __device__ inline WinSmem win_smem_map(int G) {
    WinSmem m;
    uint32_t o = 0;
    m.kw = o;
    o += kWinTokensPerSplit * kHeadDim * 2;
    m.vw = o;
    o += kWinTokensPerSplit * kHeadDim * 2;
    m.bar = o;
    o += 16;
    m.qs = o;  // float [G][16 chunks of 8 channels, pitch 9]: see win_q_index
    o += G * kWinQPitch * 4;
    m.sw = o;  // float [G][64] scores, then probabilities
    o += G * kWinTokensPerSplit * 4;
    m.red = o;  // float [8 warps][G][128]
    o += kWinWarps * G * 128 * 4;
    m.ored = o;  // float [G][128]
    o += G * 128 * 4;
    m.ml = o;  // float m[8], l[8]; uint32 tag
    o += 64 + 16;
    m.total = o;
    return m;
}

template <int G, bool FLAGGED>
__device__ __forceinline__ void window_split(const DecodeArgs& a, uint8_t* smem, int unit, int split, int n_split, int wchunk) {
    const mfb200_decode_params& p = a.p;
    const WinSmem sm = win_smem_map(G);
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t lane = lane_id();
    const bool active = warp < kWinWarps;
    const int t0 = wchunk * kWinTokensPerSplit;
    const int b = unit / p.kv_heads;
    const bool ref_round = (p.flags & MFB200_F_REF_SCORE_ROUNDING) != 0;

    const __half* kw_s = reinterpret_cast<const __half*>(smem + sm.kw);
    const __half* vw_s = reinterpret_cast<const __half*>(smem + sm.vw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + sm.bar);
    float* qs = reinterpret_cast<float*>(smem + sm.qs);
    float* sw = reinterpret_cast<float*>(smem + sm.sw);
    float* red = reinterpret_cast<float*>(smem + sm.red);
    float* ored = reinterpret_cast<float*>(smem + sm.ored);
    float* ml = reinterpret_cast<float*>(smem + sm.ml);

    pdl_launch_dependents();
    pdl_wait_prior_grids();  // the window is written by the previous step's launch
    const int win_len = cur_win_len(a);
    if (t0 >= win_len) return;  // static steps: the launch is planned for the window's capacity, this chunk is still empty
    const int nt = min(kWinTokensPerSplit, win_len - t0);
    if (tid == 0) {
        reinterpret_cast<uint32_t*>(ml)[16] = ld_relaxed_u32(epoch_ptr(p, unit)) + 1u;  // this launch's partial tag
        mbar_init(bar, 1);
        fence_mbar_init();
        const __half* kw = static_cast<const __half*>(p.k_win) + static_cast<int64_t>(unit) * p.win_stride + static_cast<int64_t>(t0) * kHeadDim;
        const __half* vw = static_cast<const __half*>(p.v_win) + static_cast<int64_t>(unit) * p.win_stride + static_cast<int64_t>(t0) * kHeadDim;
        const uint32_t bytes = static_cast<uint32_t>(nt) * kHeadDim * 2;
        mbar_expect_tx(bar, 2 * bytes);
        bulk_g2s(smem + sm.kw, kw, bytes, bar);
        bulk_g2s(smem + sm.vw, vw, bytes, bar);
    }
    {
        const __half* q = static_cast<const __half*>(p.q) + static_cast<int64_t>(unit) * G * kHeadDim;
        const RopeRows rp = rope_rows(p, unit);
        for (int i = tid; i < G * kHeadDim; i += blockDim.x)
            qs[win_q_index(i >> 7, i & 127)] = __half2float(rope_elem(rp, q + (i & ~127), i & 127));
    }
    __syncthreads();  // q staged, barrier initialised
    mbar_wait(bar, 0);
    // fused append: the step's new K/V row belongs at window row win_len-1; the split that owns that row
    // takes it from k_new/v_new (the bulk copy fetched a stale row there) and stores it to the window.
    if (p.k_new != nullptr && win_len - 1 >= t0 && win_len - 1 < t0 + nt) {
        if (tid < 32) {
            const int r = win_len - 1 - t0;
            const bool is_v = tid >= 16;
            const int j = tid & 15;
            uint4 row = reinterpret_cast<const uint4*>(is_v ? p.v_new : p.k_new)[static_cast<int64_t>(unit) * 16 + j];
            if (!is_v && p.rope_cos != nullptr) {  // the cache holds rotated keys (llama_mustafar_kernel.py:253, :270)
                const RopeRows rp = rope_rows(p, unit);
                const uint4 other = reinterpret_cast<const uint4*>(p.k_new)[static_cast<int64_t>(unit) * 16 + (j ^ 8)];
                const uint4 c4 = reinterpret_cast<const uint4*>(rp.cs)[j], s4 = reinterpret_cast<const uint4*>(rp.sn)[j];
                const __half* x = reinterpret_cast<const __half*>(&row);
                const __half* y = reinterpret_cast<const __half*>(&other);
                const __half* cc = reinterpret_cast<const __half*>(&c4);
                const __half* ss = reinterpret_cast<const __half*>(&s4);
                uint4 rot;
                __half* r8 = reinterpret_cast<__half*>(&rot);
#pragma unroll
                for (int e = 0; e < 8; ++e) r8[e] = rope_value(x[e], y[e], cc[e], ss[e], j < 8);
                row = rot;
            }
            reinterpret_cast<uint4*>(smem + (is_v ? sm.vw : sm.kw))[r * 16 + j] = row;
            uint4* gw = reinterpret_cast<uint4*>(static_cast<__half*>(is_v ? p.v_win : p.k_win) +
                                                 static_cast<int64_t>(unit) * p.win_stride + static_cast<int64_t>(win_len - 1) * kHeadDim);
            gw[j] = row;
        }
        __syncthreads();
    }

    // ---- scores: 8 lanes per token, lane reads two 16-byte chunks (channels 8s..8s+7, 64+8s..64+8s+7);
    //      a quarter-warp reads 128 contiguous bytes -> conflict-free.  Warp w: tokens 8w..8w+7.
    if (active) {
        const int tsub = lane >> 3, seg = lane & 7;
        float qr[G][16];
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                qr[g][i] = qs[win_q_index(g, 8 * seg + i)];
                qr[g][8 + i] = qs[win_q_index(g, 64 + 8 * seg + i)];
            }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int t = warp * 8 + u * 4 + tsub;
            const uint4 ka = *reinterpret_cast<const uint4*>(kw_s + t * kHeadDim + 8 * seg);
            const uint4 kb = *reinterpret_cast<const uint4*>(kw_s + t * kHeadDim + 64 + 8 * seg);
            const uint32_t w[8] = {ka.x, ka.y, ka.z, ka.w, kb.x, kb.y, kb.z, kb.w};
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
                if (seg == 0) sw[g * kWinTokensPerSplit + t] = s;  // rows >= nt hold stale smem: masked below
            }
        }
    }
    __syncthreads();
    // ---- softmax over the nt tokens: warp g handles head g (lane owns tokens lane, lane+32) ---------------
    if (warp < G) {
        const int g = warp;
        const __half* mask = p.mask ? static_cast<const __half*>(p.mask) + static_cast<int64_t>(b) * p.mask_stride + p.comp_len + t0 : nullptr;
        float s[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int t = lane + 32 * r;
            s[r] = -INFINITY;
            if (t < nt) {
                s[r] = ref_round_score(sw[g * kWinTokensPerSplit + t], p.score_div, ref_round);
                if (mask) s[r] = fmaxf(s[r] + __half2float(mask[t]), -65504.f);
            }
        }
        const float mx = warp_max(fmaxf(s[0], s[1]));
        float e[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int t = lane + 32 * r;
            e[r] = (t < nt) ? exp2f((s[r] - mx) * kLog2e) : 0.f;
            sw[g * kWinTokensPerSplit + t] = e[r];
        }
        const float tot = warp_sum(e[0] + e[1]);
        if (lane == 0) {
            ml[g] = mx;
            ml[8 + g] = tot;
        }
    }
    __syncthreads();
    // ---- P.V: warp w takes tokens 8w..8w+7; lane owns channels 4*lane .. 4*lane+3 ---------------------------
    if (active) {
        float o[G][4];
#pragma unroll
        for (int g = 0; g < G; ++g) o[g][0] = o[g][1] = o[g][2] = o[g][3] = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int t = warp * 8 + u;
            if (t < nt) {
                const uint2 vv = *reinterpret_cast<const uint2*>(vw_s + t * kHeadDim + 4 * lane);
                const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&vv.x));
                const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&vv.y));
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
#pragma unroll
        for (int g = 0; g < G; ++g) *reinterpret_cast<float4*>(red + (warp * G + g) * 128 + 4 * lane) = make_float4(o[g][0], o[g][1], o[g][2], o[g][3]);
    }
    __syncthreads();
    for (int i = tid; i < G * 128; i += blockDim.x) {
        const int g = i >> 7, c = i & 127;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kWinWarps; ++w) s += red[(w * G + g) * 128 + c];
        ored[i] = s;
    }
    __syncthreads();
    publish_or_merge<G, FLAGGED>(a, unit, split, n_split, reinterpret_cast<const uint32_t*>(ml)[16], ored, ml, ml + 8,
                        reinterpret_cast<float*>(smem));
}

// MODE is a compile-time switch: 0 = uniform plan + ticket merge, 1 = flat plan + ticket merge, 2 = uniform plan +
// flagged merge.  (The uniform kernels carry no segment-loop state in registers; the ticket kernels carry none of
// the flagged protocol's code: sharing instantiations cost 2-3 % through register pressure.)
template <int G, int MODE>
__global__ void __launch_bounds__(cta_threads(G), ctas_per_sm(G)) sparse_decode_attn_kernel(const __grid_constant__ DecodeArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    // 1-D grid, long CTAs first: all compressed splits of all units, then the short window splits.
#ifdef MFB_POISON
    {   // debug: poison the whole dynamic shared memory to expose reads of uninitialised data
        uint32_t total;
        asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(total));
        for (uint32_t i = threadIdx.x; i < total / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = MFB_POISON;
        __syncthreads();
    }
#endif
    if (threadIdx.x == 0) {
#ifdef MFB_TRACE
        s_trace_half = (a.p.flags >> 8) & 1;
#endif
        MFB_TRACE_AT(0);
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        MFB_TRACE_VAL(11, smid);
        MFB_TRACE_VAL(12, ~0ull);  // window CTA unless overwritten
    }
    const int units = a.p.batch * a.p.kv_heads;
    const int nblk = a.p.comp_len / kBlockTokens;
    const int id = blockIdx.x;
    const int n_base = a.n_csplit * units;  // uniform mode: CTAs [0, n_base) = split-major, then one extra split for units < n_extra
    constexpr bool FLAT = MODE == 1, FLAGGED = MODE == 2;
    const int n_comp = FLAT ? a.flat_ctas : n_base + a.n_extra;
    if (id < n_comp) {
        uint32_t tmem = 0;
        if constexpr (use_tc(G)) {  // tensor-memory accumulators S[2], O[2]: allocated once per CTA by the producer warp
            __shared__ uint32_t s_tmem;
            if ((threadIdx.x >> 5) == kWarpProducerTc) tmem_alloc(&s_tmem);
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
            tmem = s_tmem;
        }
        auto run_split = [&](int unit, int split, int n_split, int b0, int b1, bool again) {
            if constexpr (use_tc(G)) compressed_split_tc<G, FLAGGED>(a, smem, tmem, unit, split, n_split, b0, b1, again);
            else compressed_split<G, FLAGGED>(a, smem, unit, split, n_split, b0, b1, again);
        };
        if constexpr (!FLAT) {  // uniform mode: split `id / units` of unit `id % units`
            const int unit = id < n_base ? id % units : id - n_base, split = id < n_base ? id / units : a.n_csplit;
            const int nc = unit_csplits(a, unit);
            run_split(unit, split, nc + a.n_wsplit, split * nblk / nc, (split + 1) * nblk / nc, false);
        } else {
            // flat mode: an even cut of all units*nblk blocks, processed as segments cut at unit boundaries (one partial
            // per segment).  The segment arithmetic is redone per segment from opaque copies of (id, j) so that nothing
            // but j stays in registers across the split body.
            for (int j = 0;; ++j) {
                int idv = id, jv = j;
                asm volatile("" : "+r"(idv), "+r"(jv));
                const uint32_t cur = flat_start(a, idv), end = cur + a.flat_q + (idv < a.flat_r ? 1 : 0);
                const int unit = cur / static_cast<uint32_t>(nblk) + jv;
                const uint32_t u0 = static_cast<uint32_t>(unit) * nblk;
                const int b0 = jv == 0 ? cur - u0 : 0;
                const int b1 = min(static_cast<uint32_t>(nblk), end - u0);
                const bool more = u0 + nblk < end;
                run_split(unit, idv - static_cast<int>(flat_owner(a, u0)), unit_csplits(a, unit) + a.n_wsplit, b0, b1, j > 0);
                if (!more) break;
            }
        }
        if constexpr (use_tc(G)) {
            tc_fence_before();
            __syncthreads();
            if ((threadIdx.x >> 5) == kWarpProducerTc) tmem_dealloc(tmem);
        }
    } else {  // dense-window chunks come after all compressed CTAs
        const int unit = (id - n_comp) % units, wchunk = (id - n_comp) / units;
        const int nc = unit_csplits(a, unit);
        window_split<G, FLAGGED>(a, smem, unit, nc + wchunk, nc + a.n_wsplit, wchunk);
    }
    if (threadIdx.x == 0) MFB_TRACE_AT(10);
}

static size_t window_smem_bytes(int G) { return win_smem_map(G).total; }

static int pick_slot_nz_bytes(const mfb200_decode_params* p) {
    // capacity of one ring slot's nonzero area; without a hint the worst case (every element kept).
    int kb = p->slot_kb;
    if (kb <= 0 || kb > 16) kb = 16;
    return kb * 1024;
}

// Ring geometry of the CUDA-core / mma.sync variants: the plan counts on ctas_per_sm(G) resident CTAs per SM, so the
// shared memory of one CTA has to fit that many times (228 KB per SM, 1 KB reserved per CTA).  Depth 3 for small slots
// when it fits, else depth 2; a slot that still does not fit is cut (the rare block that is larger than its slot takes
// the kernel's global-load path instead).
static void fit_ring(int G, int* slot_nz_bytes, int* depth) {
    const int budget = 233472 / ctas_per_sm(G) - 1024;
    int slot = *slot_nz_bytes;
    const int d = (slot <= 8 * 1024 && static_cast<int>(smem_map(G, slot, 3).total) <= budget) ? 3 : 2;
    while (slot > 1024 && static_cast<int>(smem_map(G, slot, d).total) > budget) slot -= 1024;
    *slot_nz_bytes = slot;
    *depth = d;
}

template <int G, int MODE>
static int launch_decode(const DecodeArgs& a, cudaStream_t s) {
    const size_t comp_smem = use_tc(G) ? tc_smem_map(a.slot_nz_bytes, a.depth, a.nvd).total : smem_map(G, a.slot_nz_bytes, a.depth).total;
    size_t smem = (a.n_csplit > 0 || a.n_extra > 0 || a.flat_ctas > 0) ? comp_smem : 0;
    if (a.n_wsplit > 0) smem = smem > window_smem_bytes(G) ? smem : window_smem_bytes(G);
    static size_t configured[kMaxDevices] = {0};  // per device, per instantiation
    if (smem > 0) {
        const int rc = ensure_dynamic_smem(sparse_decode_attn_kernel<G, MODE>, configured, smem);
        if (rc) return rc;
    }
    cudaLaunchConfig_t cfg = {};
    const int units = a.p.batch * a.p.kv_heads;
    cfg.gridDim = dim3((a.flat_ctas > 0 ? a.flat_ctas : a.n_csplit * units + a.n_extra) + a.n_wsplit * units);
    cfg.blockDim = dim3(cta_threads(G));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (a.p.flags & MFB200_F_PDL) ? 1 : 0;
    MFB_CUDA(cudaLaunchKernelEx(&cfg, sparse_decode_attn_kernel<G, MODE>, a));
    return launch_status("sparse_decode_attn_kernel");
}

}  // namespace mfb

using namespace mfb;

namespace mfb {
constexpr int kMinBlocksPerSplit = 4;
struct Plan {
    int n_csplit, n_extra, n_wsplit, flat_ctas, flat_q, flat_r, max_split, flagged;
};
// Work decomposition of one launch.  The compressed CTAs of a small launch should all be resident at once (a second,
// nearly empty wave doubles the time of a batch-1 launch); the short window CTAs are dispatched last.
//  * uniform mode: every unit is cut into the same number of compressed splits;
//  * flat mode (24..256 blocks per CTA): the launch's units*nblk blocks are divided evenly over exactly the
//    resident CTA slots, CTAs may cross unit boundaries -> no wave-quantisation loss for mid-size batches.
// plan_hint (mfb200_decode_params::plan_hint): 0 = automatic, n > 0 = flat plan with n compressed CTAs, < 0 = never flat.
static Plan make_plan(int batch, int kv_heads, int groups, int comp_len, int win_len, int sm_count, int plan_hint = 0) {
    const int kMaxBlocksPerSplit = use_tc(groups) ? kMaxBlocksPerSplitTc : mfb::kMaxBlocksPerSplit;  // shadows the constant
    Plan pl = {0, 0, (win_len + kWinTokensPerSplit - 1) / kWinTokensPerSplit, 0, 0, 0, 0, 0};
    const int64_t units = static_cast<int64_t>(batch) * kv_heads;
    const int nblk = comp_len / kBlockTokens;
    if (nblk > 0) {
        const int64_t slots = static_cast<int64_t>(sm_count) * ctas_per_sm(groups);
        // No slots are held back for the window CTAs: they are short, trail the compressed CTAs in the grid and take
        // the slots of the first compressed CTAs that finish (a reserve of up to slots/8 measured 2 % slower at batch 1).
        const int64_t avail = slots;
        const int64_t B = units * nblk;
        const int forced_flat = plan_hint > 0 ? plan_hint : (plan_hint < 0 ? 0 : -1);
        // flat mode: long CTAs need no reserve for the (short, trailing) window CTAs.  Large MHA launches get two
        // CTAs per slot (measured: 355 -> 320 us for 128 units x 508 blocks; the hardware scheduler evens out the tail).
        int64_t n_flat = 0;
        if (forced_flat > 0) n_flat = forced_flat;
        else if (forced_flat < 0) {
            if (groups <= 2) {
                n_flat = B / (2 * slots) >= 48 ? 2 * slots : (B / slots >= 24 ? slots : 0);
            } else {
                // G >= 4 (2 or 1 CTAs per SM): the flat cut pays when the uniform plan's busiest slot carries > 5 % more
                // blocks than an even cut would (config 3: 2 splits of 62 blocks per unit on 296 slots vs 54 blocks
                // per CTA: 168 -> 156 us); k full waves when one wave would exceed the per-CTA block limit.
                const int64_t k = ((B + slots - 1) / slots + kMaxBlocksPerSplit - 1) / kMaxBlocksPerSplit;
                const int64_t n = k * slots;
                const int64_t min_c = (nblk + kMaxBlocksPerSplit - 1) / kMaxBlocksPerSplit;
                int64_t per_unit = slots / units;
                if (per_unit < min_c) per_unit = min_c;
                if (per_unit > nblk) per_unit = nblk;
                const int64_t waves_u = (units * per_unit + slots - 1) / slots;
                const int64_t busiest_u = waves_u * ((nblk + per_unit - 1) / per_unit), busiest_f = k * ((B + n - 1) / n);
                if (B / n >= 24 && busiest_u * 20 > busiest_f * 21) n_flat = n;
            }
        }
        if (n_flat > B) n_flat = B;
        if (n_flat > 0 && B < (int64_t{1} << 31) && (B + n_flat - 1) / n_flat <= kMaxBlocksPerSplit) {
            pl.flat_ctas = static_cast<int>(n_flat);
            pl.flat_q = static_cast<int>(B / n_flat);
            pl.flat_r = static_cast<int>(B % n_flat);
            // a unit's nblk blocks intersect at most ceil(nblk / q) + 1 CTAs
            pl.max_split = (nblk + pl.flat_q - 1) / pl.flat_q + 1 + pl.n_wsplit;
            return pl;
        }
        const int64_t target = avail;  // compressed CTAs of a uniform-mode launch
        int64_t per_unit = target / units;
        const int min_c = (nblk + kMaxBlocksPerSplit - 1) / kMaxBlocksPerSplit;
        if (per_unit < min_c) per_unit = min_c;
        if (per_unit > nblk) per_unit = nblk;
        // Never cut finer than kMinBlocksPerSplit blocks per CTA: below that the per-CTA start-up (index fetch, first ring fill) and
        // the number of partials to merge cost more than the parallelism returns.  Few-unit launches used to get one or two blocks
        // per CTA (8 units x 60 blocks on 296 slots: 37 splits per unit); measured cold, isolated: G=4 batch 1 x 4K 29.7 -> 21.4 us,
        // 2K s=0.7 25.4 -> 18.4, G=8 39.9 -> 28.7, 8 MHA heads 19.5 -> 17.4; launches that already had >= 4 blocks per CTA are
        // unchanged.  (plan_hint = -k, k > 1, overrides the minimum: tuning.)
        const int min_blocks = plan_hint < -1 ? -plan_hint : kMinBlocksPerSplit;
        if (per_unit > (nblk + min_blocks - 1) / min_blocks) per_unit = (nblk + min_blocks - 1) / min_blocks;
        if (per_unit < min_c) per_unit = min_c;
        pl.n_csplit = static_cast<int>(per_unit);
        // spend the remaining slots on one more split for the first units (no unit is ever cut finer than a block)
        if (per_unit == target / units && per_unit < nblk) pl.n_extra = static_cast<int>(target % units);
    }
    pl.max_split = pl.n_csplit + (pl.n_extra > 0 ? 1 : 0) + pl.n_wsplit;
    // short launches (every compressed CTA resident from the start, <= 16 blocks each): the merge tail matters
    pl.flagged = (nblk == 0 || (nblk + pl.n_csplit - 1) / pl.n_csplit <= 16) ? 1 : 0;
    return pl;
}
}  // namespace mfb

extern "C" int mfb200_decode_plan(int batch, int kv_heads, int groups, int comp_len, int win_len, int sm_count, int plan_hint,
                                  size_t* workspace_bytes, size_t* counter_bytes) {
    MFB_REQUIRE(batch > 0 && kv_heads > 0, "decode_plan: batch/kv_heads must be positive");
    MFB_REQUIRE(groups == 1 || groups == 2 || groups == 4 || groups == 8, "decode_plan: groups=%d not in {1,2,4,8}", groups);
    MFB_REQUIRE(comp_len >= 0 && comp_len % 64 == 0, "decode_plan: comp_len=%d must be a multiple of 64", comp_len);
    MFB_REQUIRE(win_len >= 0 && comp_len + win_len >= 1, "decode_plan: empty context");
    if (sm_count <= 0) {
        const int rc = current_device_sm_count(&sm_count);
        if (rc) return rc;
    }
    const int64_t units = static_cast<int64_t>(batch) * kv_heads;
    MFB_REQUIRE(units <= (1 << 20), "decode_plan: batch*kv_heads=%lld exceeds 2^20", static_cast<long long>(units));
    const Plan pl = make_plan(batch, kv_heads, groups, comp_len, win_len, sm_count, plan_hint);
    const size_t cbytes = 2 * ws_counter_bytes(static_cast<size_t>(units));
    if (counter_bytes) *counter_bytes = cbytes;
    if (workspace_bytes) *workspace_bytes = cbytes + static_cast<size_t>(units) * pl.max_split * groups * kPartStride * 8;
    return pl.max_split;  // partial slots per unit (informational; the launch re-derives the plan itself)
}

// Host-side self-check of the work decomposition: walks every CTA of the launch exactly like the kernel entry does
// (same helpers: flat_start / flat_owner / unit_csplits) and verifies that every unit's blocks are covered exactly
// once, that no CTA exceeds the per-CTA block limit, that partial slots are unique and inside the workspace stride,
// and that the flagged merge's owner of the last slot is the unit's highest CTA index.  No GPU involved.
extern "C" int mfb200_decode_plan_check(int batch, int kv_heads, int groups, int comp_len, int win_len, int sm_count, int plan_hint) {
    size_t ws = 0;
    const int rc = mfb200_decode_plan(batch, kv_heads, groups, comp_len, win_len, sm_count > 0 ? sm_count : 148, plan_hint, &ws, nullptr);
    if (rc < 0) return rc;
    if (sm_count <= 0) sm_count = 148;
    const Plan pl = make_plan(batch, kv_heads, groups, comp_len, win_len, sm_count, plan_hint);
    DecodeArgs a = {};
    a.p.batch = batch;
    a.p.kv_heads = kv_heads;
    a.p.groups = groups;
    a.p.comp_len = comp_len;
    a.p.win_len = win_len;
    a.n_csplit = pl.n_csplit;
    a.n_extra = pl.n_extra;
    a.n_wsplit = pl.n_wsplit;
    a.flat_ctas = pl.flat_ctas;
    a.flat_q = pl.flat_q;
    a.flat_r = pl.flat_r;
    a.max_split = pl.max_split;
    a.flagged = pl.flagged;
    const int units = batch * kv_heads, nblk = comp_len / kBlockTokens;
    MFB_REQUIRE(pl.n_wsplit * kWinTokensPerSplit >= win_len && (pl.n_wsplit - 1) * kWinTokensPerSplit < win_len || win_len == 0,
                "plan_check: %d window chunks for %d window tokens", pl.n_wsplit, win_len);
    MFB_REQUIRE(!(pl.flat_ctas > 0 && pl.flagged), "plan_check: flat plan with flagged merge");
    // per unit: next expected block, bitmap of used slots (max_split <= a few hundred), highest CTA index seen
    const int n_comp = pl.flat_ctas > 0 ? pl.flat_ctas : pl.n_csplit * units + pl.n_extra;
    int* next_blk = static_cast<int*>(calloc(units, sizeof(int)));
    int* n_seg = static_cast<int*>(calloc(units, sizeof(int)));
    int* last_id = static_cast<int*>(calloc(units, sizeof(int)));
    int* last_slot_owner = static_cast<int*>(calloc(units, sizeof(int)));
    unsigned char* used = static_cast<unsigned char*>(calloc(static_cast<size_t>(units) * pl.max_split, 1));
    int err = 0;
    auto segment = [&](int id, int unit, int split, int b0, int b1) {
        const int nc = unit_csplits(a, unit);
        if (unit < 0 || unit >= units || split < 0 || split >= nc || nc + pl.n_wsplit > pl.max_split || b0 != next_blk[unit] || b1 <= b0 ||
            b1 > nblk || b1 - b0 > (use_tc(groups) ? kMaxBlocksPerSplitTc : kMaxBlocksPerSplit) || used[static_cast<size_t>(unit) * pl.max_split + split]) {
            if (!err) set_error("plan_check: CTA %d segment (unit %d, slot %d of %d, blocks [%d, %d)) is inconsistent (next block %d, max_split %d)",
                                id, unit, split, nc, b0, b1, unit >= 0 && unit < units ? next_blk[unit] : -1, pl.max_split);
            err = 1;
            return;
        }
        used[static_cast<size_t>(unit) * pl.max_split + split] = 1;
        next_blk[unit] = b1;
        n_seg[unit] += 1;
        last_id[unit] = id;
        if (split == nc + pl.n_wsplit - 1) last_slot_owner[unit] = id;
    };
    for (int id = 0; id < n_comp && !err; ++id) {
        if (pl.flat_ctas == 0) {  // mirrors the uniform branch of the kernel entry
            const int n_base = pl.n_csplit * units;
            const int unit = id < n_base ? id % units : id - n_base, split = id < n_base ? id / units : pl.n_csplit;
            const int nc = unit_csplits(a, unit);
            // uniform splits of one unit are visited in increasing split order only across ids; order them by block range
            const int b0 = split * nblk / nc, b1 = (split + 1) * nblk / nc;
            if (b0 != next_blk[unit]) {  // split-major ids visit split s of every unit before split s+1: always in order
                if (!err) set_error("plan_check: uniform split %d of unit %d starts at block %d, expected %d", split, unit, b0, next_blk[unit]);
                err = 1;
                break;
            }
            segment(id, unit, split, b0, b1);
        } else {  // mirrors the flat branch
            for (int j = 0;; ++j) {
                const uint32_t cur = flat_start(a, id), end = cur + a.flat_q + (id < a.flat_r ? 1 : 0);
                const int unit = cur / static_cast<uint32_t>(nblk) + j;
                const uint32_t u0 = static_cast<uint32_t>(unit) * nblk;
                const int b0 = j == 0 ? cur - u0 : 0;
                const int b1 = static_cast<int>(end - u0 < static_cast<uint32_t>(nblk) ? end - u0 : nblk);
                segment(id, unit, id - static_cast<int>(flat_owner(a, u0)), b0, b1);
                if (err || !(u0 + nblk < end)) break;
            }
        }
    }
    for (int w = 0; w < pl.n_wsplit * units && !err; ++w) {  // window CTAs: chunk-major after all compressed CTAs
        const int id = n_comp + w, unit = w % units, wchunk = w / units, nc = nblk > 0 ? unit_csplits(a, unit) : 0;
        const int slot = nc + wchunk;
        if (slot >= pl.max_split || used[static_cast<size_t>(unit) * pl.max_split + slot]) {
            set_error("plan_check: window chunk %d of unit %d: slot %d unusable (max_split %d)", wchunk, unit, slot, pl.max_split);
            err = 1;
            break;
        }
        used[static_cast<size_t>(unit) * pl.max_split + slot] = 1;
        last_id[unit] = id;
        if (slot == nc + pl.n_wsplit - 1) last_slot_owner[unit] = id;
    }
    for (int u = 0; u < units && !err; ++u) {
        const int nc = nblk > 0 ? unit_csplits(a, u) : 0;
        if (next_blk[u] != nblk || n_seg[u] != nc) {
            set_error("plan_check: unit %d covered up to block %d of %d with %d segments (expected %d)", u, next_blk[u], nblk, n_seg[u], nc);
            err = 1;
        } else if (pl.flagged && last_slot_owner[u] != last_id[u]) {
            set_error("plan_check: unit %d: the owner of the last slot (CTA %d) is not its highest CTA (%d)", u, last_slot_owner[u], last_id[u]);
            err = 1;
        }
    }
    free(next_blk);
    free(n_seg);
    free(last_id);
    free(last_slot_owner);
    free(used);
    return err ? MFB200_EINVAL : MFB200_OK;
}

extern "C" int mfb200_sparse_decode_attention(const mfb200_decode_params* p, mfb200_stream_t stream) {
    MFB_REQUIRE(p != nullptr, "decode: null params");
    MFB_REQUIRE(p->batch > 0 && p->kv_heads > 0, "decode: batch/kv_heads must be positive");
    MFB_REQUIRE(static_cast<int64_t>(p->batch) * p->kv_heads <= (1 << 20), "decode: too many units");
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
    MFB_REQUIRE((p->rope_cos == nullptr) == (p->rope_sin == nullptr), "decode: rope_cos and rope_sin go together");
    if (p->rope_cos)
        MFB_REQUIRE(((reinterpret_cast<uintptr_t>(p->rope_cos) | reinterpret_cast<uintptr_t>(p->rope_sin)) & 15) == 0 &&
                        p->rope_stride >= 0 && p->rope_stride % 8 == 0,
                    "decode: rope_cos/rope_sin must be 16-byte aligned fp16 rows (stride a multiple of 8 halves, 0 = shared)");
    MFB_REQUIRE((p->k_new == nullptr) == (p->v_new == nullptr), "decode: k_new and v_new must be given together");
    if (p->k_new)
        MFB_REQUIRE(p->win_len >= 1 && ((reinterpret_cast<uintptr_t>(p->k_new) | reinterpret_cast<uintptr_t>(p->v_new)) & 15) == 0,
                    "decode: k_new/v_new need win_len >= 1 and 16-byte alignment");
    DecodeArgs a;
    a.p = *p;
    a.peer = mfb200_peer_out{};
    if (p->peer != nullptr) {
        const mfb200_peer_out& pe = *p->peer;
        MFB_REQUIRE(pe.n_peers >= 2 && pe.n_peers <= MFB200_MAX_PEERS && pe.rank >= 0 && pe.rank < pe.n_peers && pe.reserved == 0,
                    "decode: peer output needs 2..%d peers and a rank inside them", MFB200_MAX_PEERS);
        MFB_REQUIRE(pe.row0 >= 0 && pe.row0 + p->kv_heads * p->groups <= pe.rows_total, "decode: peer output rows [%d, %d) outside [0, %d)",
                    pe.row0, pe.row0 + p->kv_heads * p->groups, pe.rows_total);
        for (int r = 0; r < pe.n_peers; ++r)
            MFB_REQUIRE(pe.out[r] != nullptr && pe.flags[r] != nullptr, "decode: peer %d has no output buffer / flags", r);
        a.peer = pe;
    }
    int sm_count = 0;
    {
        const int rc = current_device_sm_count(&sm_count);
        if (rc) return rc;
    }
    const Plan pl = make_plan(p->batch, p->kv_heads, p->groups, p->comp_len, p->win_len, sm_count, p->plan_hint);
    a.n_csplit = pl.n_csplit;
    a.n_extra = pl.n_extra;
    a.n_wsplit = pl.n_wsplit;
    a.flat_ctas = pl.flat_ctas;
    a.flat_q = pl.flat_q;
    a.flat_r = pl.flat_r;
    a.max_split = pl.max_split;
    a.flagged = pl.flagged;
    if (p->workspace_kb > 0) {
        const size_t need = 2 * ws_counter_bytes(static_cast<size_t>(p->batch) * p->kv_heads) +
                            static_cast<size_t>(p->batch) * p->kv_heads * pl.max_split * p->groups * kPartStride * 8;
        MFB_REQUIRE(need <= static_cast<size_t>(p->workspace_kb) * 1024, "decode: workspace of %d KB is too small for this launch (%zu bytes needed; size it with mfb200_decode_workspace_max)",
                    p->workspace_kb, need);
    }
    a.slot_nz_bytes = pick_slot_nz_bytes(p);
    fit_ring(p->groups, &a.slot_nz_bytes, &a.depth);
    a.nvd = 1;
    if (use_tc(p->groups)) {
        // two CTAs per SM: 227 KB / 2 minus the per-CTA reservation.  Ring depth 2 (the dense operand buffers are the
        // second pipeline stage); a second dense V buffer if it still fits.
        constexpr uint32_t kBudget = (233472 / 2) - 1024;  // 228 KB per SM, 1 KB reserved per CTA
        a.depth = 2;
        if (tc_smem_map(a.slot_nz_bytes, 2, 2).total <= kBudget) a.nvd = 2;
        else if (tc_smem_map(a.slot_nz_bytes - 512, 2, 2).total <= kBudget) {  // half a KB less staging per slot buys the buffer
            a.slot_nz_bytes -= 512;
            a.nvd = 2;
        }
    }
    auto s = static_cast<cudaStream_t>(stream);
    switch (p->groups) {
        case 1: return a.flat_ctas > 0 ? launch_decode<1, 1>(a, s) : (a.flagged ? launch_decode<1, 2>(a, s) : launch_decode<1, 0>(a, s));
        case 2: return a.flat_ctas > 0 ? launch_decode<2, 1>(a, s) : (a.flagged ? launch_decode<2, 2>(a, s) : launch_decode<2, 0>(a, s));
        case 4: return a.flat_ctas > 0 ? launch_decode<4, 1>(a, s) : (a.flagged ? launch_decode<4, 2>(a, s) : launch_decode<4, 0>(a, s));
        default: return a.flat_ctas > 0 ? launch_decode<8, 1>(a, s) : (a.flagged ? launch_decode<8, 2>(a, s) : launch_decode<8, 0>(a, s));
    }
}

extern "C" size_t mfb200_decode_workspace_max(int batch, int kv_heads, int groups, int max_comp_len, int max_win_len,
                                              int sm_count) {
    // The number of partial slots per unit is not monotone in the lengths (ragged splits, the switch between the
    // uniform and the flat plan, k-wave flat plans): take the maximum over EVERY (compressed length, window chunk
    // count) the cache can pass through.  Pure host arithmetic, a few microseconds per thousand tokens of capacity.
    if (sm_count <= 0 && current_device_sm_count(&sm_count) != MFB200_OK) sm_count = 148;
    size_t best = 0;
    const int max_nw = (max_win_len + kWinTokensPerSplit - 1) / kWinTokensPerSplit;
    for (int comp = 0; comp <= max_comp_len - max_comp_len % kBlockTokens; comp += kBlockTokens) {
        for (int nw = (comp > 0 ? 0 : 1); nw <= (max_nw > 0 ? max_nw : 1); ++nw) {
            int win = nw * kWinTokensPerSplit;
            if (win > max_win_len) win = max_win_len > 0 ? max_win_len : 1;
            size_t ws = 0;
            if (mfb200_decode_plan(batch, kv_heads, groups, comp, win, sm_count, 0, &ws, nullptr) > 0 && ws > best) best = ws;
        }
    }
    return best + 4096;
}

extern "C" int mfb200_decode_step(mfb200_decode_params* p, const void* q, const void* k_new, const void* v_new, void* out,
                                  int sm_count, mfb200_stream_t stream) {
    MFB_REQUIRE(p != nullptr && q && k_new && v_new && out, "decode_step: null pointer");
    p->q = q;
    p->out = out;
    p->k_new = k_new;
    p->v_new = v_new;
    (void)sm_count;
    p->win_len += 1;
    const int rc = mfb200_sparse_decode_attention(p, stream);
    if (rc < 0) p->win_len -= 1;  // nothing was launched: the long-lived block keeps describing the cache as it is
    return rc;
}

extern "C" int mfb200_decode_step_layers(mfb200_decode_params* const* layers, int n_layers, const void* q, const void* k_new,
                                         const void* v_new, void* out, int64_t q_layer_stride, int64_t kv_layer_stride,
                                         int64_t out_layer_stride, mfb200_stream_t stream) {
    MFB_REQUIRE(layers != nullptr && n_layers >= 0 && q && k_new && v_new && out, "decode_step_layers: null pointer");
    for (int l = 0; l < n_layers; ++l) {
        MFB_REQUIRE(layers[l] != nullptr, "decode_step_layers: layer %d has no parameter block", l);
        const int rc = mfb200_decode_step(layers[l], static_cast<const __half*>(q) + l * q_layer_stride,
                                          static_cast<const __half*>(k_new) + l * kv_layer_stride,
                                          static_cast<const __half*>(v_new) + l * kv_layer_stride,
                                          static_cast<__half*>(out) + l * out_layer_stride, 0, stream);
        if (rc < 0) return rc;  // layers [0, l) were launched and advanced, layers [l, n) are untouched
    }
    return n_layers;
}

namespace mfb {
__global__ void lengths_add_kernel(int32_t* lengths, int n, int delta) {
    pdl_launch_dependents();
    pdl_wait_prior_grids();  // the previous step's launches still read the old values
    for (int i = threadIdx.x; i < n; i += blockDim.x) lengths[i] += delta;
}
}  // namespace mfb

extern "C" int mfb200_lengths_add(int32_t* lengths, int n, int delta, mfb200_stream_t stream) {
    MFB_REQUIRE(lengths != nullptr && n > 0, "lengths_add: null pointer / empty range");
    mfb::lengths_add_kernel<<<1, 128, 0, static_cast<cudaStream_t>(stream)>>>(lengths, n, delta);
    return launch_status("lengths_add_kernel");
}

extern "C" int mfb200_decode_layers_static(const mfb200_decode_params* const* layers, int n_layers, mfb200_stream_t stream) {
    MFB_REQUIRE(layers != nullptr && n_layers >= 0, "decode_layers_static: null pointer");
    for (int l = 0; l < n_layers; ++l) {
        MFB_REQUIRE(layers[l] != nullptr && layers[l]->win_len_dev != nullptr, "decode_layers_static: layer %d has no device-side window length", l);
        const int rc = mfb200_sparse_decode_attention(layers[l], stream);
        if (rc < 0) return rc;
    }
    return n_layers;
}

#ifdef MFB_TRACE
// debug build only: copies the last launch's per-CTA timeline (kTraceSlots u64 per CTA) to the host
extern "C" int mfb200_debug_trace(unsigned long long* out, int n_ctas) {
    if (n_ctas > mfb::kTraceMaxCtas) n_ctas = mfb::kTraceMaxCtas;
    MFB_CUDA(cudaMemcpyFromSymbol(out, mfb::g_trace, sizeof(unsigned long long) * mfb::kTraceSlots * n_ctas));
    MFB_CUDA(cudaMemcpyFromSymbol(out + static_cast<size_t>(mfb::kTraceSlots) * n_ctas, mfb::g_trace,
                                  sizeof(unsigned long long) * mfb::kTraceSlots * n_ctas,
                                  sizeof(unsigned long long) * mfb::kTraceSlots * mfb::kTraceMaxCtas));
    return mfb::kTraceSlots;
}
#endif
