/*
 * mustafar_b200.h — C ABI of libmustafar_b200.so (sm_100a).
 *
 * This is the drop-in boundary for the sparse-KV decode path of dhjoo98/mustafar.  Every entry
 * point takes plain device pointers, sizes and a CUDA stream handle; there are no torch types in
 * any signature.  All functions are asynchronous on `stream`, never synchronise the host, and
 * return 0 on success or a negative MFB200_E* code (the message is available from
 * mfb200_last_error()).  CUDA launch errors are reported, unlike the reference wrapper which drops
 * them (kernel/kernel_wrapper/mustafar_wrapper.cu:113, :242).
 *
 * Reference interfaces each entry point replaces (paths relative to /root/reference):
 *
 *   mfb200_prune_rows            models/llama_mustafar_kernel.py:77-113, :117-153  (dh_prune_key/value)
 *   mfb200_prune_rows_scored     models/llama_mustafar_Kt_Opa_Vt_Mag.py:98-106, :131-156 (output-aware dh_prune_key)
 *   mfb200_prune_token_groups    models/llama_mustafar_Kt_Mag_Vc_Mag.py:107-170 (channel-wise dh_prune_value)
 *   mfb200_compress_count        kernel/compression.py:9-54, :57-115   (calculate_bitmap_{key,value}_batched)
 *   mfb200_compress_scan         kernel/compression.py:294-304, :387-397 (torch.cumsum / cat glue)
 *   mfb200_compress_pack         kernel/compression.py:118-174, :178-247 (compress_{key,value}_batched)
 *   mfb200_key_formulation       kernel/build/SpMM_API.cuh:46-64  Key_SplitK_API   (csrc/SpMM_API.cu:86-139)
 *   mfb200_value_formulation     kernel/build/SpMM_API.cuh:92-110 Value_SplitK_API (csrc/SpMM_API.cu:193-254)
 *   mfb200_decode_plan / mfb200_sparse_decode_attention
 *                                the fused replacement of models/llama_mustafar_kernel.py:268-320
 *                                (pad q → Key SpMV → window matmul → softmax → pad P → Value SpMV →
 *                                window matmul → add); new, not in the reference.
 *   mfb200_window_append         models/llama_mustafar_kernel.py:270, :309 (torch.cat of the new k/v row)
 *
 * Compressed-KV format (normative, SURVEY.md App. A / kernel/compression.py): tile = 64 fp16
 * elements; bitmap bit (63-e) set iff element e != 0; nonzeros packed in ascending e, zero padded
 * to a multiple of 8 halves; idx[t] = exclusive prefix of padded sizes in units of 2 halves.
 * K tile t = token_block*128 + channel (64 consecutive tokens of one channel);
 * V tile t = token_block*128 + channel_half*64 + token_in_block (64 consecutive channels).
 */
#ifndef MUSTAFAR_B200_H
#define MUSTAFAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFB200_ABI_VERSION 5

#define MFB200_OK 0
#define MFB200_EINVAL (-1)  /* bad argument (shape, alignment, null pointer) */
#define MFB200_ECUDA (-2)   /* CUDA runtime / launch error */

#define MFB200_LAYOUT_KEY 0
#define MFB200_LAYOUT_VALUE 1

#define MFB200_HEAD_DIM 128

typedef void* mfb200_stream_t; /* cudaStream_t */

int mfb200_abi_version(void);
/* Thread-local, NUL-terminated description of the last error returned on this thread. */
const char* mfb200_last_error(void);

/* ---- a1: per-token magnitude threshold pruning -------------------------------------------------
 * y[r, :] = x[r, :] * (|x[r, :]| >= kth_smallest(|x[r, :]|, k)),  rows of 128 fp16.
 * k = max(1, int(sparsity*128)) is computed by the caller.  x == y (in place) is allowed. */
int mfb200_prune_rows(const void* x, void* y, int64_t rows, int k, mfb200_stream_t stream);

/* ---- f4: the other pruning policies that feed the same compressed format -----------------------
 * Output-aware key pruning (models/llama_mustafar_Kt_Opa_Vt_Mag.py:98-106 prefill, :131-156 decode):
 *   score[r, c] = | x[r, c] * w[r / rows_per_unit, c] |  (fp16 product),  w: fp16 [units, 128] = the unit's folded |q|;
 *   y[r, :] = x[r, :] * (score[r, :] >= kth_smallest(score[r, :], k)),  k = 128 - n_keep + 1 for n_keep survivors.
 *   rows_per_unit == 0: w is fp16 [rows, 128] and IS the (non-negative) score - the decode-time form, where the score of the
 *   oldest window row has been accumulated over group_size steps (`:131-145`).
 *   Equal to the reference's sort + scatter mask unless the n_keep-th and (n_keep+1)-th highest scores tie (it then keeps an
 *   arbitrary subset of the tied entries; here all of them survive).  x == y allowed. */
int mfb200_prune_rows_scored(const void* x, const void* w, void* y, int64_t rows, int64_t rows_per_unit, int k, mfb200_stream_t stream);
/* Channel-wise value pruning (models/llama_mustafar_Kt_Mag_Vc_Mag.py:107-170): x fp16 [units, tokens, 128], tokens % group == 0,
 * group <= 128; inside every group of `group` consecutive tokens each channel keeps |x| >= the k-th smallest magnitude of its
 * `group` values (k = max(1, int(sparsity*group)) is computed by the caller).  x == y allowed. */
int mfb200_prune_token_groups(const void* x, void* y, int64_t units, int64_t tokens, int group, int k, mfb200_stream_t stream);

/* ---- a2/a3: bitmaps + padded per-tile counts ----------------------------------------------------
 * x: fp16 [heads, tokens, 128] contiguous, tokens % 64 == 0.  If prune_k > 0 the threshold prune of
 * mfb200_prune_rows is applied on the fly (x itself is not modified).
 * bitmaps: int64 [heads, tokens*2]; counts: int32 [heads, tokens*2] in units of 2 halves. */
int mfb200_compress_count(const void* x, int64_t heads, int64_t tokens, int layout, int prune_k,
                          int64_t* bitmaps, int32_t* counts, mfb200_stream_t stream);

/* ---- a4: exclusive scan ---------------------------------------------------------------------------
 * accum: int32 [heads, accum_stride], accum[h, tile_offset + t] for t in [0, tiles]; the running
 * value starts at accum[h, tile_offset] when tile_offset > 0 (append to an existing head) and at 0
 * otherwise.  head_total (may be NULL): int32 [heads] = accum[h, tile_offset + tiles]. */
int mfb200_compress_scan(const int32_t* counts, int64_t heads, int64_t tiles, int32_t* accum,
                         int64_t accum_stride, int64_t tile_offset, int32_t* head_total,
                         mfb200_stream_t stream);

/* ---- a5/a6: pack nonzeros ---------------------------------------------------------------------------
 * Writes, for every tile, its nonzeros followed by zero padding at
 *   packed + head_base[h] + 2*accum[h, tile_offset + t]      (all in halves).
 * bitmaps/accum as produced above (bitmaps: [heads, tokens*2] of THIS chunk; accum may be the
 * long-lived per-head array of a cache, addressed through accum_stride/tile_offset).
 * head_base: int64 [heads], halves, multiple of 8.
 * head_capacity (halves, 0 = unchecked): a tile that would end beyond head_base[h]+head_capacity is
 * not written and *overflow (int32, device, may be NULL) is set to 1 — lets a preallocated slab be
 * smaller than the worst case without risking its neighbours. */
int mfb200_compress_pack(const void* x, int64_t heads, int64_t tokens, int layout,
                         const int64_t* bitmaps, const int32_t* accum, int64_t accum_stride,
                         int64_t tile_offset, const int64_t* head_base, void* packed,
                         int64_t head_capacity, int32_t* overflow, mfb200_stream_t stream);

/* ---- a7: decode-time append of one 256-token chunk (models/llama_mustafar_kernel.py:324-398) ------------
 * For every unit: prune (prune_k > 0) and compress window rows [0, 256) of K and of V, append them to the
 * unit's bitmap / idx / nonzero slabs at tile_offset (= 2 * tokens already compressed), then move window rows
 * [256, win_len) to the front.  One launch, no host sync, no temporaries; the caller then adds 256 to its
 * compressed length and subtracts 256 from its window length.  Layout of the slabs as in
 * mfb200_decode_params; head_base in halves; head_capacity / overflow as in mfb200_compress_pack. */
int mfb200_compress_append_chunk(void* k_win, void* v_win, int64_t win_stride, int64_t units, int win_len,
                                 int prune_k_key, int prune_k_value, int64_t* k_bmp, int32_t* k_idx, void* k_nz,
                                 const int64_t* k_head_base, int64_t* v_bmp, int32_t* v_idx, void* v_nz,
                                 const int64_t* v_head_base, int64_t bmp_stride, int64_t idx_stride,
                                 int64_t tile_offset, int64_t head_capacity, int32_t* overflow,
                                 mfb200_stream_t stream);

/* Prefill (models/llama_mustafar_kernel.py:416-442: dh_prune_key/value + convert_key/value_batched on the prompt):
 * prunes and compresses tokens [0, tokens) of key_states / value_states (fp16 [batch, kv_heads, >= tokens, 128],
 * arbitrary batch / head / token strides given in halves as {stride_b, stride_h, stride_t}, innermost stride 1)
 * into the cache slabs at tile_offset, K and V in ONE single-pass launch (each input element is read once; block
 * offsets are exchanged between CTAs with a decoupled look-back whose block order comes from an atomic ticket, so it does
 * not depend on the CTA dispatch order).  status_ws: 2 * batch * kv_heads * (tokens/64 + 1) * 8 bytes of scratch (zeroed
 * by the call).  Slabs / head_base / head_capacity / overflow as in
 * mfb200_compress_append_chunk.  Bit-identical to mfb200_compress_count + _scan + _pack. */
int mfb200_compress_prefill(const void* k, const void* v, const int64_t* k_strides, const int64_t* v_strides,
                            int batch, int kv_heads, int64_t tokens, int prune_k_key, int prune_k_value,
                            int64_t* k_bmp, int32_t* k_idx, void* k_nz, const int64_t* k_head_base,
                            int64_t* v_bmp, int32_t* v_idx, void* v_nz, const int64_t* v_head_base,
                            int64_t bmp_stride, int64_t idx_stride, int64_t tile_offset, int64_t head_capacity,
                            int32_t* overflow, void* status_ws, mfb200_stream_t stream);

/* ---- a8-a13: the two batched SpMV operators, reference argument order ----------------------------------
 * C[bq, n, m] (fp16 [Batch_Size, 8, M_Global]).  N_Global must be 8, K_Global 128 (key) /
 * M_Global 128 (value), Split_K is ignored (the reference hard-wires 1).  `A` is unused (NULL in the
 * reference).  workspace: value op only, >= mfb200_value_workspace_bytes(...) bytes, zero-initialised
 * once by the caller (the kernel leaves it zeroed). */
int mfb200_key_formulation(mfb200_stream_t stream, const void* A, const uint64_t* bmp, const void* NZ,
                           const uint32_t* idx, const uint32_t* NZ_offset, const void* B, void* C,
                           int M_Global, int N_Global, int K_Global, void* Reduction_Workspace,
                           int Split_K, int Batch_Size, int num_key_value_groups);
int mfb200_value_formulation(mfb200_stream_t stream, const void* A, const uint64_t* bmp, const void* NZ,
                             const uint32_t* idx, const uint32_t* NZ_offset, const void* B, void* C,
                             int M_Global, int N_Global, int K_Global, void* workspace,
                             int Split_K, int Batch_Size, int num_key_value_groups);
size_t mfb200_value_workspace_bytes(int K_Global, int Batch_Size);

/* ---- fused sparse decode attention ------------------------------------------------------------------ */
/* Head-sharded decode without a collective: every rank's launch stores its output rows straight into the gathered
 * output buffer of EVERY rank (peer-to-peer stores over NVLink / NVSwitch from the split-merge epilogue) and then raises
 * one arrival flag per (rank, unit) on every rank; mfb200_peer_wait() on the consumer side replaces the all-gather.
 * Replaces the torch.cat / all-gather a tensor-sharded caller would issue after the attention op.  All pointers are
 * device pointers valid on THIS device (own allocations, or peers' allocations opened with mfb200_ipc_open). */
#define MFB200_MAX_PEERS 8
typedef struct mfb200_peer_out {
    int32_t n_peers;    /* ranks that share the gathered output, 2..MFB200_MAX_PEERS */
    int32_t rank;       /* this rank */
    int32_t row0;       /* first row (query head) of this rank inside a sequence's rows_total rows */
    int32_t rows_total; /* rows (query heads) per sequence of the gathered output: fp16 [B, rows_total, 128] */
    uint32_t epoch;     /* the flags are set to this value; must grow from launch to launch on the same flags */
    int32_t reserved;   /* must be 0 */
    void* out[MFB200_MAX_PEERS];       /* out[r]: rank r's gathered output buffer */
    uint32_t* flags[MFB200_MAX_PEERS]; /* flags[r]: rank r's arrival flags, uint32 [n_peers][B * Hkv] (zero before first use) */
} mfb200_peer_out;

typedef struct mfb200_decode_params {
    /* geometry */
    int32_t batch;        /* B  */
    int32_t kv_heads;     /* Hkv */
    int32_t groups;       /* G = Hq / Hkv, 1..8 */
    int32_t comp_len;     /* L: tokens in the compressed cache, multiple of 64 (0 allowed) */
    int32_t win_len;      /* Lw: tokens in the dense window, >= 0; L + Lw >= 1 */
    int32_t flags;        /* MFB200_F_* */
    float score_div;      /* sqrt(head_dim): scores are divided by it (llama_mustafar_kernel.py:284) */
    int32_t n_split;      /* informational (mfb200_decode_plan's return value); the launch derives its own work
                             decomposition from the geometry, so a stale value is harmless */
    int32_t slot_kb;      /* staging capacity for one 64-token block's nonzeros, KB (1..16); 0 = 16
                             (worst case, every element kept).  Larger blocks still work: they take a
                             slower path that reads their nonzeros straight from global memory. */
    int32_t workspace_kb; /* size of `workspace` in KB (rounded down); 0 = not checked.  When set, a launch whose plan needs
                             more is refused with MFB200_EINVAL instead of writing past the buffer. */
    int32_t plan_hint;    /* work decomposition: 0 = chosen from the geometry (production), n > 0 = flat plan with n compressed
                             CTAs, < 0 = never the flat plan, -k (k > 1) additionally at least k blocks per compressed CTA
                             (tests / tuning; size the workspace with the same hint) */
    int32_t reserved0;    /* must be 0 */
    /* query / output: fp16 [B, Hq, 128] contiguous */
    const void* q;
    void* out;
    /* compressed K and V.  Per (b, hkv) unit u = b*Hkv + h:
     *   bitmaps  at  bmp + u*bmp_stride            (uint64, tile order of the layout)
     *   idx      at  idx + u*idx_stride            (uint32, units of 2 halves, relative to the unit)
     *   nonzeros at  (uint4*)nz + nz_off[u]        (nz_off in 16-byte units) */
    const uint64_t* k_bmp;
    const uint32_t* k_idx;
    const void* k_nz;
    const int64_t* k_nz_off;
    const uint64_t* v_bmp;
    const uint32_t* v_idx;
    const void* v_nz;
    const int64_t* v_nz_off;
    int64_t bmp_stride; /* in tiles */
    int64_t idx_stride; /* in entries */
    /* dense window: fp16, row t of unit u at  win + u*win_stride + t*128 (written only when k_new/v_new
     * are given) */
    void* k_win;
    void* v_win;
    int64_t win_stride; /* halves */
    /* optional fused append of the step's new token (models/llama_mustafar_kernel.py:270, :309): fp16
     * [units, 128] rows that the kernel stores at window row win_len-1 (win_len already counts the new
     * token) and attends to in the same launch; both NULL = the window already holds every row. */
    const void* k_new;
    const void* v_new;
    /* optional additive mask, fp16 [B, mask_stride], entry t (0 <= t < L+Lw); NULL = none */
    const void* mask;
    int64_t mask_stride;
    /* scratch: >= mfb200_decode_plan() bytes; the counter part must be zero before the first launch
     * (kernels leave it zeroed). */
    void* workspace;
    /* optional (host pointer, read at launch): also store the output rows into the peers' gathered buffers; NULL = off */
    const mfb200_peer_out* peer;
    /* optional (DEVICE pointer): the window length is read from *win_len_dev by the kernel instead of from win_len, which
     * then only gives the CAPACITY the launch is planned for (win_len_dev[0] <= win_len).  Makes a decode step a static
     * sequence of launches - the same parameter blocks every step - that can be captured ONCE in a CUDA graph and replayed
     * until the next compression event: see mfb200_decode_layers_static / mfb200_lengths_add.  NULL = off. */
    const int32_t* win_len_dev;
    /* optional fused rotary position embedding (models/llama_mustafar_kernel.py:238-253: apply_rotary_pos_emb in front of
     * the attention): fp16 cos / sin rows of 128 entries per SEQUENCE (the layout transformers' rotary_emb returns: entry c
     * and c+64 hold the same angle), sequence b at rope_cos + b*rope_stride (stride 0 = one row for all).  When set, q and
     * k_new are taken as UNROTATED: the kernel attends with  x*cos + rotate_half(x)*sin  (each product and the sum rounded
     * to fp16, exactly the arithmetic of the reference's fp16 tensor ops) and appends the rotated K row.  Both NULL = off. */
    const void* rope_cos;
    const void* rope_sin;
    int64_t rope_stride; /* halves */
} mfb200_decode_params;

/* Round q·k to fp16 and divide by score_div in fp16 like the reference glue does
 * (SpMM_Kernel.cuh:418, llama_mustafar_kernel.py:284). Off = keep fp32 scores. */
#define MFB200_F_REF_SCORE_ROUNDING 1
/* Launch with programmatic stream serialization (PDL): the kernel lets its successor in the stream begin
 * launching early and itself waits (griddepcontrol.wait) for its predecessors before it touches q, k_new,
 * v_new, the window or the workspace.  Always safe. */
#define MFB200_F_PDL 2
/* With MFB200_F_PDL: additionally request idx / bitmaps / nonzeros of the first blocks BEFORE that wait.
 * Only valid when the kernel immediately preceding this launch in the stream does not write this cache's
 * compressed streams (i.e. it is not this cache's own compress_scan/compress_pack). */
#define MFB200_F_PDL_EARLY_KV 4

/* Reports the workspace a launch of this geometry needs.  sm_count <= 0 → query the current device.
 * Returns the number of partial-result slots per (sequence, KV head) unit (>= 1) or a negative error.
 * Decomposition (chosen inside the launch, same rule): small launches cut every unit into the same number of
 * compressed splits so that all CTAs are resident at once; large MHA launches (G <= 2, >= 24 blocks per resident
 * CTA slot) divide all units*blocks evenly over 1x or 2x the resident CTA slots, CTAs may cross unit boundaries.
 * Workspace layout: [per-unit launch epochs][per-unit tickets] (units*4 bytes each, rounded to 256; together
 * counter_bytes) followed by the partials as 8-byte {fp32, tag} entries.
 * The WHOLE workspace must be zero before the first launch on it (and again if batch / kv_heads / groups change);
 * launches keep it consistent afterwards (tags only grow), so it is graph-replayable. */
int mfb200_decode_plan(int batch, int kv_heads, int groups, int comp_len, int win_len, int sm_count, int plan_hint,
                       size_t* workspace_bytes, size_t* counter_bytes);
/* Host-only self-check of the work decomposition the launch would use for this geometry: every unit's blocks covered
 * exactly once, per-CTA block limit, unique partial slots inside the workspace stride, merge ownership.  Returns
 * MFB200_OK or MFB200_EINVAL (mfb200_last_error() names the first inconsistency).  Needs no GPU when sm_count > 0. */
int mfb200_decode_plan_check(int batch, int kv_heads, int groups, int comp_len, int win_len, int sm_count, int plan_hint);
/* Scheduling assumption (short launches only): when every compressed CTA of a launch is resident at once and holds <= 16
 * blocks, the split merge uses a "flagged" protocol in which the CTA with the HIGHEST index of a unit waits (bounded
 * back-off polling of L2) for partials written by lower-indexed CTAs of the same launch.  This relies on the hardware
 * dispatching the CTAs of a 1-D grid in ascending blockIdx order, which every CUDA GPU to date does but CUDA does not
 * promise; under MPS time-slicing or a debugger that single-steps CTAs the wait can be long (never wrong).  All other
 * launches use an atomic ticket (last arrival merges) and wait for nothing. */
int mfb200_sparse_decode_attention(const mfb200_decode_params* p, mfb200_stream_t stream);

/* One decode step on a long-lived parameter block (the host keeps one per layer cache): sets q / out /
 * k_new / v_new, advances p->win_len by one (the new token) and
 * launches mfb200_sparse_decode_attention.  p->workspace must hold mfb200_decode_workspace_max() bytes.
 * Exists so that a per-layer decode step costs the host a single FFI call. */
int mfb200_decode_step(mfb200_decode_params* p, const void* q, const void* k_new, const void* v_new, void* out,
                       int sm_count, mfb200_stream_t stream);
/* The same for a whole decoder: layer l uses layers[l] and q + l*q_layer_stride, k_new/v_new + l*kv_layer_stride,
 * out + l*out_layer_stride (strides in halves).  One FFI call per decode step instead of one per layer; the launches are
 * issued back to back on `stream`.  Returns n_layers, or the first error (layers before it were launched and advanced,
 * the failing layer and the ones after it are untouched). */
int mfb200_decode_step_layers(mfb200_decode_params* const* layers, int n_layers, const void* q, const void* k_new,
                              const void* v_new, void* out, int64_t q_layer_stride, int64_t kv_layer_stride,
                              int64_t out_layer_stride, mfb200_stream_t stream);
/* Graph-capturable decode step: launches every layer with its parameter block AS IT IS (q / out / k_new / v_new preset,
 * win_len_dev set, win_len = the window capacity) - nothing on the host changes from step to step, so the sequence
 *     mfb200_lengths_add(lengths, n_layers, 1) ; mfb200_decode_layers_static(layers, n_layers)
 * can be captured once and replayed for every decode step between two compression events (llama_mustafar_kernel.py:324:
 * every 256 tokens the caller compresses, resets the device lengths and captures again). */
int mfb200_decode_layers_static(const mfb200_decode_params* const* layers, int n_layers, mfb200_stream_t stream);
/* lengths[i] += delta for i < n (device int32 array; one tiny launch). */
int mfb200_lengths_add(int32_t* lengths, int n, int delta, mfb200_stream_t stream);
/* Workspace size that is sufficient for every (comp_len <= max_comp_len, win_len <= max_win_len), plan_hint 0. */
size_t mfb200_decode_workspace_max(int batch, int kv_heads, int groups, int max_comp_len, int max_win_len,
                                   int sm_count);

/* ---- peer-to-peer plumbing of the head-sharded path ------------------------------------------------------
 * mfb200_peer_wait: one small launch that returns (in stream order) once flags[0 .. n) >= epoch, i.e. once every rank's
 * rows of this step have landed in the local gathered buffer.  It gives up after about two seconds and sets *timed_out
 * (device int32, may be NULL) instead of hanging the stream when a peer died; once set, later waits return at once.
 * mfb200_peer_alloc/free: cudaMalloc'ed (IPC-exportable) memory; mfb200_ipc_export/open/close: cudaIpc handles (64 bytes)
 * so that the ranks of ONE node can map each other's buffers; exchange the handles with any host-side channel. */
int mfb200_peer_wait(const uint32_t* flags, int n, uint32_t epoch, int32_t* timed_out, mfb200_stream_t stream);
int mfb200_peer_alloc(size_t bytes, void** ptr);
int mfb200_peer_free(void* ptr);
int mfb200_ipc_export(void* ptr, unsigned char handle[64]);
int mfb200_ipc_open(const unsigned char handle[64], void** ptr);
int mfb200_ipc_close(void* ptr);

/* ---- window append (new token's k and v rows) ----------------------------------------------------
 * win[u, pos, :] = row[u, :] for K and V; row: fp16 [units, 128]. */
int mfb200_window_append(void* k_win, void* v_win, int64_t win_stride, const void* k_row,
                         const void* v_row, int64_t units, int64_t pos, mfb200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MUSTAFAR_B200_H */
