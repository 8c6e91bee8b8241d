/*
 * libsummerclip_b200 — C ABI of the B200-native CLIP-search hot path.
 *
 * Replaces, for the one path BASELINE.json names, the torch calls made by
 * myrachins/summer-clip (citations are paths under the reference tree):
 *   summer_clip/clip_searcher/cache_weights_strategy.py:18-36   (normalise, Q^T K, exp(-b(1-A)))
 *   summer_clip/clip_searcher/cache_value_strategy.py:14-28     (one-hot / softmax cache values)
 *   summer_clip/clip_searcher/cache_strategy.py:48-81           (per-class top-k pseudo-labels)
 *   summer_clip/clip_searcher/image_attention.py:80-83,107-111  (zero-shot logits, W@V, Z+alpha*O)
 *   summer_clip/clip_searcher/utils.py:15-21, clip_adapter/train_adapter.py:156-159 (top-1/5)
 *   summer_clip/tip_adapter/utils.py:10-15,99-129               (Tip-Adapter head, search_hp)
 * The reference has no FFI: its boundary is Python strategy classes.  The Python mirror of those
 * classes (summer_clip_b200/clip_searcher/...) binds these symbols with ctypes; INTEGRATION.md
 * shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless stated otherwise;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default);
 *   - return value: 0 on success, negative SC_E* for argument errors, positive cudaError_t
 *     otherwise; sc_last_error() returns a thread-local message for the last failure;
 *   - no call allocates caller-visible memory or keeps global mutable state;
 *   - strides / leading dimensions are in ELEMENTS.
 */
#ifndef SUMMER_CLIP_B200_H_
#define SUMMER_CLIP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 3 = 2 + sc_rowconf_from_rows, sc_gemm_rows_nt, sc_transpose_norms (additive); sc_topk_per_class accepts an empty
 * shard (N == 0, NULL rows).  Nothing of version 2 changed its signature or meaning. */
#define SC_ABI_VERSION 3

/* element types of caller buffers.  SC_E4M3 (8-bit float, 4 exponent / 3 mantissa bits) is an OPERAND type only:
 * sc_normalize_cast can write it and sc_attn_fwd_hard[_multi] can read it (tcgen05 kind::f8f6f4, fp32 accumulate:
 * half the bank bytes and twice the contraction length per instruction of the 16-bit types).  e4m3 rows hold
 * SC_E4M3_SCALE * x, so that unit-norm feature rows (|x| <= 1, typically 1/sqrt(D)) sit in the normal range;
 * the attention kernel divides the scale back out.  Reduced precision: opt-in, see DESIGN.md for measured error. */
enum { SC_F16 = 0, SC_BF16 = 1, SC_F32 = 2, SC_E4M3 = 3 };
#define SC_E4M3_SCALE 256.0f
/* sc_rowconf modes: rank rows by the raw maximum (TopKStrategy, cache_strategy.py:67-70) or by
 * the maximum of softmax(scale * row) (TopKProbStrategy, cache_strategy.py:79-81). */
enum { SC_CONF_RAW = 0, SC_CONF_PROB = 1 };
/* sc_values_prepare modes: HardCacheStrategy (cache_value_strategy.py:14-17) or
 * SoftmaxCacheStrategy (cache_value_strategy.py:26-28). */
enum { SC_VALUES_HARD = 0, SC_VALUES_SOFTMAX = 1 };
/* error codes */
enum { SC_OK = 0, SC_EINVAL = -1, SC_EALIGN = -2, SC_ESHAPE = -3, SC_EUNSUPPORTED = -4, SC_EDRIVER = -5 };

int sc_version(void);
const char* sc_last_error(void);

/* Geometry helpers (host only, no CUDA calls).  The attention kernel consumes
 *   Qn [Nq, D_pad], Kn [Nk, D_pad] (bf16 or fp16; rows L2-normalised, K-major),
 *   Vt [C_pad, Nk_pad] (same type; cache values TRANSPOSED, zero padded),
 * where D_pad = sc_pad_dim(D), Nk_pad = sc_pad_keys(Nk), C_pad = sc_pad_classes(C). */
int64_t sc_pad_dim(int64_t D);          /* multiple of 64 */
int64_t sc_pad_dim_op(int64_t D, int op_dtype);  /* row length for an operand type: 128-byte chunks (64 16-bit / 128 e4m3) */
int64_t sc_pad_keys(int64_t Nk);        /* multiple of 8  */
int64_t sc_pad_queries(int64_t Nq);     /* multiple of 256: rows the QUERY operand of sc_attn_fwd_hard[_multi], sc_attn_softmax_hard
                                           and sc_attn_rowmax must have allocated (rows Nq .. pad-1: any finite values, e.g.
                                           zeros; they are read, never written or reported) */
int64_t sc_pad_classes(int64_t C);      /* n_slices * slice width (2 or 4k slices; width a multiple of 16, <= 256) */
int64_t sc_class_slice(int64_t C);      /* slice width the kernel uses for C classes */

/* Column L2-normalise + gather + transpose + cast (cache_weights_strategy.py:19-20 fused with the
 * column gather K[:, idx] of image_attention.py:55).
 *   src: element (d, n) at src[d*stride_d + n*stride_n], d < D, n < N, dtype src_dtype.
 *   idx: optional int64[n_out] column indices (NULL: n_out must equal N, identity).
 *   dst: [n_out, D_pad] of dst_dtype (SC_BF16 or SC_F16 — the tensor-core operand type, see
 *        sc_attn_fwd — or SC_E4M3 with D_pad = sc_pad_dim_op(D, SC_E4M3), values scaled by SC_E4M3_SCALE,
 *        round-to-nearest, for sc_attn_fwd_hard); columns D..D_pad-1 are written as zero.
 *   normalize: 1 = divide by the column's L2 norm (fp32), 0 = cast only. */
int sc_normalize_cast(const void* src, int src_dtype, int64_t D, int64_t N, int64_t stride_d,
                      int64_t stride_n, const int64_t* idx, int64_t n_out, void* dst, int dst_dtype,
                      int64_t D_pad, int normalize, void* stream);

/* Tail of Tip-Adapter's build_cache_model (tip_adapter/utils.py:59-60): cache keys = row-L2-normalised mean over
 * the E augment epochs of the encoder features.  src element (e, n, d) at src[e*stride_e + n*stride_n + d], dtype
 * SC_F16 / SC_BF16 / SC_F32; dst [N, ld_dst] of the SAME type.  The mean, the norm and the quotient are rounded to
 * the storage type one after the other, as the reference's tensor expressions are; sums in fp32.  The reference then
 * stores the [D, N] permuted VIEW of this matrix (utils.py:61): pass it to sc_normalize_cast with stride_d = 1. */
int sc_mean_normalize_rows(const void* src, int dtype, int64_t E, int64_t N, int64_t D, int64_t stride_e,
                           int64_t stride_n, void* dst, int64_t ld_dst, void* stream);

/* Per-row confidence and predicted label of a logits bank L[N, C] (leading dim ld):
 *   SC_CONF_RAW : conf = max_c L, label = first argmax          (cache_strategy.py:68)
 *   SC_CONF_PROB: conf = max_c softmax(scale * L), same label   (cache_strategy.py:80, :68)
 * Arithmetic is fp32 whatever the storage type. */
int sc_rowconf(const void* L, int dtype, int64_t N, int64_t C, int64_t ld, float scale, int mode,
               float* conf, int32_t* label, void* stream);

/* select_topk_per_label (cache_strategy.py:48-59): for every class c in [0, C) the indices of
 * the min(k, n_c) most confident rows whose label is c, most confident first; ties are broken by
 * the smaller row index.  out_idx is int64 [C, k] (unused slots = -1), out_count int32 [C].
 * The reference's concatenated result is, for c ascending, out_idx[c, :out_count[c]].  N == 0 (an empty row shard;
 * conf / label may be NULL) gives all -1.  Rows sharded over ranks: each rank's [C, k] candidates (confidence, global
 * row) are all-gathered and merged under the same order (summer_clip_b200/selection.py). */
size_t sc_topk_workspace_bytes(int64_t N, int32_t C);
int sc_topk_per_class(const float* conf, const int32_t* label, int64_t N, int32_t C, int32_t k,
                      int64_t* out_idx, int32_t* out_count, void* workspace, size_t ws_bytes,
                      void* stream);

/* Cache values V = f(L[idx]) written TRANSPOSED as Vt[C_pad, Nk_pad] (vt_dtype SC_BF16 or SC_F16,
 * zero padded):
 *   SC_VALUES_HARD   : one_hot(argmax_c L)                      (cache_value_strategy.py:15-16)
 *   SC_VALUES_SOFTMAX: softmax(scale * L, dim=1), scale = clip_scale*scale  (:27)
 * idx optional int64[n_out] row gather (image_attention.py:55, L[idx]); labels_override optional
 * int32[n_out]: if given (HARD mode) it replaces the argmax (replace_outs_with_golds,
 * image_attention.py:65-66; Tip-Adapter one-hot cache values, tip_adapter/utils.py:62).
 * ones_row >= 0 additionally sets Vt[ones_row, k] = 1 for k < n_out (row sums for softmax mode). */
int sc_values_prepare(const void* L, int dtype, int64_t N, int64_t C, int64_t ld,
                      const int64_t* idx, const int32_t* labels_override, int64_t n_out, int mode,
                      float scale, void* Vt, int vt_dtype, int64_t C_pad, int64_t Nk_pad,
                      int64_t ones_row, void* stream);

/* Fused attention (cache_weights_strategy.py:34-35 + image_attention.py:109, never
 * materialising the [Nq, Nk] matrix):
 *     O[s, q, c] = sum_{k in split s} exp(beta * (Qn[q].Kn[k] - 1)) * Vt[c, k]
 * for c < n_cols (n_cols <= C_pad).  Qn, Kn, Vt (and the on-chip weights P) all have type op_dtype:
 * SC_BF16 or SC_F16, both with fp32 accumulation at the same tensor-core rate; every operand of
 * this path lies in [-1, 1], where fp16 carries 3 more mantissa bits than bf16.
 * O is fp32 [splits, Nq, ldo]; the caller sums the splits
 * (sc_merge_partials).  splits >= 1 partitions the key tiles so that small query batches still
 * fill the GPU; splits = 0 lets the library choose (query with sc_attn_splits). */
int sc_attn_splits(int64_t Nq, int64_t Nk, int64_t C_pad, int sm_count);
/* ... and with L2 blocking (what splits = 0 uses): for banks larger than L2 scored by several waves of query tiles the
 * key range of a split is sized so that its K + Vt bytes stay L2-resident while every query tile passes over it
 * (work items launch split-major); costs one [Nq, ldo] fp32 partial tile per split. */
int sc_attn_splits_for(int64_t Nq, int64_t Nk, int64_t D_pad, int64_t C_pad, int sm_count);
int sc_attn_fwd(const void* Qn, const void* Kn, const void* Vt, int op_dtype, int64_t Nq, int64_t Nk,
                int64_t D_pad, int64_t n_cols, int64_t C_pad, int64_t Nk_pad, float beta,
                int splits, float* O, int64_t ldo, void* stream);
/* The same with a per-query reference in place of the constant 1:
 *     O[s, q, c] = sum_{k in split s} exp(beta * (Qn[q].Kn[k] - row_shift[q])) * Vt[c, k]
 * (row_shift NULL = sc_attn_fwd).  With row_shift[q] = max_k Qn[q].Kn[k] (sc_attn_rowmax) every row's largest
 * weight is exactly 1 whatever beta is, which is what the temperature-softmax mode needs before the weights are
 * rounded to 16 bits: O together with a ones row in Vt (sc_values_prepare's ones_row) is the (m, l, O) partial
 * triple of sc_merge_softmax with m = beta * log2(e) * row_shift. */
int sc_attn_fwd_shifted(const void* Qn, const void* Kn, const void* Vt, int op_dtype, int64_t Nq, int64_t Nk,
                        int64_t D_pad, int64_t n_cols, int64_t C_pad, int64_t Nk_pad, float beta, const float* row_shift,
                        int splits, float* O, int64_t ldo, void* stream);
/* rowmax[q] = max_{k < Nk} Qn[q].Kn[k] in fp32 (tensor-core GEMM-1 of the attention kernel with a running maximum
 * in place of the exponential sum; the [Nq, Nk] matrix is not materialised).  Kn is any [Nk, D_pad] bank of op_dtype
 * (SC_F16, SC_BF16 or SC_E4M3), in any key order. */
int sc_attn_rowmax(const void* Qn, const void* Kn, int op_dtype, int64_t Nq, int64_t Nk, int64_t D_pad, float* rowmax,
                   void* stream);

/* Hard-label cache values for sc_attn_fwd_hard: labels16[k] = argmax_c L[idx[k]] (or labels_override[k]) as
 * int16 for k < n_out, and -1 for n_out <= k < n_pad and for labels outside [0, C) — i.e. the one-hot cache
 * values of HardCacheStrategy (cache_value_strategy.py:14-17), of cache.replace_outs_with_golds
 * (image_attention.py:65-66) and of Tip-Adapter's cache_values (tip_adapter/utils.py:62) in 2 bytes per key.
 * n_pad = sc_pad_labels(n_out) (a multiple of the kernel's 128-key tile). */
int64_t sc_pad_labels(int64_t Nk);
int sc_hard_labels(const void* L, int dtype, int64_t N, int64_t C, int64_t ld, const int64_t* idx,
                   const int32_t* labels_override, int64_t n_out, int16_t* labels16, int64_t n_pad,
                   void* stream);

/* Layout of the label-sorted key bank sc_attn_fwd_hard consumes, from the int16 labels of sc_hard_labels:
 * a STABLE counting sort of the keys by class (original order inside a class, so results are reproducible bit
 * for bit) with every class segment padded to whole 16-key groups.  Outputs, all sized by
 * capacity = sc_hard_bank_capacity(n_keys, n_classes) (a multiple of 256, >= n_keys + 15 n_classes):
 *   perm        int64  [capacity]       original index of sorted key j, -1 = padding;
 *   group_class int16  [capacity / 16]  class of every 16-key group, -1 = none;
 *   key_bits    uint32 [capacity / 32]  bit j of word w = 1 iff sorted key 32 w + j is real;
 *   n_sorted    int64  [1] (device)     keys in the padded bank (a multiple of 16): Nks of sc_attn_fwd_hard.
 * Labels outside [0, n_classes) select no class and are dropped.  sc_gather_rows then builds the bank itself:
 * dst[j] = src[perm[j]] for rows of row_bytes (a multiple of 16), zero rows where perm[j] < 0. */
int64_t sc_hard_bank_capacity(int64_t n_keys, int32_t n_classes);
size_t sc_hard_bank_workspace_bytes(int64_t n_keys, int32_t n_classes);
int sc_hard_bank_layout(const int16_t* labels16, int64_t n_keys, int32_t n_classes, int64_t* perm,
                        int16_t* group_class, uint32_t* key_bits, int64_t capacity, int64_t* n_sorted,
                        void* workspace, size_t ws_bytes, void* stream);
int sc_gather_rows(const void* src, int64_t n_src, int64_t row_bytes, const int64_t* perm, int64_t n_out,
                   void* dst, void* stream);
/* The one-pass alternative to normalise + sc_gather_rows: inv[k] = sorted position of original key k (-1 if the
 * layout dropped it), and — when rows != NULL — zero rows at the padding positions of the sorted bank
 * rows [n_sorted_rows, row_bytes]; sc_normalize_scatter then writes every key's normalised row straight to
 * dst[dst_row[o]] (dst_row = inv, or inv gathered by the selection idx), so the bank crosses HBM once. */
int sc_hard_bank_inverse(const int64_t* perm, int64_t n_sorted_rows, int64_t n_keys, int64_t* inv, void* rows,
                         int64_t row_bytes, void* stream);
/* sc_normalize_cast with a scattered destination: output o (o < n_out) is written to row dst_row[o] of dst
 * (skipped when dst_row[o] < 0).  Same sources, types and arithmetic as sc_normalize_cast. */
int sc_normalize_scatter(const void* src, int src_dtype, int64_t D, int64_t N, int64_t stride_d, int64_t stride_n,
                         const int64_t* idx, int64_t n_out, const int64_t* dst_row, void* dst, int dst_dtype,
                         int64_t D_pad, int normalize, void* stream);
/* sc_normalize_cast that also returns the inverse column norms: inv_norm[o] = 1 / |src[:, idx[o]]| (fp32, computed
 * from the source values whether or not `normalize` is set).  normalize = 0 gives the transposed RAW bank plus the
 * factors that normalise it later (sc_rowconf_from_rows). */
int sc_transpose_norms(const void* src, int src_dtype, int64_t D, int64_t N, int64_t stride_d, int64_t stride_n,
                       const int64_t* idx, int64_t n_out, void* dst, int dst_dtype, int64_t D_pad, int normalize,
                       float* inv_norm, void* stream);

/* sc_attn_fwd for one-hot cache values on a LABEL-SORTED key bank (same result as sc_attn_fwd on
 * Vt = one_hot(label)^T; the sum over keys does not depend on their order):
 *     O[s, q, c] = sum_{k in split s, class(k) == c} exp(beta * (Qn[q].Ks[k] - 1)),   c < n_classes.
 * With one-hot values W @ V is a per-class segmented row sum of the weights.  The caller permutes the
 * normalised bank once so that keys of one class are adjacent and every class segment starts on a 16-key
 * boundary (padding rows: anything finite, e.g. zeros):
 *   Qn          [sc_pad_queries(Nq), D_pad]   the normalised queries in whole 256-row tiles (rows >= Nq: finite padding;
 *                                             a partly out-of-bounds TMA box costs a single-tile launch 30 % of its HBM rate);
 *   Ks          [Nks, D_pad]                  the permuted, padded bank (op_dtype);
 *   group_class int16  [ceil(Nks/256) * 16]   class of every 16-key group, -1 = no real key in it;
 *   key_bits    uint32 [ceil(Nks/256) * 8]    bit j of word w = 1 iff sorted key 32 w + j is a real key
 *                                             (padding keys get weight 0).
 * GEMM-1 (Q.K^T) runs as 256 x 256 CTA-pair tensor-core tiles; the exponentials are summed per class straight
 * out of tensor memory in fp32 — the weights are never rounded, stored or exchanged.  O is fp32
 * [splits, Nq, ldo]; the library zeroes it and writes only the classes each split meets (sum the splits with
 * sc_merge_partials).  splits = 0: the library chooses (sc_attn_hard_splits_for).  Any n_classes <= 32767.
 * sc_attn_hard_splits fills the SM pairs by tensor-core work alone; sc_attn_hard_splits_for also charges every
 * extra split the zeroing and reading back of its [Nq, n_classes] tile per beta, which decides for small banks
 * (Tip-Adapter's 16 000 keys: 1 split instead of 3). */
int sc_attn_hard_supported(int64_t n_classes);
int sc_attn_hard_splits(int64_t Nq, int64_t Nks, int sm_count);
int sc_attn_hard_splits_for(int64_t Nq, int64_t Nks, int64_t D_pad, int op_dtype, int64_t n_classes, int n_betas,
                            int sm_count);
int sc_attn_fwd_hard(const void* Qn, const void* Ks, const int16_t* group_class, const uint32_t* key_bits,
                     int op_dtype, int64_t Nq, int64_t Nks, int64_t D_pad, int64_t n_classes, float beta,
                     int splits, float* O, int64_t ldo, void* stream);
/* The same for 1..4 betas in ONE pass (`betas` is a HOST array): S = Q.K^T does not depend on beta, so a beta
 * sweep (image_attention.yaml: 8 per cache; tip_adapter/utils.py:99-129: 200) pays the tensor-core GEMM once
 * per group of betas and only the exponentials per beta.  O is fp32 [n_betas, splits, Nq, ldo]. */
int sc_attn_fwd_hard_multi(const void* Qn, const void* Ks, const int16_t* group_class, const uint32_t* key_bits,
                           int op_dtype, int64_t Nq, int64_t Nks, int64_t D_pad, int64_t n_classes,
                           const float* betas, int n_betas, int splits, float* O, int64_t ldo, void* stream);

/* Temperature-softmax mode on a label-sorted bank (north-star extension; the reference has only the un-normalised
 * Tip-Adapter weights): softmax over the KEYS of tau * Qn.Ks^T with an ONLINE running row
 * maximum, summed per class.  Same operands, tiles and splits as sc_attn_fwd_hard; the result is the per-class
 * log-sum-exp in base 2,
 *     LSE[s, q, c] = log2 sum_{k in split s, class(k) == c} 2^(tau * log2(e) * Qn[q].Ks[k]),   -inf if there is none,
 * i.e. the (m, l) pair of every class in one float — independent of the maximum it was accumulated against, so
 * classes, key splits and key shards combine exactly (sc_softmax_partials, sc_merge_softmax).  The library fills
 * the tile with -inf first. */
int sc_attn_softmax_hard(const void* Qn, const void* Ks, const int16_t* group_class, const uint32_t* key_bits,
                         int op_dtype, int64_t Nq, int64_t Nks, int64_t D_pad, int64_t n_classes, float tau, int splits,
                         float* LSE, int64_t ldo, void* stream);
/* LSE tiles [n_parts, Nq, ld] (parts part_stride floats apart) -> the (m, l, O) partial triple of these keys:
 *     m[q] = max_{p, c} LSE,   O[q, c] = sum_p 2^(LSE[p, q, c] - m[q]),   l[q] = sum_c O[q, c]
 * (m = -inf, O = 0, l = 0 for a row without keys).  O / l is the softmax attention output over these keys. */
int sc_softmax_partials(const float* LSE, int n_parts, int64_t part_stride, int64_t Nq, int64_t C, int64_t ld,
                        float* O, int64_t ld_out, float* m, float* l, void* stream);
/* Log-sum-exp merge of (m, l, O) partial triples (key splits, key-sharded ranks):
 *     M[q] = m_ref ? m_ref[q] : max_p m_scale * m_p[q],    w_p = 2^(m_scale * m_p[q] - M[q]),
 *     out[q, c] = sum_p w_p O_p[q, c],   L[q] = sum_p w_p l_p[q],   out /= L if normalize.
 * O_parts fp32 [n_parts, Nq, ld] (parts o_part_stride floats apart), m_parts / l_parts [n_parts, Nq] (parts
 * ml_part_stride apart).  m_scale converts the stored m to base-2 exponents: 1 for sc_softmax_partials' output,
 * tau * log2(e) for a row maximum of cosines (sc_attn_rowmax + sc_attn_fwd_shifted).  m_ref (nullable) = a common
 * maximum agreed between ranks (all-reduce MAX of M), so that the rescaled O and L can be summed by a reduce-scatter.
 * out may alias O_parts when n_parts == 1; m_out / l_out nullable. */
int sc_merge_softmax(const float* O_parts, const float* m_parts, const float* l_parts, int n_parts,
                     int64_t o_part_stride, int64_t ml_part_stride, int64_t Nq, int64_t C, int64_t ld, float m_scale,
                     const float* m_ref, int normalize, float* out, int64_t ld_out, float* m_out, float* l_out,
                     void* stream);

/* out[r, c] = sum_p parts[p, r, c]  (key splits and key-sharded ranks; with the Tip weights
 * exp(beta(A-1)) <= 1 the running maximum of an LSE merge is the constant 0, so the merge of
 * partial (m, l, O) triples is a plain sum).  parts: fp32 [n_parts, rows, ld]. `out` may alias
 * parts[0]. */
int sc_merge_partials(const float* parts, int n_parts, int64_t rows, int64_t cols, int64_t ld,
                      float* out, int64_t ld_out, void* stream);

/* The key-sharded exchange as ONE pass over peer memory: out[r, c] = sum_p parts[p][r * ld + c], where parts is a
 * HOST array of n_parts <= 16 device pointers that may belong to OTHER GPUs of the node (peer-mapped over NVLink,
 * e.g. the buffers of a symmetric-memory allocation): every rank reads the rows of its query slice straight out of
 * every rank's partial tile — no staging copy, no NCCL kernel, no shared memory, so the pass runs beside the
 * attention CTAs of the next query block.  The caller orders it after the peers' writes (a device-side barrier).
 * Parts are added in index order: every rank computes bit-identical sums. */
int sc_merge_peer_parts(const float* const* parts, int n_parts, int64_t rows, int64_t cols, int64_t ld, float* out,
                        int64_t ld_out, void* stream);

/* Zero-shot logits Z = scale * normalise_cols(X)^T @ T in fp32 (image_attention.py:80-83; with
 * scale = 1 also the pseudo-label logits bank of save_image_outs.py:25).
 *   X element (d, n) at X[d*stride_d + n*stride_n]; T is [D, C] row-major with leading dim ldt. */
int sc_zero_shot_logits(const void* X, int x_dtype, int64_t D, int64_t N, int64_t stride_d,
                        int64_t stride_n, const void* T, int t_dtype, int64_t C, int64_t ldt,
                        float scale, int normalize, float* Z, int64_t ldz, void* stream);

/* Tensor-core route of the same product (image_attention.py:80-83; the logits-bank producer
 * save_image_outs.py:25): fp32-accurate although it runs on fp16 tensor cores.
 *   sc_normalize_split: (optionally column-normalised) src, any layout / dtype as in sc_normalize_cast, written
 *     as an fp16 PAIR hi = fp16(v), lo = fp16(v - hi), both [N, D_pad] K-major (22 significant bits);
 *   sc_gemm_split_nt:   Z[m, n] = scale * sum_d (Ah[m,d] Bh[n,d] + Ah[m,d] Bl[n,d] + Al[m,d] Bh[n,d])
 *     — 256 x 256 CTA-pair tcgen05 tiles, three operand passes into one TMEM accumulator; the dropped lo.lo
 *     term is below 2^-22 |a||b|.  Z is fp32 [M, ldz]; A is [M, D_pad], B is [N, D_pad].
 * Zero-shot logits: A = split(normalised queries), B = split(T^T), scale = 100. */
int sc_normalize_split(const void* src, int src_dtype, int64_t D, int64_t N, int64_t stride_d,
                       int64_t stride_n, void* hi, void* lo, int64_t D_pad, int normalize, void* stream);
int sc_gemm_split_nt(const void* Ah, const void* Al, const void* Bh, const void* Bl, int64_t M, int64_t N,
                     int64_t D_pad, float scale, float* Z, int64_t ldz, void* stream);
/* The same product for A rows that are EXACT in fp16 (raw fp16 feature banks: sc_transpose_norms gives the transposed
 * rows and row_scale[m] = 1 / |row m|): two operand passes,
 *     Z[m, n] = scale * row_scale[m] * sum_d A[m,d] (Bh[n,d] + Bl[n,d]);   row_scale may be NULL (= 1). */
int sc_gemm_rows_nt(const void* A, const float* row_scale, const void* Bh, const void* Bl, int64_t M, int64_t N,
                    int64_t D_pad, float scale, float* Z, int64_t ldz, void* stream);

/* Pseudo-labels WITHOUT the logits bank (save_image_outs.py:25 fused with cache_strategy.py:67-70 / :79-81): the
 * rows of L = scale * A @ B^T (A = split(normalised image features) [M, D_pad], B = split(T^T) [C, D_pad], as for
 * sc_gemm_split_nt: fp32-accurate on fp16 tensor cores) are reduced inside the kernel to what sc_rowconf returns,
 *     label[m] = first argmax_c L[m, c],   conf[m] = max_c L[m, c]  (SC_CONF_RAW)
 *                                          conf[m] = max_c softmax(prob_scale * L[m, :])  (SC_CONF_PROB),
 * so that the [M, C] bank (5 GB fp32 at ImageNet scale) is never written or read.  Feed sc_topk_per_class. */
int sc_rowconf_from_split(const void* Ah, const void* Al, const void* Bh, const void* Bl, int64_t M, int64_t C,
                          int64_t D_pad, float scale, float prob_scale, int mode, float* conf, int32_t* label,
                          void* stream);
/* The same for image features that are EXACT in fp16 (the reference's cached fp16 banks): A [M, D_pad] holds the
 * RAW rows (sc_transpose_norms: transposed, not normalised) and row_scale[m] = 1 / |row m|, so there is no lo part
 * and two operand passes suffice:  L[m, c] = scale * row_scale[m] * sum_d A[m,d] (Bh[c,d] + Bl[c,d]).
 * row_scale may be NULL (rows already unit length and exact in fp16). */
int sc_rowconf_from_rows(const void* A, const float* row_scale, const void* Bh, const void* Bl, int64_t M, int64_t C,
                         int64_t D_pad, float scale, float prob_scale, int mode, float* conf, int32_t* label,
                         void* stream);

/* Epilogue (image_attention.py:111-112, clip_searcher/utils.py:15-21, tip_adapter/utils.py:10-15):
 * for every alpha a: out = Z + O * alpha  (O optionally divided by rowsum[q] first), prediction =
 * first argmax, and — if labels are given — the number of rows whose label is the top-1 / within
 * the top-5 (ties: larger value first, then smaller class index).  `alphas` is a HOST array of
 * na <= 64 floats.  Z may be NULL (treated as 0).  out_logits (nullable) is fp32 [na, Nq, C];
 * pred (nullable) int32 [na, Nq]; top1/top5 (nullable) int32 [na], ACCUMULATED into (caller zeroes). */
int sc_epilogue(const float* Z, int64_t ldz, const float* O, int64_t ldo, const float* rowsum,
                int64_t Nq, int64_t C, const float* alphas, int na, const int32_t* labels,
                float* out_logits, int32_t* pred, int32_t* top1, int32_t* top5, void* stream);

/* sc_epilogue on UNMERGED partial tiles: O is fp32 [n_parts, Nq, ldo] with the parts part_stride floats apart
 * (what sc_attn_fwd_hard / _multi write for splits > 1); every element is summed over the parts in
 * sc_merge_partials' order, so the results are bit-identical to sc_merge_partials followed by sc_epilogue —
 * without the extra pass over HBM.  Replaces the same reference lines as sc_epilogue. */
int sc_epilogue_parts(const float* Z, int64_t ldz, const float* O, int64_t ldo, int n_parts, int64_t part_stride,
                      const float* rowsum, int64_t Nq, int64_t C, const float* alphas, int na,
                      const int32_t* labels, float* out_logits, int32_t* pred, int32_t* top1, int32_t* top5,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SUMMER_CLIP_B200_H_ */
