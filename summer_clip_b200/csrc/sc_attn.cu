// C ABI of the fused CLIP-search attention kernels (include/summer_clip_b200.h) and their host-side helpers:
//     O[q, c] = sum_k exp(beta * (Qn[q].Kn[k] - shift[q])) * V[k, c]
// which is reference  cache_weights_strategy.py:34-35  (A = Q^T K ; W = exp(-beta (1 - A)))
// followed by         image_attention.py:109           (W @ V)
// and the Tip-Adapter head  tip_adapter/utils.py:114-116.   The [Nq, Nk] matrix never leaves the SM.
// The kernels live in sc_attn_seg.cu (one-hot values on a label-sorted bank; also the split-fp16 GEMM, the row-max
// pre-pass and the softmax mode) and sc_attn_t.cu (dense values, GEMM-1 + GEMM-2); this file validates arguments,
// builds the TMA descriptors and dispatches.
#include "sc_common.cuh"
#include "sc_ptx.cuh"

#include <cuda.h>
#include <cstdlib>
#include <mutex>

namespace sc {
// sc_attn_t.cu
int attn_t_launch(int (*make_tmap)(CUtensorMap*, const void*, int64_t, int64_t, int64_t, int, bool),
                  const void* Qn, const void* Kn, const void* Vt, bool f16, int64_t Nq, int64_t Nk, int64_t D_pad,
                  int64_t n_cols, int64_t C_pad, int64_t Nk_pad, int slice, int64_t n_slices, float beta,
                  const float* row_shift, int splits, float* O, int64_t ldo, cudaStream_t st);
// sc_attn_seg.cu
int attn_seg_splits(int64_t Nq, int64_t Nks, int sm_count, int64_t row_bytes, int64_t n_classes, int n_betas);
int attn_seg_launch(int (*make_tmap)(CUtensorMap*, const void*, int64_t, int64_t, int64_t, int, bool),
                    int (*make_tmap_u8)(CUtensorMap*, const void*, int64_t, int64_t, int64_t, int),
                    const void* Qn, const void* Ks, const int16_t* gcls, const uint32_t* kbits, int op_dtype, int64_t Nq,
                    int64_t Nks, int64_t D_pad, const float* betas, int n_betas, int splits, float* O, int64_t ldo,
                    int softmax, cudaStream_t st);
int attn_rowmax_launch(int (*make_tmap)(CUtensorMap*, const void*, int64_t, int64_t, int64_t, int, bool),
                       int (*make_tmap_u8)(CUtensorMap*, const void*, int64_t, int64_t, int64_t, int),
                       const void* Qn, const void* Kn, int op_dtype, int64_t Nq, int64_t Nk, int64_t D_pad, int splits,
                       float* rowmax, cudaStream_t st);
int rowconf_fused_launch(int (*make_tmap)(CUtensorMap*, const void*, int64_t, int64_t, int64_t, int, bool),
                         const void* Ah, const void* Al, const void* Bh, const void* Bl, int64_t M, int64_t N,
                         int64_t D_pad, float scale, const float* row_scale, float prob_scale, int prob, float* conf,
                         int* label, cudaStream_t st);
int gemm_split_launch(int (*make_tmap)(CUtensorMap*, const void*, int64_t, int64_t, int64_t, int, bool),
                      const void* Ah, const void* Al, const void* Bh, const void* Bl, int64_t M, int64_t N,
                      int64_t D_pad, float scale, const float* row_scale, float* Z, int64_t ldz, cudaStream_t st);
}  // namespace sc

namespace {

using namespace scptx;

constexpr int kBM = 128;            // queries per tile of the dense kernel
constexpr int kBN = 128;            // keys per tile of the dense kernel
constexpr int kBK = 64;             // 16-bit elements per swizzled smem row

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
  });
  return fn;
}

// 16-bit row-major [rows, cols] with row pitch `pitch_elems`; box = [box_rows x 64 cols], SW128.
int make_tmap(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t pitch_elems,
              int box_rows, bool is_f16) {
  EncodeTiledFn fn = get_encode_fn();
  SC_REQUIRE(fn != nullptr, SC_EDRIVER, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(pitch_elems) * 2u};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, is_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SC_REQUIRE(r == CUDA_SUCCESS, SC_EDRIVER, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return SC_OK;
}

// e4m3 operand rows: bytes as elements, 128-byte (128-element) boxes, same 128-byte swizzle
int make_tmap_u8(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t pitch_elems, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  SC_REQUIRE(fn != nullptr, SC_EDRIVER, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(pitch_elems)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(2 * kBK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SC_REQUIRE(r == CUDA_SUCCESS, SC_EDRIVER, "cuTensorMapEncodeTiled (u8) failed with CUresult %d", (int)r);
  return SC_OK;
}

// number of class slices: 2 or a multiple of 4 — the slices of one query tile are the CTAs of one cluster of the
// dense kernel (a CTA pair, or two pairs per 1024 classes); few classes still make a pair (two narrow slices)
int64_t n_class_slices(int64_t C) {
  const int64_t c16 = sc::round_up(C, 16);
  if (c16 <= 512) return 2;
  return 4 * sc::ceil_div(c16, 1024);
}
int class_slice(int64_t C) {
  const int64_t c16 = sc::round_up(C, 16);
  return static_cast<int>(sc::round_up(sc::ceil_div(c16, n_class_slices(C)), 16));
}

}  // namespace

extern "C" {

int64_t sc_pad_dim(int64_t D) { return sc::round_up(D, 64); }
int64_t sc_pad_dim_op(int64_t D, int op_dtype) { return sc::round_up(D, op_dtype == SC_E4M3 ? 128 : 64); }
int64_t sc_pad_keys(int64_t Nk) { return sc::round_up(Nk, 8); }
int64_t sc_class_slice(int64_t C) { return class_slice(C); }
int64_t sc_pad_classes(int64_t C) {
  // n evenly sized slices; idempotent: sc_class_slice(sc_pad_classes(C)) == sc_class_slice(C)
  return n_class_slices(C) * class_slice(C);
}

int sc_attn_splits(int64_t Nq, int64_t Nk, int64_t C_pad, int sm_count);

// Key splits of the dense-values kernel with L2 blocking: work items are launched split-major (blockIdx.z slowest), so
// while the query tiles pass over one key range its K and Vt bytes are read from DRAM once and then served by L2 —
// if they fit.  Query-tile clusters that start at different times (every wave after the first) otherwise stream the
// bank on their own schedule and the DRAM traffic grows ~15x (measured: profiles/r02n_dense_dram_vs_splits.log).
// Block = 64 MB (half of the 126 MB L2: the two dies cache a line both read twice).  Measured at 12.5k queries x
// 1.28M keys: 5 / 40 / 80 / 160 / 320 splits -> 150 / 126 / 41 / 10 / 10 GB of DRAM reads and 76.6 / 73.2 / 70.7 / 72.6 /
// 79.4 ms: smaller blocks read less but pay more per-item prologue and one more [Nq, C] partial tile each.
int sc_attn_splits_for(int64_t Nq, int64_t Nk, int64_t D_pad, int64_t C_pad, int sm_count) {
  const int base = sc_attn_splits(Nq, Nk, C_pad, sm_count);
  if (Nq <= 0 || Nk <= 0 || D_pad <= 0 || C_pad <= 0) return base;
  if (sm_count <= 0) sm_count = 148;
  const int64_t clusters = sc::ceil_div(Nq, kBM);                                   // query tiles
  const int64_t resident = sm_count / (n_class_slices(C_pad) >= 4 ? 4 : 2);        // clusters in flight
  const double bank_bytes = static_cast<double>(Nk) * static_cast<double>(D_pad + C_pad) * 2.0;
  if (clusters < 2 * resident || bank_bytes <= 48.0e6) return base;                 // one wave, or the bank fits anyway
  const int64_t tiles = sc::ceil_div(Nk, kBN);
  int64_t want = static_cast<int64_t>(bank_bytes / 64.0e6) + 1;
  if (want > tiles) want = tiles;
  if (want > 65535) want = 65535;
  return want > base ? static_cast<int>(want) : base;
}

int sc_attn_splits(int64_t Nq, int64_t Nk, int64_t C_pad, int sm_count) {
  if (Nq <= 0 || Nk <= 0 || C_pad <= 0) return 1;
  if (sm_count <= 0) sm_count = 148;
  const int64_t base = sc::ceil_div(Nq, kBM) * n_class_slices(C_pad);
  const int64_t tiles = sc::ceil_div(Nk, kBN);
  // cost model: waves * (key tiles per CTA + fixed prologue/epilogue expressed in tiles)
  int best = 1;
  double best_cost = 1e300;
  const int64_t smax = tiles < 64 ? tiles : 64;
  for (int64_t s = 1; s <= smax; ++s) {
    const double waves = static_cast<double>(sc::ceil_div(base * s, sm_count));
    const double cost = waves * (static_cast<double>(sc::ceil_div(tiles, s)) + 3.0);
    if (cost < best_cost * 0.995) { best_cost = cost; best = static_cast<int>(s); }
  }
  return best;
}

int64_t sc_pad_queries(int64_t Nq) { return sc::round_up(Nq > 0 ? Nq : 1, 2 * kBM); }
int64_t sc_pad_labels(int64_t Nk) { return sc::round_up(Nk > 0 ? Nk : 1, kBN); }

int sc_attn_hard_supported(int64_t n_classes) { return (n_classes > 0 && n_classes <= 32767) ? 1 : 0; }

int sc_attn_hard_splits(int64_t Nq, int64_t Nks, int sm_count) { return sc::attn_seg_splits(Nq, Nks, sm_count, 0, 0, 1); }
int sc_attn_hard_splits_for(int64_t Nq, int64_t Nks, int64_t D_pad, int op_dtype, int64_t n_classes, int n_betas,
                            int sm_count) {
  return sc::attn_seg_splits(Nq, Nks, sm_count, D_pad * (op_dtype == SC_E4M3 ? 1 : 2), n_classes, n_betas);
}

static int attn_hard_common(const void* Qn, const void* Ks, const int16_t* group_class, const uint32_t* key_bits,
                            int op_dtype, int64_t Nq, int64_t Nks, int64_t D_pad, int64_t n_classes,
                            const float* betas, int n_betas, int splits, float* O, int64_t ldo, int softmax,
                            void* stream) {
  SC_REQUIRE(Qn && Ks && group_class && key_bits && O && betas, SC_EINVAL, "sc_attn_fwd_hard: null pointer");
  SC_REQUIRE(n_betas >= 1 && n_betas <= 4, SC_ESHAPE, "sc_attn_fwd_hard_multi: n_betas=%d must be in [1, 4]", n_betas);
  SC_REQUIRE(op_dtype == SC_F16 || op_dtype == SC_BF16 || op_dtype == SC_E4M3, SC_EINVAL,
             "sc_attn_fwd_hard: op_dtype must be SC_F16, SC_BF16 or SC_E4M3");
  SC_REQUIRE(Nq > 0 && Nks > 0 && n_classes > 0, SC_ESHAPE, "sc_attn_fwd_hard: empty problem");
  SC_REQUIRE(sc_attn_hard_supported(n_classes), SC_EUNSUPPORTED, "sc_attn_fwd_hard: n_classes=%lld exceeds int16 labels",
             (long long)n_classes);
  SC_REQUIRE(D_pad > 0 && D_pad % (op_dtype == SC_E4M3 ? 128 : 64) == 0, SC_ESHAPE,
             "sc_attn_fwd_hard: D_pad=%lld must be a multiple of %d", (long long)D_pad, op_dtype == SC_E4M3 ? 128 : 64);
  SC_REQUIRE(ldo >= n_classes, SC_ESHAPE, "sc_attn_fwd_hard: ldo < n_classes");
  SC_REQUIRE((reinterpret_cast<uintptr_t>(Qn) | reinterpret_cast<uintptr_t>(Ks) |
              reinterpret_cast<uintptr_t>(group_class) | reinterpret_cast<uintptr_t>(key_bits)) % 32 == 0,
             SC_EALIGN, "sc_attn_fwd_hard: Qn, Ks, group_class and key_bits must be 32-byte aligned");
  SC_REQUIRE(Nq < (1ll << 31) && Nks < (1ll << 31) - 512, SC_ESHAPE, "sc_attn_fwd_hard: Nq/Nks exceed int32 coordinates");
  const int64_t steps_total = sc::ceil_div(Nks, 256);
  if (splits <= 0) {
    int dev = 0, sms = 148;
    SC_CUDA(cudaGetDevice(&dev));
    SC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    splits = sc::attn_seg_splits(Nq, Nks, sms, D_pad * (op_dtype == SC_E4M3 ? 1 : 2), n_classes, n_betas);
  }
  SC_REQUIRE(splits <= steps_total && splits <= 65535, SC_ESHAPE,
             "sc_attn_fwd_hard: splits=%d exceeds the %lld key steps", splits, (long long)steps_total);
  int rc = sc::attn_seg_launch(&make_tmap, &make_tmap_u8, Qn, Ks, group_class, key_bits, op_dtype, Nq, Nks, D_pad, betas,
                               n_betas, splits, O, ldo, softmax, static_cast<cudaStream_t>(stream));
  if (rc != SC_OK) return rc;
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

int sc_attn_fwd_hard_multi(const void* Qn, const void* Ks, const int16_t* group_class, const uint32_t* key_bits,
                           int op_dtype, int64_t Nq, int64_t Nks, int64_t D_pad, int64_t n_classes,
                           const float* betas, int n_betas, int splits, float* O, int64_t ldo, void* stream) {
  return attn_hard_common(Qn, Ks, group_class, key_bits, op_dtype, Nq, Nks, D_pad, n_classes, betas, n_betas, splits, O,
                          ldo, 0, stream);
}

int sc_attn_softmax_hard(const void* Qn, const void* Ks, const int16_t* group_class, const uint32_t* key_bits,
                         int op_dtype, int64_t Nq, int64_t Nks, int64_t D_pad, int64_t n_classes, float tau, int splits,
                         float* LSE, int64_t ldo, void* stream) {
  return attn_hard_common(Qn, Ks, group_class, key_bits, op_dtype, Nq, Nks, D_pad, n_classes, &tau, 1, splits, LSE, ldo, 1,
                          stream);
}

int sc_attn_rowmax(const void* Qn, const void* Kn, int op_dtype, int64_t Nq, int64_t Nk, int64_t D_pad, float* rowmax,
                   void* stream) {
  SC_REQUIRE(Qn && Kn && rowmax, SC_EINVAL, "sc_attn_rowmax: null pointer");
  SC_REQUIRE(op_dtype == SC_F16 || op_dtype == SC_BF16 || op_dtype == SC_E4M3, SC_EINVAL,
             "sc_attn_rowmax: op_dtype must be SC_F16, SC_BF16 or SC_E4M3");
  SC_REQUIRE(Nq > 0 && Nk > 0, SC_ESHAPE, "sc_attn_rowmax: empty problem");
  SC_REQUIRE(D_pad > 0 && D_pad % (op_dtype == SC_E4M3 ? 128 : 64) == 0, SC_ESHAPE,
             "sc_attn_rowmax: D_pad=%lld must be a multiple of %d", (long long)D_pad, op_dtype == SC_E4M3 ? 128 : 64);
  SC_REQUIRE((reinterpret_cast<uintptr_t>(Qn) | reinterpret_cast<uintptr_t>(Kn)) % 16 == 0, SC_EALIGN,
             "sc_attn_rowmax: Qn and Kn must be 16-byte aligned");
  SC_REQUIRE(Nq < (1ll << 31) && Nk < (1ll << 31) - 512, SC_ESHAPE, "sc_attn_rowmax: Nq/Nk exceed int32 coordinates");
  int dev = 0, sms = 148;
  SC_CUDA(cudaGetDevice(&dev));
  SC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int splits = sc::attn_seg_splits(Nq, Nk, sms, D_pad * (op_dtype == SC_E4M3 ? 1 : 2), 0, 1);
  int rc = sc::attn_rowmax_launch(&make_tmap, &make_tmap_u8, Qn, Kn, op_dtype, Nq, Nk, D_pad, splits, rowmax,
                                  static_cast<cudaStream_t>(stream));
  if (rc != SC_OK) return rc;
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

int sc_attn_fwd_hard(const void* Qn, const void* Ks, const int16_t* group_class, const uint32_t* key_bits,
                     int op_dtype, int64_t Nq, int64_t Nks, int64_t D_pad, int64_t n_classes, float beta, int splits,
                     float* O, int64_t ldo, void* stream) {
  return sc_attn_fwd_hard_multi(Qn, Ks, group_class, key_bits, op_dtype, Nq, Nks, D_pad, n_classes, &beta, 1, splits, O,
                                ldo, stream);
}

static int gemm_nt(const char* who, const void* Ah, const void* Al, const float* row_scale, const void* Bh, const void* Bl,
                   int64_t M, int64_t N, int64_t D_pad, float scale, float* Z, int64_t ldz, void* stream) {
  SC_REQUIRE(Ah && Bh && Bl && Z, SC_EINVAL, "%s: null pointer", who);
  SC_REQUIRE(M > 0 && N > 0 && ldz >= N, SC_ESHAPE, "%s: bad shape", who);
  SC_REQUIRE(D_pad > 0 && D_pad % 64 == 0, SC_ESHAPE, "%s: D_pad=%lld must be a multiple of 64", who, (long long)D_pad);
  SC_REQUIRE((reinterpret_cast<uintptr_t>(Ah) | reinterpret_cast<uintptr_t>(Al) | reinterpret_cast<uintptr_t>(Bh) |
              reinterpret_cast<uintptr_t>(Bl)) % 16 == 0,
             SC_EALIGN, "%s: operands must be 16-byte aligned", who);
  SC_REQUIRE(M < (1ll << 31) && N < (1ll << 31) - 512, SC_ESHAPE, "%s: M/N exceed int32 coordinates", who);
  int rc = sc::gemm_split_launch(&make_tmap, Ah, Al, Bh, Bl, M, N, D_pad, scale, row_scale, Z, ldz,
                                 static_cast<cudaStream_t>(stream));
  if (rc != SC_OK) return rc;
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

int sc_gemm_split_nt(const void* Ah, const void* Al, const void* Bh, const void* Bl, int64_t M, int64_t N,
                     int64_t D_pad, float scale, float* Z, int64_t ldz, void* stream) {
  SC_REQUIRE(Al != nullptr, SC_EINVAL, "sc_gemm_split_nt: null pointer");
  return gemm_nt("sc_gemm_split_nt", Ah, Al, nullptr, Bh, Bl, M, N, D_pad, scale, Z, ldz, stream);
}

int sc_gemm_rows_nt(const void* A, const float* row_scale, const void* Bh, const void* Bl, int64_t M, int64_t N,
                    int64_t D_pad, float scale, float* Z, int64_t ldz, void* stream) {
  return gemm_nt("sc_gemm_rows_nt", A, nullptr, row_scale, Bh, Bl, M, N, D_pad, scale, Z, ldz, stream);
}

static int rowconf_fused(const char* who, const void* Ah, const void* Al, const float* row_scale, const void* Bh,
                         const void* Bl, int64_t M, int64_t C, int64_t D_pad, float scale, float prob_scale, int mode,
                         float* conf, int32_t* label, void* stream) {
  SC_REQUIRE(Ah && Bh && Bl && conf && label, SC_EINVAL, "%s: null pointer", who);
  SC_REQUIRE(M > 0 && C > 0, SC_ESHAPE, "%s: bad shape", who);
  SC_REQUIRE(mode == SC_CONF_RAW || mode == SC_CONF_PROB, SC_EINVAL, "%s: bad mode %d", who, mode);
  SC_REQUIRE(D_pad > 0 && D_pad % 64 == 0, SC_ESHAPE, "%s: D_pad=%lld must be a multiple of 64", who, (long long)D_pad);
  SC_REQUIRE((reinterpret_cast<uintptr_t>(Ah) | reinterpret_cast<uintptr_t>(Al) | reinterpret_cast<uintptr_t>(Bh) |
              reinterpret_cast<uintptr_t>(Bl)) % 16 == 0,
             SC_EALIGN, "%s: operands must be 16-byte aligned", who);
  SC_REQUIRE(M < (1ll << 31) && C < (1ll << 31) - 512, SC_ESHAPE, "%s: M/C exceed int32 coordinates", who);
  int rc = sc::rowconf_fused_launch(&make_tmap, Ah, Al, Bh, Bl, M, C, D_pad, scale, row_scale, prob_scale,
                                    mode == SC_CONF_PROB ? 1 : 0, conf, label, static_cast<cudaStream_t>(stream));
  if (rc != SC_OK) return rc;
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

int sc_rowconf_from_split(const void* Ah, const void* Al, const void* Bh, const void* Bl, int64_t M, int64_t C,
                          int64_t D_pad, float scale, float prob_scale, int mode, float* conf, int32_t* label,
                          void* stream) {
  SC_REQUIRE(Al != nullptr, SC_EINVAL, "sc_rowconf_from_split: null pointer");
  return rowconf_fused("sc_rowconf_from_split", Ah, Al, nullptr, Bh, Bl, M, C, D_pad, scale, prob_scale, mode, conf, label,
                       stream);
}

int sc_rowconf_from_rows(const void* A, const float* row_scale, const void* Bh, const void* Bl, int64_t M, int64_t C,
                         int64_t D_pad, float scale, float prob_scale, int mode, float* conf, int32_t* label,
                         void* stream) {
  return rowconf_fused("sc_rowconf_from_rows", A, nullptr, row_scale, Bh, Bl, M, C, D_pad, scale, prob_scale, mode, conf,
                       label, stream);
}

int sc_attn_fwd_shifted(const void* Qn, const void* Kn, const void* Vt, int op_dtype, int64_t Nq, int64_t Nk,
                        int64_t D_pad, int64_t n_cols, int64_t C_pad, int64_t Nk_pad, float beta, const float* row_shift,
                        int splits, float* O, int64_t ldo, void* stream) {
  SC_REQUIRE(Qn && Kn && Vt && O, SC_EINVAL, "sc_attn_fwd: null pointer");
  SC_REQUIRE(op_dtype == SC_F16 || op_dtype == SC_BF16, SC_EINVAL, "sc_attn_fwd: op_dtype must be SC_F16 or SC_BF16");
  SC_REQUIRE(Nq > 0 && Nk > 0 && n_cols > 0, SC_ESHAPE, "sc_attn_fwd: empty problem");
  const bool f16 = (op_dtype == SC_F16);
  SC_REQUIRE(D_pad > 0 && D_pad % 64 == 0, SC_ESHAPE, "sc_attn_fwd: D_pad=%lld must be a multiple of 64",
             (long long)D_pad);
  SC_REQUIRE(Nk_pad >= Nk && Nk_pad % 8 == 0, SC_ESHAPE, "sc_attn_fwd: Nk_pad=%lld must be >= Nk and a multiple of 8",
             (long long)Nk_pad);
  const int slice = class_slice(C_pad);
  const int64_t n_slices = n_class_slices(C_pad);
  SC_REQUIRE(n_slices * slice == C_pad && n_cols <= C_pad, SC_ESHAPE,
             "sc_attn_fwd: C_pad=%lld is not a whole number of %d-wide class slices (use sc_pad_classes)",
             (long long)C_pad, slice);
  SC_REQUIRE(ldo >= n_cols, SC_ESHAPE, "sc_attn_fwd: ldo < n_cols");
  SC_REQUIRE((reinterpret_cast<uintptr_t>(Qn) | reinterpret_cast<uintptr_t>(Kn) |
              reinterpret_cast<uintptr_t>(Vt)) % 16 == 0,
             SC_EALIGN, "sc_attn_fwd: operand pointers must be 16-byte aligned");
  SC_REQUIRE(Nq < (1ll << 31) && Nk < (1ll << 31) - 256, SC_ESHAPE, "sc_attn_fwd: Nq/Nk exceed int32 coordinates");

  const int64_t tiles_total = sc::ceil_div(Nk, kBN);
  if (splits <= 0) {
    int dev = 0, sms = 148;
    SC_CUDA(cudaGetDevice(&dev));
    SC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    splits = sc_attn_splits_for(Nq, Nk, D_pad, C_pad, sms);
  }
  SC_REQUIRE(splits <= tiles_total && splits <= 65535, SC_ESHAPE,
             "sc_attn_fwd: splits=%d exceeds the %lld key tiles", splits, (long long)tiles_total);
  int rc = sc::attn_t_launch(&make_tmap, Qn, Kn, Vt, f16, Nq, Nk, D_pad, n_cols, C_pad, Nk_pad, slice, n_slices, beta,
                             row_shift, splits, O, ldo, static_cast<cudaStream_t>(stream));
  if (rc != SC_OK) return rc;
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

int sc_attn_fwd(const void* Qn, const void* Kn, const void* Vt, int op_dtype, int64_t Nq, int64_t Nk,
                int64_t D_pad, int64_t n_cols, int64_t C_pad, int64_t Nk_pad, float beta,
                int splits, float* O, int64_t ldo, void* stream) {
  return sc_attn_fwd_shifted(Qn, Kn, Vt, op_dtype, Nq, Nk, D_pad, n_cols, C_pad, Nk_pad, beta, nullptr, splits, O, ldo,
                             stream);
}

}  // extern "C"
