// Fused CLIP-search attention for sm_100a:
//     O[q, c] = sum_k exp(beta * (Qn[q].Kn[k] - 1)) * V[k, c]
// which is reference  cache_weights_strategy.py:34-35  (A = Q^T K ; W = exp(-beta (1 - A)))
// followed by         image_attention.py:109           (W @ V)
// and the Tip-Adapter head  tip_adapter/utils.py:114-116.   The [Nq, Nk] matrix never leaves the SM.
//
// Work decomposition.  A CTA owns a 128-query tile and one class slice (<= 256 classes: the fp32 O
// accumulator of 128 x 1000 does not fit the 512 TMEM columns of one SM).  The G (1, 2 or 4) CTAs
// that hold the class slices of the SAME query tile form a thread-block cluster and split the key
// tiles between them: CTA g computes S and the weights P only for key tiles g, g+G, g+2G, ... and
// broadcasts each bf16/fp16 P tile (32 KB) to its peers' shared memory with DSMEM bulk copies, so
// Q.K^T is computed once per (query tile, key tile) instead of once per class slice.  Per 128-key tile:
//   GEMM-1  S[128q x 128k] = Qn_tile . Kn_tile^T   tcgen05.mma, operands TMA-staged (SW128), fp32 in TMEM
//   exp     P = exp2(c1*S + c0)                     4 warps: tcgen05.ld -> ex2 -> 16-bit -> smem (+ peers)
//   GEMM-2  O[128q x slice] += P . V_tile           tcgen05.mma, A = P slot (smem), B = Vt tile (TMA),
//                                                   fp32 accumulator resident in TMEM for the whole pass
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2-5 = exp/epilogue.
// All operand tiles are K-major rows of 64 16-bit elements (128 B) with the 128-byte swizzle.
#include "sc_common.cuh"
#include "sc_ptx.cuh"

#include <cuda.h>
#include <cstdlib>
#include <mutex>

namespace sc {
// sc_attn_pair.cu
int attn_pair_launch(int (*make_tmap)(CUtensorMap*, const void*, int64_t, int64_t, int64_t, int, bool),
                     const void* Qn, const void* Kn, const void* Vt, bool f16, int64_t Nq, int64_t Nk,
                     int64_t D_pad, int64_t n_cols, int64_t C_pad, int64_t Nk_pad, int slice, int64_t n_slices,
                     float beta, int splits, float* O, int64_t ldo, cudaStream_t st);
// sc_attn_t.cu
int attn_t_launch(int (*make_tmap)(CUtensorMap*, const void*, int64_t, int64_t, int64_t, int, bool),
                  const void* Qn, const void* Kn, const void* Vt, bool f16, int64_t Nq, int64_t Nk, int64_t D_pad,
                  int64_t n_cols, int64_t C_pad, int64_t Nk_pad, int slice, int64_t n_slices, float beta,
                  int splits, float* O, int64_t ldo, cudaStream_t st);
// sc_attn_seg.cu
int attn_seg_splits(int64_t Nq, int64_t Nks, int sm_count, int64_t row_bytes, int64_t n_classes, int n_betas);
int attn_seg_launch(int (*make_tmap)(CUtensorMap*, const void*, int64_t, int64_t, int64_t, int, bool),
                    int (*make_tmap_u8)(CUtensorMap*, const void*, int64_t, int64_t, int64_t, int),
                    const void* Qn, const void* Ks, const int16_t* gcls, const uint32_t* kbits, int op_dtype, int64_t Nq,
                    int64_t Nks, int64_t D_pad, const float* betas, int n_betas, int splits, float* O, int64_t ldo,
                    cudaStream_t st);
int gemm_split_launch(int (*make_tmap)(CUtensorMap*, const void*, int64_t, int64_t, int64_t, int, bool),
                      const void* Ah, const void* Al, const void* Bh, const void* Bl, int64_t M, int64_t N,
                      int64_t D_pad, float scale, float* Z, int64_t ldz, cudaStream_t st);
}  // namespace sc

namespace {

using namespace scptx;

constexpr int kBM = 128;            // queries per CTA (UMMA M)
constexpr int kBN = 128;            // keys per S tile (UMMA N of GEMM-1 / K extent of GEMM-2)
constexpr int kBK = 64;             // 16-bit elements per swizzled smem row
constexpr int kUnits = 7;           // 32 KB smem units: (7 - G) operand ring stages + G weight slots
constexpr int kStageBytes = 32768;  // GEMM-1: Q chunk 16 KB + K chunk 16 KB; GEMM-2: Vt chunk <= 32 KB
constexpr int kPBytes = 32768;      // one 16-bit P tile [128 x 128] as two [128 x 64] swizzled halves
constexpr int kThreads = 192;
constexpr int kExpThreads = 128;
constexpr int kTmemCols = 512;      // S0 @0, S1 @128, O @256 (<= 256 columns)
constexpr int kColS = 0;
constexpr int kColO = 256;
constexpr int kMaxStages = 6;
constexpr int kMaxCluster = 4;
constexpr int kSmemBytes = kUnits * 32768 + 1024 /*align*/ + 256 /*barriers*/;

struct AttnParams {
  int Nq;
  int n_dchunks;     // D_pad / 64
  int n_cols;        // valid output columns (<= C_pad)
  int slice;         // class-slice width = UMMA N of GEMM-2 (multiple of 16, <= 256)
  int tiles_total;   // ceil(Nk / 128)
  int splits;
  int ns;            // operand ring stages in use (<= 7 - G)
  int pf_dist;       // L2 prefetch distance in rounds (0 = off)
  int dbg_skip;      // TIMING EXPERIMENTS ONLY (wrong results): bit0 skip Q loads, bit1 skip V loads, bit2 skip K loads
  float c1;          // beta * log2(e)
  float c0;          // exponent offset: -c1 (+ kPShift for fp16 operands)
  float o_scale;     // 2^-kPShift undoes the offset in the epilogue
  float* O;          // [splits, Nq, ldo]
  long long ldo;
};

struct Bars {
  uint64_t full[kMaxStages];     // TMA bytes landed in ring stage
  uint64_t empty[kMaxStages];    // MMAs reading ring stage retired
  uint64_t s_full[2];            // GEMM-1 accumulator ready
  uint64_t s_empty[2];           // exp warps drained the accumulator (128 arrivals)
  uint64_t p_full[kMaxCluster];  // weight slot of source CTA g' holds the tile of the current round
  uint64_t p_empty;              // all G consumers retired GEMM-2 on MY last tile (G arrivals)
  uint64_t o_full;
  uint32_t tmem_slot;
};

template <bool kF16>
__device__ __forceinline__ uint32_t pack_16x2(float lo, float hi) {
  uint32_t r;
  if (kF16)
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// fp16 weights are stored as 2^kPShift * exp(beta (A - 1)) <= 2^kPShift * (1 + eps): the offset moves
// small weights out of the fp16 subnormal range at no cost (folded into the exponent FMA) and is
// removed from the fp32 accumulator in the epilogue.
constexpr float kPShift = 8.0f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool kF16, int G>
__global__ void __launch_bounds__(kThreads, 1)
sc_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  constexpr int NSmax = kUnits - G;              // operand ring capacity
  const int NS = p.ns;                           // stages in use (tuning knob, <= NSmax)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t ring0 = (raw_addr + 1023u) & ~1023u;       // 1024-B aligned (SW128 atoms)
  const uint32_t pbuf0 = ring0 + NSmax * kStageBytes;          // G weight slots, one per source CTA
  Bars* bars = reinterpret_cast<Bars*>(smem_raw + (pbuf0 - raw_addr) + G * kPBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int g = (G > 1) ? static_cast<int>(cluster_ctarank()) : 0;   // == blockIdx.x % G

  const int c0 = blockIdx.x * p.slice;          // first class of this slice
  const int q0 = blockIdx.y * kBM;              // first query of this tile
  const int split = blockIdx.z;
  const int t0 = static_cast<int>((static_cast<long long>(p.tiles_total) * split) / p.splits);
  const int t1 = static_cast<int>((static_cast<long long>(p.tiles_total) * (split + 1)) / p.splits);
  const int T = t1 - t0;                        // key tiles of this split (shared by the cluster)
  const int R = (T + G - 1) / G;                // rounds: round r holds tiles r*G + g', g' < G
  const int nd = p.n_dchunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    for (int s = 0; s < NSmax; ++s) {
      mbar_init(smem_u32(&bars->full[s]), 1);
      mbar_init(smem_u32(&bars->empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bars->s_full[b]), 1);
      mbar_init(smem_u32(&bars->s_empty[b]), kExpThreads);
    }
    for (int s = 0; s < G; ++s) mbar_init(smem_u32(&bars->p_full[s]), 1);
    mbar_init(smem_u32(&bars->p_empty), G);
    mbar_init(smem_u32(&bars->o_full), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&bars->tmem_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (G > 1) cluster_sync_all();                 // peers' barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t v_bytes = static_cast<uint32_t>(p.slice) * kBK * 2u;
      auto load_v_round = [&](int rr) {
#pragma unroll 1
        for (int gp = 0; gp < G; ++gp) {
          const int i = rr * G + gp;
          if (i >= T) break;
#pragma unroll 1
          for (int c = 0; c < kBN / kBK; ++c) {
            const uint32_t fb = smem_u32(&bars->full[stage]);
            mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1u);
            if (p.dbg_skip & 2) { mbar_arrive(fb); } else {
            mbar_arrive_expect_tx(fb, v_bytes);
            tma_load_2d(ring0 + stage * kStageBytes, &tmV, fb, (t0 + i) * kBN + c * kBK, c0); }
            if (++stage == NS) { stage = 0; phase ^= 1u; }
          }
        }
      };
      // L2 prefetch of the K tile and Vt slices this CTA will load `pf_dist` rounds from now: the
      // demand loads then complete at L2-hit latency whatever the drift between clusters is.
      auto prefetch_round = [&](int pr) {
        if (pr >= R) return;
        const int ip = pr * G + g;
        if (ip < T) {
#pragma unroll 1
          for (int d = 0; d < nd; ++d) tma_prefetch_2d(&tmK, d * kBK, (t0 + ip) * kBN);
        }
#pragma unroll 1
        for (int gp = 0; gp < G; ++gp) {
          const int i = pr * G + gp;
          if (i >= T) break;
          tma_prefetch_2d(&tmV, (t0 + i) * kBN, c0);
          tma_prefetch_2d(&tmV, (t0 + i) * kBN + kBK, c0);
        }
      };
      if (p.pf_dist > 0)
        for (int pr = 0; pr < p.pf_dist; ++pr) prefetch_round(pr);
#pragma unroll 1
      for (int r = 0; r < R; ++r) {
        const int i_own = r * G + g;
        if (p.pf_dist > 0) prefetch_round(r + p.pf_dist);
        if (i_own < T) {
          const int tile = t0 + i_own;
#pragma unroll 1
          for (int d = 0; d < nd; ++d) {
            const uint32_t fb = smem_u32(&bars->full[stage]);
            const uint32_t dst = ring0 + stage * kStageBytes;
            mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1u);
            const uint32_t qb = (p.dbg_skip & 1) ? 0u : 16384u, kb = (p.dbg_skip & 4) ? 0u : 16384u;
            if (qb + kb) mbar_arrive_expect_tx(fb, qb + kb); else mbar_arrive(fb);
            if (qb) tma_load_2d(dst, &tmQ, fb, d * kBK, q0);
            if (kb) tma_load_2d(dst + 16384, &tmK, fb, d * kBK, tile * kBN);
            if (++stage == NS) { stage = 0; phase ^= 1u; }
          }
        }
        if (r > 0) load_v_round(r - 1);
      }
      if (R > 0) load_v_round(R - 1);
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (one thread)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t idesc1 = umma_idesc_16b(kBM, kBN, kF16);
      const uint32_t idesc2 = umma_idesc_16b(kBM, static_cast<uint32_t>(p.slice), kF16);
      const uint32_t tmem_o = tmem_base + kColO;
      if (G > 1) {
        // arm the slots fed by peers for round 0 (1 arrival + 32 KB of complete_tx from the peer's copy)
        for (int gp = 0; gp < G; ++gp)
          if (gp != g && gp < T) mbar_arrive_expect_tx(smem_u32(&bars->p_full[gp]), kPBytes);
      }
      auto gemm2_round = [&](int rr) {
#pragma unroll 1
        for (int gp = 0; gp < G; ++gp) {
          const int i = rr * G + gp;
          if (i >= T) break;
          mbar_wait(smem_u32(&bars->p_full[gp]), rr & 1);
          tc_fence_after();
          if (G > 1 && gp != g && i + G < T)      // the slot's next phase: arm it before it can be refilled
            mbar_arrive_expect_tx(smem_u32(&bars->p_full[gp]), kPBytes);
#pragma unroll 1
          for (int c = 0; c < kBN / kBK; ++c) {
            mbar_wait(smem_u32(&bars->full[stage]), phase);
            tc_fence_after();
            const uint32_t a_addr = pbuf0 + gp * kPBytes + c * 16384;
            const uint32_t b_addr = ring0 + stage * kStageBytes;
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              if (!(p.dbg_skip & 32)) umma_ss(tmem_o, umma_desc_k128(a_addr + k * 32), umma_desc_k128(b_addr + k * 32),
                      idesc2, (i | c | k) != 0 ? 1u : 0u);
            }
            umma_commit(smem_u32(&bars->empty[stage]));
            if (++stage == NS) { stage = 0; phase ^= 1u; }
          }
          // tell the SOURCE of this tile that one more consumer is done with it
          if (G > 1)
            umma_commit_mcast(smem_u32(&bars->p_empty), static_cast<uint16_t>(1u << gp));
          else
            umma_commit(smem_u32(&bars->p_empty));
        }
      };
      int own = 0;
#pragma unroll 1
      for (int r = 0; r < R; ++r) {
        if (r * G + g < T) {
          const int sb = own & 1;
          mbar_wait(smem_u32(&bars->s_empty[sb]), ((own >> 1) & 1) ^ 1u);
          tc_fence_after();
          const uint32_t tmem_s = tmem_base + kColS + sb * kBN;
#pragma unroll 1
          for (int d = 0; d < nd; ++d) {
            mbar_wait(smem_u32(&bars->full[stage]), phase);
            tc_fence_after();
            const uint32_t a_addr = ring0 + stage * kStageBytes;
            const uint32_t b_addr = a_addr + 16384;
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              if (!(p.dbg_skip & 16)) umma_ss(tmem_s, umma_desc_k128(a_addr + k * 32), umma_desc_k128(b_addr + k * 32),
                      idesc1, (d | k) != 0 ? 1u : 0u);
            }
            umma_commit(smem_u32(&bars->empty[stage]));
            if (++stage == NS) { stage = 0; phase ^= 1u; }
          }
          umma_commit(smem_u32(&bars->s_full[sb]));
          ++own;
        }
        if (r > 0) gemm2_round(r - 1);
      }
      if (R > 0) gemm2_round(R - 1);
      umma_commit(smem_u32(&bars->o_full));
    }
  } else {
    // ===================================================== exp warps (+ epilogue)
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;          // query row within the tile
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const float c1 = p.c1;
    const float cadd = p.c0;
    const float o_scale = p.o_scale;
    const uint32_t row_off = static_cast<uint32_t>(row) * 128u;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    const uint32_t my_slot = pbuf0 + g * kPBytes;
    int own = 0;
#pragma unroll 1
    for (int r = 0; r < R; ++r) {
      if (r * G + g >= T) break;
      const int b = own & 1;
      mbar_wait(smem_u32(&bars->s_full[b]), (own >> 1) & 1);
      tc_fence_after();
      mbar_wait(smem_u32(&bars->p_empty), (own & 1) ^ 1u);   // every consumer retired my previous tile
#pragma unroll
      for (int cc = 0; cc < ((p.dbg_skip & 8) ? 0 : kBN / 32); ++cc) {
        uint32_t rg[32];
        tmem_ld_32x32(tmem_base + lane_addr + kColS + b * kBN + cc * 32, rg);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float e0 = ex2_approx(fmaf(__uint_as_float(rg[2 * j]), c1, cadd));
          const float e1 = ex2_approx(fmaf(__uint_as_float(rg[2 * j + 1]), c1, cadd));
          pk[j] = pack_16x2<kF16>(e0, e1);
        }
        // keys cc*32 .. cc*32+31 of this row -> half (cc>>1), 16-byte chunks (cc&1)*4 .. +3
        const uint32_t half_base = my_slot + static_cast<uint32_t>(cc >> 1) * 16384u + row_off;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t chunk = static_cast<uint32_t>((cc & 1) * 4 + j);
          const uint32_t addr = half_base + ((chunk ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * j]),
                       "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                       : "memory");
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->s_empty[b]));
      fence_proxy_async_smem();                 // my P writes -> visible to UMMA and to the bulk copies
      named_bar_sync(1, kExpThreads);           // the whole tile is written
      if (threadIdx.x == 64) {
        const uint32_t pf = smem_u32(&bars->p_full[g]);
        mbar_arrive(pf);                        // local consumer
        if (G > 1) {
#pragma unroll
          for (int gp = 0; gp < G; ++gp) {
            if (gp == g) continue;
            bulk_copy_to_peer(mapa(my_slot, gp), my_slot, kPBytes, mapa(pf, gp));
          }
        }
      }
      ++own;
    }
    // ---- epilogue: O slice TMEM -> global partial
    mbar_wait(smem_u32(&bars->o_full), 0);
    tc_fence_after();
    const int q = q0 + row;
    float* orow = p.O + (static_cast<long long>(split) * p.Nq + q) * p.ldo + c0;
    const int ncol_here = min(p.slice, p.n_cols - c0);   // may be <= 0 for an all-padding slice
#pragma unroll 1
    for (int cc = 0; cc < p.slice / 16; ++cc) {
      uint32_t rg[16];
      tmem_ld_32x16(tmem_base + lane_addr + kColO + cc * 16, rg);
      tmem_ld_wait();
      if (q < p.Nq) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int c = cc * 16 + j;
          if (c < ncol_here) orow[c] = (T > 0) ? __uint_as_float(rg[j]) * o_scale : 0.0f;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (G > 1) cluster_sync_all();   // no CTA may exit while a peer's commit / copy can still target it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

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

// number of class slices: 1, 2 or a multiple of 4 (so the slices of one query tile fill whole clusters)
int64_t n_class_slices(int64_t C) {
  const int64_t c16 = sc::round_up(C, 16);
  if (c16 <= 256) return 1;
  if (c16 <= 512) return 2;
  return 4 * sc::ceil_div(c16, 1024);
}
int class_slice(int64_t C) {
  const int64_t c16 = sc::round_up(C, 16);
  return static_cast<int>(sc::round_up(sc::ceil_div(c16, n_class_slices(C)), 16));
}

template <bool kF16, int G>
int launch(dim3 grid, cudaStream_t st, const CUtensorMap& tmQ, const CUtensorMap& tmK,
           const CUtensorMap& tmV, const AttnParams& p) {
  auto kernel = sc_attn_kernel<kF16, G>;
  // per-device attribute; setting it on every call is a few hundred ns and keeps the call stateless
  SC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = G;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SC_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmQ, tmK, tmV, p));
  return SC_OK;
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

int64_t sc_pad_labels(int64_t Nk) { return sc::round_up(Nk > 0 ? Nk : 1, kBN); }

int sc_attn_hard_supported(int64_t n_classes) { return (n_classes > 0 && n_classes <= 32767) ? 1 : 0; }

int sc_attn_hard_splits(int64_t Nq, int64_t Nks, int sm_count) { return sc::attn_seg_splits(Nq, Nks, sm_count, 0, 0, 1); }
int sc_attn_hard_splits_for(int64_t Nq, int64_t Nks, int64_t D_pad, int op_dtype, int64_t n_classes, int n_betas,
                            int sm_count) {
  return sc::attn_seg_splits(Nq, Nks, sm_count, D_pad * (op_dtype == SC_E4M3 ? 1 : 2), n_classes, n_betas);
}

int sc_attn_fwd_hard_multi(const void* Qn, const void* Ks, const int16_t* group_class, const uint32_t* key_bits,
                           int op_dtype, int64_t Nq, int64_t Nks, int64_t D_pad, int64_t n_classes,
                           const float* betas, int n_betas, int splits, float* O, int64_t ldo, void* stream) {
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
                               n_betas, splits, O, ldo, static_cast<cudaStream_t>(stream));
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

int sc_gemm_split_nt(const void* Ah, const void* Al, const void* Bh, const void* Bl, int64_t M, int64_t N,
                     int64_t D_pad, float scale, float* Z, int64_t ldz, void* stream) {
  SC_REQUIRE(Ah && Al && Bh && Bl && Z, SC_EINVAL, "sc_gemm_split_nt: null pointer");
  SC_REQUIRE(M > 0 && N > 0 && ldz >= N, SC_ESHAPE, "sc_gemm_split_nt: bad shape");
  SC_REQUIRE(D_pad > 0 && D_pad % 64 == 0, SC_ESHAPE, "sc_gemm_split_nt: D_pad=%lld must be a multiple of 64",
             (long long)D_pad);
  SC_REQUIRE((reinterpret_cast<uintptr_t>(Ah) | reinterpret_cast<uintptr_t>(Al) | reinterpret_cast<uintptr_t>(Bh) |
              reinterpret_cast<uintptr_t>(Bl)) % 16 == 0,
             SC_EALIGN, "sc_gemm_split_nt: operands must be 16-byte aligned");
  SC_REQUIRE(M < (1ll << 31) && N < (1ll << 31) - 512, SC_ESHAPE, "sc_gemm_split_nt: M/N exceed int32 coordinates");
  int rc = sc::gemm_split_launch(&make_tmap, Ah, Al, Bh, Bl, M, N, D_pad, scale, Z, ldz, static_cast<cudaStream_t>(stream));
  if (rc != SC_OK) return rc;
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

int sc_attn_fwd(const void* Qn, const void* Kn, const void* Vt, int op_dtype, int64_t Nq, int64_t Nk,
                int64_t D_pad, int64_t n_cols, int64_t C_pad, int64_t Nk_pad, float beta,
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
    splits = sc_attn_splits(Nq, Nk, C_pad, sms);
  }
  SC_REQUIRE(splits <= tiles_total && splits <= 65535, SC_ESHAPE,
             "sc_attn_fwd: splits=%d exceeds the %lld key tiles", splits, (long long)tiles_total);

  // Kernel choice: the CTA-pair (cta_group::2) kernel is the default; SC_ATTN_IMPL=cluster selects the
  // single-CTA-MMA cluster kernel below (A/B runs and a cross-check in the tests).
  const char* impl = std::getenv("SC_ATTN_IMPL");
  // default: the transposed pair kernel (sc_attn_t.cu) when the class slices fill whole clusters;
  // SC_ATTN_IMPL=pair | cluster select the other two kernels (A/B runs, cross-checks in the tests)
  if ((impl == nullptr || impl[0] == 't') && (n_slices == 2 || n_slices % 4 == 0)) {
    int rct = sc::attn_t_launch(&make_tmap, Qn, Kn, Vt, f16, Nq, Nk, D_pad, n_cols, C_pad, Nk_pad, slice, n_slices,
                                beta, splits, O, ldo, static_cast<cudaStream_t>(stream));
    if (rct != SC_OK) return rct;
    SC_CUDA(cudaGetLastError());
    return SC_OK;
  }
  if (impl == nullptr || impl[0] != 'c') {
    int rcp = sc::attn_pair_launch(&make_tmap, Qn, Kn, Vt, f16, Nq, Nk, D_pad, n_cols, C_pad, Nk_pad, slice,
                                   n_slices, beta, splits, O, ldo, static_cast<cudaStream_t>(stream));
    if (rcp != SC_OK) return rcp;
    SC_CUDA(cudaGetLastError());
    return SC_OK;
  }

  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_tmap(&tmQ, Qn, Nq, D_pad, D_pad, kBM, f16)) != SC_OK) return rc;
  if ((rc = make_tmap(&tmK, Kn, Nk, D_pad, D_pad, kBN, f16)) != SC_OK) return rc;
  if ((rc = make_tmap(&tmV, Vt, C_pad, Nk_pad, Nk_pad, slice, f16)) != SC_OK) return rc;

  AttnParams p;
  p.Nq = static_cast<int>(Nq);
  p.n_dchunks = static_cast<int>(D_pad / kBK);
  p.n_cols = static_cast<int>(n_cols);
  p.slice = slice;
  p.tiles_total = static_cast<int>(tiles_total);
  p.splits = splits;
  p.c1 = beta * 1.4426950408889634f;
  p.c0 = -p.c1 + (f16 ? kPShift : 0.0f);
  p.o_scale = f16 ? exp2f(-kPShift) : 1.0f;
  p.O = O;
  p.ldo = ldo;

  // cluster size: the slices of one query tile share their weight tiles (4, 2 or 1 CTAs)
  int G = (n_slices % 4 == 0) ? 4 : (n_slices % 2 == 0 ? 2 : 1);
  if (const char* env = std::getenv("SC_ATTN_CLUSTER")) {       // tuning / A-B knob: force a smaller cluster
    const int want = std::atoi(env);
    if ((want == 1 || want == 2 || want == 4) && want <= G) G = want;
  }

  p.ns = kUnits - G;
  if (const char* env = std::getenv("SC_ATTN_STAGES")) {        // tuning knob: shallower operand ring
    const int want = std::atoi(env);
    if (want >= 2 && want < p.ns) p.ns = want;
  }

  p.dbg_skip = 0;
#ifdef SC_ATTN_TIMING_EXPERIMENTS   // never in the shipped library: skipping loads gives wrong results
  if (const char* env = std::getenv("SC_ATTN_DEBUG_SKIP")) p.dbg_skip = std::atoi(env);
#endif
  p.pf_dist = 0;
  if (const char* env = std::getenv("SC_ATTN_PREFETCH")) {      // tuning knob: L2 prefetch distance (rounds)
    const int want = std::atoi(env);
    if (want >= 0 && want <= 16) p.pf_dist = want;
  }

  dim3 grid(static_cast<unsigned>(n_slices), static_cast<unsigned>(sc::ceil_div(Nq, kBM)),
            static_cast<unsigned>(splits));
  SC_REQUIRE(grid.y <= 65535, SC_ESHAPE, "sc_attn_fwd: more than 65535 query tiles; chunk the queries");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (f16) {
    if (G == 4) rc = launch<true, 4>(grid, st, tmQ, tmK, tmV, p);
    else if (G == 2) rc = launch<true, 2>(grid, st, tmQ, tmK, tmV, p);
    else rc = launch<true, 1>(grid, st, tmQ, tmK, tmV, p);
  } else {
    if (G == 4) rc = launch<false, 4>(grid, st, tmQ, tmK, tmV, p);
    else if (G == 2) rc = launch<false, 2>(grid, st, tmQ, tmK, tmV, p);
    else rc = launch<false, 1>(grid, st, tmQ, tmK, tmV, p);
  }
  if (rc != SC_OK) return rc;
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

}  // extern "C"
