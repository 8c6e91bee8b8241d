// Hard-label ("one-hot values") variant of the transposed CTA-pair attention kernel, on a LABEL-SORTED bank.
//
// Same maths, reference lines and operand roles as sc_attn_t.cu; the cache values are one_hot(label[k])
// (HardCacheStrategy cache_value_strategy.py:14-17, Tip-Adapter cache_values tip_adapter/utils.py:62, gold-label
// caches image_attention.py:65-66), so  O[q, c] = sum_{k: label(k) = c} exp(beta (q.k - 1)).  The sum over keys
// is order-free, so the host permutes the key bank ONCE so that equal labels are adjacent and pads every class
// segment to whole 16-key groups (sc_attn_fwd_hard's contract).  Then every K=16 step of GEMM-2 sees a Vt tile
// with a single non-zero row (the group's class):
//   * that tile is a window into a static 48 KB shared-memory zone holding one 128-byte row of ones (a uniform
//     row is invariant under the 128-byte swizzle, so the window base may sit at any 128-byte offset: base =
//     &ones_row - 128 * (class row)); nothing is streamed or synthesised for GEMM-2 — the dense kernel pulled
//     1 MB of Vt per 128-query x 512-key round, 40 % of its L2 -> SM bytes, and was bound by exactly that
//     traffic (operand ring starved; board at the 1 kW cap with the tensor pipe half idle);
//   * the pair UMMA (M = 256: 128 class rows of each CTA) is issued only by the pair that owns the group's class
//     — all-zero Vt tiles are skipped, which is exact — so GEMM-2 executes 1/4 of the dense tile count;
//   * the TMA ring serves GEMM-1 alone and keeps streaming the next round's K/Q chunks during GEMM-2.
// Padding keys (inside a class segment's last group, and the ragged tail) get weight 0 in the exp warps.
#include "sc_common.cuh"
#include "sc_ptx.cuh"

#include <cuda.h>
#include <cstdlib>

namespace {

using namespace scptx;

constexpr int kBQ = 128;             // queries per cluster tile (UMMA N)
constexpr int kBN = 128;             // keys per tile = TMEM lanes of S^T
constexpr int kBK = 64;              // 16-bit elements per swizzled smem row
constexpr int kSub = 24576;          // GEMM-1 ring stage: Kn chunk 16 KB + Qn half 8 KB
constexpr int kHalf = 16384;         // half of a weight tile: [128 keys x 64 queries] MN-major SW128
// Static one-hot zone: 48 KB of zeros with ONE 128-byte row of ones per CTA — at kOnesEven in the even CTA of
// a pair, at kOnesOdd in the odd CTA.  A [128 classes x 64 keys] K-major SW128 Vt tile whose row r is ones (in
// the even / odd CTA only) is the 16 KB window starting at zone + kOnesEven/Odd - 128 r; the two offsets are
// 16 KB apart so that no window of one parity covers the other parity's row.
constexpr int kZone = 49152;
constexpr int kOnesEven = 16384 - 128;
constexpr int kOnesOdd = kOnesEven + 16384;
constexpr int kThreads = 192;
constexpr int kExpThreads = 128;
// Warp roles: 0-3 exp/epilogue (TMEM lane quadrant = warp), then the two single-warp issue loops.  The SM
// sub-partition arbiter favours higher warp ids, so the issue-bound loops come last.
constexpr int kProducerWarp = 4;
constexpr int kMmaWarp = 5;
constexpr int kTmemCols = 512;       // S^T0 @0, S^T1 @128, O^T blocks @256 (+128)
constexpr int kColO = 256;
constexpr int kMaxStages = 8;
constexpr int kMaxCluster = 4;
constexpr int kSmemPayload = 7 * 32768;
constexpr int kSmemBytes = kSmemPayload + 1024 + 2048;
constexpr float kPShift = 8.0f;      // see sc_attn.cu

struct TParams {
  int Nq;
  int n_dchunks;
  int n_cols;
  int slice;        // classes per CTA
  int n_mb;         // 128-class blocks per CTA = ceil(slice / 128)
  int tiles_total;
  int splits;
  int pf_dist;      // L2 prefetch distance in rounds (0 = off)
  int dbg;          // SC_ATTN_TIMING_EXPERIMENTS builds only (wrong results): bit0/1/2 skip Q/V/K loads,
                    // 3 exp math, 4/5 GEMM-1/2 MMAs, 6 shrink the exchange to 1 KB
  float c1, c0, o_scale;
  const int16_t* gcls;       // class of every 16-key group of the sorted bank, [tiles_total * 8]; -1 = skip
  const uint8_t* kvalid;     // 1 = real key, 0 = padding key (weight forced to 0), [tiles_total * 128]
  float* O;
  long long ldo;
  unsigned long long* clk;   // experiments builds: {sum of CTA cycles, sum of CTA ns, CTAs} -> effective SM clock
};

struct Bars {
  uint64_t full[kMaxStages];       // leader: TMA bytes of BOTH CTAs landed
  uint64_t empty[kMaxStages];      // both: pair MMAs reading the stage retired
  uint64_t s_full[2];              // both
  uint64_t s_empty[2];             // leader: 8 warp-elected arrivals (4 per CTA)
  uint64_t p_full[kMaxCluster];    // both: my half-slot for source CTA `src` is full
  uint64_t p_peer[kMaxCluster];    // leader: the odd CTA's half-slot is full (relay)
  uint64_t p_empty;                // both: every consumer pair retired GEMM-2 on MY last tile
  uint64_t o_full;                 // both
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
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool kF16, int NPAIR>
__global__ void __launch_bounds__(kThreads, 1)
sc_attn_ts_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const TParams p) {
  constexpr int CS = 2 * NPAIR;                                            // cluster size
  constexpr int kStage = kSub;
  constexpr int NS = (kSmemPayload - (CS + 1) * kHalf - kZone) / kStage;   // 4 (CS=4) / 5 (CS=2) stages
  static_assert(NS <= kMaxStages && NS >= 2, "ring depth");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t ring0 = (raw_addr + 1023u) & ~1023u;
  const uint32_t slot0 = ring0 + NS * kStage;                              // CS half-tile slots (by source)
  const uint32_t stag0 = slot0 + CS * kHalf;                               // the other half of MY tile
  const uint32_t zone0 = stag0 + kHalf;                                    // static one-hot zone
  Bars* bars = reinterpret_cast<Bars*>(smem_raw + (zone0 - raw_addr) + kZone);

  // warp index through a shuffle: the compiler then KNOWS it is warp-uniform and keeps the role loops' state
  // (ring stage, phases, descriptors) in uniform registers instead of converting it per UMMA / TMA issue
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();        // == blockIdx.x % CS
  const int h = static_cast<int>(rank & 1);       // which 64-query half of P^T this CTA feeds to its pair
  const uint32_t leader = rank & ~1u;
  const bool is_leader = (h == 0);
  const uint16_t pair_mask = static_cast<uint16_t>(3u << leader);

  const int c0 = blockIdx.x * p.slice;            // first class of this CTA
  const int q0 = blockIdx.y * kBQ;
  const int split = blockIdx.z;
  const int t0 = static_cast<int>((static_cast<long long>(p.tiles_total) * split) / p.splits);
  const int t1 = static_cast<int>((static_cast<long long>(p.tiles_total) * (split + 1)) / p.splits);
  const int T = t1 - t0;
  const int R = (T + CS - 1) / CS;
  const int nd = p.n_dchunks;
  const int n_mb = p.n_mb;
  const int pair_first = static_cast<int>(leader);   // round r is active for my pair iff r*CS + pair_first < T

#ifdef SC_ATTN_TIMING_EXPERIMENTS
  long long clk_c0 = 0;
  unsigned long long clk_t0 = 0;
  if (p.clk != nullptr && threadIdx.x == 0) {
    clk_c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(clk_t0));
  }
#endif
  if (warp == kProducerWarp && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    for (int s = 0; s < NS; ++s) {
      mbar_init(smem_u32(&bars->full[s]), 1);
      mbar_init(smem_u32(&bars->empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bars->s_full[b]), 1);
      mbar_init(smem_u32(&bars->s_empty[b]), 8);
    }
    for (int s = 0; s < CS; ++s) {
      mbar_init(smem_u32(&bars->p_full[s]), 1);
      mbar_init(smem_u32(&bars->p_peer[s]), 1);
    }
    mbar_init(smem_u32(&bars->p_empty), NPAIR);
    mbar_init(smem_u32(&bars->o_full), 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc2(smem_u32(&bars->tmem_slot), kTmemCols);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;
  const uint32_t xbytes = (p.dbg & 64) ? 1024u : static_cast<uint32_t>(kHalf);   // exchange unit

  // one-time: zero the O^T accumulators (GEMM-2 tiles are issued only for classes that occur, so every UMMA
  // accumulates) and build the static one-hot zone; then a second cluster-wide sync
  if (warp < 4) {
    uint32_t z[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) z[j] = 0u;
    const uint32_t lane_addr0 = static_cast<uint32_t>(warp * 32) << 16;
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) tmem_st_32x32(tmem_base + lane_addr0 + kColO + cc * 32, z);
    tmem_st_wait();
    for (uint32_t i = threadIdx.x; i < kZone / 16; i += kExpThreads) {
      const uint32_t off = i * 16;
      const bool ones = (off >= (h ? kOnesOdd : kOnesEven)) && (off < (h ? kOnesOdd : kOnesEven) + 128u);
      const uint32_t v = ones ? (kF16 ? 0x3C003C00u : 0x3F803F80u) : 0u;
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(zone0 + off), "r"(v) : "memory");
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();

  if (warp == kProducerWarp) {
    // ===================================================== TMA producer (both CTAs; warp-uniform, elected issue)
    // Per ring stage: (blocking wait only if the previous probe failed) -> fused [probe next slot's empty
    // barrier | expect_tx | TMA loads].
    int stage = 0;
    uint32_t phase = 0, ready = 0;
    const uint32_t full0 = smem_u32(&bars->full[0]);
    const uint32_t full0c = mapa(full0, leader);
    const uint32_t empty0 = smem_u32(&bars->empty[0]);
    const uint32_t kon = (p.dbg & 4) ? 0u : 1u, qon = (p.dbg & 1) ? 0u : 1u;
    const uint32_t tx1 = is_leader ? 2u * (kon * 16384u + qon * 8192u) : 0u;      // both CTAs' bytes
    const uint32_t plain1 = (is_leader && tx1 == 0u) ? 1u : 0u;
#pragma unroll 1
    for (int r = 0; r < R; ++r) {
      if (r * CS + pair_first >= T) break;
      const int krow = (t0 + r * CS + static_cast<int>(rank)) * kBN;   // my key tile (may be past the split: unused)
      // L2 prefetch of the key stream.  Every co-resident cluster walks the SAME key tiles at about the same
      // time, so a tile's first touch pays the HBM latency (~2000 cycles against ~800 for an L2 hit) for all
      // of them, and the 4-stage ring covers only the latter.  One cluster in 32 (by query tile) pulls the
      // tiles of round r + pf_dist into L2 ahead of the pack.
      if (p.pf_dist > 0 && ((r + p.pf_dist) & 31) == static_cast<int>(blockIdx.y & 31u)) {
        const int tile_pf = (r + p.pf_dist) * CS + static_cast<int>(rank);
        if (tile_pf < T && elect_one()) {
          for (int d = 0; d < nd; ++d) tma_prefetch_2d(&tmK, d * kBK, (t0 + tile_pf) * kBN);
        }
        __syncwarp();
      }
#pragma unroll 1
      for (int d = 0; d < nd; ++d) {            // 128 keys x 64 d  +  64 queries x 64 d
        if (!ready) mbar_wait(empty0 + stage * 8, phase ^ 1u);
        const bool wrap = (stage + 1 == NS);
        const uint32_t dst = ring0 + stage * kStage;
        ready = __all_sync(0xffffffffu, tma2_cg2_probe(dst, &tmK, d * kBK, krow, kon, dst + 16384, &tmQ, d * kBK,
                                                       q0 + h * 64, qon, full0c + stage * 8, full0 + stage * 8, tx1,
                                                       plain1, empty0 + (wrap ? 0 : stage + 1) * 8,
                                                       (wrap ? phase ^ 1u : phase) ^ 1u));
        if (wrap) { stage = 0; phase ^= 1u; } else { ++stage; }
      }
    }
  } else if (warp == kMmaWarp) {
    const uint32_t pfull0 = smem_u32(&bars->p_full[0]);
    const uint32_t ppeer0 = smem_u32(&bars->p_peer[0]);
    if (elect_one()) {                      // arm the half-slots fed by the other CTAs for round 0
      for (int src = 0; src < CS; ++src)
        if (src != static_cast<int>(rank) && src < T) mbar_arrive_expect_tx(pfull0 + src * 8, xbytes);
    }
    __syncwarp();
    if (is_leader) {
      // ===================================================== MMA issuer for the pair
      // Per ring stage: (blocking wait only if the previous probe failed) -> fused [probe next slot's full
      // barrier | 4 pair UMMAs | commit(s)].
      int stage = 0;
      uint32_t phase = 0, ready = 0;
      const uint32_t idesc1 = umma_idesc_16b(256, kBQ, kF16);                 // A, B K-major
      const uint32_t idesc2 = umma_idesc_16b(256, kBQ, kF16) | (1u << 16);    // B (P^T) MN-major
      const uint32_t tmem_o = tmem_base + kColO;
      const uint32_t full0 = smem_u32(&bars->full[0]);
      const uint32_t empty0 = smem_u32(&bars->empty[0]);
      const uint32_t pempty = smem_u32(&bars->p_empty);
      const uint32_t en1 = (p.dbg & 16) ? 0u : 1u, en2 = (p.dbg & 32) ? 0u : 1u;
      // GEMM-1 stage: operands from the TMA ring (a_step / b_step = descriptor increment per 16-element K step)
      auto issue = [&](uint32_t d_tmem, uint64_t a_desc, uint32_t a_step, uint64_t b_desc, uint32_t b_step,
                       uint32_t idesc, uint32_t acc0, uint32_t enable, uint32_t bar2, uint16_t mask2, uint32_t flag2) {
        if (!ready) mbar_wait(full0 + stage * 8, phase);
        tc_fence_after();
        const bool wrap = (stage + 1 == NS);
        ready = __all_sync(0xffffffffu,
                           umma4_cg2_probe(d_tmem, a_desc, a_desc + a_step, a_desc + 2 * a_step, a_desc + 3 * a_step,
                                           b_desc, b_desc + b_step, b_desc + 2 * b_step, b_desc + 3 * b_step, idesc,
                                           acc0, enable, empty0 + stage * 8, pair_mask, bar2, mask2, flag2,
                                           full0 + (wrap ? 0 : stage + 1) * 8, wrap ? phase ^ 1u : phase));
        if (wrap) { stage = 0; phase ^= 1u; } else { ++stage; }
      };
      // GEMM-2 of round rr: one pair UMMA (K = 16 keys) per 16-key group whose class this pair owns.  Lane l of
      // the warp holds the class of group l of the round (8 groups per source tile); the ballot of "mine" is
      // the uniform work list.
      const int cpair0 = c0;                                // leader: first class of the pair
      const int cpair1 = c0 + 2 * p.slice;
      const int16_t* gc = p.gcls + static_cast<long long>(t0) * 8;
      auto load_gcls = [&](int rr) -> int {
        const int i = rr * CS + (lane >> 3);
        return (rr < R && (lane >> 3) < CS && i < T) ? static_cast<int>(__ldg(gc + static_cast<long long>(i) * 8 + (lane & 7)))
                                                    : -1;
      };
      int g_cur = load_gcls(0), g_nxt = -1;
      auto gemm2_round = [&](int rr) {
        g_nxt = load_gcls(rr + 1);                          // prefetch: consumed a whole round later
        const int w_all = g_cur - cpair0;                   // class relative to the pair
        const uint32_t work = __ballot_sync(0xffffffffu, g_cur >= cpair0 && g_cur < cpair1);
#pragma unroll 1
        for (int src = 0; src < CS; ++src) {
          const int i = rr * CS + src;
          if (i >= T) break;
          mbar_wait(pfull0 + src * 8, rr & 1);
          mbar_wait(ppeer0 + src * 8, rr & 1);
          if (src != static_cast<int>(rank) && i + CS < T) {
            if (elect_one()) mbar_arrive_expect_tx(pfull0 + src * 8, xbytes);
            __syncwarp();
          }
          tc_fence_after();
          uint32_t m = (work >> (src * 8)) & 0xffu;
#pragma unroll 1
          while (m) {
            const int g = __ffs(m) - 1;
            m &= m - 1;
            const int w = __shfl_sync(0xffffffffu, w_all, src * 8 + g);
            const int owner = (w >= p.slice) ? 1 : 0;
            const int wi = w - owner * p.slice;             // class within the owner CTA's slice
            // A: window of the one-hot zone whose row (wi & 127) is ones in the owner CTA only
            const uint64_t a_desc = umma_desc_k128(zone0 + (owner ? kOnesOdd : kOnesEven) - 128u * static_cast<uint32_t>(wi & 127));
            // B: P^T rows 16g .. 16g+15 of source src (MN-major, 64 queries per CTA)
            const uint64_t b_desc = umma_desc_k128(slot0 + src * kHalf + (g >> 2) * 8192) + 128u * static_cast<uint32_t>(g & 3);
            if (en2 && elect_one()) umma_ss2(tmem_o + (wi >> 7) * 128, a_desc, b_desc, idesc2, 1u);
            __syncwarp();
          }
          if (elect_one()) umma_commit2_mcast(pempty, static_cast<uint16_t>(1u << src));   // release the P^T slot
          __syncwarp();
        }
        g_cur = g_nxt;
      };
      int own = 0;
#pragma unroll 1
      for (int r = 0; r < R; ++r) {
        if (r * CS + pair_first < T) {
          const int sb = own & 1;
          mbar_wait(smem_u32(&bars->s_empty[sb]), ((own >> 1) & 1) ^ 1u);
          const uint32_t tmem_s = tmem_base + sb * 128;
          const uint32_t sfull = smem_u32(&bars->s_full[sb]);
#pragma unroll 1
          for (int d = 0; d < nd; ++d) {
            const uint32_t a_addr = ring0 + stage * kStage;
            issue(tmem_s, umma_desc_k128(a_addr), 2u, umma_desc_k128(a_addr + 16384), 2u, idesc1, d != 0 ? 1u : 0u, en1,
                  sfull, pair_mask, d == nd - 1 ? 1u : 0u);
          }
          ++own;
        }
        if (r > 0) gemm2_round(r - 1);
      }
      if (R > 0) gemm2_round(R - 1);
      if (elect_one()) umma_commit2_mcast(smem_u32(&bars->o_full), pair_mask);
      __syncwarp();
    } else {
      // ===================================================== relay (odd CTA): half-slot full -> tell the leader
#pragma unroll 1
      for (int rr = 0; rr < R; ++rr) {
#pragma unroll 1
        for (int src = 0; src < CS; ++src) {
          const int i = rr * CS + src;
          if (i >= T) break;
          mbar_wait(pfull0 + src * 8, rr & 1);
          if (elect_one()) {
            if (src != static_cast<int>(rank) && i + CS < T) mbar_arrive_expect_tx(pfull0 + src * 8, xbytes);
            mbar_arrive_cluster(ppeer0 + src * 8, leader);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================================================== exp warps (+ epilogue), both CTAs
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                 // TMEM lane = key within my tile (exp) / class (epilogue)
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const float c1 = p.c1;
    const float cadd = p.c0;
    const float o_scale = p.o_scale;
    const uint32_t row_off = static_cast<uint32_t>(row) * 128u;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    const uint32_t my_slot = slot0 + rank * kHalf;    // half h of my tile (for my own pair)
    int own = 0, sent = 0;
#pragma unroll 1
    for (int r = 0; r < R; ++r) {
      if (r * CS + pair_first >= T) break;
      const bool has_tile = (r * CS + static_cast<int>(rank) < T);
      const int b = own & 1;
      // padding keys of the sorted bank (and the ragged tail) get weight 0: load the flag before the wait
      const uint32_t live = has_tile ? __ldg(p.kvalid + (static_cast<long long>(t0 + r * CS + static_cast<int>(rank)) * kBN + row)) : 0u;
      mbar_wait(smem_u32(&bars->s_full[b]), (own >> 1) & 1);
      tc_fence_after();
      if (has_tile) {
        mbar_wait(smem_u32(&bars->p_empty), (sent & 1) ^ 1u);     // both pairs retired my previous tile
#pragma unroll
        for (int cc = 0; cc < ((p.dbg & 8) ? 0 : kBQ / 32); ++cc) {
          uint32_t rg[32];
          tmem_ld_32x32(tmem_base + lane_addr + b * 128 + cc * 32, rg);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float e0 = ex2_approx(fmaf(__uint_as_float(rg[2 * j]), c1, cadd));
            const float e1 = ex2_approx(fmaf(__uint_as_float(rg[2 * j + 1]), c1, cadd));
            pk[j] = live ? pack_16x2<kF16>(e0, e1) : 0u;
          }
          // queries cc*32 .. +31 of key `row`: half (cc>>1) of P^T, 16-byte chunks (cc&1)*4 .. +3 of its row
          const uint32_t base = (((cc >> 1) == h) ? my_slot : stag0) + row_off;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t chunk = static_cast<uint32_t>((cc & 1) * 4 + j);
            const uint32_t addr = base + ((chunk ^ sw) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * j]),
                         "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                         : "memory");
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (is_leader) mbar_arrive(smem_u32(&bars->s_empty[b]));
        else mbar_arrive_cluster(smem_u32(&bars->s_empty[b]), leader);
      }
      if (has_tile) {
        fence_proxy_async_smem();
        named_bar_sync(1, kExpThreads);
        if (threadIdx.x == 0) {
          const uint32_t pf = smem_u32(&bars->p_full[rank]);
          mbar_arrive(pf);                                                   // my own half-slot
          bulk_copy_to_peer(mapa(my_slot, rank ^ 1u), stag0, xbytes, mapa(pf, rank ^ 1u));          // partner: other half
          if (CS == 4) {
            bulk_copy_to_peer(mapa(my_slot, rank ^ 2u), my_slot, xbytes, mapa(pf, rank ^ 2u));      // same half, other pair
            bulk_copy_to_peer(mapa(my_slot, rank ^ 3u), stag0, xbytes, mapa(pf, rank ^ 3u));        // other half, other pair
          }
        }
        ++sent;
      }
      ++own;
    }
    // ---- epilogue: O^T blocks (lane = class, column = query) -> O[q, class]
    mbar_wait(smem_u32(&bars->o_full), 0);
    tc_fence_after();
    const int c_end = min(c0 + p.slice, p.n_cols);
#pragma unroll 1
    for (int mb = 0; mb < n_mb; ++mb) {
      const int cls = c0 + mb * 128 + row;
#pragma unroll 1
      for (int cc = 0; cc < kBQ / 16; ++cc) {
        uint32_t rg[16];
        tmem_ld_32x16(tmem_base + lane_addr + kColO + mb * 128 + cc * 16, rg);
        tmem_ld_wait();
        if (cls < c_end) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int q = q0 + cc * 16 + j;
            if (q < p.Nq)
              p.O[(static_cast<long long>(split) * p.Nq + q) * p.ldo + cls] =
                  (T > 0) ? __uint_as_float(rg[j]) * o_scale : 0.0f;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
#ifdef SC_ATTN_TIMING_EXPERIMENTS
  if (p.clk != nullptr && threadIdx.x == 0) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    atomicAdd(p.clk, static_cast<unsigned long long>(clock64() - clk_c0));
    atomicAdd(p.clk + 1, t1 - clk_t0);
    atomicAdd(p.clk + 2, 1ull);
  }
#endif
  cluster_sync_all();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, kTmemCols);
  }
}

template <bool kF16, int NPAIR>
int launch_ts(dim3 grid, cudaStream_t st, const CUtensorMap& tmQ, const CUtensorMap& tmK, const TParams& p) {
  auto kernel = sc_attn_ts_kernel<kF16, NPAIR>;
  SC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2 * NPAIR;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SC_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmQ, tmK, p));
  return SC_OK;
}

#ifdef SC_ATTN_TIMING_EXPERIMENTS
unsigned long long* g_clk = nullptr;
#endif

}  // namespace

#ifdef SC_ATTN_TIMING_EXPERIMENTS
// experiments builds only: effective SM clock seen by the hard-label attention CTAs since the last call
extern "C" double sc_debug_attn_hard_clock_mhz(void) {
  if (g_clk == nullptr) return 0.0;
  unsigned long long h[3] = {0, 0, 0};
  cudaDeviceSynchronize();
  cudaMemcpy(h, g_clk, sizeof(h), cudaMemcpyDeviceToHost);
  cudaMemset(g_clk, 0, sizeof(h));
  return h[1] ? 1e3 * static_cast<double>(h[0]) / static_cast<double>(h[1]) : 0.0;
}
#endif

namespace sc {

// Called by sc_attn_fwd_hard (sc_attn.cu) after argument validation; needs n_slices == 2 or a multiple of 4.
int attn_ts_launch(int (*make_tmap)(CUtensorMap*, const void*, int64_t, int64_t, int64_t, int, bool),
                   const void* Qn, const void* Kn, const int16_t* gcls, const uint8_t* kvalid, bool f16, int64_t Nq, int64_t Nk,
                   int64_t D_pad, int64_t n_cols, int slice, int64_t n_slices, float beta, int splits, float* O,
                   int64_t ldo, cudaStream_t st) {
  SC_REQUIRE(n_slices == 2 || n_slices % 4 == 0, SC_EUNSUPPORTED, "hard-label kernel needs 2 or 4k class slices");
  CUtensorMap tmQ, tmK;
  int rc;
  if ((rc = make_tmap(&tmQ, Qn, Nq, D_pad, D_pad, 64, f16)) != SC_OK) return rc;        // 64 queries x 64 d
  if ((rc = make_tmap(&tmK, Kn, Nk, D_pad, D_pad, 128, f16)) != SC_OK) return rc;       // 128 keys x 64 d
  TParams p;
  p.Nq = static_cast<int>(Nq);
  p.n_dchunks = static_cast<int>(D_pad / kBK);
  p.n_cols = static_cast<int>(n_cols);
  p.slice = slice;
  p.n_mb = (slice + 127) / 128;
  p.tiles_total = static_cast<int>(ceil_div(Nk, kBN));
  p.splits = splits;
  p.c1 = beta * 1.4426950408889634f;
  p.c0 = -p.c1 + (f16 ? kPShift : 0.0f);
  p.o_scale = f16 ? exp2f(-kPShift) : 1.0f;
  p.gcls = gcls;
  p.kvalid = kvalid;
  p.O = O;
  p.ldo = ldo;
  p.dbg = 0;
  p.clk = nullptr;
  p.pf_dist = 4;
  if (const char* env = std::getenv("SC_ATTN_PREFETCH")) {      // tuning knob: L2 prefetch distance (rounds)
    const int want = std::atoi(env);
    if (want >= 0 && want <= 64) p.pf_dist = want;
  }
#ifdef SC_ATTN_TIMING_EXPERIMENTS   // never in the shipped library: skipping work gives wrong results
  if (const char* env = std::getenv("SC_ATTN_DEBUG_SKIP")) p.dbg = std::atoi(env);
  if (std::getenv("SC_ATTN_CLKPROBE")) {
    if (g_clk == nullptr) {
      SC_CUDA(cudaMalloc(&g_clk, 3 * sizeof(unsigned long long)));
      SC_CUDA(cudaMemset(g_clk, 0, 3 * sizeof(unsigned long long)));
    }
    p.clk = g_clk;
  }
#endif
  dim3 grid(static_cast<unsigned>(n_slices), static_cast<unsigned>(ceil_div(Nq, kBQ)), static_cast<unsigned>(splits));
  SC_REQUIRE(grid.y <= 65535, SC_ESHAPE, "sc_attn_fwd_hard: too many query tiles; chunk the queries");
  if (n_slices == 2) {
    rc = f16 ? launch_ts<true, 1>(grid, st, tmQ, tmK, p) : launch_ts<false, 1>(grid, st, tmQ, tmK, p);
  } else {
    rc = f16 ? launch_ts<true, 2>(grid, st, tmQ, tmK, p) : launch_ts<false, 2>(grid, st, tmQ, tmK, p);
  }
  return rc;
}

}  // namespace sc
