// Segmented hard-label attention kernel ("SEG") on a LABEL-SORTED key bank.
//
// Reference lines: cache_weights_strategy.py:33-36 (A = Q^T K, W = exp(-beta (1 - A))) + image_attention.py:109
// (O = W @ V) for the one-hot cache values of HardCacheStrategy (cache_value_strategy.py:14-17), of
// cache.replace_outs_with_golds (image_attention.py:65-66) and of Tip-Adapter (tip_adapter/utils.py:62,114-116).
// With V = one_hot(label),  O[q, c] = sum_{k: label(k) = c} exp(beta (q.k - 1))  — a per-class SEGMENTED ROW SUM
// of the weight matrix.  The sum over keys is order-free, so the host sorts the bank by label once (class
// segments padded to whole 16-key groups, sc_attn_fwd_hard's contract) and the kernel becomes:
//   GEMM-1  S[256 q x 256 keys] = Qn_tile . Ks_tile^T   one CTA pair, tcgen05.mma cta_group::2, M = 256 (128
//           queries per CTA), N = 256 (128 keys per CTA: the pair shares B), K = 16 x (D / 16); operands
//           TMA-staged through a 7-stage x 32 KB ring (Q chunk 16 KB + K chunk 16 KB per CTA), accumulators in
//           TMEM: two [128 lanes x 256 columns] fp32 buffers = all 512 columns, double buffered against
//   exp+sum four warps: thread = query (TMEM lane); tcgen05.ld 32 columns at a time, P = exp2(c1 S + c0) in
//           fp32, running sum per class in a register; when the (warp-uniform) class of a 16-key group changes
//           the finished sum is stored to O[q, class].  Padding keys are masked by a per-key bit.
// Nothing else touches memory: the weights are never rounded to 16 bits, never written to shared memory, never
// exchanged between CTAs, and there is no O accumulator in TMEM — which is what frees TMEM for N = 256 tiles
// (64 B/cycle/SM of operand traffic instead of 96 for N = 128) and shared memory for a ring deep enough to
// cover the ~2500-cycle release -> TMA -> full turnaround measured on B200 (tools/ubench).  Clusters are CTA
// pairs, so all 148 SMs are used (4-CTA clusters strand 16).  The dense-values kernel (sc_attn_t.cu) remains the
// path for soft cache values.
#include "sc_common.cuh"
#include "sc_ptx.cuh"

#include <cuda.h>
#include <cstdlib>

namespace {

using namespace scptx;

constexpr int kBQ = 128;             // queries per CTA (its half of UMMA M = 256)
constexpr int kBKeys = 128;          // keys per CTA per step (its half of UMMA N = 256)
constexpr int kStepKeys = 256;       // keys per pair step
constexpr int kBK = 64;              // 16-bit elements per swizzled smem row
constexpr int kStage = 32768;        // Q chunk [128 x 64] 16 KB + K chunk [128 x 64] 16 KB
constexpr int kNS = 7;               // ring stages: 224 KB
constexpr int kThreads = 192;        // warps 0-3 exp+sum (TMEM lane quadrant = warp), 4 TMA producer, 5 MMA issuer
constexpr int kProducerWarp = 4;
constexpr int kMmaWarp = 5;
constexpr int kTmemCols = 512;       // S0 @0, S1 @256
constexpr int kSmemBytes = kNS * kStage + 1024 + 256;

struct SParams {
  int Nq;
  int n_dchunks;
  int steps_total;           // ceil(Nks / 256)
  int splits;
  int pf_dist;               // L2 prefetch distance in steps (0 = off)
  int dbg;                   // SC_ATTN_TIMING_EXPERIMENTS builds only (wrong results): bit0/2 skip Q/K loads, 3 exp, 4 MMAs
  float c1[4], c0[4];        // per beta: exponent scale beta * log2(e) and offset -c1 (kNB betas per launch)
  long long o_beta_stride;   // elements between the O slabs of consecutive betas (= splits * Nq * ldo)
  const int16_t* gcls;       // class of every 16-key group, [steps_total * 16]; -1 = no real key
  const uint32_t* kbits;     // validity bit per key, [steps_total * 8] words (bit j of word w = key 32 w + j)
  float* O;                  // [splits, Nq, ldo], zeroed by the launcher; only the classes met are written
  long long ldo;
  unsigned long long* clk;   // experiments builds: {sum of CTA cycles, sum of CTA ns, CTAs}
  // kGemm mode (sc_gemm_split_nt): Z[m, n] = scale * S[m, n], S accumulated over 3 operand passes
  float* Z;
  long long ldz;
  int n_cols;                // valid columns of Z (rows of B)
  float scale;
  // kRowConf: per-row confidence / label of scale * S (sc_rowconf_from_features)
  float* conf;
  int* label;
  int conf_prob;             // 0: conf = row max; 1: conf = max softmax(prob_scale * row) = 1 / sum exp2(c1[0] (v - max))
  const float* row_scale;    // kRowConf: optional per-row factor on top of `scale` (1 / norm of a raw feature row)
  int passes;                // kGemm: operand passes per S tile: 3 = (A hi, B hi), (A hi, B lo), (A lo, B hi);
                             // 2 = A is exact in fp16 (no lo part): (A, B hi), (A, B lo)
};

struct Bars {
  uint64_t full[8];          // leader: TMA bytes of BOTH CTAs landed
  uint64_t empty[8];         // both: pair MMAs reading the stage retired
  uint64_t s_full[2];        // both: S buffer complete
  uint64_t s_empty[2];       // leader: 8 warp arrivals (4 per CTA): S buffer drained
  uint32_t tmem_slot;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// atomic max of a float cell that starts at -inf (sign-magnitude order: non-negative floats compare as signed
// ints, negative floats in reverse as unsigned ints; -inf loses to everything under both)
__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// what the four consumer warps do with a finished S tile
enum : int {
  kAttn = 0,      // Tip-Adapter weights exp(beta (S - 1)), summed per class                       -> O[q, class]
  kGemmOut = 1,   // scale * S stored as is (split-fp16 GEMM, sc_gemm_split_nt)                     -> Z[q, n]
  kSoftmax = 2,   // temperature softmax over the keys with a RUNNING ROW MAXIMUM: per class the
                  // log2-sum-exp of tau * S over its keys (the (m, l) pair in one float)            -> LSE[q, class]
  kRowMax = 3,    // max_k S[q, k] over the valid keys (pre-pass of the dense-values softmax mode)   -> rowmax[q]
  kRowConf = 4,   // split-fp16 GEMM (3 operand passes, as kGemmOut) whose rows are reduced on the fly to
                  // (confidence, first argmax): pseudo-labels without writing the logits bank        -> conf[m], label[m]
};

// kGemm = false: the attention kernel described above.  kGemm = true: the same TMA ring / pair-UMMA / TMEM
// pipeline used as a plain "NT" GEMM with split-fp16 operands (sc_gemm_split_nt): three operand passes
// (Ah.Bh, Ah.Bl, Al.Bh) accumulate into one S tile, and the four "exp" warps store scale * S instead.
// kNB (attention mode): betas per launch.  S = Q.K^T does not depend on beta, so a beta sweep (8 betas per cache
// in image_attention.yaml, 200 in Tip-Adapter's search_hp) pays GEMM-1 once per kNB betas; the exp warps then do
// kNB exponentials per S element — 2048 MUFU cycles per beta and step against 8192 cycles of UMMAs, so kNB = 4
// balances the two pipes.
template <int kOp, int kKind, int kNB>
__global__ void __launch_bounds__(kThreads, 1)
sc_attn_seg_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmQ2, const __grid_constant__ CUtensorMap tmK2, const SParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t ring0 = (raw_addr + 1023u) & ~1023u;
  Bars* bars = reinterpret_cast<Bars*>(smem_raw + (ring0 - raw_addr) + kNS * kStage);

  // warp index through a shuffle: provably warp-uniform for the compiler (role loops live in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader (issues the pair UMMAs), 1 = peer
  const bool is_leader = (rank == 0);
  const uint16_t pair_mask = 3;

  const int q0 = (blockIdx.y * 2 + static_cast<int>(rank)) * kBQ;      // my 128 queries
  const int split = blockIdx.z;
  const int s0 = static_cast<int>((static_cast<long long>(p.steps_total) * split) / p.splits);
  const int s1 = static_cast<int>((static_cast<long long>(p.steps_total) * (split + 1)) / p.splits);
  const int nsteps = s1 - s0;
  const int nd = p.n_dchunks;
  constexpr bool kF8 = (kOp == SC_E4M3);
  constexpr int kChunkElems = kF8 ? 2 * kBK : kBK;      // elements in a 128-byte operand row
  constexpr bool kGemm = (kKind == kGemmOut || kKind == kRowConf);      // split-fp16 operands, 3 passes per S tile
  const int n_pass = kGemm ? p.passes : 1;    // operand passes accumulated into one S tile

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
    if (kGemm) {
      prefetch_tmap(&tmQ2);
      prefetch_tmap(&tmK2);
    }
    for (int s = 0; s < kNS; ++s) {
      mbar_init(smem_u32(&bars->full[s]), 1);
      mbar_init(smem_u32(&bars->empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bars->s_full[b]), 1);
      mbar_init(smem_u32(&bars->s_empty[b]), 8);
    }
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

  if (warp == kProducerWarp) {
    // ===================================================== TMA producer (both CTAs; warp-uniform, elected issue)
    int stage = 0;
    uint32_t phase = 0, ready = 0;
    const uint32_t full0 = smem_u32(&bars->full[0]);
    const uint32_t full0c = mapa(full0, 0);
    const uint32_t empty0 = smem_u32(&bars->empty[0]);
    const uint32_t kon = (p.dbg & 4) ? 0u : 1u, qon = (p.dbg & 1) ? 0u : 1u;
    const uint32_t tx = is_leader ? 2u * 16384u * (kon + qon) : 0u;       // both CTAs' bytes
    const uint32_t plain = (is_leader && tx == 0u) ? 1u : 0u;
#pragma unroll 1
    for (int st = 0; st < nsteps; ++st) {
      const int krow = (s0 + st) * kStepKeys + static_cast<int>(rank) * kBKeys;      // my 128 keys of the step
      // L2 prefetch of the key stream: co-resident pairs walk the same key steps at about the same time, so a
      // step's first touch pays the HBM latency for all of them; one pair in 32 (by query tile) pulls the
      // chunks of step st + pf_dist into L2 ahead of the pack
      if (kKind == kAttn && p.pf_dist > 0 && ((st + p.pf_dist) & 31) == static_cast<int>(blockIdx.y & 31u) && st + p.pf_dist < nsteps) {
        if (elect_one()) {
          for (int d = 0; d < nd; ++d) tma_prefetch_2d(&tmK, d * kChunkElems, krow + p.pf_dist * kStepKeys);
        }
        __syncwarp();
      }
#pragma unroll 1
      for (int ps = 0; ps < n_pass; ++ps) {
        // kGemm passes: (A hi, B hi), (A hi, B lo), (A lo, B hi)
        const CUtensorMap* mq = (kGemm && ps == 2) ? &tmQ2 : &tmQ;
        const CUtensorMap* mk = (kGemm && ps == 1) ? &tmK2 : &tmK;
#pragma unroll 1
        for (int d = 0; d < nd; ++d) {
          if (!ready) mbar_wait(empty0 + stage * 8, phase ^ 1u);
          const bool wrap = (stage + 1 == kNS);
          const uint32_t dst = ring0 + stage * kStage;
          ready = __all_sync(0xffffffffu,
                             tma2_cg2_probe(dst, mq, d * kChunkElems, q0, qon, dst + 16384, mk, d * kChunkElems, krow, kon,
                                            full0c + stage * 8, full0 + stage * 8, tx, plain,
                                            empty0 + (wrap ? 0 : stage + 1) * 8, (wrap ? phase ^ 1u : phase) ^ 1u));
          if (wrap) { stage = 0; phase ^= 1u; } else { ++stage; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    if (is_leader) {
      // ===================================================== MMA issuer for the pair
      int stage = 0;
      uint32_t phase = 0, ready = 0;
      const uint32_t idesc = umma_idesc_16b(256, 256, kOp != SC_BF16);
      const uint32_t full0 = smem_u32(&bars->full[0]);
      const uint32_t empty0 = smem_u32(&bars->empty[0]);
      const uint32_t en = (p.dbg & 16) ? 0u : 1u;
#pragma unroll 1
      for (int st = 0; st < nsteps; ++st) {
        const int sb = st & 1;
        mbar_wait(smem_u32(&bars->s_empty[sb]), ((st >> 1) & 1) ^ 1u);
        const uint32_t tmem_s = tmem_base + sb * 256;
        const uint32_t sfull = smem_u32(&bars->s_full[sb]);
#pragma unroll 1
        for (int dd = 0; dd < nd * n_pass; ++dd) {
          if (!ready) mbar_wait(full0 + stage * 8, phase);
          tc_fence_after();
          const bool wrap = (stage + 1 == kNS);
          const uint64_t a_desc = umma_desc_k128(ring0 + stage * kStage);             // Q chunk: my 128 queries
          const uint64_t b_desc = umma_desc_k128(ring0 + stage * kStage + 16384);     // K chunk: my 128 keys
          ready = __all_sync(0xffffffffu,
                             umma4_cg2_probe<kF8>(tmem_s, a_desc, a_desc + 2, a_desc + 4, a_desc + 6, b_desc, b_desc + 2,
                                             b_desc + 4, b_desc + 6, idesc, dd != 0 ? 1u : 0u, en, empty0 + stage * 8,
                                             pair_mask, sfull, pair_mask, dd == nd * n_pass - 1 ? 1u : 0u,
                                             full0 + (wrap ? 0 : stage + 1) * 8, wrap ? phase ^ 1u : phase));
          if (wrap) { stage = 0; phase ^= 1u; } else { ++stage; }
        }
      }
    }
  } else {
    // ===================================================== exp + segmented sum warps: thread = query
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    if constexpr (kKind == kRowConf) {
      // ---- logits never leave the SM: thread = image row; running first-argmax over the class columns (strictly
      // greater replaces, so equal values keep the smaller class like torch.max) and, for the softmax confidence,
      // an online sum of exp2(c1 (v - running max)) rescaled whenever the maximum grows
      const int qg = q0 + row;
      float best = -INFINITY, m_run = -1e30f, acc = 0.f;
      int bidx = 0;
      const float c1 = p.c1[0];
      const float rs = (p.row_scale != nullptr && qg < p.Nq) ? p.scale * p.row_scale[qg] : p.scale;
#pragma unroll 1
      for (int st = 0; st < nsteps; ++st) {
        const int b = st & 1;
        mbar_wait(smem_u32(&bars->s_full[b]), (st >> 1) & 1);
        tc_fence_after();
        const int n0 = (s0 + st) * kStepKeys;
#pragma unroll 1
        for (int cc = 0; cc < kStepKeys / 32; ++cc) {
          const int nb = n0 + cc * 32;
          if (nb >= p.n_cols) break;
          uint32_t rg[32];
          tmem_ld_32x32(tmem_base + lane_addr + b * 256 + cc * 32, rg);
          tmem_ld_wait();
          const int nvalid = min(32, p.n_cols - nb);
          float cm = -1e30f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float v = (j < nvalid) ? __uint_as_float(rg[j]) * rs : -INFINITY;
            if (v > best) { best = v; bidx = nb + j; }
            cm = fmaxf(cm, v);
          }
          if (p.conf_prob) {
            const float mn = fmaxf(m_run, cm);
            acc *= ex2_approx(c1 * (m_run - mn));
            m_run = mn;
            float sacc = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float e = ex2_approx(c1 * (__uint_as_float(rg[j]) * rs - mn));
              sacc += (j < nvalid) ? e : 0.f;
            }
            acc += sacc;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (is_leader) mbar_arrive(smem_u32(&bars->s_empty[b]));
          else mbar_arrive_cluster_relaxed(smem_u32(&bars->s_empty[b]), 0);
        }
      }
      if (qg < p.Nq) {
        p.conf[qg] = p.conf_prob ? 1.0f / acc : best;
        p.label[qg] = bidx;
      }
    } else if constexpr (kGemm) {
      // ---- GEMM mode: Z[q, n] = scale * S[q, n]; thread = output row, 32 consecutive columns per TMEM load
      const int qg = q0 + row;
      float* zrow = p.Z + static_cast<long long>(qg) * p.ldz;
      const float rs = (p.row_scale != nullptr && qg < p.Nq) ? p.scale * p.row_scale[qg] : p.scale;
      const bool vec4 = (p.ldz % 4 == 0) && (reinterpret_cast<uintptr_t>(p.Z) % 16 == 0);
#pragma unroll 1
      for (int st = 0; st < nsteps; ++st) {
        const int b = st & 1;
        mbar_wait(smem_u32(&bars->s_full[b]), (st >> 1) & 1);
        tc_fence_after();
        const int n0 = (s0 + st) * kStepKeys;
#pragma unroll 1
        for (int cc = 0; cc < kStepKeys / 32; ++cc) {
          uint32_t rg[32];
          tmem_ld_32x32(tmem_base + lane_addr + b * 256 + cc * 32, rg);
          tmem_ld_wait();
          const int nb = n0 + cc * 32;
          if (qg < p.Nq && nb < p.n_cols) {
            if (vec4 && nb + 32 <= p.n_cols) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(zrow + nb + j) =
                    make_float4(__uint_as_float(rg[j]) * rs, __uint_as_float(rg[j + 1]) * rs,
                                __uint_as_float(rg[j + 2]) * rs, __uint_as_float(rg[j + 3]) * rs);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nb + j < p.n_cols) zrow[nb + j] = __uint_as_float(rg[j]) * rs;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (is_leader) mbar_arrive(smem_u32(&bars->s_empty[b]));
          else mbar_arrive_cluster_relaxed(smem_u32(&bars->s_empty[b]), 0);
        }
      }
    } else if constexpr (kKind == kRowMax) {
      // ---- row maximum of S over the valid keys (key index < n_cols); thread = query
      const int qg = q0 + row;
      float m = -INFINITY;
#pragma unroll 1
      for (int st = 0; st < nsteps; ++st) {
        const int b = st & 1;
        mbar_wait(smem_u32(&bars->s_full[b]), (st >> 1) & 1);
        tc_fence_after();
        const int n0 = (s0 + st) * kStepKeys;
#pragma unroll 1
        for (int cc = 0; cc < kStepKeys / 32; ++cc) {
          uint32_t rg[32];
          tmem_ld_32x32(tmem_base + lane_addr + b * 256 + cc * 32, rg);
          tmem_ld_wait();
          const int nb = n0 + cc * 32;
          if (nb + 32 <= p.n_cols) {
#pragma unroll
            for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(rg[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb + j < p.n_cols) m = fmaxf(m, __uint_as_float(rg[j]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (is_leader) mbar_arrive(smem_u32(&bars->s_empty[b]));
          else mbar_arrive_cluster_relaxed(smem_u32(&bars->s_empty[b]), 0);
        }
      }
      if (qg < p.Nq && nsteps > 0) atomic_max_f32(p.Z + qg, m * p.scale);
    } else if constexpr (kKind == kSoftmax) {
      // ---- temperature softmax over the keys, online: m = running maximum of S over the valid keys seen so far,
      // acc = sum of exp2(c1 (S - m)) over the keys of the current class.  When m grows acc is rescaled (one ex2 per
      // 32-key chunk); a finished class leaves as its log2-sum-exp  c1 m + log2(acc), which does not depend on the m
      // it was accumulated against — so classes flushed under different maxima (and different key splits, and
      // different ranks) combine exactly: sc_softmax_partials.
      const float c1 = p.c1[0];
      const int q = q0 + row;
      float* orow = p.O + (static_cast<long long>(split) * p.Nq + q) * p.ldo;
      const bool q_ok = q < p.Nq;
      int cur = -1;
      float m = -1e30f, acc = 0.f;
      auto flush = [&]() {
        if (cur >= 0 && q_ok && acc > 0.f) orow[cur] = fmaf(c1, m, lg2_approx(acc));
      };
#pragma unroll 1
      for (int st = 0; st < nsteps; ++st) {
        const int b = st & 1;
        const uint4* gp = reinterpret_cast<const uint4*>(p.gcls + static_cast<long long>(s0 + st) * 16);
        const uint4 ga = __ldg(gp), gb = __ldg(gp + 1);
        const uint4* kp = reinterpret_cast<const uint4*>(p.kbits + static_cast<long long>(s0 + st) * 8);
        const uint4 ka = __ldg(kp), kb = __ldg(kp + 1);
        uint32_t myw = 0;                      // lane l < 8: group-class word l, lane 8 + l: validity word l
        {
          const uint32_t gw[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
          const uint32_t kw[8] = {ka.x, ka.y, ka.z, ka.w, kb.x, kb.y, kb.z, kb.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            myw = lane == i ? gw[i] : myw;
            myw = lane == 8 + i ? kw[i] : myw;
          }
        }
        mbar_wait(smem_u32(&bars->s_full[b]), (st >> 1) & 1);
        tc_fence_after();
        const uint32_t tcol = tmem_base + lane_addr + b * 256;
#pragma unroll 1
        for (int cc = 0; cc < kStepKeys / 32; ++cc) {
          uint32_t rg[32];
          tmem_ld_32x32(tcol + cc * 32, rg);
          const uint32_t gwc = __shfl_sync(0xffffffffu, myw, cc);
          const uint32_t bits = __shfl_sync(0xffffffffu, myw, 8 + cc);
          tmem_ld_wait();
          float cm = -1e30f;
#pragma unroll
          for (int j = 0; j < 32; ++j) cm = fmaxf(cm, ((bits >> j) & 1u) ? __uint_as_float(rg[j]) : -1e30f);
          const float mn = fmaxf(m, cm);
          acc *= ex2_approx(c1 * (m - mn));
          m = mn;
          const float off = -c1 * m;
          float sa = 0.f, sb = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float ea = ex2_approx(fmaf(__uint_as_float(rg[j]), c1, off));
            const float eb = ex2_approx(fmaf(__uint_as_float(rg[16 + j]), c1, off));
            sa += ((bits >> j) & 1u) ? ea : 0.f;
            sb += ((bits >> (16 + j)) & 1u) ? eb : 0.f;
          }
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int cls = static_cast<int>(static_cast<int16_t>((gwc >> (16 * hh)) & 0xffffu));
            if (cls != cur) {
              flush();
              cur = cls;
              acc = 0.f;
            }
            acc += hh == 0 ? sa : sb;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (is_leader) mbar_arrive(smem_u32(&bars->s_empty[b]));
          else mbar_arrive_cluster_relaxed(smem_u32(&bars->s_empty[b]), 0);
        }
      }
      flush();
    } else {
    float c1[kNB], cadd[kNB];
#pragma unroll
    for (int bi = 0; bi < kNB; ++bi) { c1[bi] = p.c1[bi]; cadd[bi] = p.c0[bi]; }
    const int q = q0 + row;
    float* orow = p.O + (static_cast<long long>(split) * p.Nq + q) * p.ldo;     // beta bi: + bi * o_beta_stride
    const bool q_ok = q < p.Nq;
    int cur = -1;          // class of the running sums (warp-uniform)
    float acc[kNB];
#pragma unroll
    for (int bi = 0; bi < kNB; ++bi) acc[bi] = 0.f;
    auto flush = [&]() {
      if (cur >= 0 && q_ok) {
#pragma unroll
        for (int bi = 0; bi < kNB; ++bi) orow[bi * p.o_beta_stride + cur] = acc[bi];
      }
    };
#pragma unroll 1
    for (int st = 0; st < nsteps; ++st) {
      const int b = st & 1;
      // classes of the 16 groups and validity bits of the 256 keys of this step (uniform loads, issued before the wait)
      const uint4* gp = reinterpret_cast<const uint4*>(p.gcls + static_cast<long long>(s0 + st) * 16);
      const uint4 ga = __ldg(gp), gb = __ldg(gp + 1);
      const uint4* kp = reinterpret_cast<const uint4*>(p.kbits + static_cast<long long>(s0 + st) * 8);
      const uint4 ka = __ldg(kp), kb = __ldg(kp + 1);
      const uint32_t gw[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};      // 2 classes per word
      const uint32_t kw[8] = {ka.x, ka.y, ka.z, ka.w, kb.x, kb.y, kb.z, kb.w};      // 32 keys per word
      mbar_wait(smem_u32(&bars->s_full[b]), (st >> 1) & 1);
      tc_fence_after();
      // 32 columns (two 16-key groups) per TMEM load, double buffered: the load of chunk cc+1 is in flight
      // while chunk cc is exponentiated and summed
      auto consume = [&](const uint32_t (&rg)[32], const uint32_t gwc, const uint32_t kwc) {
        // The exponentials of BOTH 16-key groups of the chunk first, as straight-line code (the two groups' sums
        // interleave and the MUFU pipe never drains: ~9 instead of ~15 cycles per element, ncu source page of the
        // per-group version), then the warp-uniform class bookkeeping.  Per group: s = e0 + e1 + ... + e15 in that
        // order, padding keys contribute an exact 0 — the results do not depend on where classes change.
        float sa[kNB], sb[kNB];
        if (kwc == 0xffffffffu) {                         // every key of the chunk is valid (the common case)
#pragma unroll
          for (int bi = 0; bi < kNB; ++bi) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              a += ex2_approx(fmaf(__uint_as_float(rg[j]), c1[bi], cadd[bi]));
              b += ex2_approx(fmaf(__uint_as_float(rg[16 + j]), c1[bi], cadd[bi]));
            }
            sa[bi] = a;
            sb[bi] = b;
          }
        } else {                                          // a class segment's last group / bank padding: masked
          const uint32_t bits = kwc;
#pragma unroll
          for (int bi = 0; bi < kNB; ++bi) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float ea = ex2_approx(fmaf(__uint_as_float(rg[j]), c1[bi], cadd[bi]));
              const float eb = ex2_approx(fmaf(__uint_as_float(rg[16 + j]), c1[bi], cadd[bi]));
              a += ((bits >> j) & 1u) ? ea : 0.f;
              b += ((bits >> (16 + j)) & 1u) ? eb : 0.f;
            }
            sa[bi] = a;
            sb[bi] = b;
          }
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int cls = static_cast<int>(static_cast<int16_t>((gwc >> (16 * hh)) & 0xffffu));
          if (cls != cur) {                              // warp-uniform: the finished class sums go out
            flush();
            cur = cls;
#pragma unroll
            for (int bi = 0; bi < kNB; ++bi) acc[bi] = 0.f;
          }
#pragma unroll
          for (int bi = 0; bi < kNB; ++bi) acc[bi] += hh == 0 ? sa[bi] : sb[bi];
        }
      };
      if (!(p.dbg & 8)) {
        const uint32_t tcol = tmem_base + lane_addr + b * 256;
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(tcol, ra);
        tmem_ld_wait();
        if constexpr (kNB == 1) {
#pragma unroll
          for (int cc = 0; cc < kStepKeys / 32; cc += 2) {
            tmem_ld_32x32(tcol + (cc + 1) * 32, rb);
            consume(ra, gw[cc], kw[cc]);
            tmem_ld_wait();
            if (cc + 2 < kStepKeys / 32) tmem_ld_32x32(tcol + (cc + 2) * 32, ra);
            consume(rb, gw[cc + 1], kw[cc + 1]);
            tmem_ld_wait();
          }
        } else {
          // several betas: the chunk body is kNB times longer, and eight unrolled copies of it no longer fit the
          // instruction cache (ncu: "no instruction" stalls as frequent as issues).  Roll the loop: lane l < 8 keeps
          // group-class word l, lane 8 + l validity word l, a shuffle hands the chunk's words to the warp.
          uint32_t myw = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            myw = lane == i ? gw[i] : myw;
            myw = lane == 8 + i ? kw[i] : myw;
          }
#pragma unroll 1
          for (int cc = 0; cc < kStepKeys / 32; cc += 2) {
            tmem_ld_32x32(tcol + (cc + 1) * 32, rb);
            consume(ra, __shfl_sync(0xffffffffu, myw, cc), __shfl_sync(0xffffffffu, myw, 8 + cc));
            tmem_ld_wait();
            if (cc + 2 < kStepKeys / 32) tmem_ld_32x32(tcol + (cc + 2) * 32, ra);
            consume(rb, __shfl_sync(0xffffffffu, myw, cc + 1), __shfl_sync(0xffffffffu, myw, 9 + cc));
            tmem_ld_wait();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (is_leader) mbar_arrive(smem_u32(&bars->s_empty[b]));
        else mbar_arrive_cluster_relaxed(smem_u32(&bars->s_empty[b]), 0);
      }
    }
    flush();
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

template <int kOp, int kKind, int kNB>
int launch_seg(dim3 grid, cudaStream_t st, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmQ2,
               const CUtensorMap& tmK2, const SParams& p) {
  auto kernel = sc_attn_seg_kernel<kOp, kKind, kNB>;
  SC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SC_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmQ, tmK, tmQ2, tmK2, p));
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

// key splits for the pair kernel: work items = query tiles (256 queries) x splits over sm_count / 2 pairs.
// Cost in key steps: waves x (steps per split + fixed prologue/epilogue) + what every extra split costs outside
// the tensor cores — its [Nq, n_classes] fp32 tile is zeroed before and read back after (memset + merge /
// epilogue, ~4 TB/s), per beta.  One 256-key step is 8192 UMMA cycles per 1024 bytes of row (~4.8 us at the
// ~1.7 GHz these kernels see).  n_classes = 0 ignores the tile traffic (the historical rule).
int attn_seg_splits(int64_t Nq, int64_t Nks, int sm_count, int64_t row_bytes, int64_t n_classes, int n_betas) {
  if (Nq <= 0 || Nks <= 0) return 1;
  if (sm_count <= 0) sm_count = 148;
  const int64_t pairs = sm_count / 2 > 0 ? sm_count / 2 : 1;
  const int64_t base = ceil_div(Nq, 2 * kBQ);
  const int64_t steps = ceil_div(Nks, kStepKeys);
  const double step_us = 4.8 * static_cast<double>(row_bytes > 0 ? row_bytes : 2048) / 2048.0;
  const double tile_us = static_cast<double>(Nq) * static_cast<double>(n_classes > 0 ? n_classes : 0) * 8.0 / 4.0e6;
  const double per_split = (n_betas > 0 ? n_betas : 1) * tile_us / step_us;     // in key steps
  int best = 1;
  double best_cost = 1e300;
  const int64_t smax = steps < 128 ? steps : 128;
  for (int64_t s = 1; s <= smax; ++s) {
    const double waves = static_cast<double>(ceil_div(base * s, pairs));
    const double cost = waves * (static_cast<double>(ceil_div(steps, s)) + 1.5) + per_split * static_cast<double>(s);
    if (cost < best_cost * 0.995) { best_cost = cost; best = static_cast<int>(s); }
  }
  return best;
}

// sc_softmax.cu
int fill_f32_async(float* dst, size_t n, float value, cudaStream_t st);

// Called by sc_attn_fwd_hard[_multi] / sc_attn_softmax_hard (sc_attn.cu) after argument validation: 1..4 betas per
// launch, O is [n_betas, splits, Nq, ldo] and is zeroed here.  softmax != 0 (one "beta" = the temperature): O
// receives the per-class log2-sum-exp tiles instead and is filled with -inf (= no key of that class) first.
int attn_seg_launch(int (*make_tmap)(CUtensorMap*, const void*, int64_t, int64_t, int64_t, int, bool),
                    int (*make_tmap_u8)(CUtensorMap*, const void*, int64_t, int64_t, int64_t, int),
                    const void* Qn, const void* Ks, const int16_t* gcls, const uint32_t* kbits, int op_dtype, int64_t Nq,
                    int64_t Nks, int64_t D_pad, const float* betas, int n_betas, int splits, float* O, int64_t ldo,
                    int softmax, cudaStream_t st) {
  CUtensorMap tmQ, tmK;
  int rc;
  const bool f8 = (op_dtype == SC_E4M3), f16 = (op_dtype == SC_F16);
  // The query operand is mapped over WHOLE 256-row tiles (the caller allocates sc_pad_queries(Nq) rows): a TMA box
  // that is partly out of bounds is filled row by row and a single-tile launch then streams the bank at 3.7 instead
  // of 5.5 TB/s (measured, tools/probes/probe_small_batch.py: 1 query 0.64 -> 0.49 ms, 8 queries 0.71 -> 0.47 ms)
  const int64_t nq_rows = round_up(Nq, 2 * kBQ);
  if (f8) {                                                                                 // 128 rows x 128 e4m3
    if ((rc = make_tmap_u8(&tmQ, Qn, nq_rows, D_pad, D_pad, kBQ)) != SC_OK) return rc;
    if ((rc = make_tmap_u8(&tmK, Ks, Nks, D_pad, D_pad, kBKeys)) != SC_OK) return rc;
  } else {
    if ((rc = make_tmap(&tmQ, Qn, nq_rows, D_pad, D_pad, kBQ, f16)) != SC_OK) return rc;   // 128 queries x 64 d
    if ((rc = make_tmap(&tmK, Ks, Nks, D_pad, D_pad, kBKeys, f16)) != SC_OK) return rc;    // 128 keys x 64 d
  }
  SParams p = {};
  p.Nq = static_cast<int>(Nq);
  p.n_dchunks = static_cast<int>(D_pad / (f8 ? 2 * kBK : kBK));
  // e4m3 operands are stored as SC_E4M3_SCALE * x: S comes back scaled by its square
  const float s_unscale = f8 ? 1.0f / (SC_E4M3_SCALE * SC_E4M3_SCALE) : 1.0f;
  p.steps_total = static_cast<int>(ceil_div(Nks, kStepKeys));
  p.splits = splits;
  for (int bi = 0; bi < 4; ++bi) {
    const float c = betas[bi < n_betas ? bi : n_betas - 1] * 1.4426950408889634f;
    p.c1[bi] = c * s_unscale;
    p.c0[bi] = -c;
  }
  p.o_beta_stride = static_cast<long long>(splits) * Nq * ldo;
  p.gcls = gcls;
  p.kbits = kbits;
  p.O = O;
  p.ldo = ldo;
  p.dbg = 0;
  p.clk = nullptr;
  // tuning knob, read once per process: L2 prefetch distance in key steps
  static const int pf_dist = [] {
    const char* env = std::getenv("SC_ATTN_PREFETCH");
    const int want = env ? std::atoi(env) : 2;
    return (want >= 0 && want <= 64) ? want : 2;
  }();
  p.pf_dist = pf_dist;
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
  if (softmax) {
    if ((rc = fill_f32_async(O, static_cast<size_t>(splits) * Nq * ldo, -INFINITY, st)) != SC_OK) return rc;
  } else {
    SC_CUDA(cudaMemsetAsync(O, 0, static_cast<size_t>(n_betas) * splits * Nq * ldo * sizeof(float), st));
  }
  dim3 grid(2u, static_cast<unsigned>(ceil_div(Nq, 2 * kBQ)), static_cast<unsigned>(splits));
  SC_REQUIRE(grid.y <= 65535, SC_ESHAPE, "sc_attn_fwd_hard: too many query tiles; chunk the queries");
  p.Z = nullptr;
  p.ldz = 0;
  p.n_cols = 0;
  p.scale = 1.0f;
#define SC_SEG_LAUNCH(NB)                                                          \
  (f8    ? launch_seg<SC_E4M3, kAttn, NB>(grid, st, tmQ, tmK, tmQ, tmK, p)         \
   : f16 ? launch_seg<SC_F16, kAttn, NB>(grid, st, tmQ, tmK, tmQ, tmK, p)          \
         : launch_seg<SC_BF16, kAttn, NB>(grid, st, tmQ, tmK, tmQ, tmK, p))
  if (softmax) {
    return f8    ? launch_seg<SC_E4M3, kSoftmax, 1>(grid, st, tmQ, tmK, tmQ, tmK, p)
           : f16 ? launch_seg<SC_F16, kSoftmax, 1>(grid, st, tmQ, tmK, tmQ, tmK, p)
                 : launch_seg<SC_BF16, kSoftmax, 1>(grid, st, tmQ, tmK, tmQ, tmK, p);
  }
  int rc2;
  switch (n_betas) {
    case 1: rc2 = SC_SEG_LAUNCH(1); break;
    case 2: rc2 = SC_SEG_LAUNCH(2); break;
    case 3: rc2 = SC_SEG_LAUNCH(3); break;
    default: rc2 = SC_SEG_LAUNCH(4); break;
  }
#undef SC_SEG_LAUNCH
  return rc2;
}

// (conf, label) per row of scale * A @ B^T for split-fp16 operands, the logits never written: the kGemmOut pipeline
// with ALL column steps of a row tile in one work item and the row reduction in the consumer warps.
int rowconf_fused_launch(int (*make_tmap)(CUtensorMap*, const void*, int64_t, int64_t, int64_t, int, bool),
                         const void* Ah, const void* Al, const void* Bh, const void* Bl, int64_t M, int64_t N,
                         int64_t D_pad, float scale, const float* row_scale, float prob_scale, int prob, float* conf, int* label,
                         cudaStream_t st) {
  CUtensorMap tmA, tmA2, tmB, tmB2;
  int rc;
  if ((rc = make_tmap(&tmA, Ah, M, D_pad, D_pad, kBQ, true)) != SC_OK) return rc;
  if ((rc = make_tmap(&tmA2, Al ? Al : Ah, M, D_pad, D_pad, kBQ, true)) != SC_OK) return rc;   // Al == null: 2 passes
  if ((rc = make_tmap(&tmB, Bh, N, D_pad, D_pad, kBKeys, true)) != SC_OK) return rc;
  if ((rc = make_tmap(&tmB2, Bl, N, D_pad, D_pad, kBKeys, true)) != SC_OK) return rc;
  SParams p = {};
  p.Nq = static_cast<int>(M);
  p.n_dchunks = static_cast<int>(D_pad / kBK);
  p.steps_total = static_cast<int>(ceil_div(N, kStepKeys));
  p.splits = 1;
  for (int bi = 0; bi < 4; ++bi) p.c1[bi] = p.c0[bi] = 0.f;
  p.c1[0] = prob_scale * 1.4426950408889634f;
  p.o_beta_stride = 0;
  p.gcls = nullptr;
  p.kbits = nullptr;
  p.O = nullptr;
  p.ldo = 0;
  p.dbg = 0;
  p.clk = nullptr;
  p.pf_dist = 0;
  p.Z = nullptr;
  p.ldz = 0;
  p.n_cols = static_cast<int>(N);
  p.scale = scale;
  p.conf = conf;
  p.label = label;
  p.conf_prob = prob;
  p.row_scale = row_scale;
  p.passes = Al ? 3 : 2;
  dim3 grid(2u, static_cast<unsigned>(ceil_div(M, 2 * kBQ)), 1u);
  SC_REQUIRE(grid.y <= 65535, SC_ESHAPE, "sc_rowconf_from_split: too many row tiles; chunk the rows");
  return launch_seg<SC_F16, kRowConf, 1>(grid, st, tmA, tmB, tmA2, tmB2, p);
}

// rowmax[q] = max over the Nk keys of Qn[q].Kn[k] (any key order; rows past Nk are never counted): GEMM-1 of the
// attention kernel with a max instead of the exponential sum.  Pre-pass of the dense-values softmax mode, which
// needs the exact row maximum before it rounds weights to 16 bits (sc_attn_fwd_shifted).
int attn_rowmax_launch(int (*make_tmap)(CUtensorMap*, const void*, int64_t, int64_t, int64_t, int, bool),
                       int (*make_tmap_u8)(CUtensorMap*, const void*, int64_t, int64_t, int64_t, int),
                       const void* Qn, const void* Kn, int op_dtype, int64_t Nq, int64_t Nk, int64_t D_pad, int splits,
                       float* rowmax, cudaStream_t st) {
  CUtensorMap tmQ, tmK;
  int rc;
  const bool f8 = (op_dtype == SC_E4M3), f16 = (op_dtype == SC_F16);
  const int64_t nq_rows = round_up(Nq, 2 * kBQ);          // whole query tiles, as in attn_seg_launch
  if (f8) {
    if ((rc = make_tmap_u8(&tmQ, Qn, nq_rows, D_pad, D_pad, kBQ)) != SC_OK) return rc;
    if ((rc = make_tmap_u8(&tmK, Kn, Nk, D_pad, D_pad, kBKeys)) != SC_OK) return rc;
  } else {
    if ((rc = make_tmap(&tmQ, Qn, nq_rows, D_pad, D_pad, kBQ, f16)) != SC_OK) return rc;
    if ((rc = make_tmap(&tmK, Kn, Nk, D_pad, D_pad, kBKeys, f16)) != SC_OK) return rc;
  }
  SParams p = {};
  p.Nq = static_cast<int>(Nq);
  p.n_dchunks = static_cast<int>(D_pad / (f8 ? 2 * kBK : kBK));
  p.steps_total = static_cast<int>(ceil_div(Nk, kStepKeys));
  p.splits = splits;
  for (int bi = 0; bi < 4; ++bi) p.c1[bi] = p.c0[bi] = 0.f;
  p.o_beta_stride = 0;
  p.gcls = nullptr;
  p.kbits = nullptr;
  p.O = nullptr;
  p.ldo = 0;
  p.dbg = 0;
  p.clk = nullptr;
  p.pf_dist = 0;
  p.Z = rowmax;
  p.ldz = 0;
  p.n_cols = static_cast<int>(Nk);
  p.scale = f8 ? 1.0f / (SC_E4M3_SCALE * SC_E4M3_SCALE) : 1.0f;
  if ((rc = fill_f32_async(rowmax, static_cast<size_t>(Nq), -INFINITY, st)) != SC_OK) return rc;
  dim3 grid(2u, static_cast<unsigned>(ceil_div(Nq, 2 * kBQ)), static_cast<unsigned>(splits));
  SC_REQUIRE(grid.y <= 65535, SC_ESHAPE, "sc_attn_rowmax: too many query tiles; chunk the queries");
  return f8    ? launch_seg<SC_E4M3, kRowMax, 1>(grid, st, tmQ, tmK, tmQ, tmK, p)
         : f16 ? launch_seg<SC_F16, kRowMax, 1>(grid, st, tmQ, tmK, tmQ, tmK, p)
               : launch_seg<SC_BF16, kRowMax, 1>(grid, st, tmQ, tmK, tmQ, tmK, p);
}

// Z[m, n] = scale * sum_d (Ah[m,d] Bh[n,d] + Ah[m,d] Bl[n,d] + Al[m,d] Bh[n,d]): fp32-accurate "NT" GEMM of
// operands given as fp16 (hi, lo) pairs, on the attention kernel's pipeline (one 256-row step of B per work item).
int gemm_split_launch(int (*make_tmap)(CUtensorMap*, const void*, int64_t, int64_t, int64_t, int, bool),
                      const void* Ah, const void* Al, const void* Bh, const void* Bl, int64_t M, int64_t N,
                      int64_t D_pad, float scale, const float* row_scale, float* Z, int64_t ldz, cudaStream_t st) {
  CUtensorMap tmA, tmA2, tmB, tmB2;
  int rc;
  if ((rc = make_tmap(&tmA, Ah, M, D_pad, D_pad, kBQ, true)) != SC_OK) return rc;
  if ((rc = make_tmap(&tmA2, Al ? Al : Ah, M, D_pad, D_pad, kBQ, true)) != SC_OK) return rc;   // Al == null: 2 passes
  if ((rc = make_tmap(&tmB, Bh, N, D_pad, D_pad, kBKeys, true)) != SC_OK) return rc;
  if ((rc = make_tmap(&tmB2, Bl, N, D_pad, D_pad, kBKeys, true)) != SC_OK) return rc;
  SParams p = {};
  p.Nq = static_cast<int>(M);
  p.n_dchunks = static_cast<int>(D_pad / kBK);
  p.steps_total = static_cast<int>(ceil_div(N, kStepKeys));
  p.splits = p.steps_total;                 // one step of 256 B-rows per work item
  for (int bi = 0; bi < 4; ++bi) p.c1[bi] = p.c0[bi] = 0.f;
  p.o_beta_stride = 0;
  p.gcls = nullptr;
  p.kbits = nullptr;
  p.O = nullptr;
  p.ldo = 0;
  p.dbg = 0;
  p.clk = nullptr;
  p.pf_dist = 0;
  p.Z = Z;
  p.ldz = ldz;
  p.n_cols = static_cast<int>(N);
  p.scale = scale;
  p.row_scale = row_scale;
  p.passes = Al ? 3 : 2;
  dim3 grid(2u, static_cast<unsigned>(ceil_div(M, 2 * kBQ)), static_cast<unsigned>(p.splits));
  SC_REQUIRE(grid.y <= 65535 && grid.z <= 65535, SC_ESHAPE, "sc_gemm_split_nt: too many tiles; chunk the rows");
  return launch_seg<SC_F16, kGemmOut, 1>(grid, st, tmA, tmB, tmA2, tmB2, p);
}

}  // namespace sc
