// sm_100a PTX wrappers used by the CLIP-search attention kernel: mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / fences) and the
// UMMA shared-memory + instruction descriptors.  Hand-written; no CUTLASS.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace scptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  // generic-proxy smem writes -> visible to the async proxy (UMMA / TMA reads)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
// arrive on the same-offset barrier of another CTA in the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta_rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 remote;\n\t"
      "mapa.shared::cluster.u32 remote, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [remote];\n\t"
      "}\n" ::"r"(bar),
      "r"(cta_rank)
      : "memory");
}
// the same without the release fence (MEMBAR.ALL.GPU + ERRBAR: ~20 % of an exp warp's step when global stores are
// outstanding): for hand-offs that publish no memory — "my tcgen05.ld reads of this TMEM buffer are complete",
// ordered by tcgen05.wait::ld + tcgen05.fence::before_thread_sync, not by the memory model
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t bar, uint32_t cta_rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 remote;\n\t"
      "mapa.shared::cluster.u32 remote, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [remote];\n\t"
      "}\n" ::"r"(bar),
      "r"(cta_rank)
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

// Non-blocking probe (never suspends the thread): issued BEFORE a batch of TMA / UMMA instructions for the
// current stage so that its ~150-cycle latency hides under their issue cost; a blocking wait follows only
// when the probe failed.  (try_wait on an already-complete barrier costs ~117 cycles of a single-thread
// issue loop whose whole budget is the 256 cycles four N=128 UMMAs execute in.)
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}

// Bounded wait: a pipeline bug must trap, never hang the GPU.
#ifndef SC_MBAR_TIMEOUT_CYCLES
#define SC_MBAR_TIMEOUT_CYCLES (4000000000ll)  // ~2 s at 2 GHz
#endif
template <bool kCluster = false>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (kCluster ? mbar_try_wait_cluster(bar, parity) : mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t polls = 0;
  while (!(kCluster ? mbar_try_wait_cluster(bar, parity) : mbar_try_wait(bar, parity))) {
    if ((++polls & 0x3ff) == 0 && clock64() - t0 > SC_MBAR_TIMEOUT_CYCLES) {
      printf("sc_attn: mbarrier timeout block=(%d,%d,%d) thread=%d bar=0x%x parity=%u\n",
             blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load global -> this CTA's smem, completion on an mbarrier (bytes).
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* m, uint32_t bar,
                                            int32_t c_inner, int32_t c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst_smem),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c_inner), "r"(c_outer)
      : "memory");
}
// pull a tile into L2 only (no smem destination, no completion): hides DRAM latency of a later load
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int32_t c_inner, int32_t c_outer) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(c_inner), "r"(c_outer)
               : "memory");
}
// multicast variant: the tile lands at the same smem offset in every CTA of cta_mask and
// signals the same-offset mbarrier in each.
__device__ __forceinline__ void tma_load_2d_mcast(uint32_t dst_smem, const CUtensorMap* m,
                                                  uint32_t bar, int32_t c_inner, int32_t c_outer,
                                                  uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst_smem),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c_inner), "r"(c_outer), "h"(cta_mask)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued tcgen05.mma of this thread -> one arrive on `bar` when they retire
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// commit that arrives on the same-offset barrier of every CTA in cta_mask
__device__ __forceinline__ void umma_commit_mcast(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
          "r"(bar),
      "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// store 32 consecutive fp32 columns of this thread's lane (base_lane + i); pair with tmem_st_wait()
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// ---------------------------------------------------------------- descriptors
// UMMA shared-memory matrix descriptor for a K-major operand tile whose rows are 64 bf16
// (128 B) wide and stored with the 128-byte swizzle (what TMA SWIZZLE_128B writes):
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4 (unused: 1)
//   bits [32,46) stride byte offset >> 4   (8 rows * 128 B = 1024 B between 8-row groups)
//   bits [46,48) descriptor version = 1 (sm_100)
//   bits [61,64) layout type = 2 (SWIZZLE_128B)
// Advancing along K by 16 elements (one UMMA_K step) adds 32 B to the start address.
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// UMMA instruction descriptor, kind::f16: (bf16 | fp16)^2 -> fp32, A and B K-major, dense.
//   c_format[4,6)=1 (F32)  a_format[7,10), b_format[10,13): 0 = F16, 1 = BF16
//   a_major bit15=0  b_major bit16=0  n_dim[17,23)=N>>3  m_dim[24,29)=M>>4
// kind::f8f6f4 uses the same layout with a_format / b_format 0 = E4M3, so an e4m3 x e4m3 -> fp32 descriptor has
// the bits of the fp16 one.
__host__ __device__ constexpr uint32_t umma_idesc_16b(uint32_t M, uint32_t N, bool is_f16) {
  const uint32_t fmt = is_f16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

}  // namespace scptx

// ---------------------------------------------------------------- clusters / DSMEM
namespace scptx {
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {   // every thread of every CTA in the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// bulk copy from this CTA's smem into a peer's smem; completion (bytes) on the PEER's mbarrier
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes,
                                                  uint32_t mbar_cluster) {
  asm volatile(
      "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
      "r"(src_cta), "r"(bytes), "r"(mbar_cluster)
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
}  // namespace scptx

// ---------------------------------------------------------------- cta_group::2 (CTA pair) variants
namespace scptx {
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {   // both CTAs of the pair call it
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs, M = 256] (+)= A[smem, 128 rows per CTA] * B[smem, N/2 rows per CTA]; issued by ONE
// thread of the even-ranked CTA of the pair.
__device__ __forceinline__ void umma_ss2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (when all prior pair MMAs retire) on the same-offset barrier of every CTA in cta_mask
__device__ __forceinline__ void umma_commit2_mcast(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(cta_mask)
      : "memory");
}
// TMA load whose completion bytes are credited to an mbarrier that may live in the PEER CTA of the pair
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst_smem, const CUtensorMap* m, uint32_t bar_cluster,
                                                int32_t c_inner, int32_t c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst_smem),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c_inner), "r"(c_outer)
      : "memory");
}
}  // namespace scptx

// ---------------------------------------------------------------- fused issue + probe (single-thread issue loops)
// The TMA-producer and MMA-issuer loops each have the time four N=128 UMMAs execute in (256 cycles) per ring
// stage.  Measured on B200: try_wait on a completed mbarrier 117 cycles, one UMMA issue ~36, one TMA issue
// ~70 — a naive wait-then-issue loop does not fit.  These helpers put the non-blocking probe of the NEXT
// ring slot, the elected issue of this stage's instructions and the read-back of the probe into ONE asm
// block, so that ptxas cannot place the probe's consumer before the issues and the probe latency is hidden.
// Call them warp-uniformly (all 32 lanes); they return 1 when the probed phase had completed.
namespace scptx {

// probe + 4 pair UMMAs (one 128-byte K chunk: 64 16-bit or 128 e4m3 elements) + commit to bar1 (and to bar2 iff
// flag2).  kF8: kind::f8f6f4 (K = 32 per instruction) instead of kind::f16 (K = 16) — the same 32-byte descriptor
// advance and the same cycles per instruction, twice the contraction length.
#define SC_UMMA4_CG2_PROBE_ASM(KIND)                                                                                \
  asm volatile(                                                                                                     \
      "{\n\t"                                                                                                       \
      ".reg .pred pp, pe, pm, pa, pt, p2;\n\t"                                                                      \
      "mbarrier.test_wait.parity.shared::cta.b64 pp, [%18], %19;\n\t"                                               \
      "elect.sync _|pe, 0xffffffff;\n\t"                                                                            \
      "setp.ne.and.b32 pm, %12, 0, pe;\n\t"                                                                         \
      "setp.ne.b32 pa, %11, 0;\n\t"                                                                                 \
      "setp.eq.b32 pt, %11, %11;\n\t"                                                                               \
      "@pm tcgen05.mma.cta_group::2.kind::" KIND " [%1], %2, %6, %10, pa;\n\t"                                      \
      "@pm tcgen05.mma.cta_group::2.kind::" KIND " [%1], %3, %7, %10, pt;\n\t"                                      \
      "@pm tcgen05.mma.cta_group::2.kind::" KIND " [%1], %4, %8, %10, pt;\n\t"                                      \
      "@pm tcgen05.mma.cta_group::2.kind::" KIND " [%1], %5, %9, %10, pt;\n\t"                                      \
      "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%13], %14;\n\t" \
      "setp.ne.and.b32 p2, %17, 0, pe;\n\t"                                                                         \
      "@p2 tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%15], %16;\n\t" \
      "selp.u32 %0, 1, 0, pp;\n\t"                                                                                  \
      "}\n"                                                                                                         \
      : "=r"(ok)                                                                                                    \
      : "r"(d_tmem), "l"(a0), "l"(a1), "l"(a2), "l"(a3), "l"(b0), "l"(b1), "l"(b2), "l"(b3), "r"(idesc), "r"(acc0),  \
        "r"(enable), "r"(bar1), "h"(mask1), "r"(bar2), "h"(mask2), "r"(flag2), "r"(probe_bar), "r"(probe_parity)    \
      : "memory")
template <bool kF8 = false>
__device__ __forceinline__ uint32_t umma4_cg2_probe(uint32_t d_tmem, uint64_t a0, uint64_t a1, uint64_t a2,
                                                    uint64_t a3, uint64_t b0, uint64_t b1, uint64_t b2, uint64_t b3,
                                                    uint32_t idesc, uint32_t acc0, uint32_t enable, uint32_t bar1,
                                                    uint16_t mask1, uint32_t bar2, uint16_t mask2, uint32_t flag2,
                                                    uint32_t probe_bar, uint32_t probe_parity) {
  uint32_t ok;
  if constexpr (kF8) {
    SC_UMMA4_CG2_PROBE_ASM("f8f6f4");
  } else {
    SC_UMMA4_CG2_PROBE_ASM("f16");
  }
  return ok;
}
#undef SC_UMMA4_CG2_PROBE_ASM

// probe + (expect_tx of tx_bytes on full_local iff tx_bytes != 0, plain arrive iff plain != 0) + up to two
// pair TMA loads whose bytes are credited to bar_cluster
__device__ __forceinline__ uint32_t tma2_cg2_probe(uint32_t dst0, const CUtensorMap* m0, int32_t x0, int32_t y0,
                                                   uint32_t on0, uint32_t dst1, const CUtensorMap* m1, int32_t x1,
                                                   int32_t y1, uint32_t on1, uint32_t bar_cluster, uint32_t full_local,
                                                   uint32_t tx_bytes, uint32_t plain, uint32_t probe_bar,
                                                   uint32_t probe_parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred pp, pe, px, pa, p0, p1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 pp, [%15], %16;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.and.b32 px, %13, 0, pe;\n\t"
      "@px mbarrier.arrive.expect_tx.shared::cta.b64 _, [%12], %13;\n\t"
      "setp.ne.and.b32 pa, %14, 0, pe;\n\t"
      "@pa mbarrier.arrive.shared::cta.b64 _, [%12];\n\t"
      "setp.ne.and.b32 p0, %5, 0, pe;\n\t"
      "@p0 cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%1], [%2, {%3, %4}], [%11];\n\t"
      "setp.ne.and.b32 p1, %10, 0, pe;\n\t"
      "@p1 cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%6], [%7, {%8, %9}], [%11];\n\t"
      "selp.u32 %0, 1, 0, pp;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(dst0), "l"(reinterpret_cast<uint64_t>(m0)), "r"(x0), "r"(y0), "r"(on0), "r"(dst1),
        "l"(reinterpret_cast<uint64_t>(m1)), "r"(x1), "r"(y1), "r"(on1), "r"(bar_cluster), "r"(full_local),
        "r"(tx_bytes), "r"(plain), "r"(probe_bar), "r"(probe_parity)
      : "memory");
  return ok;
}
}  // namespace scptx
