// Micro-benchmarks that size the CLIP-search attention kernel's design (run through gpurun):
//   occ   : co-resident clusters for cluster sizes 1..16 at the kernel's smem footprint
//   mma   : tcgen05.mma cta_group::2 issue rate, M=256, N=128 vs N=256, SS operands (no loads)
//   tma   : L2 -> smem TMA streaming rate per SM / chip-wide, unicast and cluster multicast
//   dsmem : cp.async.bulk shared::cta -> shared::cluster rate per CTA, fan-out 1..7
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench ubench.cu -lcuda
#include <cstdio>
#include "../../summer_clip_b200/csrc/sc_ptx.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace scptx;

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);           \
      exit(1);                                                                                  \
    }                                                                                           \
  } while (0)

// ------------------------------------------------------------------------------------------ occ
__global__ void dummy_kernel(int* p) {
  extern __shared__ uint8_t sm[];
  if (p) p[0] = sm[0];
}

// ------------------------------------------------------------------------------------------ mma
template <int N>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  // pseudo-random small fp16 operands (data toggling matters for power)
  uint32_t* w = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) {
    uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
    h ^= h >> 13;
    w[i] = (h & 0x83ff83ffu) | 0x30003000u;   // |x| in [0.125, 0.25)
  }
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc2(smem_u32(&slot), 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (rank == 0 && warp == 0) {
    const uint32_t idesc = umma_idesc_16b(256, N, true);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
        const uint64_t a = umma_desc_k128(base + (it & 1) * 32768);
        const uint64_t b = umma_desc_k128(base + (it & 1) * 32768 + 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss2(tmem + ((it >> 4) & 1) * 256, a + 2 * k, b + 2 * k, idesc, (it & 15) | k);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit2_mcast(smem_u32(&bar), 1);
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x / 2] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------ tma
struct TmaBars {
  uint64_t full[16];
  uint64_t empty[16];
};
struct TmaCfg {
  int iters, ns, cs, mcast, np, box_rows, col_chunks, row_blocks, bulk1d;
};
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__global__ void __launch_bounds__(160, 1)
tma_stream_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t* buf, const TmaCfg c, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ TmaBars bars;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int ns = c.ns, cs = c.cs, iters = c.iters;
  const uint32_t stage_bytes = (uint32_t)c.box_rows * 128u;
  if (threadIdx.x == 0) {
    prefetch_tmap(&tm);
    for (int s = 0; s < ns; ++s) {
      mbar_init(smem_u32(&bars.full[s]), 1);
      mbar_init(smem_u32(&bars.empty[s]), c.mcast ? cs : 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  cluster_sync_all();
  const int cluster_id = blockIdx.x / cs;
  const uint32_t nbox_mask = (uint32_t)(c.col_chunks * c.row_blocks) - 1u;      // power of two
  const uint32_t cc_mask = (uint32_t)c.col_chunks - 1u;
  const int cc_shift = 31 - __clz(c.col_chunks);
  const uint32_t boff = (uint32_t)(c.mcast ? cluster_id : blockIdx.x) * 101u;
  const long long t0 = clock64();
  if (warp < c.np && lane == 0) {
    for (int i = 0; i < iters; ++i) {
      if (c.mcast && (i % cs) != (int)rank) continue;
      if ((i % c.np) != warp) continue;
      const int s = i % ns;
      if (i >= ns) mbar_wait<true>(smem_u32(&bars.empty[s]), ((i / ns) - 1) & 1);
      const uint32_t box = ((uint32_t)i * 37u + boff) & nbox_mask;
      const int cc = (int)(box & cc_mask), rb = (int)(box >> cc_shift);
      if (c.bulk1d)
        bulk_load_1d(base + s * stage_bytes, buf + (size_t)box * stage_bytes, stage_bytes, smem_u32(&bars.full[s]));
      else if (c.mcast)
        tma_load_2d_mcast(base + s * stage_bytes, &tm, smem_u32(&bars.full[s]), cc * 64, rb * c.box_rows,
                          (uint16_t)((1u << cs) - 1));
      else
        tma_load_2d(base + s * stage_bytes, &tm, smem_u32(&bars.full[s]), cc * 64, rb * c.box_rows);
    }
  } else if (warp == 4 && lane == 0) {
    for (int s = 0; s < ns && s < iters; ++s) mbar_arrive_expect_tx(smem_u32(&bars.full[s]), stage_bytes);
    for (int i = 0; i < iters; ++i) {
      const int s = i % ns;
      mbar_wait(smem_u32(&bars.full[s]), (i / ns) & 1);
      if (i + ns < iters) mbar_arrive_expect_tx(smem_u32(&bars.full[s]), stage_bytes);
      if (c.mcast) mbar_arrive_cluster(smem_u32(&bars.empty[s]), (uint32_t)(i % cs));
      else mbar_arrive(smem_u32(&bars.empty[s]));
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  cluster_sync_all();
}

// ------------------------------------------------------------------------------------------ tma2
// One thread issues a batch of `nb` box loads back to back, then waits for all of them: separates the
// per-instruction issue cost from the fabric rate.
__global__ void __launch_bounds__(32, 1)
tma_batch_kernel(const __grid_constant__ CUtensorMap tm, int iters, int nb, int box_rows, int nbox_mask,
                 long long* cycles, long long* issue_cycles, int shared_stream, int skew) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = (uint32_t)box_rows * 128u;
  if (threadIdx.x == 0) {
    prefetch_tmap(&tm);
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    long long issue = 0;
    uint32_t box = shared_stream ? (blockIdx.x * (uint32_t)skew) : blockIdx.x * 101u;
    const uint32_t step = shared_stream ? 1u : 37u;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      mbar_arrive_expect_tx(smem_u32(&bar), stage_bytes * nb);
      const long long a = clock64();
      for (int b = 0; b < nb; ++b) {
        box = (box + step) & (uint32_t)nbox_mask;
        tma_load_2d(base + b * stage_bytes, &tm, smem_u32(&bar), (int)(box & 15u) * 64, (int)(box >> 4) * box_rows);
      }
      issue += clock64() - a;
      mbar_wait(smem_u32(&bar), it & 1);
    }
    cycles[blockIdx.x] = clock64() - t0;
    issue_cycles[blockIdx.x] = issue;
  }
}

// ------------------------------------------------------------------------------------------ sync
// Per-instruction costs of the pipeline skeleton: try_wait on a completed barrier, tcgen05.commit issue and
// round trip, UMMA issue, elect + syncwarp.
__global__ void __launch_bounds__(128, 1) sync_cost_kernel(long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[4];
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  constexpr int N = 256;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), N);
    mbar_init(smem_u32(&bars[2]), 1);
    mbar_init(smem_u32(&bars[3]), N);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    const uint32_t b0 = smem_u32(&bars[0]), b1 = smem_u32(&bars[1]), b2 = smem_u32(&bars[2]), b3 = smem_u32(&bars[3]);
    long long t[12];
    // (0) try_wait on an already completed phase
    if (elect_one()) mbar_arrive(b0);
    __syncwarp();
    t[0] = clock64();
    for (int i = 0; i < N; ++i) mbar_wait(b0, 0);
    t[1] = clock64();
    // (1) commit issue cost (empty pipe), N commits to one barrier
    if (elect_one()) {
      for (int i = 0; i < N; ++i) umma_commit(b1);
    }
    __syncwarp();
    t[2] = clock64();
    mbar_wait(b1, 0);
    t[3] = clock64();
    // (2) commit -> wait round trip, serial
    for (int i = 0; i < 64; ++i) {
      if (elect_one()) umma_commit(b2);
      __syncwarp();
      mbar_wait(b2, i & 1);
    }
    t[4] = clock64();
    // (3) elect + syncwarp only
    int acc = 0;
    for (int i = 0; i < N; ++i) {
      if (elect_one()) acc += i;
      __syncwarp();
    }
    t[5] = clock64();
    // (4) 4 small UMMAs (M128 N16 K16: 8 cycles each) + commit per iteration: issue-rate bound
    const uint32_t idesc = umma_idesc_16b(128, 16, true);
    const uint64_t a = umma_desc_k128(base), b = umma_desc_k128(base + 16384);
    for (int i = 0; i < N; ++i) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem, a + 2 * k, b + 2 * k, idesc, 1);
        umma_commit(b3);
      }
      __syncwarp();
    }
    t[6] = clock64();
    mbar_wait(b3, 0);
    t[7] = clock64();
    if (threadIdx.x == 0) {
      out[0] = (t[1] - t[0]) / N;          // try_wait (complete)
      out[1] = (t[2] - t[1]) / N;          // commit issue
      out[2] = t[3] - t[2];                // drain after N commits
      out[3] = (t[4] - t[3]) / 64;         // commit + wait round trip
      out[4] = (t[5] - t[4]) / N + (acc == 12345 ? 1 : 0);   // elect + syncwarp
      out[5] = (t[6] - t[5]) / N;          // 4 UMMA + commit issue loop
      out[6] = t[7] - t[6];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------ dsmem
__global__ void __launch_bounds__(32, 1)
dsmem_kernel(int iters, int cs, int fan, int bytes, int nb, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t rx[8];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t src = base;                    // 32 KB
  const uint32_t dst0 = base + 32768;           // nb slots
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int b = 0; b < nb; ++b) mbar_init(smem_u32(&rx[b]), 1);
    fence_barrier_init();
  }
  __syncwarp();
  cluster_sync_all();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int b = it % nb;
    if (b == 0 && it > 0) cluster_sync_all();
    if (threadIdx.x == 0) {
      mbar_arrive_expect_tx(smem_u32(&rx[b]), (uint32_t)(fan * bytes));
      for (int f = 1; f <= fan; ++f) {
        const uint32_t peer = (rank + f) % cs;
        bulk_copy_to_peer(mapa(dst0 + b * bytes, peer), src, (uint32_t)bytes, mapa(smem_u32(&rx[b]), peer));
      }
      mbar_wait<true>(smem_u32(&rx[b]), (it / nb) & 1);
    }
    __syncwarp();
  }
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  cluster_sync_all();
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename K, typename... Args>
float launch_cluster(K kernel, int grid, int block, int smem, int cs, Args... args) {
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if (cs > 8) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  CK(cudaLaunchKernelEx(&cfg, kernel, args...));
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms;
}

static double mean(const std::vector<long long>& v, int n) {
  double s = 0;
  for (int i = 0; i < n; ++i) s += (double)v[i];
  return s / n;
}


// ------------------------------------------------------------------------------------------ exp
// The exp warps' inner loop in isolation: per element one FFMA, one ex2.approx (MUFU) and one FADD into a running sum,
// 16-element groups as in sc_attn_seg.cu.  mode 0: that loop; 1: MUFU only (independent); 2: 4 independent sums
// (shorter FADD chains); 3: half of the exponentials on the FMA pipe (Cody-Waite + degree-5 polynomial).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_poly(float x) {
  // 2^x for x <= 0 (x >= -126): split x = n + f, f in [-0.5, 0.5]; 2^f by a degree-5 minimax polynomial
  const float n = rintf(x);
  const float f = x - n;
  float p = 1.3333558146e-3f;
  p = fmaf(p, f, 9.6181291076e-3f);
  p = fmaf(p, f, 5.5504108665e-2f);
  p = fmaf(p, f, 2.4022650696e-1f);
  p = fmaf(p, f, 6.9314718056e-1f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (static_cast<int>(n) << 23));
}
__global__ void __launch_bounds__(128, 1) exp_loop_kernel(int iters, int mode, float c1, float c0, long long* cycles, float* sink) {
  float x[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] = -0.001f * (threadIdx.x + j);
  float acc = 0.f, acc2 = 0.f, acc3 = 0.f, acc4 = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) s += ex2_approx(fmaf(x[16 * h + j], c1, c0));
        acc += s;
      }
    } else if (mode == 1) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = ex2_approx(x[j]);
    } else if (mode == 2) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        acc += ex2_approx(fmaf(x[j], c1, c0));
        acc2 += ex2_approx(fmaf(x[j + 1], c1, c0));
        acc3 += ex2_approx(fmaf(x[j + 2], c1, c0));
        acc4 += ex2_approx(fmaf(x[j + 3], c1, c0));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        acc += ex2_approx(fmaf(x[j], c1, c0));
        acc2 += ex2_poly(fmaf(x[j + 1], c1, c0));
      }
    }
    if (mode != 1) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] += 1e-7f * acc;       // keep the loop body live and dependent
    }
  }
  const long long t1 = clock64();
  float r = acc + acc2 + acc3 + acc4;
#pragma unroll
  for (int j = 0; j < 32; ++j) r += x[j];
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (r == 12345.678f) sink[0] = r;
}

int main(int argc, char** argv) {
  const char* what = argc > 1 ? argv[1] : "all";
  auto want = [&](const char* w) { return !strcmp(what, "all") || !strcmp(what, w); };
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  printf("device %s sms=%d clock=%d MHz\n", prop.name, prop.multiProcessorCount, clk_khz / 1000);
  long long* d_cycles;
  CK(cudaMalloc(&d_cycles, 4096 * sizeof(long long)));
  std::vector<long long> h(4096);

  if (want("exp")) {
    float* d_sink;
    CK(cudaMalloc(&d_sink, 4));
    const int iters = 4096;
    for (int mode = 0; mode < 4; ++mode) {
      exp_loop_kernel<<<148, 128>>>(iters, mode, 7.9f, -7.9f, d_cycles, d_sink);
      CK(cudaDeviceSynchronize());
      exp_loop_kernel<<<148, 128>>>(iters, mode, 7.9f, -7.9f, d_cycles, d_sink);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h.data(), d_cycles, sizeof(long long) * 148, cudaMemcpyDeviceToHost));
      double cyc = 0;
      for (int i = 0; i < 148; ++i) cyc += (double)h[i] / 148;
      printf("exp mode=%d (0 loop, 1 mufu only, 2 four sums, 3 half polynomial): %.2f cycles per element per warp "
             "(4 warps, one per SM sub-partition)\n", mode, cyc / iters / 32);
    }
    CK(cudaFree(d_sink));
  }

  if (want("occ")) {
    for (int cs : {1, 2, 4, 8, 16}) {
      for (int smem : {100 * 1024, 225 * 1024}) {
        CK(cudaFuncSetAttribute(dummy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (cs > 8) CK(cudaFuncSetAttribute(dummy_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs * 1024);
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy_kernel, &cfg);
        printf("occ cluster=%d smem=%dKB -> max active clusters %d (%d SMs) %s\n", cs, smem / 1024, n, n * cs,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
        cudaGetLastError();
      }
    }
  }

  if (want("mma")) {
    const int iters = 20000;
    for (int pass = 0; pass < 2; ++pass)
      for (int n : {128, 256}) {
        for (int grid : {2, 148}) {
          float ms = n == 128 ? launch_cluster(mma_rate_kernel<128>, grid, 128, 66 * 1024, 2, iters, d_cycles)
                              : launch_cluster(mma_rate_kernel<256>, grid, 128, 66 * 1024, 2, iters, d_cycles);
          CK(cudaMemcpy(h.data(), d_cycles, sizeof(long long) * (grid / 2), cudaMemcpyDeviceToHost));
          const double cyc = mean(h, grid / 2);
          const double flop = 2.0 * 256 * n * 16 * 4.0 * iters * (grid / 2);
          printf("mma M=256 N=%d grid=%d: %.1f cycles per MMA (ideal %d), %.3f ms, %.1f TFLOP/s, eff clock %.0f MHz\n", n,
                 grid, cyc / (4.0 * iters), n / 2, ms, flop / ms * 1e-9, cyc / ms * 1e-3);
        }
      }
  }

  if (want("tma")) {
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    struct Cfg { int cs, mcast, ns, grid, np, box_rows, pitch, bulk1d; size_t mb; };
    const Cfg cfgs[] = {
        {1, 0, 8, 148, 1, 128, 128, 0, 32},  {1, 0, 8, 148, 2, 128, 128, 0, 32},  {1, 0, 8, 148, 4, 128, 128, 0, 32},
        {1, 0, 8, 148, 1, 128, 2048, 0, 32}, {1, 0, 8, 148, 2, 128, 2048, 0, 32}, {1, 0, 8, 148, 4, 128, 2048, 0, 32},
        {1, 0, 4, 148, 1, 128, 2048, 0, 32}, {1, 0, 2, 148, 1, 128, 2048, 0, 32}, {1, 0, 12, 148, 1, 128, 2048, 0, 32},
        {1, 0, 4, 148, 1, 256, 128, 0, 32},  {1, 0, 4, 148, 2, 256, 2048, 0, 32}, {1, 0, 12, 148, 4, 64, 2048, 0, 32},
        {1, 0, 8, 148, 1, 128, 128, 1, 32},  {1, 0, 8, 148, 4, 128, 128, 1, 32},  {1, 0, 8, 1, 1, 128, 2048, 0, 32},
        {1, 0, 8, 148, 1, 128, 2048, 0, 8},  {1, 0, 8, 148, 1, 128, 2048, 0, 2048}, {1, 0, 12, 148, 2, 128, 2048, 0, 2048},
        {2, 1, 8, 148, 1, 128, 2048, 0, 32},
        {4, 1, 8, 132, 1, 128, 2048, 0, 32}, {8, 1, 8, 120, 1, 128, 2048, 0, 32}, {4, 0, 8, 132, 1, 128, 2048, 0, 32}};
    void* buf;
    CK(cudaMalloc(&buf, 2048ull << 20));
    CK(cudaMemset(buf, 1, 2048ull << 20));
    for (int pass = 0; pass < 2; ++pass)
      for (const Cfg& c : cfgs) {
        const size_t buf_bytes = c.mb << 20;
        CUtensorMap tm;
        cuuint64_t dims[2] = {(cuuint64_t)c.pitch / 2, buf_bytes / c.pitch};
        cuuint64_t strides[1] = {(cuuint64_t)c.pitch};
        cuuint32_t box[2] = {64, (cuuint32_t)c.box_rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        TmaCfg k;
        k.iters = 4000; k.ns = c.ns; k.cs = c.cs; k.mcast = c.mcast; k.np = c.np; k.box_rows = c.box_rows;
        k.col_chunks = c.pitch / 128; k.row_blocks = (int)(buf_bytes / c.pitch / c.box_rows); k.bulk1d = c.bulk1d;
        const int stage = c.box_rows * 128;
        const int smem = c.ns * stage + 2048;
        float ms = launch_cluster(tma_stream_kernel, c.grid, 160, smem, c.cs, tm, (const uint8_t*)buf, k, d_cycles);
        CK(cudaMemcpy(h.data(), d_cycles, sizeof(long long) * c.grid, cudaMemcpyDeviceToHost));
        const double cyc = mean(h, c.grid);
        const double delivered = (double)k.iters * stage;
        printf("tma cs=%d mcast=%d stages=%d grid=%d producers=%d box_rows=%d pitch=%d bulk1d=%d ws=%zuMB: %.1f B/cycle/SM delivered, chip %.2f TB/s delivered, %.2f TB/s L2 reads (%.3f ms)\n",
               c.cs, c.mcast, c.ns, c.grid, c.np, c.box_rows, c.pitch, c.bulk1d, c.mb, delivered / cyc,
               delivered * c.grid / ms * 1e-9, delivered * c.grid / (c.mcast ? c.cs : 1) / ms * 1e-9, ms);
      }
    CK(cudaFree(buf));
  }

  if (want("tma2")) {
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    void* buf;
    const size_t buf_bytes = 64ull << 20;
    CK(cudaMalloc(&buf, buf_bytes));
    CK(cudaMemset(buf, 1, buf_bytes));
    long long* d_issue;
    CK(cudaMalloc(&d_issue, 4096 * sizeof(long long)));
    std::vector<long long> hi(4096);
    struct Cfg { int nb, box_rows, grid; };
    const Cfg cfgs[] = {{1, 128, 148}, {2, 128, 148}, {4, 128, 148}, {8, 128, 148}, {12, 128, 148}, {12, 128, 1},
                        {6, 256, 148}, {12, 64, 148}, {12, 32, 148}, {3, 256, 148}, {1, 256, 148}};
    for (int pass = 0; pass < 2; ++pass)
      for (const Cfg& c : cfgs) {
        CUtensorMap tm;
        cuuint64_t dims[2] = {1024, buf_bytes / 2048};
        cuuint64_t strides[1] = {2048};
        cuuint32_t box[2] = {64, (cuuint32_t)c.box_rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int iters = 1000;
        const int stage = c.box_rows * 128;
        const int nboxes = (int)(buf_bytes / stage);     // power of two
        float ms = launch_cluster(tma_batch_kernel, c.grid, 32, c.nb * stage + 2048, 1, tm, iters, c.nb, c.box_rows,
                                  nboxes - 1, d_cycles, d_issue, 0, 0);
        CK(cudaMemcpy(h.data(), d_cycles, sizeof(long long) * c.grid, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hi.data(), d_issue, sizeof(long long) * c.grid, cudaMemcpyDeviceToHost));
        const double cyc = mean(h, c.grid), icyc = mean(hi, c.grid);
        printf("tma2 batch=%d box_rows=%d grid=%d: %.0f cycles per batch (%.0f issuing = %.0f per TMA), %.1f B/cycle/SM, chip %.2f TB/s (%.3f ms)\n",
               c.nb, c.box_rows, c.grid, cyc / iters, icyc / iters, icyc / iters / c.nb,
               (double)iters * c.nb * stage / cyc, (double)iters * c.nb * stage * c.grid / ms * 1e-9, ms);
      }
    CK(cudaFree(buf));
  }

  if (want("sync")) {
    CK(cudaFuncSetAttribute(sync_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024));
    for (int pass = 0; pass < 2; ++pass) {
      sync_cost_kernel<<<1, 128, 40 * 1024>>>(d_cycles);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h.data(), d_cycles, sizeof(long long) * 8, cudaMemcpyDeviceToHost));
      printf("sync: try_wait(complete)=%lld  commit issue=%lld (drain %lld)  commit+wait round trip=%lld  elect+syncwarp=%lld  4xUMMA(N16)+commit loop=%lld (drain %lld)\n",
             h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
    }
  }

  if (want("tma3")) {
    // all SMs stream the SAME box sequence of a 2 GB buffer (the attention kernel's key-bank pattern: HBM is
    // read once, L2 serves every SM), optionally skewed by a few boxes per CTA
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    void* buf;
    const size_t buf_bytes = 2048ull << 20;
    CK(cudaMalloc(&buf, buf_bytes));
    CK(cudaMemset(buf, 1, buf_bytes));
    long long* d_issue;
    CK(cudaMalloc(&d_issue, 4096 * sizeof(long long)));
    struct Cfg { int nb, grid, shared, skew; };
    const Cfg cfgs[] = {{4, 148, 1, 0}, {6, 148, 1, 0}, {12, 148, 1, 0}, {6, 148, 1, 1}, {6, 148, 1, 16}, {6, 148, 1, 256},
                        {6, 148, 0, 0}, {6, 132, 1, 0}, {6, 74, 1, 0}};
    for (int pass = 0; pass < 2; ++pass)
      for (const Cfg& c : cfgs) {
        CUtensorMap tm;
        cuuint64_t dims[2] = {1024, buf_bytes / 2048};
        cuuint64_t strides[1] = {2048};
        cuuint32_t box[2] = {64, 128};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int iters = 3000;
        const int stage = 128 * 128;
        const int nboxes = (int)(buf_bytes / stage);
        float ms = launch_cluster(tma_batch_kernel, c.grid, 32, c.nb * stage + 2048, 1, tm, iters, c.nb, 128,
                                  nboxes - 1, d_cycles, d_issue, c.shared, c.skew);
        CK(cudaMemcpy(h.data(), d_cycles, sizeof(long long) * c.grid, cudaMemcpyDeviceToHost));
        const double cyc = mean(h, c.grid);
        printf("tma3 batch=%d grid=%d shared_stream=%d skew=%d: %.0f cycles per batch, %.1f B/cycle/SM, chip %.2f TB/s (%.3f ms)\n",
               c.nb, c.grid, c.shared, c.skew, cyc / iters, (double)iters * c.nb * stage / cyc,
               (double)iters * c.nb * stage * c.grid / ms * 1e-9, ms);
      }
    CK(cudaFree(buf));
  }

  if (want("dsmem")) {
    const int iters = 2000;
    struct Cfg { int cs, fan, bytes, nb, grid; };
    const Cfg cfgs[] = {{2, 1, 16384, 4, 2},   {2, 1, 16384, 4, 148}, {4, 3, 16384, 4, 4},  {4, 3, 16384, 4, 144},
                        {8, 7, 16384, 4, 8},   {8, 7, 16384, 4, 128}, {8, 7, 32768, 2, 128}, {8, 3, 16384, 4, 128},
                        {8, 1, 16384, 4, 128}, {8, 7, 4096, 4, 128}};
    for (const Cfg& c : cfgs) {
      const int smem = 32768 + c.nb * c.bytes + 2048;
      float ms = launch_cluster(dsmem_kernel, c.grid, 32, smem, c.cs, iters, c.cs, c.fan, c.bytes, c.nb, d_cycles);
      CK(cudaMemcpy(h.data(), d_cycles, sizeof(long long) * c.grid, cudaMemcpyDeviceToHost));
      const double cyc = mean(h, c.grid);
      printf("dsmem cs=%d fan=%d bytes=%d nb=%d grid=%d: %.1f B/cycle/CTA sent (=received), %.0f cycles per iteration (%.3f ms)\n",
             c.cs, c.fan, c.bytes, c.nb, c.grid, (double)iters * c.fan * c.bytes / cyc, cyc / iters, ms);
    }
  }
  return 0;
}
