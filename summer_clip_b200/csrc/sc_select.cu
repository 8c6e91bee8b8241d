// Pseudo-label selection kernels.
//   sc_rowconf        : per-row (confidence, predicted label) of the logits bank, one HBM pass
//                       (TopKStrategy.select cache_strategy.py:67-70, TopKProbStrategy :79-81).
//   sc_topk_per_class : select_topk_per_label (cache_strategy.py:48-59) without the Python loop:
//                       histogram -> scan -> bucket scatter -> per-class 64-bit radix select +
//                       bitonic sort.  Deterministic total order: confidence descending, row index
//                       ascending, so the integer output is reproducible bit for bit.
#include "sc_common.cuh"
#include "sc_rowops.cuh"

namespace {

using sc::MaxIdx;
using sc::row_argmax;
using sc::row_expsum;

template <typename T, bool kVec>
__global__ void __launch_bounds__(256)
rowconf_kernel(const T* __restrict__ L, int64_t N, int64_t C, int64_t ld, float scale, int mode,
               float* __restrict__ conf, int32_t* __restrict__ label) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_per_grid = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < N;
       r += warps_per_grid) {
    const T* row = L + r * ld;
    const MaxIdx m = row_argmax<T, kVec>(row, C, lane);
    float cf = m.v;
    if (mode == SC_CONF_PROB) {
      // softmax(scale*l)[argmax] = exp(0) / sum_c exp(scale*l_c - max)  (torch multiplies the
      // numerator 1.0 by the reciprocal of the sum)
      const float tmax = __fmul_rn(m.v, scale);
      const float s = row_expsum<T, kVec>(row, C, lane, scale, tmax);
      cf = 1.0f / s;
    }
    if (lane == 0) {
      conf[r] = cf;
      label[r] = m.i;
    }
  }
}

// Register-resident variant (sc::RegRow) for rows of at most 32 * NV 16-byte vectors (C <= 1024 fp16 / bf16 with
// NV = 4, C <= 1024 fp32 with NV = 8) whose length is a whole number of vectors: all loads of a lane are in
// flight before the first use, the argmax costs ~4 instructions per element, and the PROB pass sums exp()
// from the registers, so the row is read from memory exactly once.
template <typename T, int NV>
__global__ void __launch_bounds__(256, sizeof(T) == 2 ? 4 : 3)
rowconf_reg_kernel(const T* __restrict__ L, int64_t N, int64_t C, int64_t ld, float scale, int mode,
                   float* __restrict__ conf, int32_t* __restrict__ label) {
  constexpr int kN = 16 / sizeof(T);
  constexpr int kRows = sizeof(T) == 2 ? 2 : 1;      // 16-bit rows are 2 KB: keep two of them in flight per warp
  const int lane = threadIdx.x & 31;
  const int nv = static_cast<int>(C / kN);
  const int64_t warps_per_grid = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t r0 = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r0 < N;
       r0 += warps_per_grid * kRows) {
    sc::RegRow<T, NV> row[kRows];
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
      const int64_t r = r0 + i * warps_per_grid;
      if (r < N) row[i].load(L + r * ld, nv, lane);
    }
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
      const int64_t r = r0 + i * warps_per_grid;
      if (r < N) {
        MaxIdx m;
        float cf;
        if (mode == SC_CONF_PROB) {
          float sum;
          m = row[i].argmax_expsum(scale, sum);
          cf = 1.0f / sum;
        } else {
          m = row[i].argmax();
          cf = m.v;
        }
        if (lane == 0) {
          conf[r] = cf;
          label[r] = m.i;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ per-class top-k
struct TopkWs {
  int32_t* counts;    // [C]
  int32_t* offsets;   // [C + 1]
  int32_t* cursor;    // [C]
  uint64_t* keys;     // [N]
};
__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }
inline size_t topk_ws_bytes(int64_t N, int32_t C) {
  return align256(sizeof(int32_t) * C) + align256(sizeof(int32_t) * (C + 1)) +
         align256(sizeof(int32_t) * C) + align256(sizeof(uint64_t) * (N > 0 ? N : 1));
}
inline TopkWs carve(void* ws, int64_t N, int32_t C) {
  uint8_t* p = static_cast<uint8_t*>(ws);
  TopkWs w;
  w.counts = reinterpret_cast<int32_t*>(p); p += align256(sizeof(int32_t) * C);
  w.offsets = reinterpret_cast<int32_t*>(p); p += align256(sizeof(int32_t) * (C + 1));
  w.cursor = reinterpret_cast<int32_t*>(p); p += align256(sizeof(int32_t) * C);
  w.keys = reinterpret_cast<uint64_t*>(p);
  (void)N;
  return w;
}

__global__ void __launch_bounds__(256)
hist_kernel(const int32_t* __restrict__ label, int64_t N, int32_t C, int32_t* __restrict__ counts) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < N;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int32_t l = label[i];
    if (l >= 0 && l < C) atomicAdd(&counts[l], 1);
  }
}

__global__ void __launch_bounds__(1024)
scan_kernel(const int32_t* __restrict__ counts, int32_t C, int32_t* __restrict__ offsets) {
  // single block exclusive scan, chunk by chunk
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int32_t base = 0; base < C; base += 1024) {
    const int32_t i = base + threadIdx.x;
    const int32_t v = (i < C) ? counts[i] : 0;
    int32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int32_t w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warp_tot[lane] = w;   // inclusive
    }
    __syncthreads();
    const int32_t before = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + (x - v);
    if (i < C) offsets[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[C] = carry;
}

__global__ void __launch_bounds__(256)
scatter_kernel(const float* __restrict__ conf, const int32_t* __restrict__ label, int64_t N,
               int32_t C, const int32_t* __restrict__ offsets, int32_t* __restrict__ cursor,
               uint64_t* __restrict__ keys) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < N;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int32_t l = label[i];
    if (l < 0 || l >= C) continue;
    const int32_t pos = offsets[l] + atomicAdd(&cursor[l], 1);
    // larger key = more confident, then smaller row index
    keys[pos] = (static_cast<uint64_t>(sc::float_order_key(conf[i])) << 32) |
                static_cast<uint64_t>(0xFFFFFFFFu - static_cast<uint32_t>(i));
  }
}

constexpr int kSelThreads = 256;
constexpr int kMaxK = 1024;

// One block per class: the kk = min(k, n_c) largest 64-bit keys of the class bucket, sorted
// descending.  Keys are unique (row index in the low word), so an 8 x 8-bit MSB-first radix select
// finds the exact kk-th largest key; everything >= it is collected and bitonic-sorted in smem.
__global__ void __launch_bounds__(kSelThreads)
select_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ offsets, int32_t C,
              int32_t k, int64_t* __restrict__ out_idx, int32_t* __restrict__ out_count) {
  __shared__ uint32_t hist[256];
  __shared__ uint64_t sel[kMaxK];
  __shared__ uint64_t s_prefix;
  __shared__ int32_t s_remaining;
  __shared__ int32_t s_nsel;
  const int c = blockIdx.x;
  const int32_t beg = offsets[c], end = offsets[c + 1];
  const int32_t n = end - beg;
  const int32_t kk = n < k ? n : k;
  const uint64_t* bucket = keys + beg;
  const int tid = threadIdx.x;

  uint64_t thresh = 0;
  if (n > kk) {
    if (tid == 0) { s_prefix = 0; s_remaining = kk; }
    __syncthreads();
    for (int pass = 7; pass >= 0; --pass) {
      hist[tid] = 0;   // kSelThreads == 256
      __syncthreads();
      const uint64_t prefix = s_prefix;
      const int shift = pass * 8;
      const uint64_t himask = (pass == 7) ? 0ull : (~0ull << (shift + 8));
      for (int32_t i = tid; i < n; i += kSelThreads) {
        const uint64_t key = bucket[i];
        if ((key & himask) == prefix) atomicAdd(&hist[(key >> shift) & 0xFF], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        int32_t rem = s_remaining;
        int b = 255;
        for (; b > 0; --b) {
          const int32_t cnt = static_cast<int32_t>(hist[b]);
          if (cnt >= rem) break;
          rem -= cnt;
        }
        s_remaining = rem;
        s_prefix = prefix | (static_cast<uint64_t>(b) << shift);
      }
      __syncthreads();
    }
    thresh = s_prefix;
  }
  if (tid == 0) s_nsel = 0;
  __syncthreads();
  for (int32_t i = tid; i < n; i += kSelThreads) {
    const uint64_t key = bucket[i];
    if (key >= thresh) {
      const int32_t pos = atomicAdd(&s_nsel, 1);
      if (pos < kMaxK) sel[pos] = key;
    }
  }
  __syncthreads();
  // pad to a power of two with 0 (smaller than any real key: order keys of real floats are > 0
  // unless conf == -NaN pattern; index word keeps real keys non-zero for i < 2^32 - 1)
  int32_t m = 1;
  while (m < kk) m <<= 1;
  for (int32_t i = kk + tid; i < m; i += kSelThreads) sel[i] = 0;
  __syncthreads();
  // bitonic sort, descending
  for (int32_t size = 2; size <= m; size <<= 1) {
    for (int32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (int32_t i = tid; i < m; i += kSelThreads) {
        const int32_t j = i ^ stride;
        if (j > i) {
          const bool desc = ((i & size) == 0);
          const uint64_t a = sel[i], b = sel[j];
          if (desc ? (a < b) : (a > b)) { sel[i] = b; sel[j] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int32_t i = tid; i < k; i += kSelThreads) {
    int64_t v = -1;
    if (i < kk) v = static_cast<int64_t>(0xFFFFFFFFu - static_cast<uint32_t>(sel[i] & 0xFFFFFFFFull));
    out_idx[static_cast<int64_t>(c) * k + i] = v;
  }
  if (tid == 0) out_count[c] = kk;
}

template <typename T>
int launch_rowconf(const void* L, int64_t N, int64_t C, int64_t ld, float scale, int mode,
                   float* conf, int32_t* label, cudaStream_t st) {
  constexpr int kN = 16 / sizeof(T);
  const bool vec = (reinterpret_cast<uintptr_t>(L) % 16 == 0) && (ld % kN == 0);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = sc::ceil_div(N, 8);
  const unsigned blocks = static_cast<unsigned>(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
  constexpr int NV = (sizeof(T) == 4) ? 8 : 4;          // 32 lanes * NV vectors * kN elements = 1024
  if (vec && C % kN == 0 && C / kN <= 32 * NV && (mode != SC_CONF_PROB || scale > 0.f))
    rowconf_reg_kernel<T, NV><<<blocks, 256, 0, st>>>(static_cast<const T*>(L), N, C, ld, scale, mode, conf, label);
  else if (vec)
    rowconf_kernel<T, true><<<blocks, 256, 0, st>>>(static_cast<const T*>(L), N, C, ld, scale, mode, conf, label);
  else
    rowconf_kernel<T, false><<<blocks, 256, 0, st>>>(static_cast<const T*>(L), N, C, ld, scale, mode, conf, label);
  return 0;
}

}  // namespace

extern "C" {

int sc_rowconf(const void* L, int dtype, int64_t N, int64_t C, int64_t ld, float scale, int mode,
               float* conf, int32_t* label, void* stream) {
  SC_REQUIRE(L && conf && label, SC_EINVAL, "sc_rowconf: null pointer");
  SC_REQUIRE(N >= 0 && C > 0 && ld >= C && C < (1ll << 31), SC_ESHAPE, "sc_rowconf: bad shape");
  SC_REQUIRE(mode == SC_CONF_RAW || mode == SC_CONF_PROB, SC_EINVAL, "sc_rowconf: bad mode %d", mode);
  if (N == 0) return SC_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_DISPATCH_DTYPE(dtype, T, (launch_rowconf<T>(L, N, C, ld, scale, mode, conf, label, st)));
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

size_t sc_topk_workspace_bytes(int64_t N, int32_t C) { return topk_ws_bytes(N, C); }

int sc_topk_per_class(const float* conf, const int32_t* label, int64_t N, int32_t C, int32_t k,
                      int64_t* out_idx, int32_t* out_count, void* workspace, size_t ws_bytes,
                      void* stream) {
  SC_REQUIRE(((conf && label) || N == 0) && out_idx && out_count && workspace, SC_EINVAL,
             "sc_topk_per_class: null pointer");       // an empty shard (N == 0) has no rows to point at
  SC_REQUIRE(N >= 0 && N < 0x7FFFFFFFll && C > 0 && k > 0, SC_ESHAPE, "sc_topk_per_class: bad shape");
  SC_REQUIRE(k <= kMaxK, SC_EUNSUPPORTED, "sc_topk_per_class: k=%d exceeds %d", k, kMaxK);
  SC_REQUIRE(ws_bytes >= topk_ws_bytes(N, C), SC_EINVAL, "sc_topk_per_class: workspace too small");
  SC_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 256 == 0, SC_EALIGN, "sc_topk_per_class: workspace must be 256-B aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TopkWs w = carve(workspace, N, C);
  SC_CUDA(cudaMemsetAsync(w.counts, 0, sizeof(int32_t) * C, st));
  SC_CUDA(cudaMemsetAsync(w.cursor, 0, sizeof(int32_t) * C, st));
  if (N > 0) {
    const int64_t want = sc::ceil_div(N, 256);
    const unsigned blocks = static_cast<unsigned>(want < 148 * 16 ? want : 148 * 16);
    hist_kernel<<<blocks, 256, 0, st>>>(label, N, C, w.counts);
    scan_kernel<<<1, 1024, 0, st>>>(w.counts, C, w.offsets);
    scatter_kernel<<<blocks, 256, 0, st>>>(conf, label, N, C, w.offsets, w.cursor, w.keys);
  } else {
    SC_CUDA(cudaMemsetAsync(w.offsets, 0, sizeof(int32_t) * (C + 1), st));
  }
  select_kernel<<<static_cast<unsigned>(C), kSelThreads, 0, st>>>(w.keys, w.offsets, C, k, out_idx, out_count);
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

}  // extern "C"
