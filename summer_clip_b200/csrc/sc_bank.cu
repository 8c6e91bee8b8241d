// Label-sorted key bank for sc_attn_fwd_hard (one-hot cache values: HardCacheStrategy cache_value_strategy.py:14-17,
// gold-label caches image_attention.py:65-66, Tip-Adapter cache_values tip_adapter/utils.py:62).
//   sc_hard_bank_layout : stable counting sort of the keys by label with every class segment padded to whole
//                         16-key groups -> perm (sorted position -> original key, -1 = padding), the class of
//                         every 16-key group and one validity bit per key.  Stable (original order within a
//                         class), hence bit-reproducible: chunk histograms -> column scan over chunks + padded
//                         class scan -> ordered scatter (one warp walks a chunk in 32-key batches, ranks inside a
//                         batch from __match_any_sync).
//   sc_gather_rows      : Ks[j] = Kn[perm[j]] (zero rows for padding), one warp per 16-byte-vectorised row.
// Both are HBM-light (8-10 bytes per key + one pass over the bank) and run once per cache.
#include "sc_common.cuh"

namespace {

constexpr int kChunk = 1024;     // keys per chunk (one warp walks it in order)

struct BankWs {
  int32_t* counts;      // [n_chunks, C]  keys of class c in chunk b  ->  exclusive scan over chunks (in place)
  int64_t* seg_start;   // [C + 1]        first sorted position of class c (multiples of 16)
};

size_t bank_ws_bytes(int64_t n_keys, int32_t C) {
  const int64_t n_chunks = sc::ceil_div(n_keys > 0 ? n_keys : 1, kChunk);
  return static_cast<size_t>(sc::round_up(n_chunks * C * 4, 256) + sc::round_up((static_cast<int64_t>(C) + 1) * 8, 256));
}

BankWs carve_bank(void* ws, int64_t n_keys, int32_t C) {
  const int64_t n_chunks = sc::ceil_div(n_keys > 0 ? n_keys : 1, kChunk);
  BankWs w;
  w.counts = static_cast<int32_t*>(ws);
  w.seg_start = reinterpret_cast<int64_t*>(static_cast<char*>(ws) + sc::round_up(n_chunks * C * 4, 256));
  return w;
}

__global__ void __launch_bounds__(256)
bank_hist_kernel(const int16_t* __restrict__ labels, int64_t n_keys, int32_t C, int32_t* __restrict__ counts) {
  const int64_t b = blockIdx.x;
  const int64_t k0 = b * kChunk;
  int32_t* row = counts + b * C;
  for (int i = threadIdx.x; i < kChunk; i += blockDim.x) {
    const int64_t k = k0 + i;
    if (k < n_keys) {
      const int c = labels[k];
      if (c >= 0 && c < C) atomicAdd(&row[c], 1);        // counts only: order-free
    }
  }
}

// exclusive scan of every class column over the chunks (in place), class total out.  Block = 32 classes x 32 slices
// of the chunk range: a thread sums its slice (independent, coalesced 128-byte loads), the 32 slice sums of a class
// are scanned through shared memory, and the slice is rewritten with its running prefix.  (A thread per class walking
// all 1252 chunks of the ImageNet bank serially took 0.32 ms on four SMs.)
__global__ void __launch_bounds__(1024)
bank_colscan_kernel(int32_t* __restrict__ counts, int64_t n_chunks, int32_t C, int64_t* __restrict__ totals) {
  __shared__ int32_t part[32][33];
  const int cl = threadIdx.x & 31, j = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const int64_t per = (n_chunks + 31) / 32;
  const int64_t b0 = j * per, b1 = (b0 + per < n_chunks) ? b0 + per : n_chunks;
  int32_t sum = 0;
  if (c < C) {
#pragma unroll 8
    for (int64_t b = b0; b < b1; ++b) sum += counts[b * C + c];
  }
  part[j][cl] = sum;
  __syncthreads();
  if (j == 0) {
    int32_t run = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int32_t v = part[i][cl];
      part[i][cl] = run;
      run += v;
    }
    if (c < C) totals[c] = run;
  }
  __syncthreads();
  if (c < C) {
    int32_t run = part[j][cl];
    for (int64_t b = b0; b < b1; b += 8) {       // eight loads in flight, then the eight prefixes
      int32_t v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (b + i < b1) ? counts[(b + i) * C + c] : 0;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (b + i < b1) {
          counts[(b + i) * C + c] = run;
          run += v[i];
        }
    }
  }
}

// one block: seg_start[c] = sum_{c' < c} pad16(total[c']) (in place over totals), seg_start[C] = n_sorted
__global__ void __launch_bounds__(1024)
bank_segscan_kernel(int64_t* __restrict__ seg, int32_t C, int64_t* __restrict__ n_sorted) {
  __shared__ int64_t part[1024];
  const int t = threadIdx.x;
  const int per = (C + 1023) / 1024;
  const int lo = t * per, hi = min(C, lo + per);
  int64_t s = 0;
  for (int c = lo; c < hi; ++c) s += (seg[c] + 15) / 16 * 16;
  part[t] = s;
  __syncthreads();
  if (t == 0) {
    int64_t run = 0;
    for (int i = 0; i < 1024; ++i) {
      const int64_t v = part[i];
      part[i] = run;
      run += v;
    }
    seg[C] = run;
    *n_sorted = run;
  }
  __syncthreads();
  int64_t run = part[t];
  for (int c = lo; c < hi; ++c) {
    const int64_t v = (seg[c] + 15) / 16 * 16;
    seg[c] = run;
    run += v;
  }
}

// one warp per chunk, keys in order: stable positions
__global__ void __launch_bounds__(32)
bank_scatter_kernel(const int16_t* __restrict__ labels, int64_t n_keys, int32_t C, int32_t* __restrict__ counts,
                    const int64_t* __restrict__ seg_start, int64_t* __restrict__ perm, int16_t* __restrict__ gcls) {
  const int lane = threadIdx.x;
  const int64_t b = blockIdx.x;
  int32_t* base = counts + b * C;           // running position of every class inside its segment (this chunk's row)
  for (int i0 = 0; i0 < kChunk; i0 += 32) {
    const int64_t k = b * kChunk + i0 + lane;
    int c = -1;
    if (k < n_keys) {
      c = labels[k];
      if (c < 0 || c >= C) c = -1;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    if (c >= 0) {
      const int rank = __popc(peers & ((1u << lane) - 1u));
      const int leader = __ffs(peers) - 1;
      int32_t pos0 = 0;
      if (lane == leader) {
        pos0 = base[c];
        base[c] = pos0 + __popc(peers);
      }
      pos0 = __shfl_sync(peers, pos0, leader);
      const int64_t dest = seg_start[c] + pos0 + rank;
      perm[dest] = k;
      gcls[dest >> 4] = static_cast<int16_t>(c);      // every key of a group writes the same class
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256)
bank_bits_kernel(const int64_t* __restrict__ perm, int64_t n_words, uint32_t* __restrict__ bits) {
  const int64_t w = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  uint32_t v = 0;
#pragma unroll 8
  for (int j = 0; j < 32; ++j) v |= (perm[w * 32 + j] >= 0 ? 1u : 0u) << j;
  bits[w] = v;
}

// warp per output row of row_bytes (multiple of 16)
__global__ void __launch_bounds__(256)
gather_rows_kernel(const uint4* __restrict__ src, int64_t n_src, const int64_t* __restrict__ perm, int64_t n_out,
                   int64_t vec_per_row, uint4* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_per_grid = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t o = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); o < n_out;
       o += warps_per_grid) {
    const int64_t r = perm[o];
    const bool ok = r >= 0 && r < n_src;
    const uint4* s = src + (ok ? r : 0) * vec_per_row;
    uint4* d = dst + o * vec_per_row;
    for (int64_t j = lane; j < vec_per_row; j += 32) d[j] = ok ? __ldg(s + j) : make_uint4(0u, 0u, 0u, 0u);
  }
}

// inverse of the bank permutation (original key -> sorted position; -1 for keys the layout dropped) and zero rows
// for the padding positions of the sorted bank: what a normalise pass needs to write each key's row straight to its
// sorted place (sc_normalize_scatter), instead of a second pass over the bank (sc_gather_rows).
__global__ void __launch_bounds__(256)
bank_inverse_kernel(const int64_t* __restrict__ perm, int64_t n_sorted_rows, int64_t n_keys, int64_t* __restrict__ inv,
                    uint4* __restrict__ rows, int64_t row_vecs) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t j = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); j < n_sorted_rows; j += warps) {
    const int64_t k = perm[j];
    if (k >= 0) {
      if (lane == 0 && k < n_keys) inv[k] = j;
    } else if (rows != nullptr) {
      for (int64_t v = lane; v < row_vecs; v += 32) rows[j * row_vecs + v] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

}  // namespace

extern "C" {

int64_t sc_hard_bank_capacity(int64_t n_keys, int32_t n_classes) {
  return sc::round_up((n_keys > 0 ? n_keys : 0) + 15ll * n_classes + 1, 256);
}

size_t sc_hard_bank_workspace_bytes(int64_t n_keys, int32_t n_classes) { return bank_ws_bytes(n_keys, n_classes); }

int sc_hard_bank_layout(const int16_t* labels16, int64_t n_keys, int32_t n_classes, int64_t* perm,
                        int16_t* group_class, uint32_t* key_bits, int64_t capacity, int64_t* n_sorted,
                        void* workspace, size_t ws_bytes, void* stream) {
  SC_REQUIRE(perm && group_class && key_bits && n_sorted && workspace, SC_EINVAL, "sc_hard_bank_layout: null pointer");
  SC_REQUIRE(labels16 || n_keys == 0, SC_EINVAL, "sc_hard_bank_layout: null labels");
  SC_REQUIRE(n_keys >= 0 && n_classes > 0 && n_classes <= 32767, SC_ESHAPE, "sc_hard_bank_layout: bad shape");
  SC_REQUIRE(capacity % 256 == 0 && capacity >= sc_hard_bank_capacity(n_keys, n_classes), SC_ESHAPE,
             "sc_hard_bank_layout: capacity must be a multiple of 256 and >= sc_hard_bank_capacity()");
  SC_REQUIRE(ws_bytes >= bank_ws_bytes(n_keys, n_classes), SC_EINVAL, "sc_hard_bank_layout: workspace too small");
  SC_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 256 == 0, SC_EALIGN, "sc_hard_bank_layout: workspace must be 256-B aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t n_chunks = sc::ceil_div(n_keys > 0 ? n_keys : 1, kChunk);
  BankWs w = carve_bank(workspace, n_keys, n_classes);
  SC_CUDA(cudaMemsetAsync(w.counts, 0, static_cast<size_t>(n_chunks) * n_classes * 4, st));
  SC_CUDA(cudaMemsetAsync(perm, 0xFF, static_cast<size_t>(capacity) * 8, st));            // -1
  SC_CUDA(cudaMemsetAsync(group_class, 0xFF, static_cast<size_t>(capacity / 16) * 2, st)); // -1
  if (n_keys > 0) bank_hist_kernel<<<static_cast<unsigned>(n_chunks), 256, 0, st>>>(labels16, n_keys, n_classes, w.counts);
  bank_colscan_kernel<<<static_cast<unsigned>(sc::ceil_div(n_classes, 32)), 1024, 0, st>>>(w.counts, n_chunks, n_classes,
                                                                                          w.seg_start);
  bank_segscan_kernel<<<1, 1024, 0, st>>>(w.seg_start, n_classes, n_sorted);
  if (n_keys > 0)
    bank_scatter_kernel<<<static_cast<unsigned>(n_chunks), 32, 0, st>>>(labels16, n_keys, n_classes, w.counts, w.seg_start,
                                                                        perm, group_class);
  bank_bits_kernel<<<static_cast<unsigned>(sc::ceil_div(capacity / 32, 256)), 256, 0, st>>>(perm, capacity / 32, key_bits);
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

int sc_hard_bank_inverse(const int64_t* perm, int64_t n_sorted_rows, int64_t n_keys, int64_t* inv, void* rows,
                         int64_t row_bytes, void* stream) {
  SC_REQUIRE(perm && inv, SC_EINVAL, "sc_hard_bank_inverse: null pointer");
  SC_REQUIRE(n_sorted_rows >= 0 && n_keys >= 0 && (rows == nullptr || (row_bytes > 0 && row_bytes % 16 == 0)), SC_ESHAPE,
             "sc_hard_bank_inverse: bad shape");
  SC_REQUIRE(reinterpret_cast<uintptr_t>(rows) % 16 == 0, SC_EALIGN, "sc_hard_bank_inverse: rows must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_keys > 0) SC_CUDA(cudaMemsetAsync(inv, 0xff, static_cast<size_t>(n_keys) * sizeof(int64_t), st));     // -1: dropped
  if (n_sorted_rows == 0) return SC_OK;
  const int64_t want = sc::ceil_div(n_sorted_rows, 8);
  const unsigned blocks = static_cast<unsigned>(want < 148 * 16 ? want : 148 * 16);
  bank_inverse_kernel<<<blocks, 256, 0, st>>>(perm, n_sorted_rows, n_keys, inv, static_cast<uint4*>(rows),
                                              rows ? row_bytes / 16 : 0);
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

int sc_gather_rows(const void* src, int64_t n_src, int64_t row_bytes, const int64_t* perm, int64_t n_out, void* dst,
                   void* stream) {
  SC_REQUIRE(src && perm && dst, SC_EINVAL, "sc_gather_rows: null pointer");
  SC_REQUIRE(n_src >= 0 && n_out >= 0 && row_bytes > 0 && row_bytes % 16 == 0, SC_ESHAPE,
             "sc_gather_rows: row_bytes must be a positive multiple of 16");
  SC_REQUIRE((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) % 16 == 0, SC_EALIGN,
             "sc_gather_rows: buffers must be 16-byte aligned");
  if (n_out == 0) return SC_OK;
  const int64_t want = sc::ceil_div(n_out, 8);
  const unsigned blocks = static_cast<unsigned>(want < 148 * 16 ? want : 148 * 16);
  gather_rows_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(src), n_src, perm, n_out, row_bytes / 16, static_cast<uint4*>(dst));
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

}  // extern "C"
