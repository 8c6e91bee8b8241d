// sc_normalize_cast: column L2-normalise + (optional) column gather + transpose + bf16/fp16 cast.
// Reference: cache_weights_strategy.py:19-20 (x / x.norm(dim=0, keepdim=True)) fused with the
// cache gather K[:, idx] of image_attention.py:55; Tip-Adapter's row normalisation
// tip_adapter/utils.py:60,84 is the stride_d == 1 case.  HBM-bound: one read of the source bank,
// one write of the K-major bf16 bank the attention kernel's TMA loads consume.
#include "sc_common.cuh"

namespace {

// ---- source is feature-major ([D, N], stride_n == 1 fast path; any strides are correct).
// Block = 32 bank columns.  Pass 1 accumulates the column sums of squares, pass 2 re-reads the
// 32 x D strip (L2-resident: <= 128 KB per block) and writes it transposed through smem so that
// both the global reads (along n) and the global writes (along d) are coalesced.
template <typename T, typename TO>
__global__ void __launch_bounds__(256)
norm_transpose_kernel(const T* __restrict__ src, int64_t D, int64_t N, int64_t stride_d,
                      int64_t stride_n, const int64_t* __restrict__ idx, int64_t n_out,
                      TO* __restrict__ dst, int64_t D_pad, int normalize, const int64_t* __restrict__ dst_row,
                      float* __restrict__ inv_out) {
  __shared__ float tile[64][33];
  __shared__ float red[8][32];
  __shared__ float inv_norm[32];
  __shared__ int64_t col_off[32];
  const int tx = threadIdx.x & 31;   // column within the strip
  const int ty = threadIdx.x >> 5;   // 0..7
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * 32;

  if (ty == 0) {
    const int64_t o = n0 + tx;
    int64_t n = -1;
    if (o < n_out) n = idx ? idx[o] : o;
    col_off[tx] = (n >= 0 && n < N) ? n * stride_n : -1;
  }
  __syncthreads();
  const int64_t my_off = col_off[tx];

  if (normalize || inv_out != nullptr) {
    float ss = 0.f;
    if (my_off >= 0) {
#pragma unroll 8
      for (int64_t d = ty; d < D; d += 8) {
        const float v = sc::to_f32<T>(src[d * stride_d + my_off]);
        ss = fmaf(v, v, ss);
      }
    }
    red[ty][tx] = ss;
    __syncthreads();
    if (ty == 0) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += red[j][tx];
      inv_norm[tx] = sqrtf(s);   // the NORM; we divide below like the reference does
      if (inv_out != nullptr && n0 + tx < n_out) inv_out[n0 + tx] = 1.0f / sqrtf(s);
    }
    __syncthreads();
  }
  const float nrm = normalize ? inv_norm[tx] : 1.0f;

  for (int64_t d0 = 0; d0 < D_pad; d0 += 64) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t d = d0 + ty + 8 * j;
      float v = 0.f;
      if (d < D && my_off >= 0) {
        v = sc::to_f32<T>(src[d * stride_d + my_off]);
        if (normalize) v = v / nrm;
      }
      tile[ty + 8 * j][tx] = v;
    }
    __syncthreads();
    // write: 32 rows (n) x 64 d; thread handles two adjacent d of one row per step
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = ty + 8 * j;            // strip column -> output row
      const int64_t o = n0 + r;
      if (o < n_out) {
        const int64_t orow = dst_row ? dst_row[o] : o;      // scattered output (the label-sorted bank); < 0 = dropped
        const int d = 2 * tx;
        if (orow >= 0) sc::store2<TO>(dst + orow * D_pad + d0 + d, tile[d][r], tile[d + 1][r]);
      }
    }
    __syncthreads();
  }
}

// ---- 16-bit feature-major source with contiguous columns and no gather (the key-bank build): ONE pass over HBM.
// Block = 32 bank columns x all D rows, 512 threads.  Phase A streams the [D x 32] strip into shared memory
// stored COLUMN-major (64 contiguous bytes per warp-row from HBM, 16 independent loads in flight per lane; the
// odd half-pitch makes the 2-byte column stores conflict free) while every thread accumulates the sum of
// squares of its column; phase B writes the 32 output rows (one 2*D_pad-byte contiguous row per column) with
// one 4-byte shared load, two multiplies by the reciprocal norm and one 4-byte store per pair.  The old kernel
// read the strip twice through a 64 x 33 tile with two block barriers per 64 rows and reached 1.5 TB/s.
template <typename T, typename TO>
__global__ void __launch_bounds__(512)
norm_strip_kernel(const T* __restrict__ src, int64_t D, int64_t N, int64_t stride_d, TO* __restrict__ dst,
                  int64_t D_pad, int normalize, const int64_t* __restrict__ dst_row, float* __restrict__ inv_out) {
  extern __shared__ uint16_t strip[];            // [32 columns][pitch]: column-major, pitch = D_even + 2 halves
  __shared__ float red[16][32];
  __shared__ float inv_s[32];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;             // 0..15
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * 32;
  const bool col_ok = n0 + lane < N;
  const int pitch = static_cast<int>((D + 1) / 2 * 2 + 2);      // pitch / 2 is odd: lane * pitch / 2 hits 32 banks
  const uint16_t* s16 = reinterpret_cast<const uint16_t*>(src) + n0 + lane;
  uint16_t* mycol = strip + lane * pitch;
  float ss = 0.f;
  // phase A: warp w reads rows d = w, w + 16, ...: 64 contiguous bytes per warp-row, 16 loads in flight per lane
  constexpr int kU = 16;                         // loads in flight per lane (32 measured the same: not latency bound)
  for (int64_t d0 = warp; d0 < D; d0 += 16 * kU) {
    uint16_t raw[kU];
    const uint16_t* p = s16 + d0 * stride_d;
#pragma unroll
    for (int j = 0; j < kU; ++j) {
      raw[j] = (col_ok && d0 + 16 * j < D) ? __ldg(p) : static_cast<uint16_t>(0);
      p += 16 * stride_d;
    }
#pragma unroll
    for (int j = 0; j < kU; ++j) {
      const int64_t d = d0 + 16 * j;
      if (d < D) {
        mycol[d] = raw[j];
        const float v = sc::to_f32<T>(*reinterpret_cast<const T*>(&raw[j]));
        ss = fmaf(v, v, ss);
      }
    }
  }
  if ((D & 1) && warp == 0) mycol[D] = 0;        // odd D: the pair partner of the last element
  red[warp][lane] = ss;
  __syncthreads();
  if (warp == 0) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) t += red[j][lane];
    const float inv = 1.0f / sqrtf(t);
    inv_s[lane] = normalize ? inv : 1.0f;
    if (inv_out != nullptr && col_ok) inv_out[n0 + lane] = inv;
  }
  __syncthreads();
  // phase B: warp w writes the output rows of columns w and w + 16; lane handles the pair d = d0 + 2 lane, +1
  // (one 4-byte shared load, one 4-byte global store: 128 contiguous bytes per warp)
  for (int c = warp; c < 32; c += 16) {
    int64_t o = n0 + c;
    if (o >= N) continue;
    if (dst_row != nullptr) o = dst_row[o];      // scattered output (the label-sorted bank); < 0 = dropped key
    if (o < 0) continue;
    const float inv = inv_s[c];
    const uint32_t* col32 = reinterpret_cast<const uint32_t*>(strip + c * pitch);
    TO* orow = dst + o * D_pad;
    const int64_t pairs = (D + 1) / 2;
    for (int64_t pr = lane; pr < D_pad / 2; pr += 32) {
      float a = 0.f, b = 0.f;
      if (pr < pairs) {
        const uint32_t w = col32[pr];
        const uint16_t lo16 = static_cast<uint16_t>(w & 0xffffu), hi16 = static_cast<uint16_t>(w >> 16);
        a = sc::to_f32<T>(*reinterpret_cast<const T*>(&lo16)) * inv;
        b = sc::to_f32<T>(*reinterpret_cast<const T*>(&hi16)) * inv;
      }
      sc::store2<TO>(orow + 2 * pr, a, b);
    }
  }
}

template <typename T, typename TO>
int launch_norm_strip(const void* src, int64_t D, int64_t N, int64_t stride_d, void* dst, int64_t D_pad, int normalize,
                      const int64_t* dst_row, float* inv_out, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(32) * ((D + 1) / 2 * 2 + 2) * 2;
  SC_CUDA(cudaFuncSetAttribute(norm_strip_kernel<T, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  norm_strip_kernel<T, TO><<<static_cast<unsigned>(sc::ceil_div(N, 32)), 512, smem, st>>>(
      static_cast<const T*>(src), D, N, stride_d, static_cast<TO*>(dst), D_pad, normalize, dst_row, inv_out);
  return SC_OK;
}

// ---- source rows are contiguous along d (stride_d == 1): one warp per output row.
template <typename T, typename TO>
__global__ void __launch_bounds__(256)
norm_rows_kernel(const T* __restrict__ src, int64_t D, int64_t N, int64_t stride_n,
                 const int64_t* __restrict__ idx, int64_t n_out, TO* __restrict__ dst,
                 int64_t D_pad, int normalize, const int64_t* __restrict__ dst_row, float* __restrict__ inv_out) {
  const int lane = threadIdx.x & 31;
  int64_t o = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (o >= n_out) return;
  const int64_t o_in = o;
  const int64_t n = idx ? idx[o] : o;
  if (dst_row != nullptr) o = dst_row[o];
  if (o < 0) return;
  const bool ok = (n >= 0 && n < N);
  const T* row = src + (ok ? n : 0) * stride_n;
  float nrm = 1.f;
  if (normalize || inv_out != nullptr) {
    float ss = 0.f;
    if (ok)
      for (int64_t d = lane; d < D; d += 32) {
        const float v = sc::to_f32<T>(row[d]);
        ss = fmaf(v, v, ss);
      }
    const float full = sqrtf(sc::warp_sum(ss));
    if (inv_out != nullptr && lane == 0) inv_out[o_in] = 1.0f / full;
    if (normalize) nrm = full;
  }
  for (int64_t d = 2 * lane; d < D_pad; d += 64) {
    float a = 0.f, b = 0.f;
    if (ok && d < D) a = sc::to_f32<T>(row[d]) / nrm;
    if (ok && d + 1 < D) b = sc::to_f32<T>(row[d + 1]) / nrm;
    sc::store2<TO>(dst + o * D_pad + d, a, b);
  }
}

// ---- split variants: every (optionally normalised) fp32 value v is written as an fp16 PAIR hi = fp16(v),
// lo = fp16(v - hi), 22 significant bits in all, so that a tensor-core GEMM over (hi, lo) operands
// (sc_gemm_split_nt) reproduces the fp32 product.  Any layout / dtype; the banks these run on are small
// (queries, text classifier), so they keep the simple tile / warp-per-row structure.
__device__ __forceinline__ void split_f16(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

template <typename T>
__global__ void __launch_bounds__(256)
split_transpose_kernel(const T* __restrict__ src, int64_t D, int64_t N, int64_t stride_d, int64_t stride_n,
                       __half* __restrict__ hi, __half* __restrict__ lo, int64_t D_pad, int normalize) {
  __shared__ float tile[64][33];
  __shared__ float red[8][32];
  __shared__ float nrm_s[32];
  const int tx = threadIdx.x & 31;
  const int ty = threadIdx.x >> 5;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int64_t n = n0 + tx;
  const bool ok = n < N;
  float ss = 0.f;
  if (normalize && ok)
    for (int64_t d = ty; d < D; d += 8) {
      const float v = sc::to_f32<T>(src[d * stride_d + n * stride_n]);
      ss = fmaf(v, v, ss);
    }
  red[ty][tx] = ss;
  __syncthreads();
  if (ty == 0) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][tx];
    nrm_s[tx] = normalize ? sqrtf(t) : 1.0f;
  }
  __syncthreads();
  const float nrm = nrm_s[tx];
  for (int64_t d0 = 0; d0 < D_pad; d0 += 64) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t d = d0 + ty + 8 * j;
      float v = 0.f;
      if (d < D && ok) {
        v = sc::to_f32<T>(src[d * stride_d + n * stride_n]);
        if (normalize) v = v / nrm;
      }
      tile[ty + 8 * j][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = ty + 8 * j;
      const int64_t o = n0 + r;
      if (o < N) {
        const int d = 2 * tx;
        __half h0, l0, h1, l1;
        split_f16(tile[d][r], h0, l0);
        split_f16(tile[d + 1][r], h1, l1);
        *reinterpret_cast<__half2*>(hi + o * D_pad + d0 + d) = __halves2half2(h0, h1);
        *reinterpret_cast<__half2*>(lo + o * D_pad + d0 + d) = __halves2half2(l0, l1);
      }
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
split_rows_kernel(const T* __restrict__ src, int64_t D, int64_t N, int64_t stride_n, __half* __restrict__ hi,
                  __half* __restrict__ lo, int64_t D_pad, int normalize) {
  const int lane = threadIdx.x & 31;
  const int64_t o = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (o >= N) return;
  const T* row = src + o * stride_n;
  float nrm = 1.f;
  if (normalize) {
    float ss = 0.f;
    for (int64_t d = lane; d < D; d += 32) {
      const float v = sc::to_f32<T>(row[d]);
      ss = fmaf(v, v, ss);
    }
    nrm = sqrtf(sc::warp_sum(ss));
  }
  for (int64_t d = 2 * lane; d < D_pad; d += 64) {
    float a = 0.f, b = 0.f;
    if (d < D) a = sc::to_f32<T>(row[d]) / nrm;
    if (d + 1 < D) b = sc::to_f32<T>(row[d + 1]) / nrm;
    __half h0, l0, h1, l1;
    split_f16(a, h0, l0);
    split_f16(b, h1, l1);
    *reinterpret_cast<__half2*>(hi + o * D_pad + d) = __halves2half2(h0, h1);
    *reinterpret_cast<__half2*>(lo + o * D_pad + d) = __halves2half2(l0, l1);
  }
}

// ---- Tip-Adapter cache keys (tip_adapter/utils.py:59-60): mean over the augment epochs, then row L2-normalise,
// in the storage type's own rounding steps (the reference runs `torch.cat(...).mean(dim=0)` and
// `cache_keys /= cache_keys.norm(dim=-1, keepdim=True)` on fp16 tensors: the mean, the norm and the quotient are
// each rounded to fp16; sums are fp32 as in torch's reductions).  Warp per key row, the row read once per epoch.
template <typename T>
__global__ void __launch_bounds__(256)
mean_normalize_rows_kernel(const T* __restrict__ src, int64_t E, int64_t N, int64_t D, int64_t stride_e, int64_t stride_n,
                           T* __restrict__ dst, int64_t ld_dst) {
  const int64_t n = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const float inv_e = 1.0f / static_cast<float>(E);
  float ss = 0.f;
  for (int64_t d = lane; d < D; d += 32) {
    float acc = 0.f;
    for (int64_t e = 0; e < E; ++e) acc += sc::to_f32<T>(src[e * stride_e + n * stride_n + d]);
    const T mean = sc::from_f32<T>(E == 1 ? acc : acc * inv_e);       // rounded like the reference's stored mean
    dst[n * ld_dst + d] = mean;
    const float mf = sc::to_f32<T>(mean);
    ss += mf * mf;
  }
  ss = sc::warp_sum(ss);
  const float norm = sc::to_f32<T>(sc::from_f32<T>(sqrtf(ss)));        // .norm() returns the storage type
  __syncwarp();
  for (int64_t d = lane; d < D; d += 32)
    dst[n * ld_dst + d] = sc::from_f32<T>(sc::to_f32<T>(dst[n * ld_dst + d]) / norm);
}

}  // namespace

extern "C" int sc_normalize_split(const void* src, int src_dtype, int64_t D, int64_t N, int64_t stride_d,
                                  int64_t stride_n, void* hi, void* lo, int64_t D_pad, int normalize,
                                  void* stream) {
  SC_REQUIRE(src && hi && lo, SC_EINVAL, "sc_normalize_split: null pointer");
  SC_REQUIRE(D > 0 && N >= 0, SC_ESHAPE, "sc_normalize_split: bad shape");
  SC_REQUIRE(D_pad >= D && D_pad % 64 == 0, SC_ESHAPE, "sc_normalize_split: D_pad=%lld must be >= D and a multiple of 64",
             (long long)D_pad);
  SC_REQUIRE((reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) % 4 == 0, SC_EALIGN,
             "sc_normalize_split: outputs misaligned");
  if (N == 0) return SC_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (stride_d == 1 && stride_n != 1) {
    const unsigned blocks = static_cast<unsigned>(sc::ceil_div(N, 8));
    SC_DISPATCH_DTYPE(src_dtype, T,
                      (split_rows_kernel<T><<<blocks, 256, 0, st>>>(static_cast<const T*>(src), D, N, stride_n,
                                                                    static_cast<__half*>(hi), static_cast<__half*>(lo),
                                                                    D_pad, normalize)));
  } else {
    const unsigned blocks = static_cast<unsigned>(sc::ceil_div(N, 32));
    SC_DISPATCH_DTYPE(src_dtype, T,
                      (split_transpose_kernel<T><<<blocks, 256, 0, st>>>(static_cast<const T*>(src), D, N, stride_d,
                                                                         stride_n, static_cast<__half*>(hi),
                                                                         static_cast<__half*>(lo), D_pad, normalize)));
  }
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

static int normalize_impl(const void* src, int src_dtype, int64_t D, int64_t N, int64_t stride_d, int64_t stride_n,
                          const int64_t* idx, int64_t n_out, void* dst, int dst_dtype, int64_t D_pad, int normalize,
                          const int64_t* dst_row, float* inv_out, void* stream) {
  SC_REQUIRE(src && dst, SC_EINVAL, "sc_normalize_cast: null pointer");
  SC_REQUIRE(D > 0 && N >= 0 && n_out >= 0, SC_ESHAPE, "sc_normalize_cast: bad shape");
  SC_REQUIRE(D_pad >= D && D_pad % 64 == 0, SC_ESHAPE,
             "sc_normalize_cast: D_pad=%lld must be >= D and a multiple of 64", (long long)D_pad);
  SC_REQUIRE(dst_dtype != SC_E4M3 || D_pad % 128 == 0, SC_ESHAPE,
             "sc_normalize_cast: D_pad=%lld must be a multiple of 128 for SC_E4M3 rows", (long long)D_pad);
  SC_REQUIRE(idx != nullptr || n_out == N, SC_ESHAPE, "sc_normalize_cast: n_out must equal N without idx");
  SC_REQUIRE(reinterpret_cast<uintptr_t>(dst) % 4 == 0, SC_EALIGN, "sc_normalize_cast: dst misaligned");
  if (n_out == 0) return SC_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (idx == nullptr && stride_n == 1 && stride_d != 1 && (src_dtype == SC_F16 || src_dtype == SC_BF16) &&
      static_cast<size_t>(D) * 33 * 2 <= 200 * 1024) {
    int rc = SC_OK;
    SC_DISPATCH_OP8(dst_dtype, TO, {
      if (src_dtype == SC_F16) rc = launch_norm_strip<__half, TO>(src, D, N, stride_d, dst, D_pad, normalize, dst_row, inv_out, st);
      else rc = launch_norm_strip<__nv_bfloat16, TO>(src, D, N, stride_d, dst, D_pad, normalize, dst_row, inv_out, st);
    });
    if (rc != SC_OK) return rc;
  } else if (stride_d == 1 && stride_n != 1) {
    const unsigned blocks = static_cast<unsigned>(sc::ceil_div(n_out, 8));
    SC_DISPATCH_OP8(dst_dtype, TO, {
      SC_DISPATCH_DTYPE(src_dtype, T,
                        (norm_rows_kernel<T, TO><<<blocks, 256, 0, st>>>(
                            static_cast<const T*>(src), D, N, stride_n, idx, n_out,
                            static_cast<TO*>(dst), D_pad, normalize, dst_row, inv_out)));
    });
  } else {
    const unsigned blocks = static_cast<unsigned>(sc::ceil_div(n_out, 32));
    SC_DISPATCH_OP8(dst_dtype, TO, {
      SC_DISPATCH_DTYPE(src_dtype, T,
                        (norm_transpose_kernel<T, TO><<<blocks, 256, 0, st>>>(
                            static_cast<const T*>(src), D, N, stride_d, stride_n, idx, n_out,
                            static_cast<TO*>(dst), D_pad, normalize, dst_row, inv_out)));
    });
  }
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

extern "C" int sc_normalize_cast(const void* src, int src_dtype, int64_t D, int64_t N,
                                 int64_t stride_d, int64_t stride_n, const int64_t* idx,
                                 int64_t n_out, void* dst, int dst_dtype, int64_t D_pad,
                                 int normalize, void* stream) {
  return normalize_impl(src, src_dtype, D, N, stride_d, stride_n, idx, n_out, dst, dst_dtype, D_pad, normalize, nullptr,
                        nullptr, stream);
}

extern "C" int sc_transpose_norms(const void* src, int src_dtype, int64_t D, int64_t N, int64_t stride_d,
                                  int64_t stride_n, const int64_t* idx, int64_t n_out, void* dst, int dst_dtype,
                                  int64_t D_pad, int normalize, float* inv_norm, void* stream) {
  SC_REQUIRE(inv_norm != nullptr, SC_EINVAL, "sc_transpose_norms: null inv_norm");
  return normalize_impl(src, src_dtype, D, N, stride_d, stride_n, idx, n_out, dst, dst_dtype, D_pad, normalize, nullptr,
                        inv_norm, stream);
}

extern "C" int sc_normalize_scatter(const void* src, int src_dtype, int64_t D, int64_t N, int64_t stride_d,
                                    int64_t stride_n, const int64_t* idx, int64_t n_out, const int64_t* dst_row,
                                    void* dst, int dst_dtype, int64_t D_pad, int normalize, void* stream) {
  SC_REQUIRE(dst_row != nullptr, SC_EINVAL, "sc_normalize_scatter: null dst_row");
  return normalize_impl(src, src_dtype, D, N, stride_d, stride_n, idx, n_out, dst, dst_dtype, D_pad, normalize, dst_row,
                        nullptr, stream);
}

extern "C" int sc_mean_normalize_rows(const void* src, int dtype, int64_t E, int64_t N, int64_t D, int64_t stride_e,
                                      int64_t stride_n, void* dst, int64_t ld_dst, void* stream) {
  SC_REQUIRE(src && dst, SC_EINVAL, "sc_mean_normalize_rows: null pointer");
  SC_REQUIRE(E >= 1 && N >= 0 && D >= 1 && ld_dst >= D, SC_ESHAPE, "sc_mean_normalize_rows: bad shape");
  if (N == 0) return SC_OK;
  const unsigned blocks = static_cast<unsigned>(sc::ceil_div(N, 8));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_DISPATCH_DTYPE(dtype, T, (mean_normalize_rows_kernel<T><<<blocks, 256, 0, st>>>(
                                  static_cast<const T*>(src), E, N, D, stride_e, stride_n, static_cast<T*>(dst), ld_dst)));
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}
