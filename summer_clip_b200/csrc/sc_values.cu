// sc_values_prepare: cache values V = f(L[idx]) written TRANSPOSED (Vt[C_pad, Nk_pad], bf16) so
// that the attention kernel's GEMM-2 B operand is K-major like every other operand.
//   HARD    : one_hot(argmax_c L).half()            cache_value_strategy.py:15-16 (0/1 exact in bf16)
//   SOFTMAX : softmax(clip_scale*scale*L, dim=1)    cache_value_strategy.py:27
//   labels_override : one_hot(gold labels)          image_attention.py:65-66, tip_adapter/utils.py:62
#include "sc_common.cuh"
#include "sc_rowops.cuh"

namespace {

// HARD: Vt is pre-zeroed; one warp per cache row scatters a single 1.0.
template <typename T, typename TO>
__global__ void __launch_bounds__(256)
values_hard_kernel(const T* __restrict__ L, int64_t N, int64_t C, int64_t ld,
                   const int64_t* __restrict__ idx, const int32_t* __restrict__ labels_override,
                   int64_t n_out, TO* __restrict__ Vt, int64_t Nk_pad) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_per_grid = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t o = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
       o < n_out; o += warps_per_grid) {
    int lab;
    if (labels_override) {
      lab = labels_override[o];
    } else {
      const int64_t r = idx ? idx[o] : o;
      if (r < 0 || r >= N) continue;
      lab = sc::row_argmax<T, false>(L + r * ld, C, lane).i;
    }
    if (lane == 0 && lab >= 0 && lab < C) Vt[static_cast<int64_t>(lab) * Nk_pad + o] = sc::from_f32<TO>(1.0f);
  }
}

// SOFTMAX: block = 32 cache rows.  Each warp computes 4 row softmaxes into a [C][32] bf16 smem
// tile; the tile is then written out class by class (64 contiguous bytes per class row).
template <typename T, typename TO>
__global__ void __launch_bounds__(256)
values_softmax_kernel(const T* __restrict__ L, int64_t N, int64_t C, int64_t ld,
                      const int64_t* __restrict__ idx, int64_t n_out, float scale,
                      TO* __restrict__ Vt, int64_t Nk_pad) {
  extern __shared__ uint16_t tile_raw[];    // [C][32 + 2] (pad: conflict-free column writes)
  TO* tile = reinterpret_cast<TO*>(tile_raw);
  constexpr int kLd = 34;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t o0 = static_cast<int64_t>(blockIdx.x) * 32;
  for (int j = 0; j < 4; ++j) {
    const int col = warp * 4 + j;
    const int64_t o = o0 + col;
    int64_t r = -1;
    if (o < n_out) r = idx ? idx[o] : o;
    if (r < 0 || r >= N) {
      for (int64_t c = lane; c < C; c += 32) tile[c * kLd + col] = sc::from_f32<TO>(0.f);
      continue;
    }
    const T* row = L + r * ld;
    const sc::MaxIdx m = sc::row_argmax<T, false>(row, C, lane);
    const float tmax = __fmul_rn(m.v, scale);
    const float s = sc::row_expsum<T, false>(row, C, lane, scale, tmax);
    const float inv = 1.0f / s;
    for (int64_t c = lane; c < C; c += 32) {
      const float e = expf(__fmul_rn(sc::to_f32<T>(row[c]), scale) - tmax);
      tile[c * kLd + col] = sc::from_f32<TO>(e * inv);
    }
  }
  __syncthreads();
  const int64_t ncol = (n_out - o0) < 32 ? (n_out - o0) : 32;
  for (int64_t e = threadIdx.x; e < C * 32; e += blockDim.x) {
    const int64_t c = e >> 5;
    const int col = static_cast<int>(e & 31);
    if (col < ncol) Vt[c * Nk_pad + o0 + col] = tile[c * kLd + col];
  }
}

// SOFTMAX, register-resident rows (C <= 1024 in whole 16-byte vectors, scale > 0): block = 32 cache rows, warp w
// scans rows 4w .. 4w+3 from registers (one HBM read of the row: max, exp-sum and the normalised values all come
// from the same registers) and parks the results ROW-major in a [32][pitch] shared tile (pitch / 2 odd: the
// transposed 2-byte reads below are conflict free); the tile then leaves class by class as 64-byte runs, one packed
// pair of adjacent cache rows per thread.  The generic kernel above re-reads every row three times with 2-byte loads
// (7.6 ms for 1.28 M x 1000 fp16; this one is bound by the 5.1 GB it has to move).
template <typename T, typename TO, int NV>
__global__ void __launch_bounds__(256)
values_softmax_reg_kernel(const T* __restrict__ L, int64_t N, int64_t C, int64_t ld, const int64_t* __restrict__ idx,
                          int64_t n_out, float scale, TO* __restrict__ Vt, int64_t Nk_pad, int pitch) {
  extern __shared__ uint16_t tile_raw[];    // [32 cache rows][pitch] of TO
  constexpr int kN = 16 / sizeof(T);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nv = static_cast<int>(C / kN);
  const int64_t o0 = static_cast<int64_t>(blockIdx.x) * 32;
#pragma unroll 1
  for (int j0 = 0; j0 < 4; j0 += 2) {
    sc::RegRow<T, NV> row[2];
    bool have[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t o = o0 + warp * 4 + j0 + i;
      int64_t r = -1;
      if (o < n_out) r = idx ? idx[o] : o;
      have[i] = (r >= 0 && r < N);
      if (have[i]) row[i].load(L + r * ld, nv, lane);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      uint16_t* trow = tile_raw + (warp * 4 + j0 + i) * pitch;
      if (!have[i]) {
        for (int c = lane * 2; c < C; c += 64) *reinterpret_cast<uint32_t*>(trow + c) = 0u;
        continue;
      }
      const float mx = row[i].max_value();
      const float tmax = __fmul_rn(mx, scale);
      const float inv = 1.0f / row[i].expsum(scale, tmax);
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        const int c0 = (lane + 32 * u) * kN;
        if (c0 < C) {
#pragma unroll
          for (int t = 0; t < kN; t += 2) {
            const float e0 = sc::exp_neg_fast(__fmul_rn(row[i].x(u, t), scale) - tmax) * inv;
            const float e1 = sc::exp_neg_fast(__fmul_rn(row[i].x(u, t + 1), scale) - tmax) * inv;
            *reinterpret_cast<uint32_t*>(trow + c0 + t) = sc::pack2<TO>(e0, e1);
          }
        }
      }
    }
  }
  __syncthreads();
  // write-out: thread -> (class c, pair of adjacent cache rows): two 2-byte shared loads, one 4-byte global store;
  // 16 threads cover the 64 contiguous bytes of one class row
  const int64_t ncol = (n_out - o0) < 32 ? (n_out - o0) : 32;
  const int pr = threadIdx.x & 15;
  for (int64_t c = threadIdx.x >> 4; c < C; c += 16) {
    const uint32_t lo = tile_raw[(2 * pr) * pitch + c], hi = tile_raw[(2 * pr + 1) * pitch + c];
    TO* dst = Vt + c * Nk_pad + o0 + 2 * pr;
    if (2 * pr + 1 < ncol) *reinterpret_cast<uint32_t*>(dst) = lo | (hi << 16);
    else if (2 * pr < ncol) *reinterpret_cast<uint16_t*>(dst) = static_cast<uint16_t>(lo);
  }
}

// Hard labels as the attention kernel's synthesised-values operand: int16 argmax (or override) per selected
// key, -1 for padding keys and for labels outside [0, C).
template <typename T>
__global__ void __launch_bounds__(256)
hard_labels_kernel(const T* __restrict__ L, int64_t N, int64_t C, int64_t ld, const int64_t* __restrict__ idx,
                   const int32_t* __restrict__ labels_override, int64_t n_out, int16_t* __restrict__ out,
                   int64_t n_pad) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_per_grid = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t o = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); o < n_pad;
       o += warps_per_grid) {
    int lab = -1;
    if (o < n_out) {
      if (labels_override) {
        lab = labels_override[o];
      } else {
        const int64_t r = idx ? idx[o] : o;
        if (r >= 0 && r < N) lab = sc::row_argmax<T, false>(L + r * ld, C, lane).i;
      }
    }
    if (lane == 0) out[o] = (lab >= 0 && lab < C) ? static_cast<int16_t>(lab) : static_cast<int16_t>(-1);
  }
}

// register-resident fast path of hard_labels_kernel (rows of whole 16-byte vectors, C <= 1024)
template <typename T, int NV>
__global__ void __launch_bounds__(256, sizeof(T) == 2 ? 5 : 3)
hard_labels_reg_kernel(const T* __restrict__ L, int64_t N, int64_t C, int64_t ld, const int64_t* __restrict__ idx,
                       int64_t n_out, int16_t* __restrict__ out, int64_t n_pad) {
  constexpr int kN = 16 / sizeof(T);
  const int lane = threadIdx.x & 31;
  const int nv = static_cast<int>(C / kN);
  const int64_t warps_per_grid = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t o = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); o < n_pad;
       o += warps_per_grid) {
    int lab = -1;
    if (o < n_out) {
      const int64_t r = idx ? idx[o] : o;
      if (r >= 0 && r < N) {
        sc::RegRow<T, NV> row;
        row.load(L + r * ld, nv, lane);
        lab = row.argmax().i;
      }
    }
    if (lane == 0) out[o] = (lab >= 0 && lab < C) ? static_cast<int16_t>(lab) : static_cast<int16_t>(-1);
  }
}

template <typename TO>
__global__ void ones_row_kernel(TO* __restrict__ row, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) row[i] = sc::from_f32<TO>(1.0f);
}

}  // namespace

extern "C" int sc_values_prepare(const void* L, int dtype, int64_t N, int64_t C, int64_t ld,
                                 const int64_t* idx, const int32_t* labels_override, int64_t n_out,
                                 int mode, float scale, void* Vt, int vt_dtype, int64_t C_pad,
                                 int64_t Nk_pad, int64_t ones_row, void* stream) {
  SC_REQUIRE(Vt, SC_EINVAL, "sc_values_prepare: null Vt");
  SC_REQUIRE(L || (labels_override && mode == SC_VALUES_HARD), SC_EINVAL, "sc_values_prepare: null L");
  SC_REQUIRE(!(labels_override && mode != SC_VALUES_HARD), SC_EINVAL,
             "sc_values_prepare: labels_override replaces the argmax of SC_VALUES_HARD only (softmax values come from L[idx])");
  SC_REQUIRE(C > 0 && C_pad >= C && Nk_pad >= n_out && n_out >= 0, SC_ESHAPE, "sc_values_prepare: bad shape");
  SC_REQUIRE(idx || labels_override || n_out == N, SC_ESHAPE, "sc_values_prepare: n_out must equal N without idx");
  SC_REQUIRE(mode == SC_VALUES_HARD || mode == SC_VALUES_SOFTMAX, SC_EINVAL, "sc_values_prepare: bad mode %d", mode);
  SC_REQUIRE(ones_row < C_pad, SC_ESHAPE, "sc_values_prepare: ones_row outside Vt");
  SC_REQUIRE(L == nullptr || ld >= C, SC_ESHAPE, "sc_values_prepare: ld < C");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_REQUIRE(vt_dtype == SC_F16 || vt_dtype == SC_BF16, SC_EINVAL, "sc_values_prepare: vt_dtype must be SC_F16 or SC_BF16");
  // register-resident softmax path: rows of whole 16-byte vectors, at most 1024 classes, positive scale
  const int64_t kn = (dtype == SC_F32) ? 4 : 8;
  const bool soft_reg = mode == SC_VALUES_SOFTMAX && L != nullptr && n_out > 0 && scale > 0.f && C % kn == 0 &&
                        C / kn <= 32 * (dtype == SC_F32 ? 8 : 4) && ld % kn == 0 && reinterpret_cast<uintptr_t>(L) % 16 == 0 &&
                        reinterpret_cast<uintptr_t>(Vt) % 4 == 0 && Nk_pad % 2 == 0;
  if (soft_reg) {
    // every (class < C, key < n_out) element is written by the kernel: zero only the padding
    if (C_pad > C) SC_CUDA(cudaMemsetAsync(static_cast<char*>(Vt) + static_cast<size_t>(C) * Nk_pad * 2, 0,
                                           static_cast<size_t>(C_pad - C) * Nk_pad * 2, st));
    if (Nk_pad > n_out) SC_CUDA(cudaMemset2DAsync(static_cast<char*>(Vt) + static_cast<size_t>(n_out) * 2, static_cast<size_t>(Nk_pad) * 2,
                                                  0, static_cast<size_t>(Nk_pad - n_out) * 2, static_cast<size_t>(C), st));
  } else {
    SC_CUDA(cudaMemsetAsync(Vt, 0, static_cast<size_t>(C_pad) * Nk_pad * 2, st));
  }
  if (n_out > 0) {
    if (L == nullptr) dtype = SC_F32;
    SC_DISPATCH_OP(vt_dtype, TO, {
      TO* vt = static_cast<TO*>(Vt);
      if (mode == SC_VALUES_HARD) {
        const int64_t want = sc::ceil_div(n_out, 8);
        const unsigned blocks = static_cast<unsigned>(want < 148 * 8 ? want : 148 * 8);
        SC_DISPATCH_DTYPE(dtype, T,
                          (values_hard_kernel<T, TO><<<blocks, 256, 0, st>>>(
                              static_cast<const T*>(L), N, C, ld, idx, labels_override, n_out, vt, Nk_pad)));
      } else if (soft_reg) {
        const int pitch = static_cast<int>(((C / 2) & 1) ? C : C + 2);
        const size_t smem = static_cast<size_t>(32) * pitch * 2;
        const unsigned blocks = static_cast<unsigned>(sc::ceil_div(n_out, 32));
        if (dtype == SC_F32) {
          SC_CUDA(cudaFuncSetAttribute(values_softmax_reg_kernel<float, TO, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
          values_softmax_reg_kernel<float, TO, 8><<<blocks, 256, smem, st>>>(static_cast<const float*>(L), N, C, ld, idx, n_out, scale, vt, Nk_pad, pitch);
        } else if (dtype == SC_F16) {
          SC_CUDA(cudaFuncSetAttribute(values_softmax_reg_kernel<__half, TO, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
          values_softmax_reg_kernel<__half, TO, 4><<<blocks, 256, smem, st>>>(static_cast<const __half*>(L), N, C, ld, idx, n_out, scale, vt, Nk_pad, pitch);
        } else {
          SC_CUDA(cudaFuncSetAttribute(values_softmax_reg_kernel<__nv_bfloat16, TO, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
          values_softmax_reg_kernel<__nv_bfloat16, TO, 4><<<blocks, 256, smem, st>>>(static_cast<const __nv_bfloat16*>(L), N, C, ld, idx, n_out, scale, vt, Nk_pad, pitch);
        }
      } else {
        const size_t smem = static_cast<size_t>(C) * 34 * 2;
        SC_REQUIRE(smem <= 200 * 1024, SC_EUNSUPPORTED, "sc_values_prepare: C=%lld too large for the softmax tile", (long long)C);
        const unsigned blocks = static_cast<unsigned>(sc::ceil_div(n_out, 32));
        SC_DISPATCH_DTYPE(dtype, T, {
          SC_CUDA(cudaFuncSetAttribute(values_softmax_kernel<T, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
          values_softmax_kernel<T, TO><<<blocks, 256, smem, st>>>(static_cast<const T*>(L), N, C, ld, idx,
                                                                  n_out, scale, vt, Nk_pad);
        });
      }
      if (ones_row >= 0) {
        ones_row_kernel<TO><<<static_cast<unsigned>(sc::ceil_div(n_out, 256)), 256, 0, st>>>(
            vt + ones_row * Nk_pad, n_out);
      }
    });
  }
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

extern "C" int sc_hard_labels(const void* L, int dtype, int64_t N, int64_t C, int64_t ld, const int64_t* idx,
                              const int32_t* labels_override, int64_t n_out, int16_t* labels16, int64_t n_pad,
                              void* stream) {
  SC_REQUIRE(labels16, SC_EINVAL, "sc_hard_labels: null output");
  SC_REQUIRE(L || labels_override, SC_EINVAL, "sc_hard_labels: null L");
  SC_REQUIRE(C > 0 && C <= 32767 && n_out >= 0 && n_pad >= n_out, SC_ESHAPE, "sc_hard_labels: bad shape");
  SC_REQUIRE(idx || labels_override || n_out == N, SC_ESHAPE, "sc_hard_labels: n_out must equal N without idx");
  SC_REQUIRE(L == nullptr || ld >= C, SC_ESHAPE, "sc_hard_labels: ld < C");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_pad > 0) {
    if (L == nullptr) dtype = SC_F32;
    const int64_t want = sc::ceil_div(n_pad, 8);
    const unsigned blocks = static_cast<unsigned>(want < 148 * 8 ? want : 148 * 8);
    SC_DISPATCH_DTYPE(dtype, T, {
      constexpr int kN = 16 / sizeof(T);
      constexpr int NV = (sizeof(T) == 4) ? 8 : 4;
      const bool vec = L != nullptr && labels_override == nullptr && reinterpret_cast<uintptr_t>(L) % 16 == 0 &&
                       ld % kN == 0 && C % kN == 0 && C / kN <= 32 * NV;
      if (vec)
        hard_labels_reg_kernel<T, NV><<<blocks, 256, 0, st>>>(static_cast<const T*>(L), N, C, ld, idx, n_out,
                                                              labels16, n_pad);
      else
        hard_labels_kernel<T><<<blocks, 256, 0, st>>>(static_cast<const T*>(L), N, C, ld, idx, labels_override,
                                                      n_out, labels16, n_pad);
    });
  }
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}
