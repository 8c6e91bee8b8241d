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
  SC_CUDA(cudaMemsetAsync(Vt, 0, static_cast<size_t>(C_pad) * Nk_pad * 2, st));
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
