// Shared host/device helpers for libsummerclip_b200.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/summer_clip_b200.h"

namespace sc {

// thread-local last error (sc_api.cu)
void set_error(const char* fmt, ...);

#define SC_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      ::sc::set_error(__VA_ARGS__);  \
      return (code);                 \
    }                                \
  } while (0)

#define SC_CUDA(expr)                                                                    \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::sc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                      __LINE__);                                                         \
      return static_cast<int>(_e);                                                       \
    }                                                                                    \
  } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// two fp32 -> one 32-bit word holding two 16-bit operands (lo in the low half)
template <typename T>
__device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// store two adjacent operand elements (p 4-byte aligned for the 16-bit types, 2-byte for e4m3; e4m3 stores
// SC_E4M3_SCALE * x, round-to-nearest-even, saturating)
template <typename T>
__device__ __forceinline__ void store2(T* p, float lo, float hi) {
  *reinterpret_cast<uint32_t*>(p) = pack2<T>(lo, hi);
}
template <>
__device__ __forceinline__ void store2<__nv_fp8_e4m3>(__nv_fp8_e4m3* p, float lo, float hi) {
  *reinterpret_cast<__nv_fp8x2_storage_t*>(p) =
      __nv_cvt_float2_to_fp8x2(make_float2(lo * SC_E4M3_SCALE, hi * SC_E4M3_SCALE), __NV_SATFINITE, __NV_E4M3);
}

// operand types a normalised bank can be written in (16-bit, or e4m3 for sc_attn_fwd_hard)
#define SC_DISPATCH_OP8(dtype, T, ...)                                         \
  switch (dtype) {                                                             \
    case SC_F16: { using T = __half; __VA_ARGS__; break; }                     \
    case SC_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }             \
    case SC_E4M3: { using T = __nv_fp8_e4m3; __VA_ARGS__; break; }             \
    default:                                                                   \
      ::sc::set_error("operand dtype %d must be SC_F16, SC_BF16 or SC_E4M3", (int)(dtype)); \
      return SC_EINVAL;                                                        \
  }

// dispatch on the tensor-core operand type (fp16 or bf16)
#define SC_DISPATCH_OP(dtype, T, ...)                                          \
  switch (dtype) {                                                             \
    case SC_F16: { using T = __half; __VA_ARGS__; break; }                     \
    case SC_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }             \
    default:                                                                   \
      ::sc::set_error("operand dtype %d must be SC_F16 or SC_BF16", (int)(dtype)); \
      return SC_EINVAL;                                                        \
  }

// dispatch on the ABI dtype enum
#define SC_DISPATCH_DTYPE(dtype, T, ...)                                 \
  switch (dtype) {                                                       \
    case SC_F16: { using T = __half; __VA_ARGS__; break; }               \
    case SC_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }       \
    case SC_F32: { using T = float; __VA_ARGS__; break; }                \
    default:                                                             \
      ::sc::set_error("unsupported dtype %d", (int)(dtype));             \
      return SC_EINVAL;                                                  \
  }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// order-preserving float -> uint32 map (larger float => larger key); NaN sorts above +inf,
// matching torch.topk which treats NaN as the largest value.
__device__ __forceinline__ uint32_t float_order_key(float f) {
  if (f != f) return 0xFFFFFFFFu;
  if (f == 0.0f) return 0x80000000u;   // -0.0 and +0.0 compare equal in torch
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

}  // namespace sc
