// Warp-per-row scans of a logits bank shared by sc_select.cu and sc_values.cu.
#pragma once
#include "sc_common.cuh"

namespace sc {

// value/index pair ordered like torch.max(dim=1): larger value wins, NaN is largest, and among
// equal values the smaller index wins.
struct MaxIdx {
  float v;
  int i;
};
__device__ __forceinline__ bool better(float v, int i, const MaxIdx& b) {
  if (b.i < 0) return true;
  const bool vn = (v != v), bn = (b.v != b.v);
  if (vn || bn) return vn && (!bn || i < b.i);
  return v > b.v || (v == b.v && i < b.i);
}
__device__ __forceinline__ MaxIdx warp_argmax(MaxIdx m) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MaxIdx other;
    other.v = __shfl_xor_sync(0xffffffffu, m.v, o);
    other.i = __shfl_xor_sync(0xffffffffu, m.i, o);
    if (other.i >= 0 && better(other.v, other.i, m)) m = other;
  }
  return m;
}

template <typename T>
struct Vec {  // 16-byte vector of T
  static constexpr int kN = 16 / sizeof(T);
  T e[kN];
};

// One warp per row.  kVec: 16-byte loads (requires 16-B aligned base and ld % kN == 0).
template <typename T, bool kVec>
__device__ __forceinline__ MaxIdx row_argmax(const T* __restrict__ row, int64_t C, int lane) {
  MaxIdx m{0.f, -1};
  if (kVec) {
    constexpr int kN = Vec<T>::kN;
    const int64_t nv = C / kN;
    const uint4* vrow = reinterpret_cast<const uint4*>(row);
    for (int64_t j = lane; j < nv; j += 32) {
      uint4 raw = __ldg(vrow + j);
      const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
      for (int t = 0; t < kN; ++t) {
        const float v = to_f32<T>(e[t]);
        const int c = static_cast<int>(j * kN + t);
        if (better(v, c, m)) { m.v = v; m.i = c; }
      }
    }
    for (int64_t c = nv * kN + lane; c < C; c += 32) {
      const float v = to_f32<T>(row[c]);
      if (better(v, static_cast<int>(c), m)) { m.v = v; m.i = static_cast<int>(c); }
    }
  } else {
    for (int64_t c = lane; c < C; c += 32) {
      const float v = to_f32<T>(row[c]);
      if (better(v, static_cast<int>(c), m)) { m.v = v; m.i = static_cast<int>(c); }
    }
  }
  return warp_argmax(m);
}

// sum_c exp(scale*l_c - scale*l_max) over the row (second read is L1/L2 resident)
template <typename T, bool kVec>
__device__ __forceinline__ float row_expsum(const T* __restrict__ row, int64_t C, int lane,
                                            float scale, float tmax) {
  float s = 0.f;
  if (kVec) {
    constexpr int kN = Vec<T>::kN;
    const int64_t nv = C / kN;
    const uint4* vrow = reinterpret_cast<const uint4*>(row);
    for (int64_t j = lane; j < nv; j += 32) {
      uint4 raw = __ldg(vrow + j);
      const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
      for (int t = 0; t < kN; ++t) s += expf(__fmul_rn(to_f32<T>(e[t]), scale) - tmax);
    }
    for (int64_t c = nv * kN + lane; c < C; c += 32)
      s += expf(__fmul_rn(to_f32<T>(row[c]), scale) - tmax);
  } else {
    for (int64_t c = lane; c < C; c += 32)
      s += expf(__fmul_rn(to_f32<T>(row[c]), scale) - tmax);
  }
  return warp_sum(s);
}

}  // namespace sc
