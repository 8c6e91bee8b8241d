// Warp-per-row scans of a logits bank shared by sc_select.cu and sc_values.cu.
#pragma once
#include "sc_common.cuh"

#include <type_traits>

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

// max that PROPAGATES NaN (torch.max treats NaN as the largest value): one instruction per element, and a
// NaN result sends the caller to the exact slow path
__device__ __forceinline__ float max_nan(float a, float b) {
  float d;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}

// exp(t) for t <= 0 as one multiply and one MUFU: ex2.approx(t * log2(e)).  Against expf this differs by ~1e-7
// relative on the terms that matter (|t| < 1), far inside the 2e-6 the selection goldens allow, and costs 2
// instructions instead of ~8 — the PROB row scan was bound by them, not by HBM.
__device__ __forceinline__ float exp_neg_fast(float t) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(t * 1.4426950408889634f));
  return y;
}

// packed max of two 16-bit pairs that PROPAGATES NaN (one instruction per two elements)
template <typename T>
__device__ __forceinline__ uint32_t max_nan_x2(uint32_t a, uint32_t b);
template <>
__device__ __forceinline__ uint32_t max_nan_x2<__half>(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("max.NaN.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
template <>
__device__ __forceinline__ uint32_t max_nan_x2<__nv_bfloat16>(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("max.NaN.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
template <>
__device__ __forceinline__ uint32_t max_nan_x2<float>(uint32_t a, uint32_t) { return a; }   // never used

// Register-resident row of at most 32 * NV 16-byte vectors (one warp per row).  load(): every lane issues all
// of its 16-byte loads before the first use; vectors past the end of the row are filled with -inf, which is the
// identity of every pass below (never the maximum unless the whole row is -inf, never equal to a finite
// maximum, exp(scale * -inf - tmax) = 0 for scale > 0), so the passes run without per-vector predicates.
// argmax(): max.NaN for the value (packed pairs for 16-bit storage: 0.5 instructions per element), then "last hit
// in reverse order" for the FIRST index at which it occurs (3 per element) — instead of the ~20-instruction
// ordered compare per element of sc::better(); rows that contain a NaN (the max comes back NaN) take the exact
// path.  Semantics are those of row_argmax / torch.max(dim=1): larger value wins, NaN is largest, equal values
// -> smaller index.  argmax_expsum(): the index pass and sum_c exp(scale l_c - scale l_max) in ONE pass over the
// registers (one conversion per element).
template <typename T, int NV>
struct RegRow {
  static constexpr int kN = 16 / sizeof(T);
  uint4 v[NV];   // the raw row: kept in storage type (registers decide the occupancy that hides the HBM latency)
  int nv;        // 16-byte vectors in the row
  int lane;

  static __device__ __forceinline__ uint32_t ninf_word() {
    return sizeof(T) == 4 ? 0xff800000u : (std::is_same<T, __half>::value ? 0xfc00fc00u : 0xff80ff80u);
  }
  __device__ __forceinline__ void load(const T* __restrict__ row, int nv_, int lane_) {
    nv = nv_;
    lane = lane_;
    const uint4* vrow = reinterpret_cast<const uint4*>(row);
    const uint32_t ni = ninf_word();
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const int j = lane + 32 * u;
      v[u] = (j < nv) ? __ldg(vrow + j) : make_uint4(ni, ni, ni, ni);
    }
  }
  __device__ __forceinline__ int col(int u, int t) const { return (lane + 32 * u) * kN + t; }
  __device__ __forceinline__ float x(int u, int t) const { return to_f32<T>(reinterpret_cast<const T*>(&v[u])[t]); }

  // row maximum, NaN-propagating, reduced over the warp
  __device__ __forceinline__ float max_value() const {
    float mx = __int_as_float(0xff800000);   // -inf
    if constexpr (sizeof(T) == 2) {
      uint32_t m2 = ninf_word();             // the maximum is exact in the storage type: take it pairwise there
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        m2 = max_nan_x2<T>(m2, v[u].x);
        m2 = max_nan_x2<T>(m2, v[u].y);
        m2 = max_nan_x2<T>(m2, v[u].z);
        m2 = max_nan_x2<T>(m2, v[u].w);
      }
      const T* pair = reinterpret_cast<const T*>(&m2);
      mx = max_nan(to_f32<T>(pair[0]), to_f32<T>(pair[1]));
    } else {
#pragma unroll
      for (int u = 0; u < NV; ++u) {
#pragma unroll
        for (int t = 0; t < kN; ++t) mx = max_nan(mx, x(u, t));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max_nan(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    return mx;
  }
  // exact ordered compare for rows that contain a NaN (warp-uniform slow path)
  __device__ __forceinline__ MaxIdx argmax_exact() const {
    MaxIdx m{0.f, -1};
#pragma unroll
    for (int u = 0; u < NV; ++u)
      if (lane + 32 * u < nv) {
#pragma unroll
        for (int t = 0; t < kN; ++t)
          if (better(x(u, t), col(u, t), m)) { m.v = x(u, t); m.i = col(u, t); }
      }
    return warp_argmax(m);
  }
  __device__ __forceinline__ MaxIdx argmax() const {
    const float mx = max_value();
    if (mx != mx) return argmax_exact();
    int idx = 0x7fffffff;
#pragma unroll
    for (int u = NV - 1; u >= 0; --u) {
#pragma unroll
      for (int t = kN - 1; t >= 0; --t) idx = (x(u, t) == mx) ? col(u, t) : idx;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) idx = min(idx, __shfl_xor_sync(0xffffffffu, idx, o));
    return MaxIdx{mx, idx};
  }
  // argmax() plus sum_c exp(scale * l_c - scale * l_max) (scale > 0) in one pass over the registers.  The sum
  // runs in reverse vector order within a lane, then the xor tree.
  __device__ __forceinline__ MaxIdx argmax_expsum(float scale, float& sum) const {
    const float mx = max_value();
    if (mx != mx) {
      sum = mx;
      return argmax_exact();
    }
    const float tmax = __fmul_rn(mx, scale);
    int idx = 0x7fffffff;
    float s = 0.f;
#pragma unroll
    for (int u = NV - 1; u >= 0; --u) {
#pragma unroll
      for (int t = kN - 1; t >= 0; --t) {
        const float xv = x(u, t);
        idx = (xv == mx) ? col(u, t) : idx;
        s += exp_neg_fast(__fmul_rn(xv, scale) - tmax);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) idx = min(idx, __shfl_xor_sync(0xffffffffu, idx, o));
    sum = warp_sum(s);
    return MaxIdx{mx, idx};
  }
  // sum_c exp(scale * l_c - tmax), scale > 0
  __device__ __forceinline__ float expsum(float scale, float tmax) const {
    float s = 0.f;
#pragma unroll
    for (int u = NV - 1; u >= 0; --u) {
#pragma unroll
      for (int t = kN - 1; t >= 0; --t) s += exp_neg_fast(__fmul_rn(x(u, t), scale) - tmax);
    }
    return warp_sum(s);
  }
};

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
      for (int t = 0; t < kN; ++t) s += exp_neg_fast(__fmul_rn(to_f32<T>(e[t]), scale) - tmax);
    }
    for (int64_t c = nv * kN + lane; c < C; c += 32)
      s += exp_neg_fast(__fmul_rn(to_f32<T>(row[c]), scale) - tmax);
  } else {
    for (int64_t c = lane; c < C; c += 32)
      s += exp_neg_fast(__fmul_rn(to_f32<T>(row[c]), scale) - tmax);
  }
  return warp_sum(s);
}

}  // namespace sc
