// Temperature-softmax mode of the CLIP-search attention (north-star extension; the reference's weights are the
// un-normalised exp(-beta (1 - A)) of cache_weights_strategy.py:33-36, for which the running maximum is the constant
// 0 and these kernels are not needed):
//     out[q, c] = sum_k softmax_k(tau * Qn[q].Kn[k]) V[k, c]
// The attention kernels emit PARTIAL results per key split / key shard as (m, l, O) triples,
//     O[q, c] = sum_k 2^(e(q,k) - m[q]) V[k, c],  l[q] = sum_k 2^(e(q,k) - m[q]),  e = tau * log2(e) * A,
// and the kernels below turn the one-hot kernel's per-class log-sum-exp tiles into such triples and merge triples
// (log-sum-exp merge: rescale every part to the common maximum, add, divide) — SURVEY.md §8b/§8e.
// HBM-bound, warp per query row, two passes over a row that stays in L1/L2.
#include "sc_common.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;

__global__ void fill_f32_kernel(float* __restrict__ dst, size_t n, float value) {
  const size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 4 <= n && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    *reinterpret_cast<float4*>(dst + i) = make_float4(value, value, value, value);
  } else {
    for (size_t j = i; j < n && j < i + 4; ++j) dst[j] = value;
  }
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// LSE [n_parts, Nq, ld] (log2 units, -inf = class absent) -> m[q] = max, O[q, c] = sum_p 2^(LSE - m), l = sum_c O
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
softmax_partials_kernel(const float* __restrict__ lse, int n_parts, long long part_stride, long long nq, int C, long long ld,
                        float* __restrict__ O, long long ld_out, float* __restrict__ m_out, float* __restrict__ l_out) {
  const long long q = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= nq) return;
  const float* row = lse + q * ld;
  float m = -INFINITY;
  for (int p = 0; p < n_parts; ++p)
    for (int c = lane; c < C; c += 32) m = fmaxf(m, row[p * part_stride + c]);
  m = sc::warp_max(m);
  float l = 0.f;
  for (int c = lane; c < C; c += 32) {
    float s = 0.f;
    for (int p = 0; p < n_parts; ++p) {
      const float v = row[p * part_stride + c];
      s += (v == -INFINITY) ? 0.f : ex2f(v - m);
    }
    O[q * ld_out + c] = s;
    l += s;
  }
  l = sc::warp_sum(l);
  if (lane == 0) {
    m_out[q] = m;
    l_out[q] = l;
  }
}

// (m, l, O) parts -> one triple against the common maximum M = max_p m_p (or the given m_ref), optionally normalised
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
merge_softmax_kernel(const float* O_parts, const float* __restrict__ m_parts, const float* __restrict__ l_parts,
                     int n_parts, long long o_part_stride, long long ml_part_stride, long long nq, int C, long long ld,
                     float m_scale, const float* __restrict__ m_ref, int normalize, float* out,
                     long long ld_out, float* __restrict__ m_out, float* __restrict__ l_out) {
  const long long q = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= nq) return;
  float M = -INFINITY;
  if (m_ref != nullptr) {
    M = m_ref[q];
  } else {
    for (int p = 0; p < n_parts; ++p) M = fmaxf(M, m_scale * m_parts[p * ml_part_stride + q]);
  }
  float L = 0.f;
  for (int p = 0; p < n_parts; ++p) {
    const float mp = m_scale * m_parts[p * ml_part_stride + q];
    const float w = (mp == -INFINITY) ? 0.f : ex2f(mp - M);
    L += w * l_parts[p * ml_part_stride + q];
  }
  const float inv = (normalize && L > 0.f) ? 1.0f / L : (normalize ? 0.f : 1.0f);
  for (int c = lane; c < C; c += 32) {
    float s = 0.f;
    for (int p = 0; p < n_parts; ++p) {
      const float mp = m_scale * m_parts[p * ml_part_stride + q];
      const float w = (mp == -INFINITY) ? 0.f : ex2f(mp - M);
      s += w * O_parts[p * o_part_stride + q * ld + c];
    }
    out[q * ld_out + c] = s * inv;
  }
  if (lane == 0) {
    if (m_out != nullptr) m_out[q] = M;
    if (l_out != nullptr) l_out[q] = L;
  }
}

}  // namespace

namespace sc {

int fill_f32_async(float* dst, size_t n, float value, cudaStream_t st) {
  if (n == 0) return SC_OK;
  const size_t threads = (n + 3) / 4;
  fill_f32_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, st>>>(dst, n, value);
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

}  // namespace sc

extern "C" {

int sc_softmax_partials(const float* lse, int n_parts, int64_t part_stride, int64_t Nq, int64_t C, int64_t ld,
                        float* O, int64_t ld_out, float* m, float* l, void* stream) {
  SC_REQUIRE(lse && O && m && l, SC_EINVAL, "sc_softmax_partials: null pointer");
  SC_REQUIRE(n_parts >= 1 && Nq >= 0 && C > 0 && C <= 65535 * 32 && ld >= C && ld_out >= C, SC_ESHAPE,
             "sc_softmax_partials: bad shape");
  if (Nq == 0) return SC_OK;
  const unsigned blocks = static_cast<unsigned>(sc::ceil_div(Nq, kWarpsPerBlock));
  softmax_partials_kernel<<<blocks, kWarpsPerBlock * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      lse, n_parts, part_stride, Nq, static_cast<int>(C), ld, O, ld_out, m, l);
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

int sc_merge_softmax(const float* O_parts, const float* m_parts, const float* l_parts, int n_parts,
                     int64_t o_part_stride, int64_t ml_part_stride, int64_t Nq, int64_t C, int64_t ld, float m_scale,
                     const float* m_ref, int normalize, float* out, int64_t ld_out, float* m_out, float* l_out,
                     void* stream) {
  SC_REQUIRE(O_parts && m_parts && l_parts && out, SC_EINVAL, "sc_merge_softmax: null pointer");
  SC_REQUIRE(n_parts >= 1 && Nq >= 0 && C > 0 && ld >= C && ld_out >= C, SC_ESHAPE, "sc_merge_softmax: bad shape");
  if (Nq == 0) return SC_OK;
  const unsigned blocks = static_cast<unsigned>(sc::ceil_div(Nq, kWarpsPerBlock));
  merge_softmax_kernel<<<blocks, kWarpsPerBlock * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      O_parts, m_parts, l_parts, n_parts, o_part_stride, ml_part_stride, Nq, static_cast<int>(C), ld, m_scale, m_ref,
      normalize, out, ld_out, m_out, l_out);
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

}  // extern "C"
