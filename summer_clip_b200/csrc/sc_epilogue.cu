// Epilogue-side kernels of the CLIP-search path (all HBM-bound):
//   sc_merge_partials   : sum of per-split / per-rank partial O tiles (the LSE merge with m == 0)
//   sc_zero_shot_logits : Z = scale * normalise_cols(X)^T @ T in fp32   image_attention.py:80-83
//   sc_epilogue         : out = Z + O*alpha, argmax, top-1/top-5 counts  image_attention.py:111-117,
//                         clip_searcher/utils.py:15-21, clip_adapter/train_adapter.py:156-159,
//                         tip_adapter/utils.py:10-15
#include "sc_common.cuh"
#include "sc_rowops.cuh"

namespace {

constexpr int kMaxAlphas = 64;
struct AlphaList {
  float a[kMaxAlphas];
};

__global__ void __launch_bounds__(256)
merge_kernel(const float* __restrict__ parts, int n_parts, int64_t rows, int64_t cols, int64_t ld,
             float* __restrict__ out, int64_t ld_out) {
  const int64_t total = rows * cols;
  const int64_t part_stride = rows * ld;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = e / cols, c = e - r * cols;
    const float* src = parts + r * ld + c;
    float s = src[0];
    for (int p = 1; p < n_parts; ++p) s += src[p * part_stride];
    out[r * ld_out + c] = s;
  }
}

// cols, ld, ld_out multiples of 4 and 16-byte aligned bases: 16-byte loads, one row segment per thread, all parts
// of a vector in flight before the adds (same summation order as merge_kernel)
template <int kParts>
__global__ void __launch_bounds__(256)
merge_vec_kernel(const float4* __restrict__ parts, int n_parts, int64_t rows, int cols4, int64_t ld4,
                 float4* __restrict__ out, int64_t ld_out4) {
  const int64_t part_stride = rows * ld4;
  const int64_t total = rows * cols4;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = e / cols4;
    const int c = static_cast<int>(e - r * cols4);
    const float4* src = parts + r * ld4 + c;
    float4 s;
    if (kParts > 0) {
      float4 v[kParts > 0 ? kParts : 1];
#pragma unroll
      for (int p = 0; p < kParts; ++p) v[p] = __ldcs(src + p * part_stride);
      s = v[0];
#pragma unroll
      for (int p = 1; p < kParts; ++p) { s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w; }
    } else {
      s = __ldcs(src);
      for (int p = 1; p < n_parts; ++p) {
        const float4 v = __ldcs(src + p * part_stride);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    }
    out[r * ld_out4 + c] = s;
  }
}

// fp32 SIMT GEMM, 64 x 64 output tile per block, 4 x 4 per thread, K step 16.
//   Z[n, c] = scale / ||X[:, n]|| * sum_d X[d, n] * T[d, c]
template <typename TX, typename TT>
__global__ void __launch_bounds__(256)
zero_shot_kernel(const TX* __restrict__ X, int64_t D, int64_t N, int64_t stride_d, int64_t stride_n,
                 const TT* __restrict__ Tm, int64_t C, int64_t ldt, float scale, int normalize,
                 float* __restrict__ Z, int64_t ldz) {
  __shared__ float Xs[16][64 + 4];
  __shared__ float Ts[16][64 + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15;   // class group
  const int ty = tid >> 4;   // query group
  const int64_t n0 = static_cast<int64_t>(blockIdx.y) * 64;
  const int64_t c0 = static_cast<int64_t>(blockIdx.x) * 64;
  const bool n_contig = (stride_n == 1) || (stride_d != 1);
  float acc[4][4] = {};
  float ss[4] = {};
  for (int64_t d0 = 0; d0 < D; d0 += 16) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = tid + 256 * j;
      int dd, nn;
      if (n_contig) { dd = e >> 6; nn = e & 63; } else { nn = e >> 4; dd = e & 15; }
      const int64_t d = d0 + dd, n = n0 + nn;
      Xs[dd][nn] = (d < D && n < N) ? sc::to_f32<TX>(X[d * stride_d + n * stride_n]) : 0.f;
      const int td = e >> 6, tc = e & 63;
      const int64_t d2 = d0 + td, c = c0 + tc;
      Ts[td][tc] = (d2 < D && c < C) ? sc::to_f32<TT>(Tm[d2 * ldt + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float xv[4], tv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { xv[i] = Xs[k][ty * 4 + i]; tv[i] = Ts[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ss[i] = fmaf(xv[i], xv[i], ss[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], tv[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t n = n0 + ty * 4 + i;
    if (n >= N) continue;
    const float f = normalize ? scale / sqrtf(ss[i]) : scale;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t c = c0 + tx * 4 + j;
      if (c < C) Z[n * ldz + c] = acc[i][j] * f;
    }
  }
}

// sum of the n_parts partial tiles of one O element, in merge_kernel's order (bit-identical to a merged O)
__device__ __forceinline__ float o_sum(const float* __restrict__ src, int n_parts, int64_t part_stride) {
  float s = src[0];
  for (int p = 1; p < n_parts; ++p) s += src[p * part_stride];
  return s;
}

// Generic path (C > 1024): one warp per query row; the row is re-read per alpha from L1.
__global__ void __launch_bounds__(256)
epilogue_kernel(const float* __restrict__ Z, int64_t ldz, const float* __restrict__ O, int64_t ldo, int n_parts,
                int64_t part_stride, const float* __restrict__ rowsum, int64_t Nq, int64_t C, AlphaList alphas, int na,
                const int32_t* __restrict__ labels, float* __restrict__ out_logits,
                int32_t* __restrict__ pred, int32_t* __restrict__ top1, int32_t* __restrict__ top5) {
  __shared__ int32_t s_top1[kMaxAlphas];
  __shared__ int32_t s_top5[kMaxAlphas];
  for (int i = threadIdx.x; i < na; i += blockDim.x) { s_top1[i] = 0; s_top5[i] = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warps_per_grid = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t q = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); q < Nq;
       q += warps_per_grid) {
    const float* zrow = Z ? Z + q * ldz : nullptr;
    const float* orow = O + q * ldo;
    const float inv = rowsum ? 1.0f / rowsum[q] : 1.0f;
    const int lab = labels ? labels[q] : -1;
    for (int ai = 0; ai < na; ++ai) {
      const float alpha = alphas.a[ai];
      float vlab = 0.f;
      if (lab >= 0 && lab < C) {
        float o = o_sum(orow + lab, n_parts, part_stride);
        if (rowsum) o *= inv;
        vlab = __fadd_rn(zrow ? zrow[lab] : 0.f, __fmul_rn(o, alpha));
      }
      float best = 0.f;
      int besti = -1;
      int ahead = 0;   // classes ranked strictly before the label (value desc, index asc)
      for (int64_t c = lane; c < C; c += 32) {
        float o = o_sum(orow + c, n_parts, part_stride);
        if (rowsum) o *= inv;
        const float v = __fadd_rn(zrow ? zrow[c] : 0.f, __fmul_rn(o, alpha));
        if (out_logits) out_logits[(static_cast<int64_t>(ai) * Nq + q) * C + c] = v;
        if (besti < 0 || v > best) { best = v; besti = static_cast<int>(c); }
        if (lab >= 0 && (v > vlab || (v == vlab && c < lab))) ++ahead;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (oi >= 0 && (besti < 0 || ov > best || (ov == best && oi < besti))) { best = ov; besti = oi; }
        ahead += __shfl_xor_sync(0xffffffffu, ahead, o);
      }
      if (lane == 0) {
        if (pred) pred[static_cast<int64_t>(ai) * Nq + q] = besti;
        if (lab >= 0 && lab < C) {
          if (ahead == 0) atomicAdd(&s_top1[ai], 1);
          if (ahead < 5) atomicAdd(&s_top5[ai], 1);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < na; i += blockDim.x) {
    if (top1 && s_top1[i]) atomicAdd(&top1[i], s_top1[i]);
    if (top5 && s_top5[i]) atomicAdd(&top5[i], s_top5[i]);
  }
}

// Register-resident path (C <= 32 * NV <= 1024): one warp per query row, Z and the (summed) O row loaded ONCE
// into registers, then every alpha is ~5 instructions per class: out = Z + O * alpha (separately rounded, as
// torch), "classes ahead of the label" counted as (v > v_label) plus, only when a value ties with the label's
// (warp-uniform, rare), the exact index-ordered recount.  kPred adds the running max.NaN and one pass for the
// FIRST index of the maximum; rows whose maximum is NaN take the generic kernel's ordered compare.
template <int NV, bool kPred>
__global__ void __launch_bounds__(256)
epilogue_reg_kernel(const float* __restrict__ Z, int64_t ldz, const float* __restrict__ O, int64_t ldo, int n_parts,
                    int64_t part_stride, const float* __restrict__ rowsum, int64_t Nq, int C, AlphaList alphas, int na,
                    const int32_t* __restrict__ labels, float* __restrict__ out_logits,
                    int32_t* __restrict__ pred, int32_t* __restrict__ top1, int32_t* __restrict__ top5) {
  __shared__ int32_t s_top1[kMaxAlphas];
  __shared__ int32_t s_top5[kMaxAlphas];
  for (int i = threadIdx.x; i < na; i += blockDim.x) { s_top1[i] = 0; s_top5[i] = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const float kNegInf = __int_as_float(0xff800000);
  const int64_t warps_per_grid = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t q = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); q < Nq;
       q += warps_per_grid) {
    const float* zrow = Z ? Z + q * ldz : nullptr;
    const float* orow = O + q * ldo;
    const float inv = rowsum ? 1.0f / rowsum[q] : 1.0f;
    float z[NV], o[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = j * 32 + lane;
      o[j] = c < C ? __ldg(orow + c) : 0.f;
    }
    for (int p = 1; p < n_parts; ++p) {                  // a whole part row in flight before the first add
      float t[NV];
      const float* prow = orow + p * part_stride;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = j * 32 + lane;
        t[j] = c < C ? __ldg(prow + c) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) o[j] += t[j];
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {                       // padding classes: -inf + 0 * alpha never wins
      const int c = j * 32 + lane;
      z[j] = c < C ? (zrow ? __ldg(zrow + c) : 0.f) : kNegInf;
    }
    if (rowsum) {
#pragma unroll
      for (int j = 0; j < NV; ++j) o[j] *= inv;
    }
    const int lab = labels ? labels[q] : -1;
    const bool lab_ok = lab >= 0 && lab < C;             // warp-uniform
    float zlab = 0.f, olab = 0.f;
    if (lab_ok) {
      zlab = zrow ? zrow[lab] : 0.f;
      olab = o_sum(orow + lab, n_parts, part_stride);
      if (rowsum) olab *= inv;
    }
    for (int ai = 0; ai < na; ++ai) {
      const float alpha = alphas.a[ai];
      const float vlab = __fadd_rn(zlab, __fmul_rn(olab, alpha));
      float m = kNegInf;
      int cnt = 0;                                       // (v > vlab) in the low half, (v == vlab) in the high half
      if (out_logits) {
        float* dst = out_logits + (static_cast<int64_t>(ai) * Nq + q) * C;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const int c = j * 32 + lane;
          if (c < C) dst[c] = __fadd_rn(z[j], __fmul_rn(o[j], alpha));
        }
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float v = __fadd_rn(z[j], __fmul_rn(o[j], alpha));
        if (kPred) m = sc::max_nan(m, v);
        cnt += (v > vlab ? 1 : 0) + (v == vlab ? 0x10000 : 0);
      }
      bool slow = false;
      if (kPred) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) m = sc::max_nan(m, __shfl_xor_sync(0xffffffffu, m, s));
        slow = m != m;
      }
      if (lab_ok) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
        slow = slow || (cnt >> 16) != 1;                 // a tie with the label's value (or a NaN label value)
      }
      int ahead = cnt & 0xffff, besti = -1;
      if (!slow) {
        if (kPred) {
          int idx = 0x7fffffff;
#pragma unroll
          for (int j = NV - 1; j >= 0; --j)
            if (__fadd_rn(z[j], __fmul_rn(o[j], alpha)) == m) idx = j * 32 + lane;
#pragma unroll
          for (int s = 16; s > 0; s >>= 1) idx = min(idx, __shfl_xor_sync(0xffffffffu, idx, s));
          besti = idx < C ? idx : -1;                    // all -inf rows: the first class, as the ordered compare
          if (idx >= C) slow = true;
        }
      }
      if (slow) {                                        // the generic kernel's ordered compare, from registers
        float best = 0.f;
        besti = -1;
        ahead = 0;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const int c = j * 32 + lane;
          if (c < C) {
            const float v = __fadd_rn(z[j], __fmul_rn(o[j], alpha));
            if (besti < 0 || v > best) { best = v; besti = c; }
            if (lab_ok && (v > vlab || (v == vlab && c < lab))) ++ahead;
          }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, best, s);
          const int oi = __shfl_xor_sync(0xffffffffu, besti, s);
          if (oi >= 0 && (besti < 0 || ov > best || (ov == best && oi < besti))) { best = ov; besti = oi; }
          ahead += __shfl_xor_sync(0xffffffffu, ahead, s);
        }
      }
      if (lane == 0) {
        if (kPred && pred) pred[static_cast<int64_t>(ai) * Nq + q] = besti;
        if (lab_ok) {
          if (ahead == 0) atomicAdd(&s_top1[ai], 1);
          if (ahead < 5) atomicAdd(&s_top5[ai], 1);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < na; i += blockDim.x) {
    if (top1 && s_top1[i]) atomicAdd(&top1[i], s_top1[i]);
    if (top5 && s_top5[i]) atomicAdd(&top5[i], s_top5[i]);
  }
}

template <int NV>
void launch_epilogue_reg(unsigned blocks, cudaStream_t st, const float* Z, int64_t ldz, const float* O, int64_t ldo,
                         int n_parts, int64_t part_stride, const float* rowsum, int64_t Nq, int C, const AlphaList& al,
                         int na, const int32_t* labels, float* out_logits, int32_t* pred, int32_t* top1, int32_t* top5) {
  if (pred)
    epilogue_reg_kernel<NV, true><<<blocks, 256, 0, st>>>(Z, ldz, O, ldo, n_parts, part_stride, rowsum, Nq, C, al, na,
                                                          labels, out_logits, pred, top1, top5);
  else
    epilogue_reg_kernel<NV, false><<<blocks, 256, 0, st>>>(Z, ldz, O, ldo, n_parts, part_stride, rowsum, Nq, C, al, na,
                                                           labels, out_logits, pred, top1, top5);
}

// The same sum over tiles that live at UNRELATED addresses: the partial tiles of the key-sharded ranks, read in
// place from every peer's memory over NVLink (peer-mapped pointers; the caller orders the read after the peers'
// writes).  Parts are added in index order on every rank, so all ranks agree bit for bit.  16-byte loads when the
// shapes allow; every part of a vector is in flight before the adds (~2 us of NVLink latency per dependent load).
constexpr int kMaxPeerParts = 16;
struct PeerParts {
  const float* p[kMaxPeerParts];
};

template <bool kVec>
__global__ void __launch_bounds__(256)
merge_peers_kernel(PeerParts parts, int n_parts, int64_t rows, int64_t cols, int64_t ld, float* __restrict__ out,
                   int64_t ld_out) {
  if (kVec) {
    const int64_t c4 = cols / 4, total = rows * c4;
    for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      const int64_t r = e / c4, c = (e - r * c4) * 4;
      float4 v[kMaxPeerParts];
#pragma unroll
      for (int p = 0; p < kMaxPeerParts; ++p)
        if (p < n_parts) v[p] = *reinterpret_cast<const float4*>(parts.p[p] + r * ld + c);
      float4 s = v[0];
#pragma unroll
      for (int p = 1; p < kMaxPeerParts; ++p)
        if (p < n_parts) { s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w; }
      *reinterpret_cast<float4*>(out + r * ld_out + c) = s;
    }
  } else {
    const int64_t total = rows * cols;
    for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      const int64_t r = e / cols, c = e - r * cols;
      float s = parts.p[0][r * ld + c];
      for (int p = 1; p < n_parts; ++p) s += parts.p[p][r * ld + c];
      out[r * ld_out + c] = s;
    }
  }
}

}  // namespace

extern "C" {

int sc_merge_peer_parts(const float* const* parts, int n_parts, int64_t rows, int64_t cols, int64_t ld, float* out,
                        int64_t ld_out, void* stream) {
  SC_REQUIRE(parts && out, SC_EINVAL, "sc_merge_peer_parts: null pointer");
  SC_REQUIRE(n_parts >= 1 && n_parts <= kMaxPeerParts, SC_ESHAPE, "sc_merge_peer_parts: n_parts=%d must be in [1, %d]",
             n_parts, kMaxPeerParts);
  SC_REQUIRE(rows >= 0 && cols >= 0 && ld >= cols && ld_out >= cols, SC_ESHAPE, "sc_merge_peer_parts: bad shape");
  if (rows * cols == 0) return SC_OK;
  PeerParts pp;
  uintptr_t bits = reinterpret_cast<uintptr_t>(out);
  for (int p = 0; p < kMaxPeerParts; ++p) {
    pp.p[p] = parts[p < n_parts ? p : 0];
    SC_REQUIRE(pp.p[p] != nullptr, SC_EINVAL, "sc_merge_peer_parts: null part pointer");
    bits |= reinterpret_cast<uintptr_t>(pp.p[p]);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec = cols % 4 == 0 && ld % 4 == 0 && ld_out % 4 == 0 && bits % 16 == 0;
  const int64_t want = sc::ceil_div(rows * (vec ? cols / 4 : cols), 256);
  const unsigned blocks = static_cast<unsigned>(want < 148 * 8 ? want : 148 * 8);
  if (vec) merge_peers_kernel<true><<<blocks, 256, 0, st>>>(pp, n_parts, rows, cols, ld, out, ld_out);
  else merge_peers_kernel<false><<<blocks, 256, 0, st>>>(pp, n_parts, rows, cols, ld, out, ld_out);
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

int sc_merge_partials(const float* parts, int n_parts, int64_t rows, int64_t cols, int64_t ld,
                      float* out, int64_t ld_out, void* stream) {
  SC_REQUIRE(parts && out, SC_EINVAL, "sc_merge_partials: null pointer");
  SC_REQUIRE(n_parts >= 1 && rows >= 0 && cols >= 0 && ld >= cols && ld_out >= cols, SC_ESHAPE,
             "sc_merge_partials: bad shape");
  if (rows * cols == 0) return SC_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec = cols % 4 == 0 && ld % 4 == 0 && ld_out % 4 == 0 && cols / 4 <= 0x7fffffff &&
                   (reinterpret_cast<uintptr_t>(parts) | reinterpret_cast<uintptr_t>(out)) % 16 == 0;
  if (vec) {
    const int64_t want = sc::ceil_div(rows * (cols / 4), 256);
    const unsigned blocks = static_cast<unsigned>(want < 148 * 16 ? want : 148 * 16);
    const float4* p4 = reinterpret_cast<const float4*>(parts);
    float4* o4 = reinterpret_cast<float4*>(out);
    const int c4 = static_cast<int>(cols / 4);
    switch (n_parts) {
      case 2: merge_vec_kernel<2><<<blocks, 256, 0, st>>>(p4, n_parts, rows, c4, ld / 4, o4, ld_out / 4); break;
      case 3: merge_vec_kernel<3><<<blocks, 256, 0, st>>>(p4, n_parts, rows, c4, ld / 4, o4, ld_out / 4); break;
      case 4: merge_vec_kernel<4><<<blocks, 256, 0, st>>>(p4, n_parts, rows, c4, ld / 4, o4, ld_out / 4); break;
      case 8: merge_vec_kernel<8><<<blocks, 256, 0, st>>>(p4, n_parts, rows, c4, ld / 4, o4, ld_out / 4); break;
      default: merge_vec_kernel<0><<<blocks, 256, 0, st>>>(p4, n_parts, rows, c4, ld / 4, o4, ld_out / 4); break;
    }
  } else {
    const int64_t want = sc::ceil_div(rows * cols, 256);
    const unsigned blocks = static_cast<unsigned>(want < 148 * 16 ? want : 148 * 16);
    merge_kernel<<<blocks, 256, 0, st>>>(parts, n_parts, rows, cols, ld, out, ld_out);
  }
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

int sc_zero_shot_logits(const void* X, int x_dtype, int64_t D, int64_t N, int64_t stride_d,
                        int64_t stride_n, const void* T, int t_dtype, int64_t C, int64_t ldt,
                        float scale, int normalize, float* Z, int64_t ldz, void* stream) {
  SC_REQUIRE(X && T && Z, SC_EINVAL, "sc_zero_shot_logits: null pointer");
  SC_REQUIRE(D > 0 && N >= 0 && C > 0 && ldt >= C && ldz >= C, SC_ESHAPE, "sc_zero_shot_logits: bad shape");
  if (N == 0) return SC_OK;
  dim3 grid(static_cast<unsigned>(sc::ceil_div(C, 64)), static_cast<unsigned>(sc::ceil_div(N, 64)));
  SC_REQUIRE(grid.y <= 65535, SC_ESHAPE, "sc_zero_shot_logits: too many rows; chunk the queries");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_DISPATCH_DTYPE(x_dtype, TX, {
    SC_DISPATCH_DTYPE(t_dtype, TT,
                      (zero_shot_kernel<TX, TT><<<grid, 256, 0, st>>>(
                          static_cast<const TX*>(X), D, N, stride_d, stride_n,
                          static_cast<const TT*>(T), C, ldt, scale, normalize, Z, ldz)));
  });
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

int sc_epilogue_parts(const float* Z, int64_t ldz, const float* O, int64_t ldo, int n_parts, int64_t part_stride,
                      const float* rowsum, int64_t Nq, int64_t C, const float* alphas, int na, const int32_t* labels,
                      float* out_logits, int32_t* pred, int32_t* top1, int32_t* top5, void* stream) {
  SC_REQUIRE(O && alphas, SC_EINVAL, "sc_epilogue: null pointer");
  SC_REQUIRE(na >= 1 && na <= kMaxAlphas, SC_ESHAPE, "sc_epilogue: na=%d must be in [1, %d]", na, kMaxAlphas);
  SC_REQUIRE(Nq >= 0 && C > 0 && ldo >= C && (Z == nullptr || ldz >= C), SC_ESHAPE, "sc_epilogue: bad shape");
  SC_REQUIRE(n_parts >= 1 && (n_parts == 1 || part_stride >= Nq * ldo), SC_ESHAPE,
             "sc_epilogue: n_parts=%d needs part_stride >= Nq * ldo", n_parts);
  if (Nq == 0) return SC_OK;
  AlphaList al;
  for (int i = 0; i < na; ++i) al.a[i] = alphas[i];   // alphas is a HOST array
  const int64_t want = sc::ceil_div(Nq, 8);
  const unsigned blocks = static_cast<unsigned>(want < 148 * 8 ? want : 148 * 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int c = static_cast<int>(C);
#define SC_EPI_REG(NV)                                                                                              \
  launch_epilogue_reg<NV>(blocks, st, Z, ldz, O, ldo, n_parts, part_stride, rowsum, Nq, c, al, na, labels, out_logits, \
                          pred, top1, top5)
  if (C <= 128) SC_EPI_REG(4);
  else if (C <= 256) SC_EPI_REG(8);
  else if (C <= 512) SC_EPI_REG(16);
  else if (C <= 1024) SC_EPI_REG(32);
  else
    epilogue_kernel<<<blocks, 256, 0, st>>>(Z, ldz, O, ldo, n_parts, part_stride, rowsum, Nq, C, al, na, labels,
                                            out_logits, pred, top1, top5);
#undef SC_EPI_REG
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

int sc_epilogue(const float* Z, int64_t ldz, const float* O, int64_t ldo, const float* rowsum,
                int64_t Nq, int64_t C, const float* alphas, int na, const int32_t* labels,
                float* out_logits, int32_t* pred, int32_t* top1, int32_t* top5, void* stream) {
  return sc_epilogue_parts(Z, ldz, O, ldo, 1, 0, rowsum, Nq, C, alphas, na, labels, out_logits, pred, top1, top5, stream);
}

}  // extern "C"
