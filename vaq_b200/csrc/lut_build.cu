// LUT build — replaces VAQ::CreateLUT (reference bitvecengine/VAQ.hpp:128-167) and the
// query projection VAQ::ProjectOnEigenVectors (VAQ.hpp:198-201).
//
// lut(q, s, c) = sum_j (q[s*L + j] - C_s[c][j])^2.
// The reference's AVX2 path (K_s >= 8) accumulates one fused multiply-add per dimension in
// dimension order (vfmadd231ps, utils/AVXUtils.hpp:11-15); the same fmaf chain here makes the
// table bit-identical.  K_s < 8 goes through fvec_L2sqr_ny (utils/Math.hpp:147-171) whose SSE
// specialisations for L in {1,2,4,8,12} use fixed summation trees, restated below.
//
// FP32 FMA only: the contraction depth is L (3..15) and the LUT must match the reference to
// 1e-5 relative, which TF32/BF16 tensor-core inputs (~1e-3) cannot deliver; the build is < 1 %
// of a search, so tcgen05 would buy nothing here (DESIGN.md "LUT build").
#include <cuda_fp16.h>

#include "common.cuh"

namespace vaqgpu {

__device__ __forceinline__ float sq_diff(float x, float y) {
  const float d = __fsub_rn(x, y);
  return __fmul_rn(d, d);
}

// utils/Math.hpp:38-128,147-171 with ElementOpL2 (:130-145)
__device__ __forceinline__ float l2sqr_small(const float *__restrict__ x, const float *__restrict__ y, int d) {
  switch (d) {
    case 1: return sq_diff(x[0], y[0]);
    case 2: return __fadd_rn(sq_diff(x[0], y[0]), sq_diff(x[1], y[1]));
    case 4:
      return __fadd_rn(__fadd_rn(sq_diff(x[0], y[0]), sq_diff(x[1], y[1])),
                       __fadd_rn(sq_diff(x[2], y[2]), sq_diff(x[3], y[3])));
    case 8: {
      float t[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float b = __fsub_rn(x[i + 4], y[i + 4]);
        t[i] = __fmaf_rn(b, b, sq_diff(x[i], y[i]));
      }
      return __fadd_rn(__fadd_rn(t[0], t[1]), __fadd_rn(t[2], t[3]));
    }
    case 12: {
      float t[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float b = __fsub_rn(x[i + 4], y[i + 4]);
        const float c = __fsub_rn(x[i + 8], y[i + 8]);
        t[i] = __fmaf_rn(c, c, __fmaf_rn(b, b, sq_diff(x[i], y[i])));
      }
      return __fadd_rn(__fadd_rn(t[0], t[1]), __fadd_rn(t[2], t[3]));
    }
    default: {
      float res = 0.f;
      for (int i = 0; i < d; i++) res = __fadd_rn(res, sq_diff(x[i], y[i]));
      return res;
    }
  }
}

// One thread per table entry and query TILE: the centroid is read once, the T queries of the tile are
// scored against it and the T values are written as one contiguous T*4-byte store — the interleaved
// [entry][T] layout the scan kernels read — so the stores of a warp are fully coalesced.  Optionally the
// same thread also writes the fp16 lower-bound entries for the tile (T == 8 only; scales from lut_scale_kernel):
//     e16 = round_toward_zero(scale_t * value),   scale_t = 2^floor(log2(16000 / ub_t)),
//     ub_t = max_s (|q_t,s| + max_c |C_s[c]|)^2  >= every entry of query t (triangle inequality),
// so scale_t * entry <= 16000 and four entries still sum below the fp16 maximum.  fp16 keeps 11 significant
// bits at any scale, so the looseness of ub only costs exponent range.
// Per-query scale of the fp16 lower-bound tables, one warp per query (run once per batch, before lut_build_kernel<8>).
__global__ void __launch_bounds__(256) lut_scale_kernel(const float *__restrict__ q_proj, int nq, int nq_pad, int D, int M, int L,
                                                        const float *__restrict__ cent_rmax, float *__restrict__ scale) {
  const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), ln = threadIdx.x & 31;
  if (qi >= nq_pad) return;
  const int q = min(qi, nq - 1);               // slots past nq (tile padding) repeat the last query
  float ub = 0.f;
  for (int s = ln; s < M; s += 32) {
    const float *qs = q_proj + (size_t)q * D + (size_t)s * L;
    float n2 = 0.f;
    for (int j = 0; j < L; j++) n2 = fmaf(qs[j], qs[j], n2);
    const float r = sqrtf(n2) + cent_rmax[s];
    ub = fmaxf(ub, r * r);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ub = fmaxf(ub, __shfl_xor_sync(0xffffffffu, ub, o));
  if (ln == 0) {
    float sc = 1.f;
    ub *= 1.0001f;                       // the entries are computed with rounding; keep a hair of slack
    if (ub > 0.f && ub < 3.0e38f) {
      int ex = (int)floorf(log2f(16000.f / ub));
      ex = max(-100, min(100, ex));
      sc = exp2f((float)ex);
      while (ub * sc > 16000.f) sc *= 0.5f;
    }
    scale[qi] = sc;
  }
}

// LC > 0: the subspace length is the compile-time constant LC — the centroid sits in registers and the dimension
// loop is unrolled (ncu, round 2: the generic kernel issued 625 instructions per warp and was issue-bound at 61 %,
// not memory-bound); LC == 0: any length.
template <int T, int LC>
__global__ void __launch_bounds__(256) lut_build_kernel(const float *__restrict__ q_proj, int nq, int D, const float *__restrict__ cent,
                                                        const __grid_constant__ LutPlan p,
                                                        float *__restrict__ lut, __half *__restrict__ lut16,
                                                        const float *__restrict__ scale) {
  const int qt = blockIdx.y;
  const int L = LC > 0 ? LC : p.L;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.total_entries) return;
  int lo = 0, hi = p.M;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (p.ent_off[mid] <= e) lo = mid; else hi = mid;
  }
  const int s = lo;
  const int c = e - p.ent_off[s];
  const int K = p.ent_off[s + 1] - p.ent_off[s];
  const float *cp = cent + p.cent_off[s] + (size_t)c * L;
  float cv[LC > 0 ? LC : 1];
  if constexpr (LC > 0) {
#pragma unroll
    for (int j = 0; j < LC; j++) cv[j] = __ldg(cp + j);
  }
  float acc[T];
#pragma unroll
  for (int t = 0; t < T; t++) {
    const int q = min(qt * T + t, nq - 1);      // slots past nq (tile padding) repeat the last query
    const float *qs = q_proj + (size_t)q * D + (size_t)s * L;
    if (K >= 8) {
      float a = 0.f;
      if constexpr (LC > 0) {
#pragma unroll
        for (int j = 0; j < LC; j++) {
          const float d = __fsub_rn(__ldg(qs + j), cv[j]);
          a = __fmaf_rn(d, d, a);
        }
      } else {
        for (int j = 0; j < L; j++) {
          const float d = __fsub_rn(__ldg(qs + j), __ldg(cp + j));
          a = __fmaf_rn(d, d, a);
        }
      }
      acc[t] = a;
    } else {
      acc[t] = l2sqr_small(qs, cp, L);
    }
  }
  const size_t o = ((size_t)qt * p.row_stride + p.pos[s] + c) * T;
  if constexpr (T == 1) {
    lut[o] = acc[0];
  } else if constexpr (T == 2) {
    *reinterpret_cast<float2 *>(lut + o) = make_float2(acc[0], acc[1]);
  } else {
#pragma unroll
    for (int i = 0; i < T / 4; i++)
      *reinterpret_cast<float4 *>(lut + o + 4 * i) = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
  }
  if constexpr (T == 8) {
    if (lut16 != nullptr) {
      const float4 s0 = __ldg(reinterpret_cast<const float4 *>(scale + (size_t)qt * T));
      const float4 s1 = __ldg(reinterpret_cast<const float4 *>(scale + (size_t)qt * T + 4));
      const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
      __half2 h[4];
#pragma unroll
      for (int i = 0; i < 4; i++)
        h[i] = __halves2half2(__float2half_rz(acc[2 * i] * sc[2 * i]), __float2half_rz(acc[2 * i + 1] * sc[2 * i + 1]));
      *reinterpret_cast<uint4 *>(lut16 + o) = make_uint4(*reinterpret_cast<uint32_t *>(&h[0]), *reinterpret_cast<uint32_t *>(&h[1]),
                                                        *reinterpret_cast<uint32_t *>(&h[2]), *reinterpret_cast<uint32_t *>(&h[3]));
    }
  }
}

// nq_launch (a multiple of plan.T) >= nq query slots are written.  lut16/scale may be NULL (fp32 tables only).
cudaError_t launch_lut_build(const float *q_proj, int nq, int nq_launch, int D, const float *centroids, const float *cent_rmax,
                             const LutPlan &plan, float *lut, void *lut16, float *scale, cudaStream_t st) {
  if (nq <= 0) return cudaSuccess;
  const int threads = 256;
  const int T = plan.T;
  dim3 grid((unsigned)((plan.total_entries + threads - 1) / threads), (unsigned)((nq_launch + T - 1) / T));
  __half *l16 = reinterpret_cast<__half *>(lut16);
  if (T == 8 && l16 != nullptr) {
    const int nq_pad = (nq_launch + T - 1) / T * T;
    lut_scale_kernel<<<(nq_pad + 7) / 8, 256, 0, st>>>(q_proj, nq, nq_pad, D, plan.M, plan.L, cent_rmax, scale);
  }
  switch (T) {
    case 1: lut_build_kernel<1, 0><<<grid, threads, 0, st>>>(q_proj, nq, D, centroids, plan, lut, nullptr, nullptr); break;
    case 2: lut_build_kernel<2, 0><<<grid, threads, 0, st>>>(q_proj, nq, D, centroids, plan, lut, nullptr, nullptr); break;
    case 4: lut_build_kernel<4, 0><<<grid, threads, 0, st>>>(q_proj, nq, D, centroids, plan, lut, nullptr, nullptr); break;
    case 8:
      switch (plan.L) {      // the subspace lengths of the BASELINE shapes get unrolled kernels
#define LUT8(LCV) case LCV: lut_build_kernel<8, LCV><<<grid, threads, 0, st>>>(q_proj, nq, D, centroids, plan, lut, l16, scale); break;
        LUT8(2) LUT8(3) LUT8(4) LUT8(6) LUT8(8) LUT8(12) LUT8(15) LUT8(16)
#undef LUT8
        default: lut_build_kernel<8, 0><<<grid, threads, 0, st>>>(q_proj, nq, D, centroids, plan, lut, l16, scale); break;
      }
      break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// out[r][j] = sum_i x[r][i] * eig[i][j]  — (X * mEigenVectors).real(), VAQ.hpp:198-201.
// Plain FP32 FMA in i order; Eigen's GEMM blocks differently, so this step is
// tolerance-only (parity runs feed host-projected queries with VAQGPU_PROJECTED).
__global__ void project_kernel(const float *__restrict__ x, int D, const float *__restrict__ eig, float *__restrict__ out) {
  extern __shared__ float xs[];
  const int r = blockIdx.x;
  for (int i = threadIdx.x; i < D; i += blockDim.x) xs[i] = x[(size_t)r * D + i];
  __syncthreads();
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < D; i++) acc = __fmaf_rn(xs[i], __ldg(eig + (size_t)i * D + j), acc);
    out[(size_t)r * D + j] = acc;
  }
}

cudaError_t launch_project(const float *x, int n, int D, const float *eig, float *out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int threads = D >= 256 ? 256 : 128;
  project_kernel<<<n, threads, D * sizeof(float), st>>>(x, D, eig, out);
  return cudaGetLastError();
}

}  // namespace vaqgpu
