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

__global__ void lut_build_kernel(const float *__restrict__ q_proj, int nq, int D, const float *__restrict__ cent,
                                 const __grid_constant__ LutPlan p, float *__restrict__ lut) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.total_entries) return;
  const int qo = blockIdx.y;            // output slot; slots past nq (tile padding) repeat the last query
  const int q = min(qo, nq - 1);
  int lo = 0, hi = p.M;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (p.ent_off[mid] <= e) lo = mid; else hi = mid;
  }
  const int s = lo;
  const int c = e - p.ent_off[s];
  const int K = p.ent_off[s + 1] - p.ent_off[s];
  const int L = p.L;
  const float *cp = cent + p.cent_off[s] + (size_t)c * L;
  const float *qs = q_proj + (size_t)q * D + (size_t)s * L;
  float acc;
  if (K >= 8) {
    acc = 0.f;
    for (int j = 0; j < L; j++) {
      const float d = __fsub_rn(__ldg(qs + j), __ldg(cp + j));
      acc = __fmaf_rn(d, d, acc);
    }
  } else {
    acc = l2sqr_small(qs, cp, L);
  }
  lut[((size_t)(qo / p.T) * p.row_stride + p.pos[s] + c) * p.T + (qo % p.T)] = acc;
}

cudaError_t launch_lut_build(const float *q_proj, int nq, int nq_launch, int D, const float *centroids,
                             const LutPlan &plan, float *lut, cudaStream_t st) {
  if (nq <= 0) return cudaSuccess;
  const int threads = 256;
  dim3 grid((unsigned)((plan.total_entries + threads - 1) / threads), (unsigned)nq_launch);
  lut_build_kernel<<<grid, threads, 0, st>>>(q_proj, nq, D, centroids, plan, lut);
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
