// Adjacent steps of the search API: exact re-rank (VAQ::refine, reference VAQ.cpp:849-876) and
// the per-query cluster ranking that turns the `visit` parameter into row ranges for the scan
// (prologue of VAQ::search in TI mode, VAQ.cpp:799-827, and the visiting rule of
// VAQ::searchTriangleInequality, VAQ.cpp:1548-1555, 1616-1618).
#include <float.h>
#include "common.cuh"

namespace vaqgpu {

// One CTA per query.  Warp per candidate: squared L2 between the raw query and the raw row
// (Eigen's squaredNorm reduction order is library-chosen, so distances are tolerance-only);
// then the k smallest (distance, label) pairs by rank counting.
__global__ void refine_kernel(const float *__restrict__ xtrain, int64_t n, int D, const float *__restrict__ queries,
                              const int32_t *__restrict__ in_labels, int R, int k, int32_t *__restrict__ labels,
                              float *__restrict__ dists) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t *keys = reinterpret_cast<uint64_t *>(smem_raw);   // [R]
  const int q = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float *qv = queries + (size_t)q * D;
  for (int i = warp; i < R; i += nwarps) {
    const int32_t id = in_labels[(size_t)q * R + i];
    uint64_t key = kEmptyKey;
    if (id >= 0 && (int64_t)id < n) {
      const float *x = xtrain + (size_t)id * D;
      float acc = 0.f;
      for (int j = lane; j < D; j += 32) {
        const float d = qv[j] - __ldg(x + j);
        acc = fmaf(d, d, acc);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      key = make_key_f32(acc, id);
    }
    if (lane == 0) keys[i] = key;
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    labels[(size_t)q * k + i] = -1;
    dists[(size_t)q * k + i] = FLT_MAX;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < R; i += blockDim.x) {
    const uint64_t key = keys[i];
    if (key == kEmptyKey) continue;
    int rank = 0;
    for (int j = 0; j < R && rank < k; j++) rank += (keys[j] < key) || (keys[j] == key && j < i);
    if (rank < k) {
      labels[(size_t)q * k + rank] = (int32_t)(uint32_t)key;
      dists[(size_t)q * k + rank] = __uint_as_float((uint32_t)(key >> 32));
    }
  }
}

cudaError_t launch_refine(const float *xtrain, int64_t n, int D, const float *queries, int nq,
                          const int32_t *in_labels, int refine_num, int k, int32_t *labels, float *dists,
                          cudaStream_t st) {
  if (nq <= 0) return cudaSuccess;
  const size_t smem = (size_t)refine_num * sizeof(uint64_t);
  static SmemOptIn optin;
  {
    cudaError_t e = optin.ensure(refine_kernel, smem);
    if (e != cudaSuccess) return e;
  }
  refine_kernel<<<nq, 256, smem, st>>>(xtrain, n, D, queries, in_labels, refine_num, k, labels, dists);
  return cudaGetLastError();
}

// One CTA per query: distance to every cluster centre over the first `segdims` projected dims
// (sqrt of the sequential sum of squares — fvec_L2sqr_ny's generic path, utils/Math.hpp:8-35),
// clusters ranked ascending (ties by index), then the reference's visiting rule: the nearest
// floor(C*visit) clusters (all if visit >= 1), continuing past that while fewer than k rows were
// covered; empty clusters are skipped.  Emits (row_begin, row_end) ranges in visiting order.
__global__ void rank_clusters_kernel(const float *__restrict__ q_proj, int D, const float *__restrict__ clusters,
                                     int C, int segdims, const int64_t *__restrict__ start,
                                     const int64_t *__restrict__ size, const int64_t *__restrict__ rule_size, float visit, int k,
                                     int2 *__restrict__ ranges, int32_t *__restrict__ n_ranges) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *dist = reinterpret_cast<float *>(smem_raw);          // [C]
  int32_t *order = reinterpret_cast<int32_t *>(dist + C);      // [C]
  const int q = blockIdx.x;
  const float *qv = q_proj + (size_t)q * D;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float *cc = clusters + (size_t)c * segdims;
    float acc = 0.f;
    for (int j = 0; j < segdims; j++) {
      const float d = __fsub_rn(qv[j], __ldg(cc + j));
      acc = __fadd_rn(acc, __fmul_rn(d, d));
    }
    dist[c] = sqrtf(acc);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float d = dist[c];
    int rank = 0;
    for (int j = 0; j < C; j++) rank += (dist[j] < d) || (dist[j] == d && j < c);
    order[rank] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int maxVisit = C;
    if (visit < 1.f) maxVisit = (int)((float)C * visit);
    int nr = 0;
    int64_t seen = 0;
    bool enough = false;
    for (int cc = 0; (cc < maxVisit) || (!enough && cc < C); cc++) {
      const int cl = order[cc];
      if (rule_size[cl] == 0) continue;          // the rule counts the rows of the whole index (a row shard holds a part)
      if (size[cl] > 0) {
        ranges[(size_t)q * C + nr] = make_int2((int)start[cl], (int)(start[cl] + size[cl]));
        nr++;
      }
      seen += rule_size[cl];
      if (seen >= k) enough = true;
    }
    n_ranges[q] = nr;
  }
}

cudaError_t launch_rank_clusters(const float *q_proj, int nq, int D, const float *clusters, int C, int segdims,
                                 const int64_t *start, const int64_t *size, const int64_t *rule_size, float visit, int k,
                                 int2 *ranges, int32_t *n_ranges, cudaStream_t st) {
  if (nq <= 0) return cudaSuccess;
  const size_t smem = (size_t)C * 8;
  static SmemOptIn optin;
  {
    cudaError_t e = optin.ensure(rank_clusters_kernel, smem);
    if (e != cudaSuccess) return e;
  }
  rank_clusters_kernel<<<nq, 256, smem, st>>>(q_proj, D, clusters, C, segdims, start, size, rule_size, visit, k, ranges, n_ranges);
  return cudaGetLastError();
}

}  // namespace vaqgpu
