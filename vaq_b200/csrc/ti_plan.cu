// TI / `visit` searches on the filter kernels: which clusters each query visits, how the queries are grouped into
// tiles, and which (query tile, cluster) pairs the scan may skip.
//
// Reference semantics (VAQ::search prologue VAQ.cpp:799-827 + searchTriangleInequality VAQ.cpp:1540-1692): the TI
// clusters are ranked by ||q[0:segdims] - cc|| (sqrt of the sequential sum of squares, ties by index), the nearest
// floor(C * visit) are visited (all when visit >= 1) and the visit continues past that while fewer than k rows were
// covered; inside a visited cluster the reference's triangle-inequality `break` is exact, so the answer is the exact
// top-k over the rows of the visited clusters.  Here:
//   ti_visit_kernel     one warp per query: cluster distances, the visited set as a byte map, the nearest cluster
//   ti_order_kernel     queries sorted by nearest cluster, so the eight queries of a scan tile visit similar sets
//                       (the scan streams the union of a tile's clusters; unrelated queries would visit ~90 % of the
//                       clusters between them at visit = 25 %)
//   ti_tiles_kernel     gathers the projected queries into tile order and ORs the visited maps of a tile into one
//                       byte per (tile, cluster): bit j = query j of the tile visits the cluster
//   tile_cluster_kernel (index time) cluster of each 32-row tile, 0xFFFF where a tile straddles clusters
#include "common.cuh"

namespace vaqgpu {

namespace {
constexpr int kVisitWarps = 4;
}

// dist[c] in shared memory (C floats per warp).  visited: [nq][C] bytes, nearest: [nq].
__global__ void __launch_bounds__(kVisitWarps * 32) ti_visit_kernel(const float *__restrict__ q_proj, int nq, int D,
                                                                    const float *__restrict__ clusters_t, int C, int segdims,
                                                                    const int64_t *__restrict__ rule_size, float visit, int k,
                                                                    uint8_t *__restrict__ visited, int32_t *__restrict__ nearest) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.x * kVisitWarps + warp;
  if (q >= nq) return;
  float *dist = reinterpret_cast<float *>(smem_raw) + (size_t)warp * C;
  const float *qv = q_proj + (size_t)q * D;
  uint8_t *vis = visited + (size_t)q * C;
  // distances: the reference's generic fvec_L2sqr_ny path (utils/Math.hpp:8-35): sequential sum of (x - y)^2, then sqrt
  uint64_t best = 0xFFFFFFFFFFFFFFFFull;
  for (int c = lane; c < C; c += 32) {
    float acc = 0.f;
    for (int j = 0; j < segdims; j++) {          // clusters_t is dimension-major [segdims][C]: lanes read consecutive floats
      const float d = __fsub_rn(qv[j], __ldg(clusters_t + (size_t)j * C + c));
      acc = __fadd_rn(acc, __fmul_rn(d, d));
    }
    const float dd = sqrtf(acc);
    dist[c] = dd;
    vis[c] = 0;
    const uint64_t key = ((uint64_t)__float_as_uint(dd) << 32) | (uint32_t)c;
    best = key < best ? key : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other < best ? other : best;
  }
  if (lane == 0) nearest[q] = (int32_t)(uint32_t)best;
  __syncwarp();
  int maxVisit = C;
  if (visit < 1.f) maxVisit = (int)((float)C * visit);
  int64_t seen = 0;
  int n_vis = 0;
  if (maxVisit >= C) {
    for (int c = lane; c < C; c += 32) { vis[c] = 1; seen += rule_size[c]; }
    n_vis = C;
  } else if (maxVisit > 0) {
    // smallest bit pattern d* with count(dist <= d*) >= maxVisit (distances are non-negative: patterns order like values)
    uint32_t lo = 0u, hi = 0x7F800000u;
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      int cnt = 0;
      for (int c = lane; c < C; c += 32) cnt += __float_as_uint(dist[c]) <= mid;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if (cnt >= maxVisit) hi = mid; else lo = mid + 1;
    }
    int below = 0;
    for (int c = lane; c < C; c += 32)
      if (__float_as_uint(dist[c]) < lo) { vis[c] = 1; seen += rule_size[c]; below++; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
    // clusters at exactly d*: the lowest indices fill the remaining places (ties rank by index)
    int need = maxVisit - below;
    for (int c0 = 0; c0 < C && need > 0; c0 += 32) {
      const int c = c0 + lane;
      const bool eq = c < C && __float_as_uint(dist[c]) == lo;
      const unsigned m = __ballot_sync(0xffffffffu, eq);
      if (eq && __popc(m & ((1u << lane) - 1u)) < need) { vis[c] = 1; seen += rule_size[c]; }
      need -= __popc(m);
    }
    n_vis = maxVisit;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) seen += __shfl_xor_sync(0xffffffffu, seen, o);
  __syncwarp();
  // the visit continues while fewer than k rows were covered (VAQ.cpp:1555, 1616-1618): next-nearest clusters one by one
  while (seen < k && n_vis < C) {
    uint64_t nb = 0xFFFFFFFFFFFFFFFFull;
    for (int c = lane; c < C; c += 32) {
      if (vis[c]) continue;
      const uint64_t key = ((uint64_t)__float_as_uint(dist[c]) << 32) | (uint32_t)c;
      nb = key < nb ? key : nb;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const uint64_t other = __shfl_xor_sync(0xffffffffu, nb, o);
      nb = other < nb ? other : nb;
    }
    const int c = (int)(uint32_t)nb;
    if (lane == 0) vis[c] = 1;
    __syncwarp();
    seen += rule_size[c];
    n_vis++;
  }
}

// perm[i] = query that takes slot i: queries ordered by nearest cluster (counting sort; the order inside a bucket is
// arbitrary — tile composition only affects speed, every query's answer is exact whatever its tile mates are).
__global__ void __launch_bounds__(1024) ti_order_kernel(const int32_t *__restrict__ nearest, int nq, int C, int32_t *__restrict__ perm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int32_t *hist = reinterpret_cast<int32_t *>(smem_raw);       // [C + 1]
  for (int c = threadIdx.x; c <= C; c += blockDim.x) hist[c] = 0;
  __syncthreads();
  for (int q = threadIdx.x; q < nq; q += blockDim.x) atomicAdd(hist + nearest[q], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int c = 0; c < C; c++) { const int n = hist[c]; hist[c] = acc; acc += n; }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < nq; q += blockDim.x) perm[atomicAdd(hist + nearest[q], 1)] = q;
}

// One CTA per query tile (8 slots): qperm[slot] = q_proj[perm[slot]] (slots past nq repeat the last query and visit
// nothing), tmask[tile][c] = OR_j visited[perm[8 tile + j]][c] << j.
__global__ void __launch_bounds__(256) ti_tiles_kernel(const float *__restrict__ q_proj, int nq, int D, int C,
                                                        const int32_t *__restrict__ perm, const uint8_t *__restrict__ visited,
                                                        float *__restrict__ qperm, uint8_t *__restrict__ tmask) {
  const int tile = blockIdx.x;
  __shared__ int32_t src[8];
  if (threadIdx.x < 8) src[threadIdx.x] = (tile * 8 + (int)threadIdx.x < nq) ? perm[tile * 8 + threadIdx.x] : -1;
  __syncthreads();
  const int last = perm[nq - 1];
  for (int i = threadIdx.x; i < 8 * D; i += blockDim.x) {
    const int j = i / D, d = i - j * D;
    qperm[((size_t)tile * 8 + j) * D + d] = q_proj[(size_t)(src[j] >= 0 ? src[j] : last) * D + d];
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (src[j] >= 0 && visited[(size_t)src[j] * C + c]) m |= 1u << j;
    tmask[(size_t)tile * C + c] = (uint8_t)m;
  }
}

// cluster of each 32-row tile of the cluster-grouped matrix; 0xFFFF when the tile holds rows of more than one cluster
__global__ void tile_cluster_kernel(const int64_t *__restrict__ start, int C, int64_t n_rows, uint16_t *__restrict__ tile_cl,
                                    int64_t n_tiles) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  const int c0 = cluster_of_row(start, C, t * kTileRows);
  const int c1 = cluster_of_row(start, C, min(n_rows - 1, t * kTileRows + kTileRows - 1));
  tile_cl[t] = (uint16_t)(c0 == c1 ? c0 : 0xFFFF);
}

cudaError_t launch_ti_plan(const float *q_proj, int nq, int D, const float *clusters_t, int C, int segdims, const int64_t *rule_size,
                           float visit, int k, uint8_t *visited, int32_t *nearest, int32_t *perm, float *qperm, uint8_t *tmask,
                           cudaStream_t st) {
  if (nq <= 0) return cudaSuccess;
  {
    const size_t smem = (size_t)kVisitWarps * C * sizeof(float);
    static SmemOptIn optin;
    cudaError_t e = optin.ensure(ti_visit_kernel, smem);
    if (e != cudaSuccess) return e;
    ti_visit_kernel<<<(nq + kVisitWarps - 1) / kVisitWarps, kVisitWarps * 32, smem, st>>>(q_proj, nq, D, clusters_t, C, segdims, rule_size,
                                                                                        visit, k, visited, nearest);
  }
  {
    const size_t smem = (size_t)(C + 1) * sizeof(int32_t);
    static SmemOptIn optin;
    cudaError_t e = optin.ensure(ti_order_kernel, smem);
    if (e != cudaSuccess) return e;
    ti_order_kernel<<<1, 1024, smem, st>>>(nearest, nq, C, perm);
  }
  ti_tiles_kernel<<<(nq + 7) / 8, 256, 0, st>>>(q_proj, nq, D, C, perm, visited, qperm, tmask);
  return cudaGetLastError();
}

// rot_tile[t] = row tile in which the cluster nearest to the first query of query tile t begins
__global__ void rot_tiles_kernel(const int32_t *__restrict__ nearest, const int32_t *__restrict__ perm, int n_qtiles,
                                 const int64_t *__restrict__ cl_start, int32_t *__restrict__ rot_tile) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_qtiles) rot_tile[t] = (int32_t)(cl_start[nearest[perm[8 * t]]] >> 5);
}

cudaError_t launch_rot_tiles(const int32_t *nearest, const int32_t *perm, int nq, const int64_t *cl_start, int32_t *rot_tile, cudaStream_t st) {
  const int n_qtiles = (nq + 7) / 8;
  if (n_qtiles <= 0) return cudaSuccess;
  rot_tiles_kernel<<<(n_qtiles + 255) / 256, 256, 0, st>>>(nearest, perm, n_qtiles, cl_start, rot_tile);
  return cudaGetLastError();
}

cudaError_t launch_tile_clusters(const int64_t *start, int C, int64_t n_rows, uint16_t *tile_cl, cudaStream_t st) {
  const int64_t n_tiles = (n_rows + kTileRows - 1) / kTileRows;
  if (n_tiles <= 0) return cudaSuccess;
  tile_cluster_kernel<<<(unsigned)((n_tiles + 255) / 256), 256, 0, st>>>(start, C, n_rows, tile_cl, n_tiles);
  return cudaGetLastError();
}

}  // namespace vaqgpu

namespace vaqgpu {
__global__ void transpose_kernel(const float *__restrict__ in, int rows, int cols, float *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
  out[(size_t)c * rows + r] = in[i];
}
cudaError_t launch_transpose(const float *in, int rows, int cols, float *out, cudaStream_t st) {
  const int64_t n = (int64_t)rows * cols;
  if (n <= 0) return cudaSuccess;
  transpose_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, rows, cols, out);
  return cudaGetLastError();
}
}  // namespace vaqgpu
