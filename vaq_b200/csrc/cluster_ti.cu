// Device-side VAQ::clusterTI (reference VAQ.cpp:878-999): k-means over the rows DECODED in their leading subspaces
// (mTISegmentNum segments), rows regrouped by cluster, per-cluster row ranges + the original id of every regrouped row.
// The reference runs arma::kmeans on the decoded matrix (an un-vendored third-party solver: clustering parity is
// unpinned, like training); what is pinned is what a search does with the clusters, and that is exact.
//
// Nothing is decoded to memory.  A decoded row is a tuple of codebook centroids, so
//     || decode(row) - cc ||^2  =  sum_s  T[s][code_s][cluster],     T[s][c][cl] = || C_s[c] - cc_cl[sL:(s+1)L] ||^2
// and the mean of a cluster's decoded rows is  sum_s sum_c count[cl][s][c] * C_s[c] / size[cl]  — integer code
// histograms (atomicAdd on ints: order-independent), so Lloyd's iterations are deterministic.
//   ti_table_kernel      T for the current centres                       (sum_s K_s x C floats)
//   ti_assign_kernel     warp per row, lanes over clusters, argmin (lowest index on ties) + code histograms
//   ti_update_kernel     new centres from the histograms (an empty cluster keeps its centre)
//   regroup              stable counting sort of the rows by cluster: per-block histograms, column scan, in-order
//                        placement (rows keep their relative order inside a cluster -> deterministic tie order)
#include "common.cuh"

namespace vaqgpu {

namespace {

constexpr int kRegroupRows = 2048;      // rows per warp of the stable counting sort

__device__ __forceinline__ uint32_t row_code(const uint4 *__restrict__ codes, int W, const ScanLayout &lay, int64_t row, int f) {
  const uint32_t *rp = reinterpret_cast<const uint32_t *>(codes) + (((size_t)(row >> 5) * W) * kTileRows + (row & 31)) * 4;
  const uint32_t meta = lay.fmeta[f];
  const uint32_t lo = __ldg(rp + lay.fw_lo[f]), hi = __ldg(rp + lay.fw_hi[f]);
  return __funnelshift_r(lo, hi, meta & 31u) & (meta >> 16);
}

}  // namespace

// centres[cl][s*L + j] = centroid of subspace s at the code of row (cl * n / C): the reference seeds arma::kmeans with
// static_subset (VAQ.cpp:1315); evenly spaced rows are this implementation's fixed subset.
__global__ void ti_seed_kernel(const uint4 *__restrict__ codes, int W, const __grid_constant__ ScanLayout lay, int64_t n, int C, int seg, int L,
                               const float *__restrict__ cent, const int32_t *__restrict__ cent_off, float *__restrict__ centres) {
  const int cl = blockIdx.x;
  const int64_t row = (int64_t)cl * n / C;
  for (int i = threadIdx.x; i < seg * L; i += blockDim.x) {
    const int s = i / L, j = i - s * L;
    centres[(size_t)cl * seg * L + i] = cent[cent_off[s] + (size_t)row_code(codes, W, lay, row, s) * L + j];
  }
}

// T[(ent_off[s] + c) * C + cl]
__global__ void ti_table_kernel(const float *__restrict__ cent, const int32_t *__restrict__ cent_off, const int32_t *__restrict__ ent_off,
                                int seg, int L, const float *__restrict__ centres, int C, float *__restrict__ T) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)ent_off[seg] * C;
  if (i >= total) return;
  const int e = (int)(i / C), cl = (int)(i - (int64_t)e * C);
  int s = 0;
  while (ent_off[s + 1] <= e) s++;
  const int c = e - ent_off[s];
  const float *cp = cent + cent_off[s] + (size_t)c * L;
  const float *cc = centres + (size_t)cl * seg * L + (size_t)s * L;
  float acc = 0.f;
  for (int j = 0; j < L; j++) {
    const float d = cp[j] - cc[j];
    acc = fmaf(d, d, acc);
  }
  T[i] = acc;
}

// one warp per row: assign[row] = argmin_cl sum_s T[s][code_s][cl]; hist[(cl * ent + ent_off[s] + code_s)]++ when hist != NULL
__global__ void __launch_bounds__(256) ti_assign_kernel(const uint4 *__restrict__ codes, int W, const __grid_constant__ ScanLayout lay, int64_t n,
                                                         int C, int seg, const int32_t *__restrict__ ent_off, const float *__restrict__ T,
                                                         int32_t *__restrict__ assign, int32_t *__restrict__ hist, int32_t *__restrict__ sizes) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  __shared__ int32_t s_code[8][kMaxSubspaces];
  int32_t *mycodes = s_code[threadIdx.x >> 5];
  for (int s = lane; s < seg; s += 32) mycodes[s] = ent_off[s] + (int32_t)row_code(codes, W, lay, row, s);
  __syncwarp();
  uint64_t best = 0xFFFFFFFFFFFFFFFFull;
  for (int cl = lane; cl < C; cl += 32) {
    float d = 0.f;
    for (int s = 0; s < seg; s++) d += __ldg(T + (size_t)mycodes[s] * C + cl);
    const uint64_t key = ((uint64_t)__float_as_uint(d) << 32) | (uint32_t)cl;      // d >= 0: bit patterns order like values
    best = key < best ? key : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other < best ? other : best;
  }
  const int cl = (int)(uint32_t)best;
  if (lane == 0) {
    assign[row] = cl;
    if (sizes) atomicAdd(sizes + cl, 1);
  }
  if (hist) {
    const int ent = ent_off[seg];
    for (int s = lane; s < seg; s += 32) atomicAdd(hist + (size_t)cl * ent + mycodes[s], 1);
  }
}

// centres[cl][s*L + j] = sum_c hist[cl][s][c] * C_s[c][j] / sizes[cl]  (double accumulation in code order: deterministic)
__global__ void ti_update_kernel(const float *__restrict__ cent, const int32_t *__restrict__ cent_off, const int32_t *__restrict__ ent_off,
                                 int seg, int L, const int32_t *__restrict__ hist, const int32_t *__restrict__ sizes, int C,
                                 float *__restrict__ centres) {
  const int cl = blockIdx.x;
  const int n = sizes[cl];
  if (n == 0) return;
  const int ent = ent_off[seg];
  for (int i = threadIdx.x; i < seg * L; i += blockDim.x) {
    const int s = i / L, j = i - s * L;
    const int K = ent_off[s + 1] - ent_off[s];
    const int32_t *hs = hist + (size_t)cl * ent + ent_off[s];
    double acc = 0.0;
    for (int c = 0; c < K; c++) acc += (double)hs[c] * (double)cent[cent_off[s] + (size_t)c * L + j];
    centres[(size_t)cl * seg * L + i] = (float)(acc / (double)n);
  }
}

// ---- stable regroup: rows sorted by cluster, original order kept inside a cluster ---------------------------------------
// pass 1: per-block (kRegroupRows rows, one warp) cluster histogram -> bh[block][C]
__global__ void __launch_bounds__(128) regroup_hist_kernel(const int32_t *__restrict__ assign, int64_t n, int C, int32_t *__restrict__ bh) {
  const int lane = threadIdx.x & 31;
  const int64_t blk = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t r0 = blk * kRegroupRows;
  if (r0 >= n) return;
  int32_t *mine = bh + (size_t)blk * C;
  const int64_t r1 = min(n, r0 + kRegroupRows);
  for (int64_t r = r0; r < r1; r += 32) {
    const bool ok = r + lane < r1;
    const int cl = ok ? assign[r + lane] : -1 - lane;
    const unsigned m = __match_any_sync(0xffffffffu, cl);
    if (ok && lane == __ffs(m) - 1) mine[cl] += __popc(m);      // only this warp touches its histogram row
    __syncwarp();
  }
}

// pass 2: one thread per cluster: start[cl] is given; bh[block][cl] becomes the first position of that block's rows
__global__ void regroup_scan_kernel(int32_t *__restrict__ bh, int64_t n_blocks, int C, const int64_t *__restrict__ start) {
  const int cl = blockIdx.x * blockDim.x + threadIdx.x;
  if (cl >= C) return;
  int64_t acc = start[cl];
  for (int64_t b = 0; b < n_blocks; b++) {
    const int32_t c = bh[(size_t)b * C + cl];
    bh[(size_t)b * C + cl] = (int32_t)acc;
    acc += c;
  }
}

// pass 3: placement in row order; the packed words of a row move with it, id_map[new position] = old row
__global__ void __launch_bounds__(128) regroup_place_kernel(const int32_t *__restrict__ assign, int64_t n, int C, int32_t *__restrict__ bh,
                                                            const uint4 *__restrict__ src, uint4 *__restrict__ dst, int W,
                                                            int32_t *__restrict__ id_map) {
  const int lane = threadIdx.x & 31;
  const int64_t blk = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t r0 = blk * kRegroupRows;
  if (r0 >= n) return;
  int32_t *mine = bh + (size_t)blk * C;
  const int64_t r1 = min(n, r0 + kRegroupRows);
  for (int64_t r = r0; r < r1; r += 32) {
    const int64_t row = r + lane;
    const bool ok = row < r1;
    const int cl = ok ? assign[row] : -1 - lane;
    const unsigned m = __match_any_sync(0xffffffffu, cl);
    int32_t base = 0;
    if (ok) base = mine[cl];
    __syncwarp();
    if (ok && lane == __ffs(m) - 1) mine[cl] = base + __popc(m);
    __syncwarp();
    if (ok) {
      const int64_t pos = (int64_t)base + __popc(m & ((1u << lane) - 1u));
      id_map[pos] = (int32_t)row;
      for (int j = 0; j < W; j++)
        dst[((size_t)(pos >> 5) * W + j) * kTileRows + (pos & 31)] = src[((size_t)(row >> 5) * W + j) * kTileRows + (row & 31)];
    }
  }
}

// exclusive scan of the cluster sizes -> start (C is small: one thread)
__global__ void ti_starts_kernel(const int32_t *__restrict__ sizes, int C, int64_t *__restrict__ start, int64_t *__restrict__ size64) {
  if (blockIdx.x || threadIdx.x) return;
  int64_t acc = 0;
  for (int c = 0; c < C; c++) { start[c] = acc; size64[c] = sizes[c]; acc += sizes[c]; }
}

size_t cluster_ti_table_floats(const LutPlan &plan, int seg, int C) { return (size_t)plan.ent_off[seg] * C; }

// k-means (iters Lloyd iterations from evenly spaced rows) + final assignment.  Workspaces: T and hist hold
// ent_off[seg] * C elements each, assign n, sizes C.  On return centres / assign / sizes describe the final clustering.
cudaError_t launch_cluster_ti_kmeans(const uint4 *codes, const ScanLayout &lay, int64_t n, const LutPlan &plan, int seg,
                                     const float *cent, const int32_t *d_cent_off, const int32_t *d_ent_off, int C, int iters,
                                     float *centres, float *T, int32_t *hist, int32_t *assign, int32_t *sizes, cudaStream_t st) {
  const int L = plan.L;
  const int64_t tbl = (int64_t)plan.ent_off[seg] * C;
  ti_seed_kernel<<<C, 128, 0, st>>>(codes, lay.W, lay, n, C, seg, L, cent, d_cent_off, centres);
  for (int it = 0; it <= iters; it++) {
    const bool last = it == iters;
    ti_table_kernel<<<(unsigned)((tbl + 255) / 256), 256, 0, st>>>(cent, d_cent_off, d_ent_off, seg, L, centres, C, T);
    cudaError_t e = cudaMemsetAsync(sizes, 0, (size_t)C * sizeof(int32_t), st);
    if (e != cudaSuccess) return e;
    if (!last) {
      e = cudaMemsetAsync(hist, 0, (size_t)tbl * sizeof(int32_t), st);
      if (e != cudaSuccess) return e;
    }
    ti_assign_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(codes, lay.W, lay, n, C, seg, d_ent_off, T, assign, last ? nullptr : hist, sizes);
    if (!last) ti_update_kernel<<<C, 128, 0, st>>>(cent, d_cent_off, d_ent_off, seg, L, hist, sizes, C, centres);
  }
  return cudaGetLastError();
}

size_t regroup_hist_ints(int64_t n, int C) { return (size_t)((n + kRegroupRows - 1) / kRegroupRows) * C; }

cudaError_t launch_regroup(const int32_t *assign, const int32_t *sizes, int64_t n, int C, const uint4 *src, uint4 *dst, int W,
                           int32_t *id_map, int64_t *start, int64_t *size64, int32_t *bh, cudaStream_t st) {
  const int64_t n_blocks = (n + kRegroupRows - 1) / kRegroupRows;
  cudaError_t e = cudaMemsetAsync(bh, 0, (size_t)n_blocks * C * sizeof(int32_t), st);
  if (e != cudaSuccess) return e;
  ti_starts_kernel<<<1, 32, 0, st>>>(sizes, C, start, size64);
  regroup_hist_kernel<<<(unsigned)((n_blocks + 3) / 4), 128, 0, st>>>(assign, n, C, bh);
  regroup_scan_kernel<<<(C + 127) / 128, 128, 0, st>>>(bh, n_blocks, C, start);
  regroup_place_kernel<<<(unsigned)((n_blocks + 3) / 4), 128, 0, st>>>(assign, n, C, bh, src, dst, W, id_map);
  return cudaGetLastError();
}

}  // namespace vaqgpu
