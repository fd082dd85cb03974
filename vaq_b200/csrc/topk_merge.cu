// Top-k list merge + result formatting.  Combines the per-CTA lists of one GPU and, after an
// all-gather, the per-GPU lists of a row-sharded index (the reference's precedent for merging
// partial answers is concatenate + sort + resize(k), BitVecEngine.cpp:1599-1611).  Lists are
// ascending 64-bit keys (distance bits << 32 | id); the result is the k smallest keys, so it is
// independent of how the rows were split (shard invariance, SURVEY.md Appendix B rule 7).
#include <float.h>
#include "common.cuh"

namespace vaqgpu {

struct MergeArgs {
  const uint64_t *in;
  int64_t stride_l, stride_q;   // element strides between lists / queries of `in`
  int32_t G, fan, nq, k;
  uint64_t *mid;                // non-final level: [nq][Gp][k]
  int32_t Gp;
  int32_t final_level, sqrt_flag, hamming;
  int32_t *ids;                 // final outputs (any may be NULL)
  void *dist;
  uint64_t *keys_out;
  const int32_t *id_map;
  int64_t id_base;
  const int32_t *qmap;          // final outputs of input query q go to row qmap[q] (NULL: q)
};

__device__ __forceinline__ void emit(const MergeArgs &a, int q, int slot, uint64_t key) {
  const size_t o = (size_t)(a.qmap ? a.qmap[q] : q) * a.k + slot;
  if (key == kEmptyKey) {
    if (a.ids) a.ids[o] = -1;
    if (a.dist) {
      if (a.hamming) reinterpret_cast<uint32_t *>(a.dist)[o] = 0xFFFFFFFFu;
      else reinterpret_cast<float *>(a.dist)[o] = FLT_MAX;
    }
    if (a.keys_out) a.keys_out[o] = kEmptyKey;
    return;
  }
  const uint32_t low = (uint32_t)key, hi = (uint32_t)(key >> 32);
  const int32_t id = a.id_map ? a.id_map[low] : (int32_t)((int64_t)low + a.id_base);
  if (a.ids) a.ids[o] = id;
  if (a.dist) {
    if (a.hamming) reinterpret_cast<uint32_t *>(a.dist)[o] = hi;
    else {
      const float d = __uint_as_float(hi);
      reinterpret_cast<float *>(a.dist)[o] = a.sqrt_flag ? sqrtf(d) : d;
    }
  }
  if (a.keys_out) a.keys_out[o] = ((uint64_t)hi << 32) | (uint32_t)id;
}

__global__ void merge_level_kernel(const __grid_constant__ MergeArgs a) {
  const int q = blockIdx.x, g = blockIdx.y;      // queries on grid.x: no 65535 limit on the batch
  const int l0 = g * a.fan, l1 = min(a.G, l0 + a.fan);
  const int nl = l1 - l0, k = a.k;
  const uint64_t *base = a.in + (size_t)q * a.stride_q;
  // pre-fill
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    if (a.final_level) emit(a, q, i, kEmptyKey);
    else a.mid[((size_t)q * a.Gp + g) * k + i] = kEmptyKey;
  }
  __syncthreads();
  const int total = nl * k;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int l = e / k, i = e - l * k;
    const uint64_t *mine = base + (size_t)(l0 + l) * a.stride_l;
    const uint64_t key = mine[i];
    if (key == kEmptyKey) continue;
    int rank = i;
    for (int o = 0; o < nl && rank < k; o++) {
      if (o == l) continue;
      rank += lower_bound_u64(base + (size_t)(l0 + o) * a.stride_l, k, key);
    }
    if (rank < k) {
      if (a.final_level) emit(a, q, rank, key);
      else a.mid[((size_t)q * a.Gp + g) * k + rank] = key;
    }
  }
}

// single list per query: pure formatting, one thread per element
__global__ void format_kernel(const __grid_constant__ MergeArgs a) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)a.nq * a.k) return;
  const int q = (int)(e / a.k), i = (int)(e - (int64_t)q * a.k);
  emit(a, q, i, a.in[(size_t)q * a.stride_q + i]);
}

// keys_in: list l of query q starts at keys_in + l*stride_l + q*stride_q.  `scratch` must hold
// 2 * nq * ceil(G/16) * k keys when G > 16 (may be NULL otherwise).
cudaError_t launch_merge_keys(const uint64_t *keys_in, int64_t stride_l, int64_t stride_q, int G, int nq, int k,
                              int sqrt_flag, int hamming, int32_t *ids, void *dist, uint64_t *keys_out,
                              const int32_t *id_map, int64_t id_base, uint64_t *scratch, cudaStream_t st, const int32_t *qmap) {
  if (nq <= 0 || k <= 0) return cudaSuccess;
  MergeArgs a{};
  a.qmap = qmap;
  a.nq = nq; a.k = k; a.sqrt_flag = sqrt_flag; a.hamming = hamming;
  a.ids = ids; a.dist = dist; a.keys_out = keys_out; a.id_map = id_map; a.id_base = id_base;
  const int kFan = 16;
  const uint64_t *in = keys_in;
  int64_t sl = stride_l, sq = stride_q;
  uint64_t *buf[2] = {scratch, nullptr};
  int which = 0;
  while (true) {
    a.in = in; a.stride_l = sl; a.stride_q = sq; a.G = G;
    if (G == 1) {
      const int threads = 256;
      const int64_t total = (int64_t)nq * k;
      format_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, st>>>(a);
      return cudaGetLastError();
    }
    const int Gp = (G + kFan - 1) / kFan;
    a.fan = kFan; a.Gp = Gp; a.final_level = (Gp == 1);
    if (!a.final_level) {
      if (!scratch) return cudaErrorInvalidValue;
      if (!buf[1]) buf[1] = scratch + (size_t)nq * ((G + kFan - 1) / kFan) * k;
      a.mid = buf[which];
    }
    const int threads = (k * min(G, kFan) >= 256) ? 256 : 64;
    merge_level_kernel<<<dim3((unsigned)nq, (unsigned)Gp), threads, 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || a.final_level) return e;
    in = a.mid; sl = k; sq = (int64_t)Gp * k; G = Gp; which ^= 1;
  }
}

}  // namespace vaqgpu
