// ADC scan — replaces the per-query row loops of VAQ::searchHeap (reference
// bitvecengine/VAQ.cpp:1729-1758), VAQ::searchEarlyAbandon (:1694-1727) and, through
// per-query row ranges, the cluster-ordered scan of VAQ::searchTriangleInequality
// (:1540-1692).
//
// One CTA = one query x one share of the rows.  The query's lookup tables are staged
// once into shared memory with a 1-D TMA bulk copy (tables that do not fit stay in
// global memory and are read through L1/L2).  Each warp streams 32-row tiles of the
// bit-packed code matrix with coalesced 128-bit loads (layout [tile][word][lane]), one
// row per lane, next tile prefetched into registers while the current one is scored.
// A row's distance is accumulated in the reference's order and grouping —
// dism = ((l0+l1)+l2)+l3 ; dist += dism over subspaces in variance-descending order —
// and after every group of four the warp votes: if every lane's partial distance already
// exceeds the running k-th best, the tile is abandoned (warp-uniform early abandon; the
// reference abandons per row on dist >= bsfK, VAQ.cpp:1708).  Survivors are inserted
// into a per-warp sorted top-k list in shared memory keyed by (distance bits, row);
// the smallest k-th key over the CTA's warps is shared so every warp prunes with the
// tightest bound.  At the end the warp lists are merged and the CTA writes its k best
// keys; launch_merge_keys combines the CTAs (and, across GPUs, the shards).
//
// Tie rule: the result is the k lexicographically smallest (distance, row) pairs, which
// is what the reference's strict `heap_dis[0] > dist` insertion (VAQ.cpp:1718,1750)
// yields whenever distances are distinct; see DESIGN.md.
#include "common.cuh"

namespace vaqgpu {

template <int W>
__device__ __forceinline__ void load_tile(uint4 (&dst)[W], const uint4 *__restrict__ codes, int tile, int lane) {
  const uint4 *p = codes + ((size_t)tile * W) * kTileRows + lane;
#pragma unroll
  for (int j = 0; j < W; j++) dst[j] = ldg_stream_u4(p + j * kTileRows);
}

// Scores one tile.  Returns false if the whole warp abandoned (no lane can qualify).
template <int W>
__device__ __forceinline__ bool score_tile(const uint4 (&cw)[W], const AdcScanArgs &a, const float *__restrict__ slut,
                                           const float *__restrict__ gspill, float thr, bool lane_dead, bool ea,
                                           float &dist_out) {
  uint32_t wd[4 * W + 1];
#pragma unroll
  for (int j = 0; j < W; j++) {
    wd[4 * j + 0] = cw[j].x; wd[4 * j + 1] = cw[j].y; wd[4 * j + 2] = cw[j].z; wd[4 * j + 3] = cw[j].w;
  }
  wd[4 * W] = 0u;
  float dist = 0.f, dism = 0.f;
  int f = 0;
#pragma unroll
  for (int w = 0; w < 4 * W; w++) {
    const int fe = a.lay.fbeg[w + 1];
    const uint32_t lo = wd[w], hi = wd[w + 1];
    for (; f < fe; f++) {
      const uint32_t meta = a.lay.fmeta[f];
      const uint32_t off = a.lay.foff[f];
      const uint32_t code = __funnelshift_r(lo, hi, meta & 31u) & (meta >> 16);
      const float v = (meta & kFieldSpill) ? __ldg(gspill + off + code) : slut[off + code];
      dism += v;
      if ((f & 3) == 3) {
        dist += dism;
        dism = 0.f;
        if (ea && __all_sync(0xffffffffu, lane_dead || (dist > thr))) return false;
      }
    }
  }
  if (a.lay.M & 3) dist += dism;
  dist_out = dist;
  return true;
}

template <int W>
__global__ void adc_scan_kernel(const __grid_constant__ AdcScanArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int k = a.k;
  const int q = blockIdx.y, split = blockIdx.x;

  float *slut = reinterpret_cast<float *>(smem_raw);
  uint64_t *lists = reinterpret_cast<uint64_t *>(smem_raw + (((size_t)a.smem_lut_floats * 4 + 15) & ~(size_t)15));
  uint64_t *merged = lists + (size_t)nwarps * k;
  uint64_t *blk_thr = merged + k;
  uint64_t *bar = blk_thr + 1;

  const float *glut = a.lut + (size_t)q * a.lut_stride;
  const float *gspill = glut + a.smem_lut_floats;

  for (int i = tid; i < nwarps * k; i += blockDim.x) lists[i] = kEmptyKey;
  if (tid == 0) *blk_thr = kEmptyKey;

  const uint32_t lut_bytes = (uint32_t)a.smem_lut_floats * 4u;
  if (a.use_tma && lut_bytes) {
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (tid == 0) {
      mbar_arrive_expect_tx(bar, lut_bytes);
      for (uint32_t off = 0; off < lut_bytes; off += 32768u) {
        const uint32_t n = min(32768u, lut_bytes - off);
        tma_bulk_g2s(reinterpret_cast<unsigned char *>(slut) + off, reinterpret_cast<const unsigned char *>(glut) + off, n, bar);
      }
    }
    mbar_wait(bar, 0);
  } else {
    for (int i = tid * 4; i < a.smem_lut_floats; i += blockDim.x * 4)
      *reinterpret_cast<float4 *>(slut + i) = __ldg(reinterpret_cast<const float4 *>(glut + i));
    __syncthreads();
  }

  volatile uint64_t *mylist = lists + (size_t)warp * k;
  const int g = split * nwarps + warp, gstride = a.splits * nwarps;
  const bool ea = a.early_abandon != 0;
  const int nr = a.ranges ? a.n_ranges[q] : 1;

  for (int r = 0; r < nr; r++) {
    int rb = 0, re = (int)a.n_rows;
    if (a.ranges) { const int2 rg = a.ranges[(size_t)q * a.max_ranges + r]; rb = rg.x; re = rg.y; }
    if (re <= rb) continue;
    const int tb = rb >> 5, te = (re + kTileRows - 1) >> 5;
    int t = tb + g;
    uint4 cur[W], nxt[W];
    if (t < te) load_tile<W>(cur, a.codes, t, lane);
    for (; t < te; t += gstride) {
      const int tn = t + gstride;
      if (tn < te) load_tile<W>(nxt, a.codes, tn, lane);

      uint64_t thrkey = mylist[k - 1];
      const uint64_t bthr = *reinterpret_cast<volatile uint64_t *>(blk_thr);
      thrkey = bthr < thrkey ? bthr : thrkey;
      // empty key -> NaN threshold -> (dist > thr) is false -> never abandons
      const float thr = __uint_as_float((uint32_t)(thrkey >> 32));
      const int row = t * kTileRows + lane;
      const bool lane_dead = (row < rb) || (row >= re);
      float dist = 0.f;
      const bool alive = score_tile<W>(cur, a, slut, gspill, thr, lane_dead, ea, dist);
      if (alive) {
        const uint64_t key = lane_dead ? kEmptyKey : make_key_f32(dist, a.rowid ? (int32_t)__ldg(a.rowid + row) : row);
        unsigned m = __ballot_sync(0xffffffffu, key < thrkey);
        if (m) {
          uint64_t kth = thrkey;
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const uint64_t kk = __shfl_sync(0xffffffffu, key, src);
            kth = warp_list_insert(mylist, k, kk, lane);
          }
          if (lane == 0 && kth != kEmptyKey) atomicMin(reinterpret_cast<unsigned long long *>(blk_thr), (unsigned long long)kth);
        }
      }
#pragma unroll
      for (int j = 0; j < W; j++) cur[j] = nxt[j];
    }
  }

  __syncthreads();
  block_merge_lists(lists, nwarps, k, merged);
  uint64_t *out = a.out_keys + ((size_t)q * a.splits + split) * k;
  for (int i = tid; i < k; i += blockDim.x) out[i] = merged[i];
}

template <int W>
static cudaError_t launch_w(const AdcScanArgs &a, int threads, size_t smem_bytes, cudaStream_t st) {
  static SmemOptIn optin;
  {
    cudaError_t e = optin.ensure(adc_scan_kernel<W>, smem_bytes);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((unsigned)a.splits, (unsigned)a.nq);
  adc_scan_kernel<W><<<grid, threads, smem_bytes, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_adc_scan(const AdcScanArgs &a, int threads, size_t smem_bytes, cudaStream_t st) {
  switch (a.lay.W) {
    case 1: return launch_w<1>(a, threads, smem_bytes, st);
    case 2: return launch_w<2>(a, threads, smem_bytes, st);
    case 3: return launch_w<3>(a, threads, smem_bytes, st);
    case 4: return launch_w<4>(a, threads, smem_bytes, st);
    case 5: return launch_w<5>(a, threads, smem_bytes, st);
    case 6: return launch_w<6>(a, threads, smem_bytes, st);
    case 7: return launch_w<7>(a, threads, smem_bytes, st);
    case 8: return launch_w<8>(a, threads, smem_bytes, st);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t adc_scan_occupancy(int W, int threads, size_t smem_bytes, int *ctas_per_sm) {
  cudaError_t e;
#define OCC(WW)                                                                                               \
  case WW:                                                                                                    \
    e = cudaFuncSetAttribute(adc_scan_kernel<WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes); \
    if (e != cudaSuccess) return e;                                                                           \
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, adc_scan_kernel<WW>, threads, smem_bytes);
  switch (W) {
    OCC(1) OCC(2) OCC(3) OCC(4) OCC(5) OCC(6) OCC(7) OCC(8)
    default: return cudaErrorInvalidValue;
  }
#undef OCC
}

}  // namespace vaqgpu
