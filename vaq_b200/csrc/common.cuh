// Shared device/host definitions for the vaqgpu kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vaqgpu {

constexpr int kMaxSubspaces = 128;   // M
constexpr int kMaxRowWords = 32;     // 32-bit words per packed row (<= 1024 bits)
constexpr int kTileRows = 32;        // rows per tile == warp width
constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;
constexpr int kDbgSlots = 24;        // development: per-CTA slots of the filter kernel's phase clocks / event counts
constexpr int kLayoutWin = 4096;     // rows per window of the conflict-aware row order (layout.cu); a multiple of kTileRows

// fmeta bit layout: [4:0] shift inside the starting 32-bit word, [5] table spilled to
// global/L2, [31:16] value mask ((1<<bits)-1).
constexpr uint32_t kFieldSpill = 1u << 5;

// Where each subspace code sits in the packed row, and where its LUT lives.
// Rows are little-endian bit strings: subspace s occupies bits [bitoff_s, bitoff_s+bits_s),
// bitoff_s = sum_{t<s} bits_t; bit i lives in 32-bit word i/32 at position i%32.
struct ScanLayout {
  int32_t M;                          // subspaces
  int32_t W;                          // uint4 words per row
  uint16_t fbeg[kMaxRowWords + 1];    // fields starting in word w: [fbeg[w], fbeg[w+1])
  uint32_t fmeta[kMaxSubspaces];      // see above
  uint32_t foff[kMaxSubspaces];       // entry offset of table s inside the smem LUT or the spill area
  uint8_t fword[kMaxSubspaces];       // 32-bit word in which field s starts
  uint16_t fw_lo[kMaxSubspaces];      // offset (32-bit units, from the row's first word in the tiled layout) of that
  uint16_t fw_hi[kMaxSubspaces];      //   word and of the following one (== fw_lo when the row ends there)
};

// Where the LUT build kernel writes table s inside a query's LUT row.
struct LutPlan {
  int32_t M, L;
  int32_t T;                          // query-tile interleave: entry e of query q lives at
                                      //   ((q / T) * row_stride + pos + e) * T + q % T
  int32_t total_entries;              // sum K_s
  int32_t row_stride;                 // floats per query row in the LUT workspace
  int32_t ent_off[kMaxSubspaces + 1]; // compact entry offsets (prefix sum of K_s)
  int32_t pos[kMaxSubspaces];         // float position of table s in the workspace row
  int32_t cent_off[kMaxSubspaces];    // float offset of centroid block s
};

__host__ __device__ inline uint64_t make_key_f32(float d, int32_t id) {
#ifdef __CUDA_ARCH__
  return ((uint64_t)__float_as_uint(d) << 32) | (uint32_t)id;
#else
  union { float f; uint32_t u; } c; c.f = d;
  return ((uint64_t)c.u << 32) | (uint32_t)id;
#endif
}


// Opt-in dynamic shared memory (> 48 KB) is a per-DEVICE attribute of a kernel function: a handle on GPU 1 of the
// same process needs its own cudaFuncSetAttribute.  One instance per kernel instantiation; remembers the largest
// size configured on each device (relaxed atomics: racing host threads at worst repeat an idempotent call).
struct SmemOptIn {
  static constexpr int kMaxDevices = 64;
  unsigned long long per_dev[kMaxDevices] = {};
  template <typename K>
  cudaError_t ensure(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= kMaxDevices) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if ((unsigned long long)bytes <= __atomic_load_n(&per_dev[dev], __ATOMIC_RELAXED)) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    unsigned long long cur = __atomic_load_n(&per_dev[dev], __ATOMIC_RELAXED);
    while (cur < bytes && !__atomic_compare_exchange_n(&per_dev[dev], &cur, (unsigned long long)bytes, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return cudaSuccess;
  }
};

#ifdef __CUDACC__

__device__ __forceinline__ uint4 ldg_stream_u4(const uint4 *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// L2 prefetch of the line holding p: hides DRAM latency of a streaming scan several tiles ahead at no register cost
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier + 1-D TMA bulk copy (global -> shared) -------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// bytes must be a multiple of 16, src/dst 16-byte aligned
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- per-warp sorted top-k list in shared memory --------------------------------------
// list[0..k) ascending; insert `key` (warp-uniform) if it is smaller than list[k-1].
// All 32 lanes must call.  Returns the (possibly new) k-th key.
__device__ __forceinline__ uint64_t warp_list_insert(volatile uint64_t *list, int k, uint64_t key, int lane) {
  uint64_t kth = list[k - 1];
  if (!(key < kth)) return kth;
  int cnt = 0;
  for (int i = lane; i < k; i += 32) cnt += (list[i] < key) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  const int rank = cnt;
  for (int base = ((k - 1) >> 5) << 5; base >= 0; base -= 32) {
    const int i = base + lane;
    uint64_t v = 0;
    const bool mv = (i < k) && (i > rank);
    if (mv) v = list[i - 1];
    __syncwarp();
    if (mv) list[i] = v;
    else if (i == rank) list[i] = key;
    __syncwarp();
    if (base <= rank) break;   // everything below `rank` is unchanged
  }
  return list[k - 1];
}

// Cluster of a row of the cluster-grouped matrix: the last cluster whose first row is <= row (starts ascend; an
// empty cluster shares its start with its successor and is never returned for a real row).
__device__ __forceinline__ int cluster_of_row(const int64_t *__restrict__ start, int C, int64_t row) {
  int lo = 0, hi = C;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (start[mid] <= row) lo = mid + 1; else hi = mid;
  }
  return lo - 1;
}

// count of entries < key in an ascending list (binary search)
__device__ __forceinline__ int lower_bound_u64(const uint64_t *a, int n, uint64_t key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Merge `nlists` ascending lists of k keys (contiguous: lists[l*k + i]) into the k smallest,
// written ascending to out[0..k).  Keys are unique except kEmptyKey padding.  Block-collective.
__device__ __forceinline__ void block_merge_lists(const uint64_t *lists, int nlists, int k, uint64_t *out) {
  for (int i = threadIdx.x; i < k; i += blockDim.x) out[i] = kEmptyKey;
  __syncthreads();
  const int total = nlists * k;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const uint64_t key = lists[e];
    if (key == kEmptyKey) continue;
    const int l = e / k, i = e - l * k;
    int rank = i;
    for (int o = 0; o < nlists && rank < k; o++) {
      if (o == l) continue;
      rank += lower_bound_u64(lists + o * k, k, key);
    }
    if (rank < k) out[rank] = key;
  }
  __syncthreads();
}

#endif  // __CUDACC__

// ---- cross-shard bound exchange ----------------------------------------------------------
// Row-sharded search: every shard owns a per-query array of the best k-th distance proven so far (float bits,
// 0xFFFFFFFF = none).  A scan publishes each new bound into its own array AND, through NVLink peer memory, into the
// arrays of the other shards, so every GPU prunes with the tightest bound found anywhere on the box.  Exact: a
// shard's k-th best distance is an upper bound of the k-th best over all shards, and a row of the global top-k is in
// its own shard's top-k with a distance <= that bound, so no shard ever drops it.
constexpr int kMaxPeers = 15;
struct PeerBounds {
  uint32_t *p[kMaxPeers];    // the same array on the other shards (peer-mapped device pointers)
  int32_t n;
};

#ifdef __CUDACC__
__device__ __forceinline__ void publish_global_bound(uint32_t *own, const PeerBounds &peers, int q, uint32_t bits) {
  if (atomicMin(own + q, bits) > bits) {            // only a bound that is new here travels
    for (int i = 0; i < peers.n; i++) atomicMin(peers.p[i] + q, bits);      // fire-and-forget RED.MIN over NVLink
  }
}
#endif

// ---- launchers (one per .cu) ----------------------------------------------------------
struct AdcScanArgs {
  const uint4 *codes;        // packed tiles [tile][w][lane]
  int64_t n_rows;            // valid rows
  const float *lut;          // LUT workspace [nq][row_stride]
  int32_t lut_stride;        // floats
  int32_t smem_lut_floats;   // resident floats (multiple of 4)
  int32_t nq, k, splits;
  int32_t early_abandon;     // 1 = EA votes on
  int32_t use_tma;           // 1 = cp.async.bulk LUT staging
  uint64_t *out_keys;        // [nq][splits][k]; low word = LOCAL row index
  // TI / visit mode: per-query list of row ranges (NULL = whole index)
  const int2 *ranges;        // [nq][max_ranges] (row_begin, row_end) in visiting order
  const int32_t *n_ranges;   // [nq]
  int32_t max_ranges;
  const uint32_t *rowid;     // original row index of each storage row (layout.cu), or NULL = identity
  ScanLayout lay;
};

cudaError_t launch_adc_scan(const AdcScanArgs &a, int threads, size_t smem_bytes, cudaStream_t st);
cudaError_t adc_scan_occupancy(int W, int threads, size_t smem_bytes, int *ctas_per_sm);

// filter-and-refine scan (adc_filter_scan.cu)
struct AdcFilterArgs {
  const uint4 *codes;        // packed tiles [tile][w][lane]
  int64_t n_rows;
  const float *lut;          // interleaved LUT workspace [query tile][row_stride][T]
  int32_t lut_stride;        // entries per query (row_stride)
  int32_t smem_lut_floats;   // resident entries per query (multiple of 4)
  int32_t nq, k;
  int64_t tile_lo, tile_hi;  // 32-row tiles scanned by this launch
  int32_t chunk_tiles;       // tiles per row chunk (grid.y = ceil((tile_hi - tile_lo) / chunk_tiles))
  int32_t out_slots;         // key lists per query in out_keys (over all launches of the search)
  int32_t slot_base;         // this launch's first list
  uint64_t *out_keys;        // [nq][out_slots][k]; low word = LOCAL row index
  uint32_t *thr_global;      // [nq] float bits of the best known k-th distance (0xFFFFFFFF = none)
  PeerBounds peers;          // the same array on the other row shards (n = 0: single shard)
  int32_t seed;              // 1 = CTAs seed their bounds from sample rows of their chunk
  const uint32_t *rowid;     // original row index of each storage row (layout.cu), or NULL = identity
  ScanLayout lay;
};
size_t adc_filter_smem_bytes(int smem_lut_floats, int T, int k, int threads);
cudaError_t launch_adc_filter_scan(const AdcFilterArgs &a, int T, int threads, size_t smem_bytes, cudaStream_t st);
cudaError_t launch_fill_u32(uint32_t *p, int n, uint32_t v, cudaStream_t st);

// filter-and-refine scan on fp16 lower-bound tables, query tiles of 8 (adc_filter16_scan.cu)
struct AdcFilter16Args {
  const uint4 *codes;
  int64_t n_rows;
  const void *lut16;         // [query tile][lut_stride][8] fp16: round-toward-zero(scale * entry)
  const float *lut32;        // [query tile][lut_stride][8] fp32: the exact tables (read through L2 by the exact level)
  const float *scale;        // [padded nq] power-of-two scale per query
  int32_t lut_stride;        // entries per query
  int32_t nq, k;
  int64_t tile_lo, tile_hi;
  int32_t chunk_tiles, out_slots, slot_base;
  uint64_t *out_keys;
  uint32_t *thr_global;
  PeerBounds peers;          // the same array on the other row shards (n = 0: single shard)
  int32_t seed;              // 1 = CTAs seed their bounds from sample rows of their chunk
  int32_t q3_cap;            // rows the exact level waits for at most (1..32)
  int32_t seed_rows;         // sample rows per lane of the bound seeding (at most a quarter of the chunk in total)
  const uint32_t *rowid;     // original row index of each storage row (layout.cu), or NULL = identity
  int32_t chunks_fast;       // grid order: 0 = query tiles fastest (default), 1 = row chunks fastest
  int32_t l2_prefetch;       // > 0: stage 1 prefetches into L2 this many iterations ahead (code matrix >> L2)
  // TI / visit (NULL / 0 otherwise): cluster of each 32-row tile (0xFFFF = straddles clusters), first row of each
  // cluster, and per (query tile, cluster) the 8-bit mask of the tile's queries that visit the cluster
  const uint16_t *tile_cl;
  const int64_t *cl_start;
  const uint8_t *tmask;
  const int32_t *qmap;       // queries re-grouped into tiles (TI, scan order): position in the caller's batch of each tile slot
                             // (bound arrays are indexed by it); NULL = identity
  const int32_t *rot_tile;   // [query tiles] row tile at which this query tile's scan starts (scan order, layout.cu), or NULL
  int32_t C;
  long long *dbg;            // development: per-CTA phase clocks (NULL = off)
  ScanLayout lay;
};
size_t adc_filter16_smem_bytes(int lut_stride, int k, int threads, int ti_clusters = 0);
cudaError_t launch_adc_filter16_scan(const AdcFilter16Args &a, int threads, size_t smem_bytes, cudaStream_t st);

// nq_launch >= nq query slots are written (tile padding repeats the last query).  With plan.T == 8 and
// lut16 != NULL the fp16 lower-bound tables and the per-query scales are written by the same kernel.
cudaError_t launch_lut_build(const float *q_proj, int nq, int nq_launch, int D, const float *centroids, const float *cent_rmax,
                             const LutPlan &plan, float *lut, void *lut16, float *scale, cudaStream_t st);
cudaError_t launch_project(const float *x, int n, int D, const float *eig, float *out, cudaStream_t st);

cudaError_t launch_pack_codes(const uint16_t *codes, int64_t n, int64_t row0, const ScanLayout &lay, uint4 *packed,
                              cudaStream_t st);
// Unpacks the ORIGINAL rows [row0, row0+n): storage rows [srow_lo, srow_hi) are visited and row s goes to
// codes[rowid[s] - row0] when that is in range (rowid == NULL: identity).
cudaError_t launch_unpack_codes(const uint4 *packed, int64_t row0, int64_t n, const ScanLayout &lay, uint16_t *codes,
                                const uint32_t *rowid, int64_t srow_lo, int64_t srow_hi, cudaStream_t st);
// conflict-aware row order (layout.cu)
cudaError_t launch_layout(uint4 *codes, int64_t row_lo, int64_t n_rows, const ScanLayout &lay, uint32_t *rowid, uint16_t *src,
                          uint4 *scratch, int scratch_ctas, const int64_t *win_tab, int64_t n_windows, cudaStream_t st);
cudaError_t launch_layout_restore(const uint4 *tmp_copy, uint4 *codes, int W, const uint32_t *rowid, int64_t n, cudaStream_t st);
size_t layout_scratch_bytes(int W, int ctas);
cudaError_t launch_iota_u32(uint32_t *p, int64_t lo, int64_t hi, cudaStream_t st);
cudaError_t launch_encode(const float *x_proj, int64_t n, const float *centroids, const LutPlan &plan,
                          uint16_t *codes, cudaStream_t st);
cudaError_t launch_synth_codes(uint16_t *codes, int64_t n, int64_t global_row0, int M, const int32_t *bits,
                               const float *cdf, const int32_t *ent_off, uint64_t seed, cudaStream_t st);

// List l of query q starts at keys_in + l*stride_l + q*stride_q (G ascending lists of k keys per
// query).  Writes the k smallest per query.  Low words are remapped on output:
// id = id_map ? id_map[low] : low + id_base.  Any of ids / dist / keys_out may be NULL.
// `scratch` (2 * nq * ceil(G/16) * k keys) is needed only when G > 16.
cudaError_t launch_merge_keys(const uint64_t *keys_in, int64_t stride_l, int64_t stride_q, int G, int nq, int k,
                              int sqrt_flag, int hamming, int32_t *ids, void *dist, uint64_t *keys_out,
                              const int32_t *id_map, int64_t id_base, uint64_t *scratch, cudaStream_t st,
                              const int32_t *qmap = nullptr);     // qmap: output slot of input query q (TI tile order)

struct HamScanArgs {
  const uint4 *codes;        // packed tiles [tile][w][lane]
  int64_t n_rows;
  int32_t W;                 // uint4 per row
  const uint4 *queries;      // [nq][W]
  int32_t nq, k, splits, qt; // qt = queries per CTA
  uint64_t *out_keys;        // [nq][splits][k]; low word = LOCAL row index
};
cudaError_t launch_ham_scan(const HamScanArgs &a, int threads, size_t smem_bytes, cudaStream_t st);
cudaError_t launch_ham_pack(const uint64_t *words, int64_t n, int64_t row0, int w64, int W, uint4 *packed,
                            cudaStream_t st);
cudaError_t launch_ham_synth(uint4 *packed, int64_t n, int64_t row0, int64_t global_row0, int nbits, int W,
                             uint64_t seed, cudaStream_t st);

cudaError_t launch_refine(const float *xtrain, int64_t n, int D, const float *queries, int nq,
                          const int32_t *in_labels, int refine_num, int k, int32_t *labels, float *dists,
                          cudaStream_t st);
// TI / visit planning for the filter kernels (ti_plan.cu)
cudaError_t launch_transpose(const float *in, int rows, int cols, float *out, cudaStream_t st);
cudaError_t launch_ti_plan(const float *q_proj, int nq, int D, const float *clusters_t, int C, int segdims, const int64_t *rule_size,
                           float visit, int k, uint8_t *visited, int32_t *nearest, int32_t *perm, float *qperm, uint8_t *tmask,
                           cudaStream_t st);
cudaError_t launch_tile_clusters(const int64_t *start, int C, int64_t n_rows, uint16_t *tile_cl, cudaStream_t st);
// device-side clusterTI (cluster_ti.cu)
cudaError_t launch_rot_tiles(const int32_t *nearest, const int32_t *perm, int nq, const int64_t *cl_start, int32_t *rot_tile, cudaStream_t st);
size_t cluster_ti_table_floats(const LutPlan &plan, int seg, int C);
size_t regroup_hist_ints(int64_t n, int C);
cudaError_t launch_cluster_ti_kmeans(const uint4 *codes, const ScanLayout &lay, int64_t n, const LutPlan &plan, int seg,
                                     const float *cent, const int32_t *d_cent_off, const int32_t *d_ent_off, int C, int iters,
                                     float *centres, float *T, int32_t *hist, int32_t *assign, int32_t *sizes, cudaStream_t st);
cudaError_t launch_regroup(const int32_t *assign, const int32_t *sizes, int64_t n, int C, const uint4 *src, uint4 *dst, int W,
                           int32_t *id_map, int64_t *start, int64_t *size64, int32_t *bh, cudaStream_t st);
cudaError_t launch_rank_clusters(const float *q_proj, int nq, int D, const float *clusters, int C, int segdims,
                                 const int64_t *start, const int64_t *size, const int64_t *rule_size, float visit, int k,
                                 int2 *ranges, int32_t *n_ranges, cudaStream_t st);

}  // namespace vaqgpu
