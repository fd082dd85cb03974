// ADC scan, filter-and-refine form — the kernel behind VAQ::searchEarlyAbandon (reference
// bitvecengine/VAQ.cpp:1694-1727) for whole-index scans.
//
// The reference abandons a row as soon as its partial distance over the leading subspaces
// reaches the running k-th best (VAQ.cpp:1708); the subspaces are in variance-descending order,
// so on the reference's data > 95 % of the rows die after the first group of four.  A lane-per-row
// GPU loop cannot exploit that (a warp only stops when all 32 rows are dead), so the scan is split:
//
//  stage 1 (every row):   one CTA owns a tile of T queries and a chunk of rows.  The T lookup
//      tables sit in shared memory interleaved per entry ([entry][T]), so ONE 16/32-byte shared load
//      returns the table values of all T queries for a row's code.  Each lane takes a row, reads only
//      the first 128-bit word of it (coalesced, prefetched three tiles ahead), extracts the first
//      group's codes once and accumulates the first group for all T queries in the reference's
//      order, dism = ((l0+l1)+l2)+l3.  (row, query) pairs whose partial distance already exceeds the
//      query's running k-th best are dropped — exactly the reference's first abandon test.
//  stage 2 (survivors):   surviving pairs are compacted into a per-warp queue; whenever 32 are
//      pending the warp scores them with full lanes — reloads the row's words, walks ALL subspaces
//      in the reference's order and grouping (so the distance is bit-identical to searchHeap's),
//      abandoning when every lane is dead, and inserts the finishers into the per-(warp, query)
//      sorted top-k lists.
//
// Thresholds only ever come from k rows that were really scored, so the pruning is exact: the
// result is the k lexicographically smallest (distance, row) keys, independent of scheduling.
// A per-query threshold in global memory lets later row chunks (and later launches) start from
// the bound earlier chunks reached.
#include "common.cuh"

namespace vaqgpu {

constexpr int kQueueCap = 64;   // per-warp survivor queue (worst case 31 pending + 32 pushed)
constexpr int kPrefetch = 3;    // tiles in flight per warp

template <int T> struct LutVec;
template <> struct LutVec<1> { float v[1]; };
template <> struct LutVec<2> { float v[2]; };
template <> struct LutVec<4> { float v[4]; };
template <> struct LutVec<8> { float v[8]; };

template <int T>
__device__ __forceinline__ void lds_vec(float (&out)[T], const float *p) {
  if constexpr (T == 1) {
    out[0] = *p;
  } else if constexpr (T == 2) {
    const float2 a = *reinterpret_cast<const float2 *>(p);
    out[0] = a.x; out[1] = a.y;
  } else {
#pragma unroll
    for (int i = 0; i < T / 4; i++) {
      const float4 a = *reinterpret_cast<const float4 *>(p + 4 * i);
      out[4 * i] = a.x; out[4 * i + 1] = a.y; out[4 * i + 2] = a.z; out[4 * i + 3] = a.w;
    }
  }
}

template <int T>
__device__ __forceinline__ void ldg_vec(float (&out)[T], const float *p) {
  if constexpr (T == 1) {
    out[0] = __ldg(p);
  } else if constexpr (T == 2) {
    const float2 a = __ldg(reinterpret_cast<const float2 *>(p));
    out[0] = a.x; out[1] = a.y;
  } else {
#pragma unroll
    for (int i = 0; i < T / 4; i++) {
      const float4 a = __ldg(reinterpret_cast<const float4 *>(p + 4 * i));
      out[4 * i] = a.x; out[4 * i + 1] = a.y; out[4 * i + 2] = a.z; out[4 * i + 3] = a.w;
    }
  }
}

// thresholds are updated by other warps: re-read every tile (asm volatile keeps the load in the loop)
template <int T>
__device__ __forceinline__ void lds_thr(float (&out)[T], const uint32_t *p) {
  const uint32_t addr = smem_u32(p);
  if constexpr (T == 1) {
    asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(out[0]) : "r"(addr));
  } else if constexpr (T == 2) {
    asm volatile("ld.volatile.shared.v2.f32 {%0,%1}, [%2];" : "=f"(out[0]), "=f"(out[1]) : "r"(addr));
  } else {
#pragma unroll
    for (int i = 0; i < T / 4; i++)
      asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=f"(out[4 * i]), "=f"(out[4 * i + 1]), "=f"(out[4 * i + 2]), "=f"(out[4 * i + 3])
                   : "r"(addr + 16 * i));
  }
}

// Full-precision score of one (row, query) pair per lane; returns false when every lane abandoned.
template <int W, int T>
__device__ __forceinline__ bool score_pair(const uint4 (&cw)[W], const ScanLayout &lay, const float *__restrict__ slut,
                                           const float *__restrict__ gspill, int t, float thr, bool active,
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
    const int fe = lay.fbeg[w + 1];
    const uint32_t lo = wd[w], hi = wd[w + 1];
    for (; f < fe; f++) {
      const uint32_t meta = lay.fmeta[f];
      const uint32_t code = __funnelshift_r(lo, hi, meta & 31u) & (meta >> 16);
      const uint32_t idx = (lay.foff[f] + code) * T + t;
      const float v = (meta & kFieldSpill) ? __ldg(gspill + idx) : slut[idx];
      dism += v;
      if ((f & 3) == 3) {
        dist += dism;
        dism = 0.f;
        if (__all_sync(0xffffffffu, !active || (dist > thr))) return false;
      }
    }
  }
  if (lay.M & 3) dist += dism;
  dist_out = dist;
  return true;
}

template <int W, int T>
__global__ void __launch_bounds__(512, 1) adc_filter_scan_kernel(const __grid_constant__ AdcFilterArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int k = a.k;
  const int qt = blockIdx.x, chunk = blockIdx.y;
  const int q0 = qt * T;

  const size_t lut_bytes = (size_t)a.smem_lut_floats * T * sizeof(float);
  float *slut = reinterpret_cast<float *>(smem_raw);
  uint32_t *thr_f = reinterpret_cast<uint32_t *>(smem_raw + ((lut_bytes + 15) & ~(size_t)15));   // [8] distance bits of blk_thr
  uint64_t *lists = reinterpret_cast<uint64_t *>(thr_f + 8);                                   // [nwarps][T][k]
  uint64_t *merged = lists + (size_t)nwarps * T * k;                                           // [k]
  uint64_t *blk_thr = merged + k;                                                              // [T]
  uint64_t *bar = blk_thr + T;
  uint32_t *queues = reinterpret_cast<uint32_t *>(bar + 1);                                    // [nwarps][kQueueCap]

  const float *glut = a.lut + (size_t)qt * a.lut_stride * T;       // this tile's interleaved tables
  const float *gspill = glut + (size_t)a.smem_lut_floats * T;

  for (int i = tid; i < nwarps * T * k; i += blockDim.x) lists[i] = kEmptyKey;
  if (tid < T) {
    const int q = min(q0 + tid, a.nq - 1);
    const uint32_t g = a.thr_global[q];
    blk_thr[tid] = ((uint64_t)g << 32) | 0xFFFFFFFFull;
    thr_f[tid] = g;
  }
  if (lut_bytes) {
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (tid == 0) {
      mbar_arrive_expect_tx(bar, (uint32_t)lut_bytes);
      for (size_t off = 0; off < lut_bytes; off += 32768) {
        const uint32_t n = (uint32_t)min((size_t)32768, lut_bytes - off);
        tma_bulk_g2s(reinterpret_cast<unsigned char *>(slut) + off, reinterpret_cast<const unsigned char *>(glut) + off, n, bar);
      }
    }
    mbar_wait(bar, 0);
  } else {
    __syncthreads();
  }

  // stage-1 program: the first group (<= 4 fields, <= 60 bits, i.e. inside 32-bit words 0..2)
  const int G1 = min(4, a.lay.M);
  uint32_t s1_sh[4], s1_mask[4], s1_off[4];
  bool s1_hi[4], s1_spill[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int f = min(i, G1 - 1);
    const uint32_t meta = a.lay.fmeta[f];
    s1_sh[i] = meta & 31u;
    s1_mask[i] = meta >> 16;
    s1_off[i] = a.lay.foff[f] * T;
    s1_hi[i] = a.lay.fword[f] != 0;
    s1_spill[i] = (meta & kFieldSpill) != 0;
  }

  uint32_t *myq = queues + warp * kQueueCap;
  int qn = 0;
  const unsigned lt_mask = (1u << lane) - 1u;

  const int64_t tile_begin = (int64_t)chunk * a.chunk_tiles;
  const int64_t n_tiles = (a.n_rows + kTileRows - 1) >> 5;
  const int64_t tile_end = min(n_tiles, tile_begin + a.chunk_tiles);
  const int64_t row_base = tile_begin << 5;

  // ---- stage 2 on up to 32 queued pairs -------------------------------------------------------
  auto drain = [&](int take) {
    const bool active = lane < take;
    const uint32_t e = active ? myq[qn - take + lane] : 0u;
    qn -= take;
    const int t = (int)(e & 7u);
    const int64_t row = row_base + (e >> 3);
    uint4 cw[W];
    {
      const uint4 *p = a.codes + ((size_t)(row >> 5) * W) * kTileRows + (row & 31);
#pragma unroll
      for (int j = 0; j < W; j++) cw[j] = active ? __ldg(p + j * kTileRows) : make_uint4(0, 0, 0, 0);
    }
    volatile uint64_t *mylist = lists + ((size_t)warp * T + t) * k;
    uint64_t thrkey = kEmptyKey;
    if (active) {
      thrkey = mylist[k - 1];
      const uint64_t b = *reinterpret_cast<volatile uint64_t *>(blk_thr + t);
      thrkey = b < thrkey ? b : thrkey;
    }
    const float thr = __uint_as_float((uint32_t)(thrkey >> 32));
    float dist = 0.f;
    if (!score_pair<W, T>(cw, a.lay, slut, gspill, t, thr, active, dist)) return;
    const uint64_t key = active ? make_key_f32(dist, (int32_t)row) : kEmptyKey;
    unsigned m = __ballot_sync(0xffffffffu, key < thrkey);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const uint64_t kk = __shfl_sync(0xffffffffu, key, src);
      const int tt = __shfl_sync(0xffffffffu, t, src);
      volatile uint64_t *lst = lists + ((size_t)warp * T + tt) * k;
      const uint64_t before = lst[k - 1];
      const uint64_t kth = warp_list_insert(lst, k, kk, lane);
      if (lane == 0 && kth != before && kth != kEmptyKey) {
        atomicMin(reinterpret_cast<unsigned long long *>(blk_thr + tt), (unsigned long long)kth);
        atomicMin(thr_f + tt, (uint32_t)(kth >> 32));
      }
    }
  };

  // ---- stage 1 --------------------------------------------------------------------------------
  const int64_t t0 = tile_begin + warp;
  uint4 buf[kPrefetch];
#pragma unroll
  for (int i = 0; i < kPrefetch; i++) {
    const int64_t tl = t0 + (int64_t)i * nwarps;
    buf[i] = make_uint4(0, 0, 0, 0);
    if (tl < tile_end) buf[i] = ldg_stream_u4(a.codes + ((size_t)tl * W) * kTileRows + lane);
  }
  for (int64_t base = t0; base < tile_end; base += (int64_t)kPrefetch * nwarps) {
#pragma unroll
    for (int i = 0; i < kPrefetch; i++) {
      const int64_t tl = base + (int64_t)i * nwarps;
      if (tl >= tile_end) break;
      const uint4 w0 = buf[i];
      const int64_t tn = tl + (int64_t)kPrefetch * nwarps;
      if (tn < tile_end) buf[i] = ldg_stream_u4(a.codes + ((size_t)tn * W) * kTileRows + lane);

      // thresholds of the T queries (distance part of the block-wide k-th keys)
      float thr[T];
      lds_thr<T>(thr, thr_f);

      float dism[T];
#pragma unroll
      for (int i1 = 0; i1 < 4; i1++) {
        if (i1 < G1) {
          const uint32_t lo = s1_hi[i1] ? w0.y : w0.x, hi = s1_hi[i1] ? w0.z : w0.y;
          const uint32_t code = __funnelshift_r(lo, hi, s1_sh[i1]) & s1_mask[i1];
          float v[T];
          if (s1_spill[i1]) ldg_vec<T>(v, gspill + s1_off[i1] + code * T);
          else lds_vec<T>(v, slut + s1_off[i1] + code * T);
#pragma unroll
          for (int t = 0; t < T; t++) dism[t] = (i1 == 0) ? v[t] : dism[t] + v[t];
        }
      }
      const int64_t row = (tl << 5) + lane;
      const bool valid = row < a.n_rows;
      unsigned sb = 0;
#pragma unroll
      for (int t = 0; t < T; t++) sb |= (valid && !(dism[t] > thr[t]) && (q0 + t < a.nq)) ? (1u << t) : 0u;
      if (__any_sync(0xffffffffu, sb != 0)) {
        const uint32_t rel = (uint32_t)(row - row_base) << 3;
#pragma unroll
        for (int t = 0; t < T; t++) {
          const unsigned m = __ballot_sync(0xffffffffu, (sb >> t) & 1u);
          if (m) {
            if ((sb >> t) & 1u) myq[qn + __popc(m & lt_mask)] = rel | (uint32_t)t;
            qn += __popc(m);
            __syncwarp();
            if (qn >= 32) drain(32);
          }
        }
      }
    }
  }
  __syncwarp();
  if (qn > 0) drain(qn);

  // ---- CTA epilogue: merge the warps' lists per query, publish keys and the new bound -----------
  __syncthreads();
  for (int t = 0; t < T; t++) {
    const int q = q0 + t;
    if (q >= a.nq) break;
    for (int i = tid; i < k; i += blockDim.x) merged[i] = kEmptyKey;
    __syncthreads();
    const int total = nwarps * k;
    for (int e = tid; e < total; e += blockDim.x) {
      const int l = e / k, i = e - l * k;
      const uint64_t key = lists[((size_t)l * T + t) * k + i];
      if (key == kEmptyKey) continue;
      int rank = i;
      for (int o = 0; o < nwarps && rank < k; o++) {
        if (o == l) continue;
        rank += lower_bound_u64(lists + ((size_t)o * T + t) * k, k, key);
      }
      if (rank < k) merged[rank] = key;
    }
    __syncthreads();
    uint64_t *out = a.out_keys + ((size_t)q * a.n_chunks + chunk) * k;
    for (int i = tid; i < k; i += blockDim.x) out[i] = merged[i];
    if (tid == 0 && merged[k - 1] != kEmptyKey) atomicMin(a.thr_global + q, (uint32_t)(merged[k - 1] >> 32));
    __syncthreads();
  }
}

size_t adc_filter_smem_bytes(int smem_lut_floats, int T, int k, int threads) {
  const int nwarps = threads / 32;
  size_t b = (((size_t)smem_lut_floats * T * 4 + 15) & ~(size_t)15) + 32;
  b += ((size_t)nwarps * T * k + k + T + 1) * sizeof(uint64_t);
  b += (size_t)nwarps * kQueueCap * sizeof(uint32_t);
  return b;
}

template <int W, int T>
static cudaError_t launch_wt(const AdcFilterArgs &a, int threads, size_t smem_bytes, cudaStream_t st) {
  static size_t configured = 0;
  if (smem_bytes > configured) {
    cudaError_t e = cudaFuncSetAttribute(adc_filter_scan_kernel<W, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    configured = smem_bytes;
  }
  dim3 grid((unsigned)((a.nq + T - 1) / T), (unsigned)a.n_chunks);
  adc_filter_scan_kernel<W, T><<<grid, threads, smem_bytes, st>>>(a);
  return cudaGetLastError();
}

template <int W>
static cudaError_t launch_w(const AdcFilterArgs &a, int T, int threads, size_t smem_bytes, cudaStream_t st) {
  switch (T) {
    case 1: return launch_wt<W, 1>(a, threads, smem_bytes, st);
    case 2: return launch_wt<W, 2>(a, threads, smem_bytes, st);
    case 4: return launch_wt<W, 4>(a, threads, smem_bytes, st);
    case 8: return launch_wt<W, 8>(a, threads, smem_bytes, st);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_adc_filter_scan(const AdcFilterArgs &a, int T, int threads, size_t smem_bytes, cudaStream_t st) {
  switch (a.lay.W) {
    case 1: return launch_w<1>(a, T, threads, smem_bytes, st);
    case 2: return launch_w<2>(a, T, threads, smem_bytes, st);
    case 3: return launch_w<3>(a, T, threads, smem_bytes, st);
    case 4: return launch_w<4>(a, T, threads, smem_bytes, st);
    case 5: return launch_w<5>(a, T, threads, smem_bytes, st);
    case 6: return launch_w<6>(a, T, threads, smem_bytes, st);
    case 7: return launch_w<7>(a, T, threads, smem_bytes, st);
    case 8: return launch_w<8>(a, T, threads, smem_bytes, st);
    default: return cudaErrorInvalidValue;
  }
}

__global__ void fill_u32_kernel(uint32_t *p, int n, uint32_t v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

cudaError_t launch_fill_u32(uint32_t *p, int n, uint32_t v, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  fill_u32_kernel<<<(n + 255) / 256, 256, 0, st>>>(p, n, v);
  return cudaGetLastError();
}

}  // namespace vaqgpu
