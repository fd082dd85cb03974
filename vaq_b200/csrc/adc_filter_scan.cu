// ADC scan, filter-and-refine on fp32 tables — the fallback kernel behind VAQ::searchEarlyAbandon (reference
// bitvecengine/VAQ.cpp:1694-1727) for the cases the fp16 form (adc_filter16_scan.cu) does not take: tables too
// large for eight queries in shared memory (T = 1..4, with the trailing tables spilled to L2 when even one
// query's tables do not fit — GIST-512-bit / 13-15-bit subspaces), very large k, or fewer than five queries.
//
// The reference abandons a row as soon as its partial distance over the leading subspaces reaches the
// running k-th best (VAQ.cpp:1708); the subspaces are in variance-descending order, so on the reference's
// data > 95 % of the rows die after the first group of four.  A lane-per-row GPU loop cannot exploit that (a
// warp only stops when all 32 rows are dead), so the scan is split:
//
//  stage 1 (every row):   one CTA owns a tile of T queries and a chunk of rows.  The T lookup tables sit in
//      shared memory interleaved per entry ([entry][T]), so ONE shared load returns the table values of all T
//      queries for a row's code.  Each lane takes a row, reads only the first 128-bit word of it (coalesced,
//      two tiles prefetched in registers), extracts the first group's codes once and accumulates the first
//      group for all T queries in the reference's order, dism = ((l0+l1)+l2)+l3.  (row, query) pairs whose
//      partial distance already exceeds the query's running k-th best are dropped — exactly the reference's
//      first abandon test.
//  stage 2 (survivors):   surviving pairs are compacted into per-warp queues (ballot + popc); whenever 32 are
//      pending the warp scores them with full lanes in two levels (groups 1-2, then the rest), the row's words
//      fetched once into registers, walking the subspaces in the reference's order and grouping (so the
//      distance is bit-identical to searchHeap's), abandoning when every lane is dead, and inserts the
//      finishers into the CTA's per-query sorted top-k list under a per-query lock.
//
// Bounds only ever come from k rows that were really scored, so the pruning is exact: the result is the k
// lexicographically smallest (distance, row) keys, independent of scheduling.  A per-query bound in global
// memory lets later row chunks start from the bound earlier chunks reached; each CTA also seeds its bounds
// from the per-warp minima of a few sample rows.
#include "common.cuh"

namespace vaqgpu {


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

// Exact score of one (row, query) pair per lane over the fields [f_begin, f_end) (f_begin a multiple of
// 4), continuing from `dist`, in the reference's order and grouping (dism = ((l0+l1)+l2)+l3 ;
// dist += dism, VAQ.cpp:1741-1748).  A compact loop (no per-word unrolling) keeps the kernel inside the
// instruction cache.  Returns false when every lane abandoned.
template <int W, int T>
__device__ __forceinline__ bool score_fields(const uint32_t *__restrict__ rp, const ScanLayout &lay, const float *__restrict__ slut,
                                             const float *__restrict__ gspill, int t, float thr, bool active, int f_begin,
                                             int f_end, float &dist) {
  // Rows of up to 256 bits are fetched once (W 16-byte loads per lane) and walked from registers: fields come
  // in increasing bit order, so a two-word window slides over the eight words and a word is picked by a select
  // tree on the warp-uniform word index (a 4-byte load per field costs 32 L1 wavefronts: every lane has its own row).
  uint32_t wd[8];
  if constexpr (W <= 2) {
    const uint4 v0 = __ldg(reinterpret_cast<const uint4 *>(rp));
    wd[0] = v0.x; wd[1] = v0.y; wd[2] = v0.z; wd[3] = v0.w;
    if constexpr (W == 2) {
      const uint4 v1 = __ldg(reinterpret_cast<const uint4 *>(rp) + kTileRows);
      wd[4] = v1.x; wd[5] = v1.y; wd[6] = v1.z; wd[7] = v1.w;
    } else {
      wd[4] = wd[5] = wd[6] = wd[7] = 0u;
    }
  }
  auto selw = [&](int i) -> uint32_t {
    const uint32_t s0 = (i & 1) ? wd[1] : wd[0], s1 = (i & 1) ? wd[3] : wd[2];
    const uint32_t s2 = (i & 1) ? wd[5] : wd[4], s3 = (i & 1) ? wd[7] : wd[6];
    const uint32_t t0 = (i & 2) ? s1 : s0, t1 = (i & 2) ? s3 : s2;
    return (i & 8) ? 0u : ((i & 4) ? t1 : t0);
  };
  int widx = -2;
  uint32_t lo = 0u, hi = 0u;
  for (int g = f_begin; g < f_end; g += 4) {
    float dism = 0.f;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int f = g + j;
      if (f < f_end) {
        const uint32_t meta = lay.fmeta[f];
        if constexpr (W <= 2) {
          const int fw = lay.fword[f];
          if (fw != widx) {
            lo = (fw == widx + 1) ? hi : selw(fw);
            hi = selw(fw + 1);
            widx = fw;
          }
        } else {
          lo = __ldg(rp + lay.fw_lo[f]);
          hi = __ldg(rp + lay.fw_hi[f]);
        }
        const uint32_t code = __funnelshift_r(lo, hi, meta & 31u) & (meta >> 16);
        const uint32_t idx = (lay.foff[f] + code) * T + t;
        const float v = (meta & kFieldSpill) ? __ldg(gspill + idx) : slut[idx];
        dism += v;
      }
    }
    dist += dism;
    if (__all_sync(0xffffffffu, !active || (dist > thr))) return false;
  }
  return true;
}

// Exact distances of one row (per lane) to all T queries of the tile, same summation order as above.
template <int T>
__device__ __forceinline__ void score_row_all(const uint32_t *__restrict__ rp, const ScanLayout &lay, const float *__restrict__ slut,
                                              const float *__restrict__ gspill, float (&dist)[T]) {
  const int M = lay.M;
#pragma unroll
  for (int t = 0; t < T; t++) dist[t] = 0.f;
  for (int g = 0; g < M; g += 4) {
    float dism[T];
#pragma unroll
    for (int t = 0; t < T; t++) dism[t] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int f = g + j;
      if (f < M) {
        const uint32_t meta = lay.fmeta[f];
        const uint32_t lo = __ldg(rp + lay.fw_lo[f]), hi = __ldg(rp + lay.fw_hi[f]);
        const uint32_t code = __funnelshift_r(lo, hi, meta & 31u) & (meta >> 16);
        const uint32_t idx = (lay.foff[f] + code) * T;
        float v[T];
        if (meta & kFieldSpill) ldg_vec<T>(v, gspill + idx);
        else lds_vec<T>(v, slut + idx);
#pragma unroll
        for (int t = 0; t < T; t++) dism[t] += v[t];
      }
    }
#pragma unroll
    for (int t = 0; t < T; t++) dist[t] += dism[t];
  }
}

constexpr int kSeedRowsPerLane = 4;   // bound seeding: nwarps * 32 * 4 sample rows per CTA
constexpr int kQ2Cap = 64;                                         // 31 pending + one level-1 drain

template <int W, int T>
__global__ void __launch_bounds__(1024, 1) adc_filter_scan_kernel(const __grid_constant__ AdcFilterArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int k = a.k;
  const int qt = blockIdx.x, chunk = blockIdx.y;
  const int q0 = qt * T;
  constexpr int kQ1Cap = 32 + 64 * T;             // 31 pending + two tiles' pushes for T queries

  const size_t lut_bytes = (size_t)a.smem_lut_floats * T * sizeof(float);
  float *slut = reinterpret_cast<float *>(smem_raw);
  uint32_t *thr_f = reinterpret_cast<uint32_t *>(smem_raw + ((lut_bytes + 15) & ~(size_t)15));   // [8] k-th distance bits per query
  uint32_t *locks = thr_f + 8;                                                                 // [8] per-query list locks
  uint64_t *lists = reinterpret_cast<uint64_t *>(locks + 8);                                   // [T][k] ascending keys
  uint64_t *bar = lists + (size_t)T * k;
  uint32_t *queues = reinterpret_cast<uint32_t *>(bar + 1);                                    // per warp: q1 | q2 entries | q2 dists

  const float *glut = a.lut + (size_t)qt * a.lut_stride * T;       // this tile's interleaved tables
  const float *gspill = glut + (size_t)a.smem_lut_floats * T;

  for (int i = tid; i < T * k; i += blockDim.x) lists[i] = kEmptyKey;
  if (tid < 8) {
    const int q = min(q0 + min(tid, T - 1), a.nq - 1);
    thr_f[tid] = a.thr_global[q];
    locks[tid] = 0u;
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
  const int M = a.lay.M;
  const int G1 = min(4, M);
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
  // stage 2 runs in two levels when there are more than two groups: level 1 re-scores groups 1-2 exactly
  // and keeps the pairs still under the bound, level 2 finishes them
  const bool two_level = M > 8;
  const int F2 = two_level ? 8 : M;

  const unsigned qmask = (a.nq - q0 >= T) ? ((1u << T) - 1u) : ((1u << (a.nq - q0)) - 1u);   // real queries of this tile
  uint32_t *q1 = queues + (size_t)warp * (kQ1Cap + 2 * kQ2Cap);
  uint32_t *q2e = q1 + kQ1Cap;
  float *q2d = reinterpret_cast<float *>(q2e + kQ2Cap);
  int q1n = 0, q2n = 0;
  const unsigned lt_mask = (1u << lane) - 1u;

  const int64_t tile_begin = a.tile_lo + (int64_t)chunk * a.chunk_tiles;
  const int64_t tile_end = min(a.tile_hi, tile_begin + a.chunk_tiles);
  const int64_t row_base = tile_begin << 5;
  const uint32_t *codes32 = reinterpret_cast<const uint32_t *>(a.codes);

  // stage 1 on one tile whose first words are in `buf`; refills `buf` with the tile 2*nwarps further on
  auto stage1 = [&](uint4 &buf, int64_t tile) {
    const uint4 w0 = buf;
    {
      const int64_t tn = tile + 2 * (int64_t)nwarps;
      if (tn < tile_end) buf = ldg_stream_u4(a.codes + ((size_t)tn * W) * kTileRows + lane);
    }
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
    const int64_t row = (tile << 5) + lane;
    const bool valid = row < a.n_rows;
    unsigned sb = 0;
#pragma unroll
    for (int t = 0; t < T; t++) sb |= !(dism[t] > thr[t]) ? (1u << t) : 0u;
    sb = valid ? (sb & qmask) : 0u;
    if (__any_sync(0xffffffffu, sb != 0)) {
      const uint32_t rel = (uint32_t)(row - row_base) << 3;
#pragma unroll
      for (int t = 0; t < T; t++) {
        const unsigned m = __ballot_sync(0xffffffffu, (sb >> t) & 1u);
        if ((sb >> t) & 1u) q1[q1n + __popc(m & lt_mask)] = rel | (uint32_t)t;
        q1n += __popc(m);
      }
    }
  };

  // ---- bound seeding ---------------------------------------------------------------------------------
  // A CTA that starts without a k-th-best bound would have to score everything it sees.  Each warp
  // scores a few sample rows of the chunk exactly for all T queries and keeps the minimum per query:
  // the k-th smallest of the nwarps minima is the distance of k distinct rows, hence a valid bound.
  if (a.seed && k <= nwarps && (tile_end - tile_begin) * kTileRows >= (int64_t)blockDim.x * kSeedRowsPerLane * 8) {
    const int64_t rows_here = min(a.n_rows, tile_end << 5) - row_base;
    const int64_t step = rows_here / ((int64_t)blockDim.x * kSeedRowsPerLane);
    float best[T];
#pragma unroll
    for (int t = 0; t < T; t++) best[t] = __uint_as_float(0x7f800000u);
    for (int j = 0; j < kSeedRowsPerLane; j++) {
      const int64_t row = row_base + ((int64_t)(j * (int)blockDim.x + tid)) * step;
      const uint32_t *rp = codes32 + (((size_t)(row >> 5) * W) * kTileRows + (row & 31)) * 4;
      float d[T];
      score_row_all<T>(rp, a.lay, slut, gspill, d);
#pragma unroll
      for (int t = 0; t < T; t++) best[t] = fminf(best[t], d[t]);
    }
#pragma unroll
    for (int t = 0; t < T; t++) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) best[t] = fminf(best[t], __shfl_xor_sync(0xffffffffu, best[t], o));
    }
    __syncthreads();
    float *allmin = reinterpret_cast<float *>(queues);   // [nwarps][T], rows of different warps' queue space: use warp 0's
    if (lane == 0) {
#pragma unroll
      for (int t = 0; t < T; t++) allmin[warp * T + t] = best[t];
    }
    __syncthreads();
    if (tid < T) {
      // k-th smallest of the nwarps minima (rank by counting; ties broken by index)
      const float *col = allmin + tid;
      float kth = __uint_as_float(0x7f800000u);
      for (int i = 0; i < nwarps; i++) {
        const float x = col[i * T];
        int rank = 0;
        for (int j = 0; j < nwarps; j++) rank += (col[j * T] < x) || (col[j * T] == x && j < i);
        if (rank == k - 1) kth = x;
      }
      atomicMin(thr_f + tid, __float_as_uint(kth));
    }
    __syncthreads();
  }

  int64_t tl = tile_begin + warp;
  uint4 bA = make_uint4(0, 0, 0, 0), bB = bA;
  if (tl < tile_end) bA = ldg_stream_u4(a.codes + ((size_t)tl * W) * kTileRows + lane);
  if (tl + nwarps < tile_end) bB = ldg_stream_u4(a.codes + ((size_t)(tl + nwarps) * W) * kTileRows + lane);
  int refresh = 0;

  while (true) {
    const bool more = tl < tile_end;
    int level = 0, take = 0;
    if (q2n >= 32) { level = 2; take = 32; }
    else if (q1n >= 32) { level = 1; take = 32; }
    else if (!more) {
      if (q1n > 0) { level = 1; take = q1n; }
      else if (q2n > 0) { level = 2; take = q2n; }
      else break;
    }
    if (level) {
      // ---- stage 2 on up to 32 queued pairs (single code site for both levels) -----------------------
      __syncwarp();
      const bool active = lane < take;
      uint32_t e = 0u;
      float dist = 0.f;
      int fb = 0, fe = F2;
      if (level == 1) {
        if (active) e = q1[q1n - take + lane];
        q1n -= take;
      } else {
        if (active) { e = q2e[q2n - take + lane]; dist = q2d[q2n - take + lane]; }
        q2n -= take;
        fb = F2; fe = M;
      }
      const int t = (int)(e & 7u);
      const int64_t row = row_base + (e >> 3);
      const uint32_t *rp = codes32 + (((size_t)(row >> 5) * W) * kTileRows + (row & 31)) * 4;
      const float thr = __uint_as_float(*reinterpret_cast<volatile uint32_t *>(thr_f + t));
      if (score_fields<W, T>(rp, a.lay, slut, gspill, t, thr, active, fb, fe, dist)) {
        if (level == 1 && two_level) {
          const bool s = active && !(dist > thr);
          const unsigned m = __ballot_sync(0xffffffffu, s);
          if (s) {
            const int pos = q2n + __popc(m & lt_mask);
            q2e[pos] = e;
            q2d[pos] = dist;
          }
          q2n += __popc(m);
        } else {
          const uint64_t key = active ? make_key_f32(dist, a.rowid ? (int32_t)__ldg(a.rowid + row) : (int32_t)row) : kEmptyKey;
          const uint64_t kth0 = active ? *reinterpret_cast<volatile uint64_t *>(lists + (size_t)t * k + (k - 1)) : 0ull;
          unsigned m = __ballot_sync(0xffffffffu, key < kth0);
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const uint64_t kk = __shfl_sync(0xffffffffu, key, src);
            const int tt = __shfl_sync(0xffffffffu, t, src);
            volatile uint64_t *lst = lists + (size_t)tt * k;
            {
              uint64_t cur = lst[k - 1];
              cur = __shfl_sync(0xffffffffu, cur, 0);      // one observer: the decision must be warp-uniform
              if (!(kk < cur)) continue;
            }
            if (lane == 0) while (atomicCAS(locks + tt, 0u, 1u) != 0u) __nanosleep(20);
            __syncwarp();
            const uint64_t before = lst[k - 1];
            const uint64_t kth = warp_list_insert(lst, k, kk, lane);
            __syncwarp();
            if (lane == 0) {
              __threadfence_block();
              atomicExch(locks + tt, 0u);
              if (kth != before && kth != kEmptyKey) {
                const uint32_t bits = (uint32_t)(kth >> 32);
                atomicMin(thr_f + tt, bits);
                if (q0 + tt < a.nq) publish_global_bound(a.thr_global, a.peers, q0 + tt, bits);
              }
            }
          }
        }
      }
      continue;
    }

    // ---- stage 1 on two tiles (two register buffers -> loads stay two iterations ahead) ---------------
    if (((++refresh) & 31) == 0 && lane < T && q0 + lane < a.nq) {
      // pick up bounds published by other row chunks of this query tile
      atomicMin(thr_f + lane, *reinterpret_cast<volatile uint32_t *>(a.thr_global + q0 + lane));
    }
    stage1(bA, tl);
    if (tl + nwarps < tile_end) stage1(bB, tl + nwarps);
    tl += 2 * (int64_t)nwarps;
  }

  // ---- CTA epilogue: publish this (query tile, chunk)'s keys -----------------------------------------
  __syncthreads();
  for (int i = tid; i < T * k; i += blockDim.x) {
    const int t = i / k, j = i - t * k;
    const int q = q0 + t;
    if (q < a.nq) a.out_keys[((size_t)q * a.out_slots + a.slot_base + chunk) * k + j] = lists[i];
  }
}

size_t adc_filter_smem_bytes(int smem_lut_floats, int T, int k, int threads) {
  const int nwarps = threads / 32;
  size_t b = (((size_t)smem_lut_floats * T * 4 + 15) & ~(size_t)15) + 64;
  b += ((size_t)T * k + 1) * sizeof(uint64_t);
  b += (size_t)nwarps * (32 + 64 * T + 2 * kQ2Cap) * sizeof(uint32_t);
  return b;
}

template <int W, int T>
static cudaError_t launch_wt(const AdcFilterArgs &a, int threads, size_t smem_bytes, cudaStream_t st) {
  static SmemOptIn optin;
  {
    cudaError_t e = optin.ensure(adc_filter_scan_kernel<W, T>, smem_bytes);
    if (e != cudaSuccess) return e;
  }
  const int64_t nt = a.tile_hi - a.tile_lo;
  if (nt <= 0) return cudaSuccess;
  dim3 grid((unsigned)((a.nq + T - 1) / T), (unsigned)((nt + a.chunk_tiles - 1) / a.chunk_tiles));
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
