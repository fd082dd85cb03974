// ADC scan, filter-and-refine with half-precision lower-bound tables — the default kernel behind
// VAQ::searchEarlyAbandon (reference bitvecengine/VAQ.cpp:1694-1727) when eight queries' tables fit in
// shared memory as fp16.
//
// The reference abandons a row once its partial distance over the leading (highest-variance) subspaces
// reaches the running k-th best (VAQ.cpp:1708); on its data ~95 % of the (row, query) pairs die after the
// first group of four.  The shared-memory tables here hold, for a tile of eight queries,
//     e16[s][c][t] = round_toward_zero_fp16( scale_t * lut[t][s][c] ),     t = 0..7
// (written by lut_build_kernel<8> together with the fp32 tables; scale_t is a per-query power of two), i.e.
// guaranteed LOWER bounds of the reference's table entries, 16 bytes per code: one LDS.128 returns the
// entries of all eight queries, and four HADD2 accumulate them.
//
//   stage 1 (every row)    each lane takes a row, reads only its first 128-bit word (coalesced, prefetched),
//                          extracts the first group's codes once and accumulates the group for the eight
//                          queries in packed half2.  Rows with at least one query still under its bound are
//                          compacted (ballot + popc) into the warp's queue as (row, query mask).
//   level 1 / level 2      32 queued rows per pass, one per lane: the row's words are fetched once into
//                          registers, the lower bound over groups 1-2 (level 1) / all subspaces (level 2) is
//                          accumulated for all eight queries the same way, and the query mask shrinks.
//   level 3 (exact)        rows that still have a query under its bound are scored EXACTLY for those queries
//                          from the fp32 tables in global memory (L2): four (row, query) pairs per round, eight
//                          lanes each, in the reference's order and grouping (dism = ((l0+l1)+l2)+l3 ;
//                          dist += dism, VAQ.cpp:1741-1748).  Only these exact distances enter the top-k lists
//                          and tighten the bounds.
// A query tile may start its scan anywhere in the chunk (scan order, vaqgpu_host.cu build_scan_order): the rows nearest
// to the tile's queries first, so that the bounds are tight after the first percent of the rows.
//
// Exactness of the pruning.  Entries are rounded toward zero, half2 additions round to nearest: after n
// additions the accumulated value is at most (1+2^-11)^n above the real sum of the (scaled) entries.  A pair
// is dropped only if its accumulated value exceeds  RU_fp16(thr * scale * (1+m))  with m = 2^-9 after the
// 3 additions of stage 1, 2^-7 after 7 (level 1), 2^-5 after <= 127 ... capped at 31 additions for M <= 32 and
// m = 2^-3 beyond (level 2).  Then the real partial sum exceeds thr by more than 2^-11 relative, far more than
// the ~M * 2^-24 by which the fp32 distance the reference computes can fall below the real sum, so the
// reference's own distance is > thr and the row cannot be among the k best.  Result: bit-identical to the fp32
// kernels (tests/test_gpu_vaq.py runs all three against the oracle).
#include <cuda_fp16.h>

#include "common.cuh"

namespace vaqgpu {

namespace {

constexpr int T8 = 8;
constexpr int kQCap = 64;                // level-2 / level-3 queues: at most 31 pending + 32 pushed by one pass
__host__ __device__ constexpr int q1_cap(int tpi) { return 32 + 32 * tpi; }      // level-1 queue: 31 pending + one stage-1 iteration (TPI tiles)

__device__ __forceinline__ uint32_t half_bits_ru(float x) { return (uint32_t)__half_as_ushort(__float2half_ru(x)); }

// shared-memory loads by 32-bit shared address (keeps the generic->shared window arithmetic out of the loop)
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 r;
  asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ uint4 lds128_volatile(uint32_t addr) {
  uint4 r;
  asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
// four 0xFF/0x00 flag bytes -> 4-bit mask (bit i = byte i set)
__device__ __forceinline__ uint32_t bytes_to_nibble(uint32_t b) { return ((b & 0x01010101u) * 0x10204080u) >> 28; }

__device__ __forceinline__ __half2 as_h2(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }

// 8-bit mask of the queries whose accumulated lower bound is above the bound stored at `s_thr`: four u32 words,
// each the fp16 bounds of two queries (2i in the low half, 2i+1 in the high half) — the layout of the accumulators
__device__ __forceinline__ uint32_t dead_mask_th(const __half2 (&acc)[4], const uint4 th) {
  const uint32_t m0 = __hgt2_mask(acc[0], as_h2(th.x));
  const uint32_t m1 = __hgt2_mask(acc[1], as_h2(th.y));
  const uint32_t m2 = __hgt2_mask(acc[2], as_h2(th.z));
  const uint32_t m3 = __hgt2_mask(acc[3], as_h2(th.w));
  return bytes_to_nibble(__byte_perm(m0, m1, 0x6420)) | (bytes_to_nibble(__byte_perm(m2, m3, 0x6420)) << 4);
}
__device__ __forceinline__ uint32_t dead_mask(const __half2 (&acc)[4], uint32_t s_thr) { return dead_mask_th(acc, lds128_volatile(s_thr)); }

// atomic min on one 16-bit half of a shared-memory word (bit patterns of non-negative halves order like integers)
__device__ __forceinline__ void atomic_min_half(uint32_t *w, int hi, uint32_t hbits) {
  uint32_t old = *reinterpret_cast<volatile uint32_t *>(w);
  while (true) {
    const uint32_t cur = hi ? (old >> 16) : (old & 0xFFFFu);
    if (cur <= hbits) return;
    const uint32_t nw = hi ? ((old & 0xFFFFu) | (hbits << 16)) : ((old & 0xFFFF0000u) | hbits);
    const uint32_t prev = atomicCAS(w, old, nw);
    if (prev == old) return;
    old = prev;
  }
}

}  // namespace

// FAST1: the first group has four fields that all start in the row's first 32-bit word (e.g. four 9- or
// 10-bit subspaces) — stage 1 then needs no per-field word selection and no group-size checks.
// TPI = tiles a warp takes per stage-1 iteration.  TPI 1: 32 warps x one row per lane (64 registers per thread).
// TPI 2: 16 warps x two rows per lane — the same rows in flight per SM, but the two independent gather chains of a
// lane hide the shared-memory latency by instruction-level parallelism, the per-iteration work (bounds, loop, prefetch
// addressing) is shared by two tiles, and 128 registers per thread keep the stage-1 program out of the loop.
// B1 > 0: the four leading fields all have width B1 (e.g. 9,9,9,9 of the 256-bit SIFT models) and their tables are
// the first four, back to back, in shared memory: shifts, masks and table offsets of stage 1 are immediates.
// TI: the query tile only visits some clusters of the (cluster-grouped) matrix: a byte per cluster says which of the
// tile's queries do; row tiles of unvisited clusters are skipped, the others start with that mask.
template <int W, bool FAST1, int TPI, int B1, bool TI>
__global__ void __launch_bounds__(TPI == 2 ? 512 : 1024, 1) adc_filter16_scan_kernel(const __grid_constant__ AdcFilter16Args a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int k = a.k;
  // grid: query tiles on x, row chunks on y (all tiles sweep a chunk while it is L2-resident); chunks_fast swaps them
  // so that the chunks of one query tile run concurrently and share their bounds (few queries, development)
  const int qt = a.chunks_fast ? blockIdx.y : blockIdx.x;
  int chunk = a.chunks_fast ? blockIdx.x : blockIdx.y;
  if (a.rot_tile) {          // scan order: the chunk that holds the query tile's nearest rows goes first
    const int nch = (int)(a.chunks_fast ? gridDim.x : gridDim.y);
    chunk += (int)(((int64_t)__ldg(a.rot_tile + qt) - a.tile_lo) / a.chunk_tiles);
    if (chunk >= nch) chunk -= nch;
  }
  const int q0 = qt * T8;

  const size_t lut_bytes = (size_t)a.lut_stride * T8 * sizeof(__half);       // multiple of 64
  uint32_t *thr_f = reinterpret_cast<uint32_t *>(smem_raw + lut_bytes);      // [8] exact k-th distance bits (fp32)
  uint32_t *thr_h = thr_f + 8;                                               // [3][4]: fp16 bits of RU(thr * scale * margin_l), two queries per word
  float *scale_s = reinterpret_cast<float *>(thr_h + 24);                    // [8] scale
  uint32_t *locks = reinterpret_cast<uint32_t *>(scale_s + 8);               // [8]
  uint64_t *lists = reinterpret_cast<uint64_t *>(locks + 8);                 // [8][k] ascending exact keys
  uint64_t *bar = lists + (size_t)T8 * k;
  uint32_t *queues = reinterpret_cast<uint32_t *>(bar + 1);                  // per warp: q1 | q2 | q3, (row << 8) | query mask
  uint8_t *smask = reinterpret_cast<uint8_t *>(queues + (size_t)nwarps * (q1_cap(TPI) + 2 * kQCap));      // [C] (TI only)

  const unsigned char *g16 = reinterpret_cast<const unsigned char *>(a.lut16) + (size_t)qt * lut_bytes;
  const float *g32 = a.lut32 + (size_t)qt * a.lut_stride * T8;
  const int M = a.lay.M;
  const float margin[3] = {1.f + 1.f / 512.f, 1.f + 1.f / 128.f, M <= 32 ? 1.f + 1.f / 32.f : 1.f + 1.f / 8.f};

  // (called by one thread per query) publish a new exact bound: the three fp16 bounds follow the fp32 one
  auto publish_bound = [&](int t, uint32_t bits) {
    atomicMin(thr_f + t, bits);
    const float v = __uint_as_float(bits) * scale_s[t];
#pragma unroll
    for (int l = 0; l < 3; l++) atomic_min_half(thr_h + l * 4 + (t >> 1), t & 1, half_bits_ru(v * margin[l]));
  };

  long long *dbg = a.dbg ? a.dbg + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kDbgSlots : nullptr;
#ifdef VAQGPU_STATS
  // development build (-DVAQGPU_STATS): per-CTA event counts (slots 8..) and warp-0 clocks per level (slots 20..)
  unsigned long long st_cnt[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long st_clk[4] = {0, 0, 0, 0};
#define STAT(i, v) st_cnt[i] += (v)
#else
#define STAT(i, v)
#endif
  if (dbg && tid == 0) dbg[0] = clock64();
  for (int i = tid; i < T8 * k; i += blockDim.x) lists[i] = kEmptyKey;
  // index of a tile slot in the per-query bound arrays: the query's position in the caller's batch (TI searches run
  // the queries in tile order — a permutation that may differ between shards, while the bound arrays are exchanged)
  auto bound_index = [&](int t) -> int {
    const int q = min(q0 + t, a.nq - 1);
    return a.qmap ? a.qmap[q] : q;
  };
  if (tid < T8) {
    const uint32_t g = a.thr_global[bound_index(tid)];
    const float sc = a.scale[q0 + tid];
    scale_s[tid] = sc;
    thr_f[tid] = g;
    // queries past nq (padding of the last tile) get the bound -1.0: every lower bound is above it, so they
    // never survive stage 1 and nothing ever updates the slot
#pragma unroll
    for (int l = 0; l < 3; l++)
      reinterpret_cast<uint16_t *>(thr_h + l * 4)[tid] =
          (uint16_t)((q0 + tid < a.nq) ? half_bits_ru(__uint_as_float(g) * sc * margin[l]) : 0xBC00u);
    locks[tid] = 0u;
  }
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(bar, (uint32_t)lut_bytes);
    for (size_t off = 0; off < lut_bytes; off += 32768) {
      const uint32_t n = (uint32_t)min((size_t)32768, lut_bytes - off);
      tma_bulk_g2s(smem_raw + off, g16 + off, n, bar);
    }
  }
  if constexpr (TI) {
    const uint8_t *gm = a.tmask + (size_t)qt * a.C;
    for (int c = tid; c < a.C; c += blockDim.x) smask[c] = gm[c];
  }
  mbar_wait(bar, 0);
  if constexpr (TI) __syncthreads();
  if (dbg && tid == 0) dbg[1] = clock64();

  // stage-1 program: the first group (<= 4 fields, <= 60 bits, i.e. inside 32-bit words 0..2).  The code is
  // extracted already multiplied by the 16-byte entry size: (row bits >> (shift - 4)) & (mask << 4); a field that
  // starts below bit 4 (only field 0 can, in the FAST1 layout) shifts left instead: funnelshift_r(0 : w, 28 + shift).
  const int G1 = min(4, M);
  uint32_t s1_sh[4], s1_mask[4], s1_addr[4];
  bool s1_hi[4];
  const uint32_t s_base = smem_u32(smem_raw);
  const uint32_t s_thr_h = smem_u32(thr_h);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int f = FAST1 ? i : min(i, G1 - 1);
    const uint32_t meta = a.lay.fmeta[f];
    s1_sh[i] = FAST1 ? (((meta & 31u) + 28u) & 31u) : (meta & 31u);          // FAST1: shift - 4 (mod 32)
    s1_mask[i] = FAST1 ? ((meta >> 16) << 4) : (meta >> 16);
    s1_addr[i] = s_base + a.lay.foff[f] * (T8 * 2);          // shared address of the table
    s1_hi[i] = a.lay.fword[f] != 0;
  }
  const bool two_level = M > 8;
  const int F2 = two_level ? 8 : M;
  // fields [0, F2) end inside the row's first 128-bit word (the field after them starts at or below bit 128)
  const bool l1_one_word = F2 < M ? ((int)a.lay.fword[F2] * 32 + (int)(a.lay.fmeta[F2] & 31u) <= 128) : (W == 1);

  uint32_t *q1 = queues + (size_t)warp * (q1_cap(TPI) + 2 * kQCap);
  uint32_t *q2 = q1 + q1_cap(TPI);
  uint32_t *q3 = q2 + kQCap;
  int q1n = 0, q2n = 0, q3n = 0;
  // Bounds only tighten when the exact level runs.  Waiting for 32 pending rows per queue is right in steady state
  // (full lanes), but under the seed bound only ~0.3 % of the rows reach the exact level: a warp would scan half of a
  // 125 K-row chunk before its first exact pass.  So the deep queues start eager — the exact level runs as soon as
  // one row is pending (four times), the full lower bound as soon as four are — and the thresholds double up to 32.
  int q2_need = 4, q3_need = 1, q3_eager = 4;
  const unsigned lt_mask = (1u << lane) - 1u;

  const int64_t tile_begin = a.tile_lo + (int64_t)chunk * a.chunk_tiles;
  const int64_t tile_end = min(a.tile_hi, tile_begin + a.chunk_tiles);
  const int64_t row_base = tile_begin << 5;
  const uint32_t *codes32 = reinterpret_cast<const uint32_t *>(a.codes);

  // Scan order.  The chunk's tiles are visited in rotated order, as two contiguous segments: [rot, n) first, then
  // [0, rot), where rot is the tile at which the rows nearest to this query tile begin (rows grouped by coarse cluster
  // at index time, queries grouped into tiles by nearest cluster: the first rows scanned already hold near neighbours,
  // so the k-th-best bounds are tight after ~1 % of the chunk instead of converging like k/n).  rot = 0 without a
  // scan order: one segment.
  int rot = 0;
  if (a.rot_tile) {
    const int64_t r = (int64_t)__ldg(a.rot_tile + qt) - tile_begin;
    rot = (r > 0 && r < tile_end - tile_begin) ? (int)r : 0;
  }
  int seg_base = rot;                                    // first tile of the current segment, relative to the chunk
  int n_it = (int)(tile_end - tile_begin) - rot;         // its tiles
  int seg_left = rot;                                    // tiles of the segment still to come
  // this warp's tiles of a segment: warp, warp + nwarps, ... ; an iteration takes TPI consecutive ones of them, the
  // next TPI are prefetched in registers
  uint4 cur[TPI], nxt[TPI];
  const size_t pstep = (size_t)nwarps * W * kTileRows;
  const uint4 *pnext;                                    // first tile of `nxt`
  uint32_t ci_cur = 0u, ci_nxt = 0u;      // TI: cluster of this warp's current / next row tile
  int it;                                 // tile index relative to the segment (a chunk has at most 32768 tiles)
  auto begin_segment = [&]() {
    it = warp;
    pnext = a.codes + ((size_t)(tile_begin + seg_base + warp) * W) * kTileRows + lane + TPI * pstep;
#pragma unroll
    for (int u = 0; u < TPI; u++) {
      cur[u] = nxt[u] = make_uint4(0, 0, 0, 0);
      if (warp + u * nwarps < n_it) cur[u] = ldg_stream_u4(pnext - (TPI - u) * pstep);
      if (warp + (TPI + u) * nwarps < n_it) nxt[u] = ldg_stream_u4(pnext + u * pstep);
    }
    if constexpr (TI) {
      if (warp < n_it) ci_cur = (uint32_t)__ldg(a.tile_cl + tile_begin + seg_base + warp);
      if (warp + nwarps < n_it) ci_nxt = (uint32_t)__ldg(a.tile_cl + tile_begin + seg_base + warp + nwarps);
    }
  };
  begin_segment();
  // bound refresh counter: steps of 2, the low bit set in the lanes that never refresh (no query of the tile in them)
  int refresh = (lane < T8 && q0 + lane < a.nq) ? 0 : 1;
  const uint32_t rows_here = (uint32_t)(min(a.n_rows, tile_end << 5) - row_base);    // valid rows of this chunk

  // ---- bound seeding ---------------------------------------------------------------------------------
  // A CTA that starts without bounds has to score everything it sees exactly.  Each lane scores a few sample
  // rows of the chunk for all eight queries from the fp16 tables and keeps its minimum per query.  Error budget:
  // e16 = RZ_fp16(scale * entry) loses < 2^-10 relative on normal values and < 2^-24 absolute on subnormal ones
  // (scaled entries below 2^-14), the M - 1 half2 additions (round to nearest) lose at most (1 - 2^-11)^(M-1) relative, so
  //     (acc * (1 + (M + 3) * 2^-11) + M * 2^-24) / scale
  // is an UPPER bound of the row's distance (M <= 64: (1 - 2^-11)^-63 * (1 + 2^-10) = 1.0323 < 1.0327).  The lanes' sample rows are
  // distinct, so the k-th smallest of the per-lane minima is the upper bound of k distinct rows' distances, hence
  // bounds the k-th best distance of the chunk.
  // Sample rows per lane: at most a quarter of the chunk.  (A smaller sample on short chunks was measured slower: the
  // looser seed costs more in the first passes than the sampling saves.)  A CTA whose queries all arrive with a bound
  // (later row chunks of a query tile, or bounds already published by other shards) skips the seeding altogether.
  // A search of one or two query tiles over many chunks (the HBM-bound regime) runs all its CTAs at once on the same
  // queries: their exact bounds pool through thr_global within microseconds, so a quarter of the sample is enough.
  const unsigned grid_chunks = a.chunks_fast ? gridDim.x : gridDim.y, grid_qtiles = a.chunks_fast ? gridDim.y : gridDim.x;
  const int spl = (int)min((grid_chunks >= 32u && grid_qtiles <= 2u) ? 1u : (unsigned)a.seed_rows, rows_here / (blockDim.x * 4u));
  int unbounded = 0;
  if (tid < T8 && q0 + tid < a.nq) unbounded = thr_f[tid] == 0xFFFFFFFFu;
  if (a.seed && k <= (int)blockDim.x && spl >= 1 && M <= 64 && __syncthreads_or(unbounded)) {
    const uint32_t step = rows_here / (blockDim.x * (uint32_t)spl);
    __half2 best[4];
    best[0] = best[1] = best[2] = best[3] = as_h2(0x7C007C00u);        // +inf
    for (int j = 0; j < spl; j++) {
      // sample: evenly spaced rows of the chunk; with a scan order the rows the scan starts with (the nearest ones)
      uint32_t rr = (uint32_t)(j * (int)blockDim.x + tid) * step;
      if (a.rot_tile) {
        rr = ((uint32_t)rot << 5) + (uint32_t)(j * (int)blockDim.x + tid);
        if (rr >= rows_here) rr -= rows_here;
      }
      const int64_t row = row_base + (int64_t)rr;
      const uint32_t *rp = codes32 + (((size_t)(row >> 5) * W) * kTileRows + (row & 31)) * 4;
      uint32_t wd[8];
      if constexpr (W <= 2) {           // row words once into registers (same sliding window as the queue passes)
        const uint4 v0 = __ldg(reinterpret_cast<const uint4 *>(rp));
        wd[0] = v0.x; wd[1] = v0.y; wd[2] = v0.z; wd[3] = v0.w;
        if constexpr (W == 2) {
          const uint4 v1 = __ldg(reinterpret_cast<const uint4 *>(rp) + kTileRows);
          wd[4] = v1.x; wd[5] = v1.y; wd[6] = v1.z; wd[7] = v1.w;
        } else {
          wd[4] = wd[5] = wd[6] = wd[7] = 0u;
        }
      }
      __half2 acc[4];
      acc[0] = acc[1] = acc[2] = acc[3] = as_h2(0u);
      if constexpr (W <= 2) {
        // word by word (compile-time register indices), the fields that start in a word in an inner loop
        int f = 0;
#pragma unroll
        for (int w = 0; w < 4 * W; w++) {
          const int few = (int)a.lay.fbeg[w + 1];
          const uint32_t wlo = wd[w], whi = (w + 1 < 8) ? wd[w + 1] : 0u;
#pragma unroll 1
          for (; f < few; f++) {
            const uint32_t meta = a.lay.fmeta[f];
            const uint32_t code = __funnelshift_r(wlo, whi, meta & 31u) & (meta >> 16);
            const uint4 v = lds128(s_base + (a.lay.foff[f] + code) * (T8 * 2));
            acc[0] = __hadd2(acc[0], as_h2(v.x)); acc[1] = __hadd2(acc[1], as_h2(v.y));
            acc[2] = __hadd2(acc[2], as_h2(v.z)); acc[3] = __hadd2(acc[3], as_h2(v.w));
          }
        }
      } else {
        for (int f = 0; f < M; f++) {
          const uint32_t meta = a.lay.fmeta[f];
          const uint32_t lo = __ldg(rp + a.lay.fw_lo[f]);
          const uint32_t hi = __ldg(rp + a.lay.fw_hi[f]);
          const uint32_t code = __funnelshift_r(lo, hi, meta & 31u) & (meta >> 16);
          const uint4 v = lds128(s_base + (a.lay.foff[f] + code) * (T8 * 2));
          acc[0] = __hadd2(acc[0], as_h2(v.x)); acc[1] = __hadd2(acc[1], as_h2(v.y));
          acc[2] = __hadd2(acc[2], as_h2(v.z)); acc[3] = __hadd2(acc[3], as_h2(v.w));
        }
      }
      if constexpr (TI) {
        // a sample row only bounds the queries that visit its cluster: the others see +inf for it
        const uint32_t cm = smask[cluster_of_row(a.cl_start, a.C, row)];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const uint32_t inf2 = (((cm >> (2 * i)) & 1u) ? 0u : 0x7C00u) | (((cm >> (2 * i + 1)) & 1u) ? 0u : 0x7C000000u);
          acc[i] = __hmax2(acc[i], as_h2(inf2));
        }
      }
#pragma unroll
      for (int i = 0; i < 4; i++) best[i] = __hmin2(best[i], acc[i]);
    }
    // per-lane minima -> shared memory as fp16 bit patterns [8][blockDim] (the queues are still empty); non-negative
    // halves order like their bit patterns
    uint16_t *lm = reinterpret_cast<uint16_t *>(queues);
    const int n = (int)blockDim.x;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const uint32_t b = *reinterpret_cast<uint32_t *>(&best[i]);
      lm[(2 * i) * n + tid] = (uint16_t)(b & 0xFFFFu);
      lm[(2 * i + 1) * n + tid] = (uint16_t)(b >> 16);
    }
    __syncthreads();
    if (warp < T8 && q0 + warp < a.nq) {
      // warp t: k-th smallest of query t's n minima by bisection over the 15-bit pattern space
      const uint32_t *col = reinterpret_cast<const uint32_t *>(lm + warp * n);
      uint32_t vals[16];                               // this lane's share of the n <= 1024 minima, two per word
#pragma unroll
      for (int i = 0; i < 16; i++) vals[i] = (lane + 32 * i < n / 2) ? col[lane + 32 * i] : 0x7C007C00u;
      uint32_t lo = 0u, hi = 0x7C00u;                // +inf: fewer than k finite values -> no seed
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) cnt += ((vals[i] & 0xFFFFu) <= mid) + ((vals[i] >> 16) <= mid);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (cnt >= k) hi = mid; else lo = mid + 1;
      }
      if (lane == 0 && lo < 0x7C00u) {
        const float x = __half2float(__ushort_as_half((unsigned short)lo));
        const float ub = (x * (1.f + (float)(M + 3) * 4.8828125e-4f) + (float)M * 5.9604645e-8f) / scale_s[warp] * (1.f + 1e-6f) + 1e-30f;
        if (ub < 3.0e38f) {
          publish_bound(warp, __float_as_uint(ub));
          publish_global_bound(a.thr_global, a.peers, bound_index(warp), __float_as_uint(ub));      // a valid bound for every chunk and shard
        }
      }
    }
    __syncthreads();
  }
  if (dbg && tid == 0) dbg[2] = clock64();
  const uint32_t q1_s = smem_u32(q1);
  int dbg_tail = 0;

  while (true) {
    if (it >= n_it && seg_left > 0) {          // on to the rows before the start tile
      seg_base = 0; n_it = seg_left; seg_left = 0;
      begin_segment();
    }
    if (dbg && it >= n_it && !dbg_tail && tid == 0) { dbg[3] = clock64(); dbg_tail = 1; }
    // ---- what next: a queue that holds a full pass goes first (deepest level first: it feeds nothing further and
    // frees the bounds), otherwise stage 1 streams tiles until the level-1 queue fills, at the end everything drains
    int level = 0, take = 0;
    if (q3n >= q3_need) { level = 3; take = min(q3n, 32); if (q3_eager > 0) q3_eager--; else q3_need = min(a.q3_cap, 2 * q3_need); }
    else if (q2n >= q2_need) { level = 2; take = min(q2n, 32); q2_need = min(32, 2 * q2_need); }
    else if (q1n >= 32) { level = 1; take = 32; }
    else if (it < n_it) {
      // ---- stage 1: tight loop over this warp's tiles ---------------------------------------------------------
      do {
        uint4 w[TPI];
#pragma unroll
        for (int u = 0; u < TPI; u++) { w[u] = cur[u]; cur[u] = nxt[u]; }
        pnext += TPI * pstep;
#pragma unroll
        for (int u = 0; u < TPI; u++)
          if (it + (2 * TPI + u) * nwarps < n_it) nxt[u] = ldg_stream_u4(pnext + u * pstep);
        if (a.l2_prefetch > 0 && (lane & 7) == 0 && it + (2 * TPI + a.l2_prefetch) * nwarps < n_it)
          prefetch_l2(pnext + (size_t)a.l2_prefetch * pstep);      // first words of a later tile, one 128-byte line per 8 lanes
        if (((refresh += 2) & (TPI == 2 ? 31 : 63)) == 0) {
          // pick up bounds published by other row chunks of this query tile (and, row-sharded, by other GPUs)
          const uint32_t g = *reinterpret_cast<volatile uint32_t *>(a.thr_global + bound_index(lane));
          if (g < *reinterpret_cast<volatile uint32_t *>(thr_f + lane)) publish_bound(lane, g);
        }
        const uint4 th = lds128_volatile(s_thr_h);      // the eight stage-1 bounds, two per word like the accumulators
        unsigned sb[TPI];
#pragma unroll
        for (int u = 0; u < TPI; u++) {
          unsigned cm = 0xFFu;          // queries of the tile that visit this row's cluster
          if constexpr (TI) {
            const int itu = it + u * nwarps;
            const uint32_t ci = ci_cur;          // loaded one iteration ago (TPI == 1 in TI mode)
            ci_cur = ci_nxt;
            ci_nxt = itu + 2 * nwarps < n_it ? (uint32_t)__ldg(a.tile_cl + tile_begin + seg_base + itu + 2 * nwarps) : 0u;
            if (ci != 0xFFFFu) cm = smask[ci];
            else cm = smask[cluster_of_row(a.cl_start, a.C, row_base + ((int64_t)(seg_base + itu) << 5) + lane)];
            if (!__any_sync(0xffffffffu, cm != 0u)) { sb[u] = 0u; continue; }          // nobody visits: no gathers
          }
          const uint4 w0 = w[u];
          __half2 acc[4];
#pragma unroll
          for (int i1 = 0; i1 < 4; i1++) {
            if (FAST1 || i1 < G1) {
              uint4 v;
              if constexpr (B1 > 0) {
                constexpr uint32_t MK = ((1u << B1) - 1u) << 4;
                const uint32_t c16 = (i1 == 0 ? (w0.x << 4) : __funnelshift_r(w0.x, w0.y, (uint32_t)(i1 * B1 - 4))) & MK;
                v = lds128(s_base + (uint32_t)i1 * (16u << B1) + c16);
              } else if constexpr (FAST1) {
                // field 0 starts at bit 0 (pair 0 : w0.x), the others at bit >= 4 inside word 0 (pair w0.x : w0.y)
                const uint32_t c16 = (i1 == 0 ? __funnelshift_r(0u, w0.x, s1_sh[0]) : __funnelshift_r(w0.x, w0.y, s1_sh[i1])) & s1_mask[i1];
                v = lds128(s1_addr[i1] + c16);
              } else {
                const uint32_t lo = s1_hi[i1] ? w0.y : w0.x, hi = s1_hi[i1] ? w0.z : w0.y;
                const uint32_t code = __funnelshift_r(lo, hi, s1_sh[i1]) & s1_mask[i1];
                v = lds128(s1_addr[i1] + code * (T8 * 2));
              }
              if (i1 == 0) { acc[0] = as_h2(v.x); acc[1] = as_h2(v.y); acc[2] = as_h2(v.z); acc[3] = as_h2(v.w); }
              else {
                acc[0] = __hadd2(acc[0], as_h2(v.x)); acc[1] = __hadd2(acc[1], as_h2(v.y));
                acc[2] = __hadd2(acc[2], as_h2(v.z)); acc[3] = __hadd2(acc[3], as_h2(v.w));
              }
            }
          }
          sb[u] = ~dead_mask_th(acc, th) & cm;
        }
#pragma unroll
        for (int u = 0; u < TPI; u++) {
          const int itu = it + u * nwarps;
          const uint32_t rel = ((uint32_t)(seg_base + itu) << 5) + lane;
          if (itu >= n_it || rel >= rows_here) sb[u] = 0u;      // past the segment / the partial last tile of the index
          // compact the rows that still have a live query into the warp's queue
          const unsigned m = __ballot_sync(0xffffffffu, sb[u] != 0);
          if (sb[u]) sts32(q1_s + (uint32_t)(q1n + __popc(m & lt_mask)) * 4u, (rel << 8) | sb[u]);
          q1n += __popc(m);
          STAT(0, lane == 0 ? __popc(m) : 0); STAT(9, __popc(sb[u]));
        }
        it += TPI * nwarps;
      } while (it < n_it && q1n < 32);
      continue;
    }
    else if (q1n > 0) { level = 1; take = q1n; }
    else if (q2n > 0) { level = 2; take = q2n; }
    else if (q3n > 0) { level = 3; take = q3n; }
    else break;
    {
      // ---- 32 queued rows, one per lane (single code site for the three levels) -----------------------
      __syncwarp();
#ifdef VAQGPU_STATS
      const long long st_t0 = clock64();
      if (lane == 0) st_cnt[level == 1 ? 1 : level == 2 ? 3 : 5] += 1;
#endif
      const bool active = lane < take;
      uint32_t e = 0u;
      if (level == 1) { if (active) e = q1[q1n - take + lane]; q1n -= take; }
      else if (level == 2) { if (active) e = q2[q2n - take + lane]; q2n -= take; }
      else { if (active) e = q3[q3n - take + lane]; q3n -= take; }
      unsigned mask = e & 0xFFu;                      // queries of the tile still alive for this row (0 when inactive)
      const int64_t row = row_base + (e >> 8);
      const uint32_t *rp = codes32 + (((size_t)(row >> 5) * W) * kTileRows + (row & 31)) * 4;
      // Rows of up to 256 bits are fetched once (W 16-byte loads per lane) and walked from registers: the
      // fields come in increasing bit order, so a two-word window (lo, hi) slides over the eight words and a
      // word is picked by a select tree on the (warp-uniform) word index.  A 4-byte load per field would
      // cost 32 L1 wavefronts each (every lane has a different row).
      uint32_t wd[8];
      if constexpr (W <= 2) {
        const uint4 v0 = __ldg(reinterpret_cast<const uint4 *>(rp));
        wd[0] = v0.x; wd[1] = v0.y; wd[2] = v0.z; wd[3] = v0.w;
        wd[4] = wd[5] = wd[6] = wd[7] = 0u;
        if constexpr (W == 2) {
          // level 1 only walks the leading fields: when they end inside the first 128-bit word, the second one
          // (32 more L1 wavefronts for the warp — every lane has its own row) is not fetched
          if (!(level == 1 && l1_one_word)) {
            const uint4 v1 = __ldg(reinterpret_cast<const uint4 *>(rp) + kTileRows);
            wd[4] = v1.x; wd[5] = v1.y; wd[6] = v1.z; wd[7] = v1.w;
          }
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
      auto field_code = [&](int f) -> uint32_t {      // code of field f of this lane's row
        const uint32_t meta = a.lay.fmeta[f];
        if constexpr (W <= 2) {
          const int fw = a.lay.fword[f];
          if (fw != widx) {
            lo = (fw == widx + 1) ? hi : selw(fw);
            hi = selw(fw + 1);
            widx = fw;
          }
        } else {
          lo = __ldg(rp + a.lay.fw_lo[f]);
          hi = __ldg(rp + a.lay.fw_hi[f]);
        }
        return __funnelshift_r(lo, hi, meta & 31u) & (meta >> 16);
      };

      if (level < 3) {
        // lower bounds of all eight queries over fields [0, fe), packed half2
        const int fe = level == 1 ? F2 : M;
        __half2 acc[4];
        acc[0] = acc[1] = acc[2] = acc[3] = as_h2(0u);
        if constexpr (W <= 2) {
          // walk the row word by word (compile-time register indices), the fields that start in a word in an inner loop
          int f = 0;
#pragma unroll
          for (int w = 0; w < 4 * W; w++) {
            if (f < fe) {
              const int few = min(fe, (int)a.lay.fbeg[w + 1]);
              const uint32_t wlo = wd[w], whi = (w + 1 < 8) ? wd[w + 1] : 0u;
#pragma unroll 1
              for (; f < few; f++) {          // 3-4 fields per word at the usual widths: not worth unrolling (code size)
                const uint32_t meta = a.lay.fmeta[f];
                const uint32_t code = __funnelshift_r(wlo, whi, meta & 31u) & (meta >> 16);
                const uint4 v = lds128(s_base + (a.lay.foff[f] + code) * (T8 * 2));
                acc[0] = __hadd2(acc[0], as_h2(v.x)); acc[1] = __hadd2(acc[1], as_h2(v.y));
                acc[2] = __hadd2(acc[2], as_h2(v.z)); acc[3] = __hadd2(acc[3], as_h2(v.w));
              }
            }
          }
        } else {
          for (int f = 0; f < fe; f++) {
            const uint32_t code = field_code(f);
            const uint4 v = lds128(s_base + (a.lay.foff[f] + code) * (T8 * 2));
            acc[0] = __hadd2(acc[0], as_h2(v.x)); acc[1] = __hadd2(acc[1], as_h2(v.y));
            acc[2] = __hadd2(acc[2], as_h2(v.z)); acc[3] = __hadd2(acc[3], as_h2(v.w));
          }
        }
        mask &= ~dead_mask(acc, s_thr_h + level * 16);
        const unsigned m = __ballot_sync(0xffffffffu, mask != 0);
        const bool to_q2 = level == 1 && two_level;
        if (mask) {
          const uint32_t ne = (e & ~0xFFu) | mask;
          if (to_q2) q2[q2n + __popc(m & lt_mask)] = ne; else q3[q3n + __popc(m & lt_mask)] = ne;
        }
        if (to_q2) q2n += __popc(m); else q3n += __popc(m);
        STAT(level == 1 ? 2 : 4, lane == 0 ? __popc(m) : 0);
      } else {
        // exact distances: four (row, query) pairs per round, eight lanes each.  Lane j of a pair's group takes the
        // subspaces 4j .. 4j+3 (and 4(j+8) .., for M > 32): code and table entry of each are fetched independently — all
        // of a pair's gathers are in flight together (one L2 / DRAM round trip; a lane-per-pair loop has M/4 dependent
        // ones) — and summed in the reference's order and grouping: dism = ((l0+l1)+l2)+l3 inside the lane, dist += dism
        // serially over the lanes of the group (VAQ.cpp:1741-1748).  Insertion stays one pair at a time (the whole warp
        // shifts the list).
        const int sub = lane >> 3, l8 = lane & 7;
        unsigned pend = __ballot_sync(0xffffffffu, mask != 0);
        while (pend) {
          int my_src = -1, my_t = 0;
#pragma unroll
          for (int sl = 0; sl < 4; sl++) {
            if (pend) {          // warp-uniform
              const int src = __ffs(pend) - 1;
              const unsigned msrc = __shfl_sync(0xffffffffu, mask, src);
              if (lane == src) mask &= mask - 1;
              if ((msrc & (msrc - 1u)) == 0u) pend &= pend - 1u;          // that was the last live query of this row
              if (sub == sl) { my_src = src; my_t = __ffs(msrc) - 1; }
            }
          }
          const bool valid = my_src >= 0;
          const int64_t prow = row_base + (__shfl_sync(0xffffffffu, e, valid ? my_src : 0) >> 8);
          const uint32_t *prp = codes32 + (((size_t)(prow >> 5) * W) * kTileRows + (prow & 31)) * 4;
          int32_t rid = (int32_t)prow;
          if (valid && l8 == 0 && a.rowid) rid = (int32_t)__ldg(a.rowid + prow);          // in flight with the gathers
          STAT(6, valid && l8 == 0 ? 1 : 0); STAT(7, lane == 0 ? 1 : 0);
          float dist = 0.f;
          for (int g0 = 0; g0 < M; g0 += 32) {          // 8 groups of 4 subspaces per round
            float dism = 0.f;
            if (valid) {
              float v[4];
#pragma unroll
              for (int j = 0; j < 4; j++) {
                const int f = g0 + 4 * l8 + j;
                v[j] = 0.f;
                if (f < M) {
                  const uint32_t meta = a.lay.fmeta[f];
                  const uint32_t code = __funnelshift_r(__ldg(prp + a.lay.fw_lo[f]), __ldg(prp + a.lay.fw_hi[f]), meta & 31u) & (meta >> 16);
                  v[j] = __ldg(g32 + (size_t)(a.lay.foff[f] + code) * T8 + my_t);
                }
              }
              dism = __fadd_rn(__fadd_rn(__fadd_rn(v[0], v[1]), v[2]), v[3]);          // a partial last group adds exact zeros
            }
            const int ng = min(8, (M - g0 + 3) >> 2);
            for (int j = 0; j < ng; j++) dist = __fadd_rn(dist, __shfl_sync(0xffffffffu, dism, (sub << 3) + j));
          }
          // insertion, one pair at a time: every decision is taken by one observer and broadcast (other warps lower the
          // bounds and the lists concurrently, and the lanes meet again in full-mask shuffles)
#pragma unroll 1
          for (int sl = 0; sl < 4; sl++) {
            const int t = __shfl_sync(0xffffffffu, valid ? my_t : -1, sl << 3);
            if (t < 0) break;          // slots fill in order
            const float d = __shfl_sync(0xffffffffu, dist, sl << 3);
            const int32_t id = __shfl_sync(0xffffffffu, rid, sl << 3);
            const float thr = __uint_as_float(__shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t *>(thr_f + t), 0));
            if (d > thr) continue;
            // keys carry the ORIGINAL row index (the storage order is conflict-aware, layout.cu): canonical (distance, id) order
            const uint64_t kk = make_key_f32(d, id);
            volatile uint64_t *lst = lists + (size_t)t * k;
            {
              uint64_t c = lst[k - 1];
              c = __shfl_sync(0xffffffffu, c, 0);
              if (!(kk < c)) continue;
            }
            STAT(8, lane == 0 ? 1 : 0);
            // take the query's list lock; give up as soon as the list has moved past this candidate
            int got = 0;
            if (lane == 0) {
              while (true) {
                if (!(kk < lst[k - 1])) break;
                if (atomicCAS(locks + t, 0u, 1u) == 0u) { got = 1; break; }
                __nanosleep(100);
              }
            }
            got = __shfl_sync(0xffffffffu, got, 0);
            if (!got) continue;
            const uint64_t before = lst[k - 1];
            const uint64_t kth = warp_list_insert(lst, k, kk, lane);
            __syncwarp();
            if (lane == 0) {
              __threadfence_block();
              atomicExch(locks + t, 0u);
              if (kth != before && kth != kEmptyKey) {
                const uint32_t bits = (uint32_t)(kth >> 32);
                publish_bound(t, bits);
                publish_global_bound(a.thr_global, a.peers, bound_index(t), bits);
              }
            }
          }
        }
      }
#ifdef VAQGPU_STATS
      st_clk[level] += clock64() - st_t0;
#endif
    }
  }
#ifdef VAQGPU_STATS
  if (dbg) {
    for (int i = 0; i < 10; i++) if (st_cnt[i]) atomicAdd(reinterpret_cast<unsigned long long *>(dbg) + 8 + i, st_cnt[i]);
    if (tid == 0) for (int i = 1; i < 4; i++) dbg[20 + i] = st_clk[i];
  }
#endif

  // ---- CTA epilogue: publish this (query tile, chunk)'s keys -----------------------------------------
  if (dbg && tid == 0) dbg[4] = clock64();
  __syncthreads();
  if (dbg && tid == 0) dbg[5] = clock64();
  for (int i = tid; i < T8 * k; i += blockDim.x) {
    const int t = i / k, j = i - t * k;
    const int q = q0 + t;
    if (q < a.nq) a.out_keys[((size_t)q * a.out_slots + a.slot_base + chunk) * k + j] = lists[i];
  }
}

size_t adc_filter16_smem_bytes(int lut_stride, int k, int threads, int ti_clusters) {
  const int nwarps = threads / 32, tpi = threads <= 512 ? 2 : 1;
  size_t b = (size_t)lut_stride * T8 * 2 + 6 * 32;          // tables + bounds/scales/locks
  b += ((size_t)T8 * k + 1) * sizeof(uint64_t);
  b += (size_t)nwarps * (q1_cap(tpi) + 2 * kQCap) * sizeof(uint32_t);
  b += ((size_t)ti_clusters + 15) & ~(size_t)15;          // TI: one mask byte per cluster
  return b;
}

template <int W, bool FAST1, int TPI, int B1, bool TI = false>
static cudaError_t launch16_wftb(const AdcFilter16Args &a, int threads, size_t smem_bytes, cudaStream_t st) {
  static SmemOptIn optin;
  {
    cudaError_t e = optin.ensure(adc_filter16_scan_kernel<W, FAST1, TPI, B1, TI>, smem_bytes);
    if (e != cudaSuccess) return e;
  }
  const int64_t nt = a.tile_hi - a.tile_lo;
  if (nt <= 0) return cudaSuccess;
  dim3 grid((unsigned)((a.nq + T8 - 1) / T8), (unsigned)((nt + a.chunk_tiles - 1) / a.chunk_tiles));
  if (a.chunks_fast) { const unsigned t = grid.x; grid.x = grid.y; grid.y = t; }
  adc_filter16_scan_kernel<W, FAST1, TPI, B1, TI><<<grid, threads, smem_bytes, st>>>(a);
  return cudaGetLastError();
}

// threads <= 512: two tiles per warp and iteration (TPI 2); more threads: one
template <int W, bool FAST1>
static cudaError_t launch16_wf(const AdcFilter16Args &a, int threads, size_t smem_bytes, cudaStream_t st, int b1) {
  if (a.tmask) {          // TI / visit
    if constexpr (FAST1 && W <= 2) {
      if (b1 == 9) return launch16_wftb<W, true, 1, 9, true>(a, 1024, smem_bytes, st);
      if (b1 == 10) return launch16_wftb<W, true, 1, 10, true>(a, 1024, smem_bytes, st);
    }
    return launch16_wftb<W, FAST1, 1, 0, true>(a, 1024, smem_bytes, st);
  }
  if (threads <= 512) return launch16_wftb<W, FAST1, 2, 0>(a, threads, smem_bytes, st);
  if constexpr (FAST1 && W <= 2) {
    switch (b1) {          // uniform leading widths met in practice (budget / subspaces around 8 bits)
      case 7: return launch16_wftb<W, true, 1, 7>(a, threads, smem_bytes, st);
      case 8: return launch16_wftb<W, true, 1, 8>(a, threads, smem_bytes, st);
      case 9: return launch16_wftb<W, true, 1, 9>(a, threads, smem_bytes, st);
      case 10: return launch16_wftb<W, true, 1, 10>(a, threads, smem_bytes, st);
      default: break;
    }
  }
  return launch16_wftb<W, FAST1, 1, 0>(a, threads, smem_bytes, st);
}

template <int W>
static cudaError_t launch16_w(const AdcFilter16Args &a, int threads, size_t smem_bytes, cudaStream_t st) {
  // FAST1: four leading fields that all start in the row's first 32-bit word, the first one at least 4 bits wide
  bool fast1 = a.lay.M >= 4;
  for (int f = 0; f < 4 && fast1; f++) fast1 = a.lay.fword[f] == 0;
  if (fast1) fast1 = (a.lay.fmeta[1] & 31u) >= 4u;
  // B1: ... all of the same width, their tables first and back to back
  int b1 = 0;
  if (fast1) {
    const uint32_t mask = a.lay.fmeta[0] >> 16;
    b1 = __builtin_popcount(mask);
    for (int f = 0; f < 4; f++)
      if ((a.lay.fmeta[f] >> 16) != mask || a.lay.foff[f] != (uint32_t)f << b1) b1 = 0;
  }
  return fast1 ? launch16_wf<W, true>(a, threads, smem_bytes, st, b1) : launch16_wf<W, false>(a, threads, smem_bytes, st, 0);
}

cudaError_t launch_adc_filter16_scan(const AdcFilter16Args &a, int threads, size_t smem_bytes, cudaStream_t st) {
  switch (a.lay.W) {
    case 1: return launch16_w<1>(a, threads, smem_bytes, st);
    case 2: return launch16_w<2>(a, threads, smem_bytes, st);
    case 3: return launch16_w<3>(a, threads, smem_bytes, st);
    case 4: return launch16_w<4>(a, threads, smem_bytes, st);
    case 5: return launch16_w<5>(a, threads, smem_bytes, st);
    case 6: return launch16_w<6>(a, threads, smem_bytes, st);
    case 7: return launch16_w<7>(a, threads, smem_bytes, st);
    case 8: return launch16_w<8>(a, threads, smem_bytes, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace vaqgpu
