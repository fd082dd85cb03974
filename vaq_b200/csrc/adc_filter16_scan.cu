// ADC scan, filter-and-refine with half-precision lower-bound tables — the default kernel behind
// VAQ::searchEarlyAbandon (reference bitvecengine/VAQ.cpp:1694-1727) when eight queries' tables fit in
// shared memory as fp16.
//
// Same structure as adc_filter_scan.cu (stage 1 over every (row, query) pair on the first group of
// subspaces, compacted survivors scored by full-lane warps), but the shared-memory tables hold
//     e16[s][c][t] = round_toward_zero_fp16( scale_t * lut[t][s][c] ),     t = 0..7 (query tile of 8)
// (written by lut_build_kernel<8> together with the fp32 tables; scale_t is a per-query power of two)
// i.e. guaranteed LOWER bounds of the reference's table entries, 16 bytes per code for 8 queries.  One
// LDS.128 therefore serves eight queries (the fp32 form serves four), which halves both the shared-memory
// wavefronts and the instructions per pair — the two resources the fp32 form saturates (ncu: LSU 86 %,
// issue 74 %).  Because the bounds are conservative, pruning stays exact:
//
//   stage 1   acc = e16_0 + e16_1 + e16_2 + e16_3 in packed half2 arithmetic (round-to-nearest: at most
//             (1+2^-11)^3 above the real sum), pruned iff acc > RU_fp16(thr * scale * (1 + 2^-9)).  Then the real
//             partial sum exceeds thr by > 2^-11 relative, far more than the 40 * 2^-24 by which the fp32
//             distance the reference computes can fall below the real sum  =>  the reference's own distance
//             is > thr and the row cannot be among the k best.
//   level 1/2 per-lane lower bound over groups 1-2, then over all subspaces, accumulated in fp32 from the
//             same fp16 entries, pruned iff LB > thr * scale * (1 + 2^-9).
//   level 3   the few pairs whose full lower bound is still under the bound are scored EXACTLY from the
//             fp32 tables in global memory (L2), in the reference's order and grouping
//             (dism = ((l0+l1)+l2)+l3 ; dist += dism, VAQ.cpp:1741-1748), and only these exact distances
//             enter the top-k lists and tighten the bounds.
//
// Result: bit-identical to the fp32 kernels (tests/test_gpu_vaq.py runs all three against the oracle).
#include <cuda_fp16.h>

#include "common.cuh"

namespace vaqgpu {

namespace {

constexpr int T8 = 8;
constexpr int kQ1Cap16 = 32 + 32 * T8;   // 31 pending + one tile's pushes for 8 queries
constexpr int kQCap = 64;                // level-2 / level-3 queues: 31 pending + one drain
constexpr float kMargin = 1.0f + 1.0f / 512.0f;

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
// four 0xFF/0x00 flag bytes -> 4-bit mask (bit i = byte i set)
__device__ __forceinline__ uint32_t bytes_to_nibble(uint32_t b) { return ((b & 0x01010101u) * 0x10204080u) >> 28; }

__device__ __forceinline__ __half2 as_h2(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }

}  // namespace

// FAST1: the first group has four fields that all start in the row's first 32-bit word (e.g. four 9- or
// 10-bit subspaces) — stage 1 then needs no per-field word selection and no group-size checks.
template <int W, bool FAST1>
__global__ void __launch_bounds__(1024, 1) adc_filter16_scan_kernel(const __grid_constant__ AdcFilter16Args a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int k = a.k;
  const int qt = blockIdx.x, chunk = blockIdx.y;
  const int q0 = qt * T8;

  const size_t lut_bytes = (size_t)a.lut_stride * T8 * sizeof(__half);       // multiple of 64
  const __half *slut = reinterpret_cast<const __half *>(smem_raw);
  uint32_t *thr_f = reinterpret_cast<uint32_t *>(smem_raw + lut_bytes);      // [8] exact k-th distance bits (fp32)
  uint32_t *thr_h = thr_f + 8;                                               // [8] fp16 bits of RU(thr * scale * margin)
  float *scale_m = reinterpret_cast<float *>(thr_h + 8);                     // [8] scale * margin
  uint32_t *locks = reinterpret_cast<uint32_t *>(scale_m + 8);               // [8]
  uint64_t *lists = reinterpret_cast<uint64_t *>(locks + 8);                 // [8][k] ascending exact keys
  uint64_t *bar = lists + (size_t)T8 * k;
  uint32_t *queues = reinterpret_cast<uint32_t *>(bar + 1);                  // per warp: q1 | q2e | q2d | q3

  const unsigned char *g16 = reinterpret_cast<const unsigned char *>(a.lut16) + (size_t)qt * lut_bytes;
  const float *g32 = a.lut32 + (size_t)qt * a.lut_stride * T8;

  for (int i = tid; i < T8 * k; i += blockDim.x) lists[i] = kEmptyKey;
  if (tid < T8) {
    const int q = min(q0 + tid, a.nq - 1);
    const float sm = a.scale[q0 + tid] * kMargin;
    const uint32_t g = a.thr_global[q];
    scale_m[tid] = sm;
    thr_f[tid] = g;
    // queries past nq (padding of the last tile) get the bound -1.0: every lower bound is above it, so they
    // never survive stage 1 and nothing ever updates the slot
    thr_h[tid] = (q0 + tid < a.nq) ? half_bits_ru(__uint_as_float(g) * sm) : 0xBC00u;
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
  mbar_wait(bar, 0);

  // stage-1 program: the first group (<= 4 fields, <= 60 bits, i.e. inside 32-bit words 0..2)
  const int M = a.lay.M;
  const int G1 = min(4, M);
  uint32_t s1_sh[4], s1_mask[4], s1_off[4];
  bool s1_hi[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int f = min(i, G1 - 1);
    const uint32_t meta = a.lay.fmeta[f];
    s1_sh[i] = meta & 31u;
    s1_mask[i] = meta >> 16;
    s1_off[i] = a.lay.foff[f] * (T8 * 2);          // byte offset of the table
    s1_hi[i] = a.lay.fword[f] != 0;
  }
  const bool two_level = M > 8;
  const int F2 = two_level ? 8 : M;

  uint32_t *q1 = queues + (size_t)warp * (kQ1Cap16 + 3 * kQCap);
  uint32_t *q2e = q1 + kQ1Cap16;
  float *q2d = reinterpret_cast<float *>(q2e + kQCap);
  uint32_t *q3 = q2e + 2 * kQCap;
  int q1n = 0, q2n = 0, q3n = 0;
  const unsigned lt_mask = (1u << lane) - 1u;

  const int64_t tile_begin = a.tile_lo + (int64_t)chunk * a.chunk_tiles;
  const int64_t tile_end = min(a.tile_hi, tile_begin + a.chunk_tiles);
  const int64_t row_base = tile_begin << 5;
  const uint32_t *codes32 = reinterpret_cast<const uint32_t *>(a.codes);

  int64_t tl = tile_begin + warp;
  uint4 cur = make_uint4(0, 0, 0, 0), nxt = cur;
  const uint4 *pnext = a.codes + ((size_t)(tl + nwarps) * W) * kTileRows + lane;     // tile whose words `nxt` holds
  const size_t pstep = (size_t)nwarps * W * kTileRows;
  if (tl < tile_end) cur = ldg_stream_u4(pnext - pstep);
  if (tl + nwarps < tile_end) nxt = ldg_stream_u4(pnext);
  int refresh = 0;
  const uint32_t rows_here = (uint32_t)(min(a.n_rows, tile_end << 5) - row_base);    // valid rows of this chunk
  const uint32_t s_base = smem_u32(smem_raw);
  const uint32_t s_thr_h = smem_u32(thr_h);
  uint32_t s1_addr[4];
#pragma unroll
  for (int i = 0; i < 4; i++) s1_addr[i] = s_base + s1_off[i];

  while (true) {
    const bool more = tl < tile_end;
    int level = 0, take = 0;
    if (((q1n | q2n | q3n) >= 32) || !more) {           // rarely true: keep the common path to one test
      if (q3n >= 32) { level = 3; take = 32; }
      else if (q2n >= 32) { level = 2; take = 32; }
      else if (q1n >= 32) { level = 1; take = 32; }
      else if (!more) {
        if (q1n > 0) { level = 1; take = q1n; }
        else if (q2n > 0) { level = 2; take = q2n; }
        else if (q3n > 0) { level = 3; take = q3n; }
        else break;
      }
    }
    if (level) {
      // ---- survivors: one code site for the two lower-bound levels and the exact level ----------------
      __syncwarp();
      const bool active = lane < take;
      uint32_t e = 0u;
      float dist = 0.f;
      int fb = 0, fe = M;
      if (level == 1) {
        if (active) e = q1[q1n - take + lane];
        q1n -= take;
        fe = F2;
      } else if (level == 2) {
        if (active) { e = q2e[q2n - take + lane]; dist = q2d[q2n - take + lane]; }
        q2n -= take;
        fb = F2;
      } else {
        if (active) e = q3[q3n - take + lane];
        q3n -= take;
      }
      const bool exact = level == 3;
      const int t = (int)(e & 7u);
      const int64_t row = row_base + (e >> 3);
      const uint32_t *rp = codes32 + (((size_t)(row >> 5) * W) * kTileRows + (row & 31)) * 4;
      float thr = __uint_as_float(*reinterpret_cast<volatile uint32_t *>(thr_f + t));
      if (!exact) thr *= scale_m[t];
      // Rows of up to 256 bits are fetched once (W 16-byte loads per lane) and walked from registers: the
      // fields come in increasing bit order, so a two-word window (lo, hi) slides over the eight words and a
      // word is picked by a select tree on the (warp-uniform) word index.  A 4-byte load per field would
      // cost 32 L1 wavefronts each (every lane has a different row).
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
      bool alive = true;
      for (int g = fb; g < fe; g += 4) {
        float dism = 0.f;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int f = g + j;
          if (f < fe) {
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
            const uint32_t code = __funnelshift_r(lo, hi, meta & 31u) & (meta >> 16);
            const uint32_t idx = (a.lay.foff[f] + code) * T8 + t;
            dism += exact ? __ldg(g32 + idx) : __half2float(slut[idx]);
          }
        }
        dist += dism;
        if (__all_sync(0xffffffffu, !active || (dist > thr))) { alive = false; break; }
      }
      if (alive) {
        if (!exact) {
          // still under the bound: next level
          const bool s = active && !(dist > thr);
          const unsigned m = __ballot_sync(0xffffffffu, s);
          const bool to_q2 = level == 1 && two_level;
          if (s) {
            if (to_q2) {
              const int pos = q2n + __popc(m & lt_mask);
              q2e[pos] = e;
              q2d[pos] = dist;
            } else {
              q3[q3n + __popc(m & lt_mask)] = e;
            }
          }
          if (to_q2) q2n += __popc(m); else q3n += __popc(m);
        } else {
          const uint64_t key = active ? make_key_f32(dist, (int32_t)row) : kEmptyKey;
          const uint64_t kth0 = active ? *reinterpret_cast<volatile uint64_t *>(lists + (size_t)t * k + (k - 1)) : 0ull;
          unsigned m = __ballot_sync(0xffffffffu, key < kth0);
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const uint64_t kk = __shfl_sync(0xffffffffu, key, src);
            const int tt = __shfl_sync(0xffffffffu, t, src);
            volatile uint64_t *lst = lists + (size_t)tt * k;
            {
              uint64_t c = lst[k - 1];
              c = __shfl_sync(0xffffffffu, c, 0);      // one observer: the decision must be warp-uniform
              if (!(kk < c)) continue;
            }
            if (lane == 0) while (atomicCAS(locks + tt, 0u, 1u) != 0u) __nanosleep(200);
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
                atomicMin(thr_h + tt, half_bits_ru(__uint_as_float(bits) * scale_m[tt]));
                if (q0 + tt < a.nq) atomicMin(a.thr_global + q0 + tt, bits);
              }
            }
          }
        }
      }
      continue;
    }

    // ---- stage 1 on one tile ---------------------------------------------------------------------------
    const uint4 w0 = cur;
    cur = nxt;
    pnext += pstep;
    if (tl + 2 * (int64_t)nwarps < tile_end) nxt = ldg_stream_u4(pnext);
    if (((++refresh) & 31) == 0 && lane < T8 && q0 + lane < a.nq) {
      // pick up bounds published by other row chunks of this query tile
      const uint32_t g = *reinterpret_cast<volatile uint32_t *>(a.thr_global + q0 + lane);
      if (g < *reinterpret_cast<volatile uint32_t *>(thr_f + lane)) {
        atomicMin(thr_f + lane, g);
        atomicMin(thr_h + lane, half_bits_ru(__uint_as_float(g) * scale_m[lane]));
      }
    }
    const uint4 th0 = lds128_volatile(s_thr_h), th1 = lds128_volatile(s_thr_h + 16);
    __half2 acc[4];
#pragma unroll
    for (int i1 = 0; i1 < 4; i1++) {
      if (FAST1 || i1 < G1) {
        const uint32_t lo = (!FAST1 && s1_hi[i1]) ? w0.y : w0.x, hi = (!FAST1 && s1_hi[i1]) ? w0.z : w0.y;
        const uint32_t code = __funnelshift_r(lo, hi, s1_sh[i1]) & s1_mask[i1];
        const uint4 v = lds128(s1_addr[i1] + code * (T8 * 2));
        if (i1 == 0) { acc[0] = as_h2(v.x); acc[1] = as_h2(v.y); acc[2] = as_h2(v.z); acc[3] = as_h2(v.w); }
        else {
          acc[0] = __hadd2(acc[0], as_h2(v.x)); acc[1] = __hadd2(acc[1], as_h2(v.y));
          acc[2] = __hadd2(acc[2], as_h2(v.z)); acc[3] = __hadd2(acc[3], as_h2(v.w));
        }
      }
    }
    // per query: is the partial lower bound above the bound?  (0xFFFF per half that is)
    const uint32_t m0 = __hgt2_mask(acc[0], as_h2(__byte_perm(th0.x, th0.y, 0x5410)));
    const uint32_t m1 = __hgt2_mask(acc[1], as_h2(__byte_perm(th0.z, th0.w, 0x5410)));
    const uint32_t m2 = __hgt2_mask(acc[2], as_h2(__byte_perm(th1.x, th1.y, 0x5410)));
    const uint32_t m3 = __hgt2_mask(acc[3], as_h2(__byte_perm(th1.z, th1.w, 0x5410)));
    const uint32_t dead = bytes_to_nibble(__byte_perm(m0, m1, 0x6420)) | (bytes_to_nibble(__byte_perm(m2, m3, 0x6420)) << 4);
    const uint32_t rel = ((uint32_t)(tl - tile_begin) << 5) + lane;
    unsigned sb = ~dead & 0xFFu;
    if (tl == tile_end - 1 && rel >= rows_here) sb = 0u;      // only the last tile of the index can be partial
    if (__any_sync(0xffffffffu, sb != 0)) {
      // compact the surviving (row, query) pairs into the warp's queue: exclusive scan of the per-lane counts
      const int cnt = __popc(sb);
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      int pos = q1n + incl - cnt;
      while (sb) {
        const int t = __ffs(sb) - 1;
        sb &= sb - 1;
        q1[pos++] = (rel << 3) | (uint32_t)t;
      }
      q1n += __shfl_sync(0xffffffffu, incl, 31);
    }
    tl += nwarps;
  }

  // ---- CTA epilogue: publish this (query tile, chunk)'s keys -----------------------------------------
  __syncthreads();
  for (int i = tid; i < T8 * k; i += blockDim.x) {
    const int t = i / k, j = i - t * k;
    const int q = q0 + t;
    if (q < a.nq) a.out_keys[((size_t)q * a.out_slots + a.slot_base + chunk) * k + j] = lists[i];
  }
}

size_t adc_filter16_smem_bytes(int lut_stride, int k, int threads) {
  const int nwarps = threads / 32;
  size_t b = (size_t)lut_stride * T8 * 2 + 4 * 32;
  b += ((size_t)T8 * k + 1) * sizeof(uint64_t);
  b += (size_t)nwarps * (kQ1Cap16 + 3 * kQCap) * sizeof(uint32_t);
  return b;
}

template <int W, bool FAST1>
static cudaError_t launch16_wf(const AdcFilter16Args &a, int threads, size_t smem_bytes, cudaStream_t st) {
  static size_t configured = 0;
  if (smem_bytes > configured) {
    cudaError_t e = cudaFuncSetAttribute(adc_filter16_scan_kernel<W, FAST1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    configured = smem_bytes;
  }
  const int64_t nt = a.tile_hi - a.tile_lo;
  if (nt <= 0) return cudaSuccess;
  dim3 grid((unsigned)((a.nq + T8 - 1) / T8), (unsigned)((nt + a.chunk_tiles - 1) / a.chunk_tiles));
  adc_filter16_scan_kernel<W, FAST1><<<grid, threads, smem_bytes, st>>>(a);
  return cudaGetLastError();
}

template <int W>
static cudaError_t launch16_w(const AdcFilter16Args &a, int threads, size_t smem_bytes, cudaStream_t st) {
  bool fast1 = a.lay.M >= 4;
  for (int f = 0; f < 4 && fast1; f++) fast1 = a.lay.fword[f] == 0;
  return fast1 ? launch16_wf<W, true>(a, threads, smem_bytes, st) : launch16_wf<W, false>(a, threads, smem_bytes, st);
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
