// Hamming scan — replaces the brute-force loops of BitVecEngine::query / queryParallel
// (reference bitvecengine/BitVecEngine.cpp:61-197, 509-519, 1264-1304) and hammingDist
// (utils/DistanceFunctions.hpp:164-172: sum_w popcount(q[w] ^ x[w])).
//
// Same streaming skeleton as the ADC scan: bit vectors are stored as 32-row tiles of uint4
// words ([tile][word][lane]), one row per lane, coalesced 128-bit loads with the next tile
// prefetched; the CTA's QT queries sit in shared memory and are read as broadcast uint4s, so a
// row fetched from HBM once is compared against QT queries with popc over its 32-bit words.
// Each (warp, query) keeps a sorted top-k list keyed by (distance << 32 | row); the reference's
// four query methods differ only in how equal distances are ordered — this
// kernel defines the order as ascending (distance, row), which reproduces the known answers
// of test/test-bitvecengine.cpp:77-79, 177-179, 258-260.
#include "common.cuh"

namespace vaqgpu {

// popcount of the XOR of two W*128-bit vectors.  POPC issues at a quarter of the ALU rate, so the words are
// first compressed with carry-save adders (Harley-Seal: 3 words -> sum + carry, two LOP3 each): per eight
// 32-bit words 4 POPC + ~14 LOP3 instead of 8 POPC.  The reference's loop is
// sum_w popcountl(q[w] ^ x[w]) (utils/DistanceFunctions.hpp:164-172); the result is the same integer.
__device__ __forceinline__ void csa(uint32_t a, uint32_t b, uint32_t c, uint32_t &sum, uint32_t &carry) {
  sum = a ^ b ^ c;
  carry = (a & b) | (c & (a ^ b));
}

__device__ __forceinline__ uint32_t popc8(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t x4, uint32_t x5,
                                          uint32_t x6, uint32_t x7) {
  uint32_t s1, c1, s2, c2, s3, c3;
  csa(x0, x1, x2, s1, c1);
  csa(x3, x4, x5, s2, c2);
  csa(x6, x7, s1, s3, c3);
  const uint32_t ones = s2 ^ s3, c4 = s2 & s3;
  uint32_t t1, d1;
  csa(c1, c2, c3, t1, d1);
  const uint32_t twos = t1 ^ c4, d2 = t1 & c4;
  const uint32_t fours = d1 ^ d2, eights = d1 & d2;
  return __popc(ones) + 2 * __popc(twos) + 4 * __popc(fours) + 8 * __popc(eights);
}

template <int W>
__device__ __forceinline__ uint32_t hamming_words(const uint4 (&r)[W], const uint4 *__restrict__ q) {
  uint32_t d = 0;
#pragma unroll
  for (int j = 0; j + 1 < W; j += 2) {
    const uint4 a = q[j], b = q[j + 1];
    d += popc8(r[j].x ^ a.x, r[j].y ^ a.y, r[j].z ^ a.z, r[j].w ^ a.w, r[j + 1].x ^ b.x, r[j + 1].y ^ b.y, r[j + 1].z ^ b.z,
               r[j + 1].w ^ b.w);
  }
  if (W & 1) {
    const uint4 a = q[W - 1];
    d += __popc(r[W - 1].x ^ a.x) + __popc(r[W - 1].y ^ a.y) + __popc(r[W - 1].z ^ a.z) + __popc(r[W - 1].w ^ a.w);
  }
  return d;
}

template <int W, int QT>
__global__ void __launch_bounds__(256, (W <= 2 ? (QT == 1 ? 6 : 4) : 2)) ham_scan_kernel(const __grid_constant__ HamScanArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int k = a.k;
  const int q0 = blockIdx.y * QT, split = blockIdx.x;
  const int nqt = min(QT, a.nq - q0);

  uint4 *sq = reinterpret_cast<uint4 *>(smem_raw);                       // [QT][W]
  uint64_t *lists = reinterpret_cast<uint64_t *>(sq + QT * W);           // [nwarps][QT][k]
  uint64_t *merged = lists + (size_t)nwarps * QT * k;                    // [k]
  uint64_t *blk_thr = merged + k;                                        // [QT]

  for (int i = tid; i < QT * W; i += blockDim.x) {
    const int t = i / W;
    sq[i] = t < nqt ? a.queries[(size_t)(q0 + t) * W + (i - t * W)] : make_uint4(0, 0, 0, 0);
  }
  for (int i = tid; i < nwarps * QT * k; i += blockDim.x) lists[i] = kEmptyKey;
  if (tid < QT) blk_thr[tid] = kEmptyKey;
  __syncthreads();

  const int n_tiles = (int)((a.n_rows + kTileRows - 1) >> 5);
  const int g = split * nwarps + warp, gstride = a.splits * nwarps;
  int t = g;
  uint4 cur[W], nxt[W];
  auto load = [&](uint4(&dst)[W], int tile) {
    const uint4 *p = a.codes + ((size_t)tile * W) * kTileRows + lane;
#pragma unroll
    for (int j = 0; j < W; j++) dst[j] = ldg_stream_u4(p + j * kTileRows);
  };
  if (t < n_tiles) load(cur, t);
  // distance part of each query's running k-th key, kept in registers: the per-row test is one compare;
  // the exact (distance, row) key test and the list update only run for rows that pass it
  uint32_t thr_d[QT];
#pragma unroll
  for (int qi = 0; qi < QT; qi++) thr_d[qi] = 0xFFFFFFFFu;
  int iter = 0;
  for (; t < n_tiles; t += gstride, iter++) {
    const int tn = t + gstride;
    if (tn < n_tiles) load(nxt, tn);
    {
      // the register prefetch covers one tile; DRAM latency needs more bytes in flight per SM than that: pull the
      // tile this warp reads six iterations from now into L2 (one 128-byte line per 8 lanes: lanes 0, 8, 16, 24 ask)
      const int tp = t + 6 * gstride;
      if (tp < n_tiles && (lane & 7) == 0) {
        const uint4 *pp = a.codes + ((size_t)tp * W) * kTileRows + lane;
#pragma unroll
        for (int j = 0; j < W; j++) prefetch_l2(pp + j * kTileRows);
      }
    }
    const int row = t * kTileRows + lane;
    const bool valid = row < a.n_rows;
    if ((iter & 15) == 15) {
      // adopt tighter bounds other warps of the CTA have reached
#pragma unroll
      for (int qi = 0; qi < QT; qi++)
        thr_d[qi] = min(thr_d[qi], (uint32_t)(*reinterpret_cast<volatile uint64_t *>(blk_thr + qi) >> 32));
    }
#pragma unroll
    for (int qi = 0; qi < QT; qi++) {
      if (qi < nqt) {
        const uint32_t d = hamming_words<W>(cur, sq + qi * W);
        if (__any_sync(0xffffffffu, d <= thr_d[qi])) {
          volatile uint64_t *mylist = lists + ((size_t)warp * QT + qi) * k;
          uint64_t thrkey = mylist[k - 1];
          const uint64_t bthr = *reinterpret_cast<volatile uint64_t *>(blk_thr + qi);
          thrkey = bthr < thrkey ? bthr : thrkey;
          const uint64_t key = valid ? (((uint64_t)d << 32) | (uint32_t)row) : kEmptyKey;
          unsigned m = __ballot_sync(0xffffffffu, key < thrkey);
          uint64_t kth = thrkey;
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const uint64_t kk = __shfl_sync(0xffffffffu, key, src);
            kth = warp_list_insert(mylist, k, kk, lane);
          }
          if (kth < thrkey && lane == 0) atomicMin(reinterpret_cast<unsigned long long *>(blk_thr + qi), (unsigned long long)kth);
          thr_d[qi] = (uint32_t)(kth >> 32);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < W; j++) cur[j] = nxt[j];
  }

  __syncthreads();
  for (int qi = 0; qi < nqt; qi++) {
    // gather this query's nwarps lists: they are strided by QT*k, merge in place via rank scatter
    for (int i = tid; i < k; i += blockDim.x) merged[i] = kEmptyKey;
    __syncthreads();
    const int total = nwarps * k;
    for (int e = tid; e < total; e += blockDim.x) {
      const int l = e / k, i = e - l * k;
      const uint64_t key = lists[((size_t)l * QT + qi) * k + i];
      if (key == kEmptyKey) continue;
      int rank = i;
      for (int o = 0; o < nwarps && rank < k; o++) {
        if (o == l) continue;
        rank += lower_bound_u64(lists + ((size_t)o * QT + qi) * k, k, key);
      }
      if (rank < k) merged[rank] = key;
    }
    __syncthreads();
    uint64_t *out = a.out_keys + ((size_t)(q0 + qi) * a.splits + split) * k;
    for (int i = tid; i < k; i += blockDim.x) out[i] = merged[i];
    __syncthreads();
  }
}

template <int W, int QT>
static cudaError_t launch_wq(const HamScanArgs &a, int threads, size_t smem_bytes, cudaStream_t st) {
  static SmemOptIn optin;
  {
    cudaError_t e = optin.ensure(ham_scan_kernel<W, QT>, smem_bytes);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((unsigned)a.splits, (unsigned)((a.nq + QT - 1) / QT));
  ham_scan_kernel<W, QT><<<grid, threads, smem_bytes, st>>>(a);
  return cudaGetLastError();
}

template <int W>
static cudaError_t launch_w(const HamScanArgs &a, int threads, size_t smem_bytes, cudaStream_t st) {
  switch (a.qt) {
    case 1: return launch_wq<W, 1>(a, threads, smem_bytes, st);
    case 2: return launch_wq<W, 2>(a, threads, smem_bytes, st);
    case 4: return launch_wq<W, 4>(a, threads, smem_bytes, st);
    case 8: return launch_wq<W, 8>(a, threads, smem_bytes, st);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_ham_scan(const HamScanArgs &a, int threads, size_t smem_bytes, cudaStream_t st) {
  switch (a.W) {
    case 1: return launch_w<1>(a, threads, smem_bytes, st);
    case 2: return launch_w<2>(a, threads, smem_bytes, st);
    case 3: return launch_w<3>(a, threads, smem_bytes, st);
    case 4: return launch_w<4>(a, threads, smem_bytes, st);
    case 6: return launch_w<6>(a, threads, smem_bytes, st);
    case 8: return launch_w<8>(a, threads, smem_bytes, st);
    default: return cudaErrorInvalidValue;
  }
}

// ---- bit-vector layout -------------------------------------------------------------------
// bitvectors (BitVector.hpp:13-19): row-major [n][w64] uint64.  Packed: uint4 word j of a row =
// 64-bit words 2j, 2j+1 (little-endian halves); words beyond w64 are zero.
__global__ void ham_pack_kernel(const uint64_t *__restrict__ words, int64_t n, int64_t row0, int w64, int W,
                                uint4 *__restrict__ packed) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t row = row0 + i;
  const int64_t tile = row >> 5;
  const int lane = (int)(row & 31);
  for (int j = 0; j < W; j++) {
    const uint64_t a = (2 * j < w64) ? words[(size_t)i * w64 + 2 * j] : 0ull;
    const uint64_t b = (2 * j + 1 < w64) ? words[(size_t)i * w64 + 2 * j + 1] : 0ull;
    packed[((size_t)tile * W + j) * kTileRows + lane] =
        make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
  }
}

cudaError_t launch_ham_pack(const uint64_t *words, int64_t n, int64_t row0, int w64, int W, uint4 *packed,
                            cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int threads = 256;
  ham_pack_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(words, n, row0, w64, W, packed);
  return cudaGetLastError();
}

__host__ __device__ inline uint64_t hmix64(uint64_t x) {
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

// word w of global row r = mix64(seed ^ (r * G1 + w * G2)), masked to nbits in the last word
__global__ void ham_synth_kernel(uint4 *__restrict__ packed, int64_t n, int64_t row0, int64_t global_row0, int nbits,
                                 int W, uint64_t seed) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t row = row0 + i;
  const int64_t tile = row >> 5;
  const int lane = (int)(row & 31);
  const int w64 = (nbits + 63) / 64;
  for (int j = 0; j < W; j++) {
    uint64_t v[2];
    for (int h = 0; h < 2; h++) {
      const int w = 2 * j + h;
      uint64_t x = 0;
      if (w < w64) {
        x = hmix64(seed ^ ((uint64_t)(global_row0 + i) * 0x9E3779B97F4A7C15ull + (uint64_t)w * 0xD1B54A32D192ED03ull));
        const int rem = nbits - w * 64;
        if (rem < 64) x &= ((1ull << rem) - 1ull);
      }
      v[h] = x;
    }
    packed[((size_t)tile * W + j) * kTileRows + lane] =
        make_uint4((uint32_t)v[0], (uint32_t)(v[0] >> 32), (uint32_t)v[1], (uint32_t)(v[1] >> 32));
  }
}

cudaError_t launch_ham_synth(uint4 *packed, int64_t n, int64_t row0, int64_t global_row0, int nbits, int W,
                             uint64_t seed, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int threads = 256;
  ham_synth_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(packed, n, row0, global_row0, nbits, W, seed);
  return cudaGetLastError();
}

}  // namespace vaqgpu
