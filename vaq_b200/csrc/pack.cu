// Code-matrix plumbing: bit-packing of the reference's uint16 code matrix (mCodebook,
// VAQ.hpp:72 / utils/Types.hpp:31 — 2 bytes per subspace regardless of width) into the
// scan layout, the inverse (round-trip check), device-side encode (VAQ::encodeImpl,
// VAQ.cpp:728-748) and the counter-based synthetic code generator for the 100M / 1B-row
// shapes the host cannot hold.
//
// HBM layout of the packed matrix: tiles of 32 rows; tile t stores, for each of the W
// 128-bit words of a row, the 32 rows' words contiguously:
//     packed[(t * W + j) * 32 + lane]   (uint4),   row = 32 t + lane
// so a warp's load of word j is one fully coalesced 512-byte request.
#include "common.cuh"

namespace vaqgpu {

__global__ void pack_codes_kernel(const uint16_t *__restrict__ codes, int64_t n, int64_t row0,
                                  const __grid_constant__ ScanLayout lay, uint4 *__restrict__ packed) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t wd[kMaxRowWords + 1];
  const int nw = lay.W * 4;
  for (int w = 0; w <= nw; w++) wd[w] = 0u;
  const uint16_t *c = codes + (size_t)i * lay.M;
  int f = 0;
  for (int w = 0; w < nw; w++) {
    const int fe = lay.fbeg[w + 1];
    for (; f < fe; f++) {
      const uint32_t meta = lay.fmeta[f];
      const uint32_t sh = meta & 31u, mask = meta >> 16;
      const uint32_t v = (uint32_t)c[f] & mask;
      wd[w] |= v << sh;
      if (sh) wd[w + 1] |= v >> (32u - sh);   // bits that straddle into the next word
    }
  }
  const int64_t row = row0 + i;
  const int64_t tile = row >> 5;
  const int lane = (int)(row & 31);
  for (int j = 0; j < lay.W; j++)
    packed[((size_t)tile * lay.W + j) * kTileRows + lane] = make_uint4(wd[4 * j], wd[4 * j + 1], wd[4 * j + 2], wd[4 * j + 3]);
}

__global__ void unpack_codes_kernel(const uint4 *__restrict__ packed, int64_t row0, int64_t n,
                                    const __grid_constant__ ScanLayout lay, uint16_t *__restrict__ codes,
                                    const uint32_t *__restrict__ rowid, int64_t srow_lo, int64_t srow_hi) {
  const int64_t row = srow_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // storage row
  if (row >= srow_hi) return;
  const int64_t orig = rowid ? (int64_t)rowid[row] : row;
  if (orig < row0 || orig >= row0 + n) return;
  const int64_t i = orig - row0;
  const int64_t tile = row >> 5;
  const int lane = (int)(row & 31);
  uint32_t wd[kMaxRowWords + 1];
  const int nw = lay.W * 4;
  for (int j = 0; j < lay.W; j++) {
    const uint4 v = packed[((size_t)tile * lay.W + j) * kTileRows + lane];
    wd[4 * j] = v.x; wd[4 * j + 1] = v.y; wd[4 * j + 2] = v.z; wd[4 * j + 3] = v.w;
  }
  wd[nw] = 0u;
  int f = 0;
  for (int w = 0; w < nw; w++) {
    const int fe = lay.fbeg[w + 1];
    for (; f < fe; f++) {
      const uint32_t meta = lay.fmeta[f];
      codes[(size_t)i * lay.M + f] = (uint16_t)(__funnelshift_r(wd[w], wd[w + 1], meta & 31u) & (meta >> 16));
    }
  }
}

cudaError_t launch_pack_codes(const uint16_t *codes, int64_t n, int64_t row0, const ScanLayout &lay, uint4 *packed,
                              cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int threads = 256;
  pack_codes_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(codes, n, row0, lay, packed);
  return cudaGetLastError();
}

cudaError_t launch_unpack_codes(const uint4 *packed, int64_t row0, int64_t n, const ScanLayout &lay, uint16_t *codes,
                                const uint32_t *rowid, int64_t srow_lo, int64_t srow_hi, cudaStream_t st) {
  if (n <= 0 || srow_hi <= srow_lo) return cudaSuccess;
  const int threads = 256;
  unpack_codes_kernel<<<(unsigned)((srow_hi - srow_lo + threads - 1) / threads), threads, 0, st>>>(packed, row0, n, lay, codes, rowid,
                                                                                              srow_lo, srow_hi);
  return cudaGetLastError();
}

// ---- synthetic codes ---------------------------------------------------------------------
// u = splitmix64(seed ^ (row * G1 + s * G2)) >> 40, scaled to [0,1); code = first c with
// cdf[c] > u (uniform when cdf == NULL).  Restated in numpy by vaq_b200/synth.py.
__host__ __device__ inline uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

__global__ void synth_codes_kernel(uint16_t *__restrict__ codes, int64_t n, int64_t global_row0, int M,
                                   const int32_t *__restrict__ bits, const float *__restrict__ cdf,
                                   const int32_t *__restrict__ ent_off, uint64_t seed) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * M) return;
  const int64_t i = idx / M;
  const int s = (int)(idx - i * M);
  const uint64_t h = mix64(seed ^ ((uint64_t)(global_row0 + i) * 0x9E3779B97F4A7C15ull + (uint64_t)s * 0xD1B54A32D192ED03ull));
  const int K = 1 << bits[s];
  uint32_t code;
  if (cdf == nullptr) {
    code = (uint32_t)(h >> 40) & (uint32_t)(K - 1);
  } else {
    const float u = (float)(uint32_t)(h >> 40) * (1.0f / 16777216.0f);
    const float *t = cdf + ent_off[s];
    int lo = 0, hi = K - 1;              // first c with t[c] > u, clamped to K-1
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (t[mid] > u) hi = mid; else lo = mid + 1;
    }
    code = (uint32_t)lo;
  }
  codes[idx] = (uint16_t)code;
}

cudaError_t launch_synth_codes(uint16_t *codes, int64_t n, int64_t global_row0, int M, const int32_t *bits,
                               const float *cdf, const int32_t *ent_off, uint64_t seed, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int threads = 256;
  const int64_t total = n * M;
  synth_codes_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, st>>>(codes, n, global_row0, M, bits, cdf,
                                                                                     ent_off, seed);
  return cudaGetLastError();
}

// ---- encode --------------------------------------------------------------------------------
// VAQ::encodeImpl (VAQ.cpp:728-748): per (row, subspace) the centroid minimising
// sum_j (x_j - c_j)^2 with strict '<' (lowest code wins ties).  One CTA handles 256 rows of one
// subspace; the row sub-vectors sit in shared memory column-per-thread, centroids are streamed
// through shared memory in chunks and broadcast.  Sum order = the oracle's sequential
// non-fused order, so the codes are bit-exact against it.
constexpr int kEncRows = 256;
constexpr int kEncChunkFloats = 8192;

__global__ void encode_kernel(const float *__restrict__ x_proj, int64_t n, int D, const float *__restrict__ cent,
                              const __grid_constant__ LutPlan p, uint16_t *__restrict__ codes) {
  extern __shared__ float sm[];
  const int L = p.L;
  float *xs = sm;                       // [L][kEncRows]
  float *cs = sm + (size_t)L * kEncRows;  // [chunk][L]
  const int s = blockIdx.y;
  const int64_t row = (int64_t)blockIdx.x * kEncRows + threadIdx.x;
  const bool valid = row < n;
  for (int j = 0; j < L; j++) xs[j * kEncRows + threadIdx.x] = valid ? x_proj[(size_t)row * D + (size_t)s * L + j] : 0.f;
  const int K = p.ent_off[s + 1] - p.ent_off[s];
  const float *cp = cent + p.cent_off[s];
  const int chunk = max(1, kEncChunkFloats / L);
  float best = 3.402823466e+38f;
  int best_c = 0;
  for (int c0 = 0; c0 < K; c0 += chunk) {
    const int nc = min(chunk, K - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < nc * L; i += blockDim.x) cs[i] = __ldg(cp + (size_t)c0 * L + i);
    __syncthreads();
    for (int c = 0; c < nc; c++) {
      float dist = 0.f;
      for (int j = 0; j < L; j++) {
        const float d = __fsub_rn(xs[j * kEncRows + threadIdx.x], cs[c * L + j]);
        dist = __fadd_rn(dist, __fmul_rn(d, d));
      }
      if (dist < best) { best = dist; best_c = c0 + c; }
    }
  }
  if (valid) codes[(size_t)row * p.M + s] = (uint16_t)best_c;
}

cudaError_t launch_encode(const float *x_proj, int64_t n, const float *centroids, const LutPlan &plan,
                          uint16_t *codes, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int L = plan.L;
  const int chunk = kEncChunkFloats / L > 0 ? kEncChunkFloats / L : 1;
  const size_t smem = ((size_t)L * kEncRows + (size_t)chunk * L) * sizeof(float);
  static SmemOptIn optin;
  {
    cudaError_t e = optin.ensure(encode_kernel, smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((unsigned)((n + kEncRows - 1) / kEncRows), (unsigned)plan.M);
  encode_kernel<<<grid, kEncRows, smem, st>>>(x_proj, n, plan.M * plan.L, centroids, plan, codes);
  return cudaGetLastError();
}

}  // namespace vaqgpu
