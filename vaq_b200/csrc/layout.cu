// Conflict-aware row order of the packed code matrix.
//
// The ADC filter kernel (adc_filter16_scan.cu) spends most of its shared-memory bandwidth on stage 1: every row
// gathers the 16-byte table entries of its first four subspace codes, one row per lane.  A 128-bit shared load is
// served a quarter-warp (8 lanes x 16 B = 128 B, all 32 banks) per wavefront; two lanes of a quarter whose entries
// fall into the same 16-byte bank group (entry index mod 8) cost an extra wavefront.  With rows in arrival order the
// eight codes of a quarter are independent, and the expected maximum multiplicity of 8 balls in 8 bins is 2.6 — the
// gathers run at 2.6x their conflict-free cost (ncu, round 1: 53 % of all shared wavefronts were bank conflicts).
//
// The order of the rows is ours to choose (ids travel with the rows), so the matrix is re-ordered inside windows of
// kLayoutWin rows such that the 8 rows that share a quarter-warp have, for each of the four stage-1 fields, (nearly)
// distinct residues `code mod 8`.  Per window, one warp builds the groups: rows are bucketed by the residue of
// field 0 (a group takes one row from each bucket: field 0 is conflict-free by construction), and every member is
// the candidate of its bucket that collides least with the residues the group already holds in fields 1-3.
// Measured in simulation: 1.26 wavefronts per quarter and field instead of 2.6 (windows of 4096 rows).
//
// The reference has no counterpart (its scan is a scalar loop over mCodebook rows, VAQ.cpp:1694-1758); results are
// unaffected: `rowid[storage row]` restores the original row index wherever a key is formed, so the canonical
// (distance, id) order is unchanged bit for bit.
#include <algorithm>

#include "common.cuh"

namespace vaqgpu {

namespace {

constexpr int kPlanWarps = 4;          // windows per CTA of the planning kernel

// residues (code mod 8) of the stage-1 fields of a row, packed 3 bits per field; fields past M read as 0
__device__ __forceinline__ uint32_t row_residues(const uint4 w0, const ScanLayout &lay, int nf) {
  const uint32_t wd[5] = {w0.x, w0.y, w0.z, w0.w, 0u};
  uint32_t key = 0;
  for (int f = 0; f < nf; f++) {
    const uint32_t meta = lay.fmeta[f];
    const int fw = lay.fword[f];          // < 2: the four leading fields end below bit 60
    const uint32_t code = __funnelshift_r(wd[fw], wd[fw + 1], meta & 31u) & (meta >> 16);
    key |= (code & 7u) << (3 * f);
  }
  return key;
}

// one-hot occupancy of a residue key: bit (8 f + r_f) for each field
__device__ __forceinline__ uint32_t key_onehot(uint32_t key, int nf) {
  uint32_t m = 0;
  for (int f = 0; f < nf; f++) m |= 1u << (8 * f + ((key >> (3 * f)) & 7u));
  return m;
}

}  // namespace

// One warp per window.  src[window * kLayoutWin + slot] = index (inside the window, current storage order) of the
// row that moves to `slot`.  Slots 8g .. 8g+7 form one quarter-warp of the scan.
// Windows: win_tab == NULL -> window w covers rows [row_lo + w * kLayoutWin, ... + kLayoutWin) clipped to n_rows;
// otherwise win_tab[w] = (first row, row count <= kLayoutWin) — TI indexes, whose windows may not cross clusters.
__device__ __forceinline__ void window_of(const int64_t *__restrict__ win_tab, int64_t win, int64_t row_lo, int64_t n_rows, int64_t &base, int &n) {
  if (win_tab) { base = win_tab[2 * win]; n = (int)win_tab[2 * win + 1]; }
  else { base = row_lo + win * kLayoutWin; n = (int)min((int64_t)kLayoutWin, n_rows - base); }
}

__global__ void __launch_bounds__(kPlanWarps * 32) layout_plan_kernel(const uint4 *__restrict__ codes, int W, int64_t row_lo,
                                                                      int64_t n_rows, const __grid_constant__ ScanLayout lay,
                                                                      uint16_t *__restrict__ src, const int64_t *__restrict__ win_tab,
                                                                      int64_t n_windows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t win = (int64_t)blockIdx.x * kPlanWarps + warp;
  if (win >= n_windows) return;
  int64_t base;
  int n;
  window_of(win_tab, win, row_lo, n_rows, base, n);
  if (n <= 0) return;
  uint16_t *keys = reinterpret_cast<uint16_t *>(smem_raw) + (size_t)warp * 2 * kLayoutWin;      // [n] residue key per row
  uint16_t *bk = keys + kLayoutWin;                                                              // [8 buckets] row lists
  const int nf = min(4, lay.M);
  uint16_t *out = src + (base - row_lo);

  // ---- residue keys + bucket sizes (bucket = residue of field 0)
  int cnt[8];
#pragma unroll
  for (int j = 0; j < 8; j++) cnt[j] = 0;
  for (int i0 = 0; i0 < n; i0 += 32) {      // the whole warp iterates together (ballots)
    const int i = i0 + lane;
    const bool ok = i < n;
    uint32_t key = 0;
    if (ok) {
      const int64_t row = base + i;
      key = row_residues(__ldg(codes + ((size_t)(row >> 5) * W) * kTileRows + (row & 31)), lay, nf);
      keys[i] = (uint16_t)key;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) cnt[j] += __popc(__ballot_sync(0xffffffffu, ok && (key & 7u) == (uint32_t)j));
  }
  int off[8], bn[8];
  {
    int acc = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { off[j] = acc; acc += cnt[j]; bn[j] = 0; }
  }
  __syncwarp();
  // ---- fill the buckets (stable: ballot order)
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    const bool ok = i < n;
    const uint32_t r0 = ok ? (keys[i] & 7u) : 8u;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const unsigned m = __ballot_sync(0xffffffffu, r0 == (uint32_t)j);
      if (r0 == (uint32_t)j) bk[off[j] + bn[j] + __popc(m & ((1u << lane) - 1u))] = (uint16_t)i;
      bn[j] += __popc(m);
    }
  }
  __syncwarp();

  // ---- groups of 8: one member per bucket while distinct non-empty buckets last
  int slot = 0;
  while (slot < n) {
    uint32_t occ = 0u;        // residues the group already holds (one-hot per field)
    unsigned used = 0u;       // buckets already taken by this group
    // a group fills one quarter-warp of the scan: 8 storage rows aligned to 8 — a window that starts off that
    // alignment (cluster windows) begins with a short group
    const int gsz = 8 - (int)((base + slot) & 7);
    for (int m = 0; m < gsz && slot < n; m++) {
      // bucket: the fullest one this group has not used yet; when none is left, the fullest one
      int best_j = -1, best_n = 0;
#pragma unroll
      for (int j = 0; j < 8; j++)
        if (!((used >> j) & 1u) && bn[j] > best_n) { best_n = bn[j]; best_j = j; }
      if (best_j < 0) {
#pragma unroll
        for (int j = 0; j < 8; j++)
          if (bn[j] > best_n) { best_n = bn[j]; best_j = j; }
      }
      used |= 1u << best_j;
      int o = 0;
#pragma unroll
      for (int j = 0; j < 8; j++) o = (j == best_j) ? off[j] : o;
      // candidate that collides least with `occ` (first one among equals); chunks of 32, stop at a collision-free one
      uint32_t pick = 0xFFFFFFFFu;       // (collisions << 16) | position
      for (int c0 = 0; c0 < best_n; c0 += 32) {
        const int c = c0 + lane;
        uint32_t mine = 0xFFFFFFFFu;
        if (c < best_n) mine = ((uint32_t)__popc(occ & key_onehot(keys[bk[o + c]], nf)) << 16) | (uint32_t)c;
        pick = min(pick, mine);
        if (__any_sync(0xffffffffu, (mine >> 16) == 0u)) break;
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) pick = min(pick, __shfl_xor_sync(0xffffffffu, pick, d));
      const int pos = (int)(pick & 0xFFFFu);
      const uint16_t row = bk[o + pos];
      __syncwarp();
      if (lane == 0) {
        bk[o + pos] = bk[o + best_n - 1];      // remove from the bucket
        out[slot] = row;
      }
      __syncwarp();
      occ |= key_onehot(keys[row], nf);
#pragma unroll
      for (int j = 0; j < 8; j++) bn[j] -= (j == best_j) ? 1 : 0;
      slot++;
    }
  }
}

// new[slot] = old[src[slot]] for the rows (all W words) and the row ids of each window.  A CTA stages its window row by
// row in a private scratch slot (global memory, L2-resident), then gathers from it.  Persistent CTAs.
__global__ void __launch_bounds__(256) layout_apply_kernel(uint4 *__restrict__ codes, int W, int64_t row_lo, int64_t n_rows,
                                                            const uint16_t *__restrict__ src, uint32_t *__restrict__ rowid,
                                                            uint4 *__restrict__ scratch, const int64_t *__restrict__ win_tab,
                                                            int64_t n_windows) {
  uint4 *mine = scratch + (size_t)blockIdx.x * ((size_t)kLayoutWin * W + kLayoutWin / 4);
  uint32_t *ids = reinterpret_cast<uint32_t *>(mine + (size_t)kLayoutWin * W);
  for (int64_t win = blockIdx.x; win < n_windows; win += gridDim.x) {
    int64_t base;
    int n;
    window_of(win_tab, win, row_lo, n_rows, base, n);
    for (int i = threadIdx.x; i < n * W; i += blockDim.x) {
      const int r = i / W, j = i - r * W;
      const int64_t row = base + r;
      mine[i] = codes[((size_t)(row >> 5) * W + j) * kTileRows + (row & 31)];
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) ids[i] = rowid[base + i];
    __threadfence_block();
    __syncthreads();
    for (int i = threadIdx.x; i < n * W; i += blockDim.x) {
      const int slot = i / W, j = i - slot * W;
      const int64_t row = base + slot;
      codes[((size_t)(row >> 5) * W + j) * kTileRows + (row & 31)] = mine[(size_t)src[(base - row_lo) + slot] * W + j];
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) rowid[base + i] = ids[src[(base - row_lo) + i]];
    __syncthreads();
  }
}

// Back to the order the row ids describe: dst[rowid[s]] = src[s] for every row s < n (rowid a permutation of [0, n)).
__global__ void layout_scatter_kernel(const uint4 *__restrict__ srcm, uint4 *__restrict__ dst, int W, const uint32_t *__restrict__ rowid, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * W) return;
  const int64_t s = i / W;
  const int j = (int)(i - s * W);
  const int64_t t = rowid[s];
  dst[((size_t)(t >> 5) * W + j) * kTileRows + (t & 31)] = srcm[((size_t)(s >> 5) * W + j) * kTileRows + (s & 31)];
}

cudaError_t launch_layout_restore(const uint4 *tmp_copy, uint4 *codes, int W, const uint32_t *rowid, int64_t n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  layout_scatter_kernel<<<(unsigned)((n * W + 255) / 256), 256, 0, st>>>(tmp_copy, codes, W, rowid, n);
  return cudaGetLastError();
}

__global__ void iota_u32_kernel(uint32_t *__restrict__ p, int64_t lo, int64_t hi) {
  const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < hi) p[i] = (uint32_t)i;
}

cudaError_t launch_iota_u32(uint32_t *p, int64_t lo, int64_t hi, cudaStream_t st) {
  if (hi <= lo) return cudaSuccess;
  iota_u32_kernel<<<(unsigned)((hi - lo + 255) / 256), 256, 0, st>>>(p, lo, hi);
  return cudaGetLastError();
}

size_t layout_scratch_bytes(int W, int ctas) { return (size_t)ctas * ((size_t)kLayoutWin * W + kLayoutWin / 4) * sizeof(uint4); }

// Re-orders the given windows (win_tab NULL: the aligned windows that cover rows [row_lo, n_rows), row_lo a multiple of
// kLayoutWin).  src: one uint16 per row of [row_lo, n_rows) of workspace.
cudaError_t launch_layout(uint4 *codes, int64_t row_lo, int64_t n_rows, const ScanLayout &lay, uint32_t *rowid, uint16_t *src,
                          uint4 *scratch, int scratch_ctas, const int64_t *win_tab, int64_t n_windows, cudaStream_t st) {
  if (n_rows <= row_lo || n_windows <= 0) return cudaSuccess;
  const size_t smem = (size_t)kPlanWarps * 2 * kLayoutWin * sizeof(uint16_t);
  static SmemOptIn optin;
  cudaError_t e = optin.ensure(layout_plan_kernel, smem);
  if (e != cudaSuccess) return e;
  layout_plan_kernel<<<(unsigned)((n_windows + kPlanWarps - 1) / kPlanWarps), kPlanWarps * 32, smem, st>>>(codes, lay.W, row_lo, n_rows, lay, src,
                                                                                                      win_tab, n_windows);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int ctas = (int)std::min<int64_t>(n_windows, scratch_ctas);
  layout_apply_kernel<<<ctas, 256, 0, st>>>(codes, lay.W, row_lo, n_rows, src, rowid, scratch, win_tab, n_windows);
  return cudaGetLastError();
}

}  // namespace vaqgpu
