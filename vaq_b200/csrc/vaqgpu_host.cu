// C-ABI implementation (include/vaqgpu.h): index state in HBM, scan planning, launch
// sequencing.  No CPU fallback: every entry point fails with VAQGPU_ECUDA when there is no
// usable sm_100 device.
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "../../include/vaqgpu.h"
#include "common.cuh"

using namespace vaqgpu;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace

namespace vaqgpu {
// same message slot, for the other translation units of the library (sharded.cu)
int set_error(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace vaqgpu

namespace {

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(e_ == cudaErrorMemoryAllocation ? VAQGPU_ENOMEM : VAQGPU_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, \
                  cudaGetErrorString(e_));                                                         \
  } while (0)

constexpr int kScanThreads = 256;
constexpr size_t kSmemCap = 227 * 1024;           // opt-in dynamic shared memory per CTA on sm_100
constexpr size_t kLutWorkspaceBytes = 1ull << 30; // LUT workspace cap -> queries per launch
constexpr int64_t kStageRows = 1 << 21;           // rows per upload / generate chunk
// scan order (ensure_layout): indexes of this many rows are grouped by coarse cluster.  Below, a search is a handful of
// microseconds per CTA anyway; above, the bounds of a query tile settle within the first percent of its rows without help.
constexpr int64_t kOrderMinRows = 1 << 15, kOrderMaxRows = 1 << 23;

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  cudaError_t ensure(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    cudaError_t e = cudaMalloc(&p, need);
    if (e == cudaSuccess) bytes = need;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int check_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(VAQGPU_ECUDA, "no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(VAQGPU_EINVAL, "device %d out of range [0,%d)", device, n);
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, device);
  if (e != cudaSuccess) return fail(VAQGPU_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (p.major != 10) return fail(VAQGPU_ECUDA, "device %d is sm_%d%d; kernels are built for sm_100a only", device, p.major, p.minor);
  return VAQGPU_OK;
}

}  // namespace

struct vaqgpu_index {
  int device = 0;
  int num_sms = 148;
  int D = 0, M = 0, L = 0;
  std::vector<int32_t> bits;
  int32_t total_entries = 0;
  int64_t id_base = 0;
  float visit = 1.f;

  ScanLayout lay{};
  LutPlan plan{};

  float *d_centroids = nullptr;
  float *d_cent_rmax = nullptr;   // [M] largest centroid norm per subspace (fp16 table scaling)
  float *d_eig = nullptr;
  int32_t *d_bits = nullptr, *d_ent_off = nullptr, *d_cent_off = nullptr;

  uint4 *d_codes = nullptr;
  int64_t n_rows = 0, cap_rows = 0;

  // TI / visit
  int32_t C = 0, segdims = 0;
  float *d_clusters = nullptr;
  int64_t *d_cl_start = nullptr, *d_cl_size = nullptr;
  int64_t *d_cl_rule = nullptr;    // cluster sizes the visiting rule counts (whole index; a row shard's own are smaller)
  int32_t *d_id_map = nullptr;
  float *d_clusters_t = nullptr;   // the centres dimension-major [segdims][C] (coalesced reads of the visit kernel)
  uint16_t *d_tile_cl = nullptr;   // cluster of each 32-row tile (filter kernels); NULL when the ranges are not an ascending partition

  // refine
  float *d_raw = nullptr;
  int64_t raw_n = 0;
  int32_t raw_D = 0;

  cudaStream_t stream = nullptr;   // used by the host-buffer entry points
  cudaEvent_t ev[5] = {};
  bool timed = false;
  int32_t cfg[12] = {};

  // conflict-aware row order (layout.cu): d_rowid[storage row] = original row; rows [0, opt_n) are re-ordered
  uint32_t *d_rowid = nullptr;
  int64_t rowid_cap = 0, rowid_n = 0, opt_n = 0;
  float layout_ms = 0.f;
  bool cluster_windows = false;                  // the re-ordering windows follow the TI / scan-order clusters (not aligned 4096-row blocks)
  // scan order (EA / HEAP searches of an index without TI clusters): rows [0, oc_n) grouped by a coarse clustering of
  // their leading subspaces so that a query tile can start its scan at the rows nearest to it (ensure_layout)
  int32_t oc_C = 0, oc_dims = 0;
  int64_t oc_n = 0;
  float *d_oc_centres_t = nullptr;               // [oc_dims][oc_C]
  int64_t *d_oc_start = nullptr, *d_oc_size = nullptr;
  std::vector<int64_t> h_cl_start, h_cl_size;    // host copies of the TI cluster ranges

  // cross-shard bound exchange (vaqgpu_bounds_*): two halves of bounds_cap entries, used alternately by
  // successive searches; the half the NEXT search will use is reset while this one runs
  uint32_t *d_bounds = nullptr;
  int32_t bounds_cap = 0, bounds_parity = 0;
  PeerBounds peers{};
  std::vector<void *> ipc_opened;

  DevBuf w_vis, w_near, w_perm, w_qperm, w_tmask, w_rot, w_lsrc, w_lscratch, w_dbg, w_lut16, w_scale, w_thr, w_q, w_qproj, w_lut, w_keys, w_scratch, w_ranges, w_nranges, w_stage, w_labels, w_dists, w_outkeys, w_cdf, w_x;
};

struct hamgpu_index {
  int device = 0;
  int num_sms = 148;
  int nbits = 0, w64 = 0, W = 0;
  int64_t id_base = 0;
  uint4 *d_codes = nullptr;
  int64_t n_rows = 0, cap_rows = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[3] = {};
  bool timed = false;
  int32_t cfg[8] = {};
  DevBuf w_q, w_qpad, w_keys, w_scratch, w_stage, w_idx, w_dist;
};

namespace {

// Grow a tiled code matrix to hold at least `rows` rows (whole 32-row tiles, zero-filled so the
// padding rows of the last tile decode to valid table indices).
cudaError_t grow_codes(uint4 **codes, int64_t *cap_rows, int64_t n_rows, int W, int64_t rows, cudaStream_t st) {
  if (rows <= *cap_rows) return cudaSuccess;
  int64_t want = std::max<int64_t>(rows, *cap_rows + *cap_rows / 2);
  want = (want + kTileRows - 1) / kTileRows * kTileRows;
  uint4 *nw = nullptr;
  const size_t bytes = (size_t)want * W * sizeof(uint4);
  cudaError_t e = cudaMalloc(&nw, bytes);
  if (e != cudaSuccess && want > rows) {          // retry without head-room
    cudaGetLastError();
    want = (rows + kTileRows - 1) / kTileRows * kTileRows;
    e = cudaMalloc(&nw, (size_t)want * W * sizeof(uint4));
  }
  if (e != cudaSuccess) return e;
  const int64_t used_rows = (n_rows + kTileRows - 1) / kTileRows * kTileRows;
  const size_t used = (size_t)used_rows * W * sizeof(uint4);
  if (*codes && used) {
    e = cudaMemcpyAsync(nw, *codes, used, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) { cudaFree(nw); return e; }
  }
  e = cudaMemsetAsync(reinterpret_cast<unsigned char *>(nw) + used, 0, (size_t)want * W * sizeof(uint4) - used, st);
  if (e != cudaSuccess) { cudaFree(nw); return e; }
  e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { cudaFree(nw); return e; }
  if (*codes) cudaFree(*codes);
  *codes = nw;
  *cap_rows = want;
  return cudaSuccess;
}

// TI state describes exactly the rows that were there when vaqgpu_set_clusters ran: appending rows drops it
// (a TI search must never silently skip the new rows), and so does a failed vaqgpu_set_clusters.
void clear_clusters(vaqgpu_index *h) {
  cudaFree(h->d_clusters); cudaFree(h->d_cl_start); cudaFree(h->d_cl_size); cudaFree(h->d_id_map);
  cudaFree(h->d_cl_rule); cudaFree(h->d_tile_cl); cudaFree(h->d_clusters_t);
  h->d_clusters_t = nullptr;
  h->d_clusters = nullptr; h->d_cl_start = h->d_cl_size = h->d_cl_rule = nullptr; h->d_id_map = nullptr; h->d_tile_cl = nullptr;
  h->C = 0; h->segdims = 0;
}

// Field placement inside the packed row + LUT placement (shared-memory resident vs spilled).
int plan_model(vaqgpu_index *h) {
  const int M = h->M;
  ScanLayout &lay = h->lay;
  LutPlan &plan = h->plan;
  memset(&lay, 0, sizeof(lay));
  memset(&plan, 0, sizeof(plan));
  int64_t bitoff = 0;
  int32_t ent = 0, coff = 0;
  std::vector<int> word_of(M);
  for (int s = 0; s < M; s++) {
    const int b = h->bits[s];
    word_of[s] = (int)(bitoff >> 5);
    lay.fmeta[s] = (uint32_t)(bitoff & 31) | ((uint32_t)((1u << b) - 1u) << 16);
    plan.ent_off[s] = ent;
    plan.cent_off[s] = coff;
    ent += 1 << b;
    coff += (1 << b) * h->L;
    bitoff += b;
  }
  plan.ent_off[M] = ent;
  plan.M = M; plan.L = h->L; plan.total_entries = ent;
  h->total_entries = ent;
  const int W = (int)((bitoff + 127) / 128);
  if (W < 1 || W > 8) return fail(VAQGPU_EINVAL, "sum(bits)=%lld needs %d uint4 words per row; supported 1..8", (long long)bitoff, W);
  lay.M = M; lay.W = W;
  int f = 0;
  for (int w = 0; w <= 4 * W; w++) {
    while (f < M && word_of[f] < w) f++;
    lay.fbeg[w] = (uint16_t)f;
  }
  lay.fbeg[4 * W] = (uint16_t)M;

  for (int s = 0; s < M; s++) {
    lay.fword[s] = (uint8_t)word_of[s];
    auto woff = [](int w) { return (uint16_t)((w >> 2) * kTileRows * 4 + (w & 3)); };   // tiled layout: uint4 j of a row is 32 uint4 apart
    lay.fw_lo[s] = woff(word_of[s]);
    lay.fw_hi[s] = word_of[s] + 1 < 4 * W ? woff(word_of[s] + 1) : lay.fw_lo[s];
  }
  plan.T = 1;
  return VAQGPU_OK;
}

// LUT residency for a given shared-memory budget (entries per query): tables are taken in scan order
// (variance-descending — with early abandon the leading tables are the ones every row touches)
// while they fit; the rest stay in global memory and are served by L1/L2 ("spill").
void apply_residency(const vaqgpu_index *h, size_t budget_entries, int T, ScanLayout &lay, LutPlan &plan, int32_t &res_floats,
                     int32_t &spill_floats) {
  const int M = h->M;
  lay = h->lay;
  plan = h->plan;
  plan.T = T;
  size_t resident = 0;
  std::vector<char> res(M, 0);
  for (int s = 0; s < M; s++) {
    const size_t K = (size_t)1 << h->bits[s];
    if (resident + K <= budget_entries) { res[s] = 1; resident += K; }
  }
  int32_t pos = 0;
  for (int s = 0; s < M; s++)
    if (res[s]) { plan.pos[s] = pos; lay.foff[s] = (uint32_t)pos; pos += 1 << h->bits[s]; }
  res_floats = (pos + 3) & ~3;
  int32_t sp = 0;
  for (int s = 0; s < M; s++)
    if (!res[s]) {
      plan.pos[s] = res_floats + sp; lay.foff[s] = (uint32_t)sp; lay.fmeta[s] |= kFieldSpill; sp += 1 << h->bits[s];
    }
  spill_floats = sp;
  plan.row_stride = (res_floats + sp + 3) & ~3;
}

size_t scan_smem_bytes(int smem_lut_floats, int k, int threads) {
  const int nwarps = threads / 32;
  size_t b = (((size_t)smem_lut_floats * 4 + 15) & ~(size_t)15);
  b += ((size_t)nwarps * k + k + 2) * sizeof(uint64_t);
  return b;
}

// Row chunks per query tile for the filter kernels (grid = query tiles x chunks, one CTA per SM).  Chunks hold at
// most 1M rows (all query tiles sweep a chunk while it is L2-resident) and leave every warp >= 16 tiles.  Within
// those limits the count minimises a simple cost model: CTAs run in waves of one per SM, and a CTA costs a fixed
// ~45 us (table staging, bound seeding, queue drain) plus ~0.9 ns per row (measured, SIFT1M shape) — so a handful of
// queries gets exactly one wave of long CTAs, many query tiles get as few chunks as the L2 rule allows, and a grid
// that is only a few waves long avoids a nearly empty last wave.
int64_t choose_chunks(int64_t n_tiles, int qtiles, int nwarps, int num_sms) {
  const int64_t c_max = std::max<int64_t>(1, n_tiles / ((int64_t)nwarps * 16));
  const int64_t c_min = std::min(c_max, std::max<int64_t>(1, (n_tiles + 32767) / 32768));
  const double t_fixed = 45.0, t_row = 0.0009;         // microseconds
  if (qtiles >= 2 * num_sms) return c_min;            // many waves either way: no reason to pay the fixed cost twice
  double best = 0.0;
  int64_t pick = c_min;
  const int64_t c_hi = std::min<int64_t>(c_max, std::max<int64_t>(c_min * 2, (4ll * num_sms + qtiles - 1) / qtiles));
  for (int64_t c = c_min; c <= c_hi; c++) {
    const int64_t ctas = c * qtiles, waves = (ctas + num_sms - 1) / num_sms;
    const double rows = 32.0 * (double)((n_tiles + c - 1) / c);
    const double cost = (double)waves * (t_fixed + rows * t_row);
    if (c == c_min || cost < best * 0.999) { best = cost; pick = c; }
  }
  return pick;
}

int ensure_stream(cudaStream_t *st, cudaEvent_t *ev, int nev) {
  if (!*st) CU(cudaStreamCreateWithFlags(st, cudaStreamNonBlocking));
  for (int i = 0; i < nev; i++)
    if (!ev[i]) CU(cudaEventCreate(&ev[i]));
  return VAQGPU_OK;
}

__global__ void ham_pad_queries_kernel(const uint64_t *__restrict__ q, int nq, int w64, int W, uint4 *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq * W) return;
  const int r = i / W, j = i - r * W;
  const uint64_t a = (2 * j < w64) ? q[(size_t)r * w64 + 2 * j] : 0ull;
  const uint64_t b = (2 * j + 1 < w64) ? q[(size_t)r * w64 + 2 * j + 1] : 0ull;
  out[i] = make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
}

// Development knobs (benchmark sweeps only): VAQGPU_TUNE="seed=0,T=2,chunks=4,threads=512"
int tune_knob(const char *name, int dflt) {
  const char *e = getenv("VAQGPU_TUNE");
  if (!e) return dflt;
  const size_t n = strlen(name);
  for (const char *p = e; *p;) {
    if (!strncmp(p, name, n) && p[n] == '=') return atoi(p + n + 1);
    const char *c = strchr(p, ',');
    if (!c) break;
    p = c + 1;
  }
  return dflt;
}

// ---- conflict-aware row order ---------------------------------------------------------------------------------
// Row ids exist as soon as any window was re-ordered; from then on they are kept valid for every row (new rows enter
// at their own index) because every scan kernel forms its keys through them.
int ensure_rowid(vaqgpu_index *h, cudaStream_t st) {
  if (!h->d_rowid) return VAQGPU_OK;
  if (h->rowid_cap < h->n_rows) {
    uint32_t *nw = nullptr;
    const int64_t cap = std::max(h->cap_rows, h->n_rows);
    CU(cudaMalloc(&nw, (size_t)cap * sizeof(uint32_t)));
    CU(cudaMemcpyAsync(nw, h->d_rowid, (size_t)h->rowid_n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    CU(cudaStreamSynchronize(st));
    cudaFree(h->d_rowid);
    h->d_rowid = nw; h->rowid_cap = cap;
  }
  if (h->rowid_n < h->n_rows) {
    CU(launch_iota_u32(h->d_rowid, h->rowid_n, h->n_rows, st));
    h->rowid_n = h->n_rows;
  }
  return VAQGPU_OK;
}

void clear_scan_order(vaqgpu_index *h) {
  cudaFree(h->d_oc_centres_t); cudaFree(h->d_oc_start); cudaFree(h->d_oc_size);
  h->d_oc_centres_t = nullptr; h->d_oc_start = h->d_oc_size = nullptr;
  h->oc_C = 0; h->oc_dims = 0; h->oc_n = 0;
}

int restore_layout(vaqgpu_index *h, cudaStream_t st);

// Scan order: groups all rows (arrival order on entry) by a coarse k-means clustering of their decoded leading
// subspaces (cluster_ti.cu, the machinery of the device-side clusterTI) and records the regrouping in rowid.
// The reference has no counterpart; the answer of a search does not depend on it (keys carry rowid).
int build_scan_order(vaqgpu_index *h, cudaStream_t st, std::vector<int64_t> &cl_start, std::vector<int64_t> &cl_size) {
  const int64_t n = h->n_rows;
  int C = 16;          // ~2000 rows per cluster, 16..64 clusters (measured flat between 32 and 256 at 1 M rows; fewer clusters plan faster)
  while (C < 64 && (int64_t)C * 2048 < n) C *= 2;
  C = tune_knob("order_c", C);
  const int seg = std::min(h->M, std::max(1, 16 / std::max(1, h->L)));          // ~16 leading dimensions
  const int iters = 6;
  const size_t tbl = cluster_ti_table_floats(h->plan, seg, C);
  const int64_t tiles = (n + kTileRows - 1) / kTileRows;
  DevBuf centres, T, hist, assign, sizes, bh, id_map, ncodes;
  float *centres_t = nullptr;
  int64_t *start = nullptr, *size64 = nullptr;
  auto cleanup = [&]() {
    for (DevBuf *b : {&centres, &T, &hist, &assign, &sizes, &bh, &id_map, &ncodes}) b->release();
    cudaFree(centres_t); cudaFree(start); cudaFree(size64);
  };
#define CO(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      cleanup();                                                                                         \
      return fail(e_ == cudaErrorMemoryAllocation ? VAQGPU_ENOMEM : VAQGPU_ECUDA, "scan order: %s: %s", #call, cudaGetErrorString(e_)); \
    }                                                                                                    \
  } while (0)
  CO(centres.ensure((size_t)C * seg * h->L * sizeof(float)));
  CO(T.ensure(tbl * sizeof(float)));
  CO(hist.ensure(tbl * sizeof(int32_t)));
  CO(assign.ensure((size_t)n * sizeof(int32_t)));
  CO(sizes.ensure((size_t)C * sizeof(int32_t)));
  CO(bh.ensure(regroup_hist_ints(n, C) * sizeof(int32_t)));
  CO(id_map.ensure((size_t)n * sizeof(int32_t)));
  CO(ncodes.ensure((size_t)tiles * kTileRows * h->lay.W * sizeof(uint4)));
  CO(cudaMalloc(&centres_t, (size_t)C * seg * h->L * sizeof(float)));
  CO(cudaMalloc(&start, (size_t)C * sizeof(int64_t)));
  CO(cudaMalloc(&size64, (size_t)C * sizeof(int64_t)));
  CO(cudaMemsetAsync(ncodes.p, 0, (size_t)tiles * kTileRows * h->lay.W * sizeof(uint4), st));
  CO(launch_cluster_ti_kmeans(h->d_codes, h->lay, n, h->plan, seg, h->d_centroids, h->d_cent_off, h->d_ent_off, C, iters, (float *)centres.p,
                              (float *)T.p, (int32_t *)hist.p, (int32_t *)assign.p, (int32_t *)sizes.p, st));
  CO(launch_regroup((const int32_t *)assign.p, (const int32_t *)sizes.p, n, C, h->d_codes, (uint4 *)ncodes.p, h->lay.W, (int32_t *)id_map.p, start,
                    size64, (int32_t *)bh.p, st));
  CO(launch_transpose((const float *)centres.p, C, seg * h->L, centres_t, st));
  // the regrouped matrix replaces the rows in place (the allocation keeps its capacity); rowid = arrival index of each row
  CO(cudaMemcpyAsync(h->d_codes, ncodes.p, (size_t)tiles * kTileRows * h->lay.W * sizeof(uint4), cudaMemcpyDeviceToDevice, st));
  CO(cudaMemcpyAsync(h->d_rowid, id_map.p, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  cl_start.resize(C); cl_size.resize(C);
  CO(cudaMemcpyAsync(cl_start.data(), start, (size_t)C * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CO(cudaMemcpyAsync(cl_size.data(), size64, (size_t)C * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CO(cudaStreamSynchronize(st));
#undef CO
  clear_scan_order(h);
  h->d_oc_centres_t = centres_t; centres_t = nullptr;
  h->d_oc_start = start; start = nullptr;
  h->d_oc_size = size64; size64 = nullptr;
  h->oc_C = C; h->oc_dims = seg * h->L; h->oc_n = n;
  cleanup();
  return VAQGPU_OK;
}

// Re-orders the windows that hold rows added since the last call (the filter kernels call this before scanning).
// With TI clusters set the windows are pieces of the clusters (<= kLayoutWin rows each), so every cluster keeps its
// row range and the cluster-of-tile map stays valid.  Without TI clusters, an index of kOrderMinRows..kOrderMaxRows rows
// is first grouped by coarse cluster (scan order, above): the windows are then pieces of those clusters, and rows
// appended later form a tail after them (re-clustered once the tail outgrows an eighth of the index).
int ensure_layout(vaqgpu_index *h, cudaStream_t st) {
  if (h->n_rows <= h->opt_n || tune_knob("layout", 1) == 0) return ensure_rowid(h, st);
  const bool order_ok = h->C == 0 && tune_knob("order", 1) != 0 && h->n_rows >= tune_knob("order_min", (int)kOrderMinRows) && h->n_rows <= kOrderMaxRows;
  if (h->oc_C > 0 && !order_ok) {          // outgrew the scan order: back to arrival order, aligned windows
    int rc = restore_layout(h, st);
    if (rc) return rc;
  }
  if (!h->d_rowid) {
    const int64_t cap = std::max(h->cap_rows, h->n_rows);
    CU(cudaMalloc(&h->d_rowid, (size_t)cap * sizeof(uint32_t)));
    h->rowid_cap = cap; h->rowid_n = 0;
  }
  int rc = ensure_rowid(h, st);
  if (rc) return rc;
  int64_t row_lo = (h->opt_n / kLayoutWin) * kLayoutWin;          // a partially filled window is planned again
  int64_t n_windows = (h->n_rows - row_lo + kLayoutWin - 1) / kLayoutWin;
  const int64_t *d_win = nullptr;
  DevBuf w_win;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  CU(cudaEventRecord(e0, st));
  std::vector<int64_t> tab;
  auto pieces = [&](int64_t start, int64_t size) {
    for (int64_t off = 0; off < size; off += kLayoutWin) {
      tab.push_back(start + off);
      tab.push_back(std::min<int64_t>(kLayoutWin, size - off));
    }
  };
  bool explicit_windows = false;
  if (h->C > 0) {
    for (int c = 0; c < h->C; c++) pieces(h->h_cl_start[c], h->h_cl_size[c]);
    row_lo = 0;
    explicit_windows = true;
  } else if (order_ok) {
    if (h->oc_C == 0 || h->n_rows - h->oc_n > h->oc_n / 8) {
      if (h->opt_n > 0) { rc = restore_layout(h, st); if (rc) return rc; rc = ensure_rowid(h, st); if (rc) return rc; }
      std::vector<int64_t> cs, cz;
      rc = build_scan_order(h, st, cs, cz);
      if (rc) return rc;
      for (size_t c = 0; c < cs.size(); c++) pieces(cs[c], cz[c]);
      row_lo = 0;
    } else {
      pieces(h->oc_n, h->n_rows - h->oc_n);          // the tail: rows appended after the clustering
      row_lo = h->oc_n;
    }
    explicit_windows = true;
  }
  if (explicit_windows) {
    n_windows = (int64_t)tab.size() / 2;
    if (n_windows > 0) {
      CU(w_win.ensure(tab.size() * sizeof(int64_t)));
      CU(cudaMemcpyAsync(w_win.p, tab.data(), tab.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
      CU(cudaStreamSynchronize(st));
      d_win = (const int64_t *)w_win.p;
    }
    h->cluster_windows = true;
  }
  const int ctas = (int)std::max<int64_t>(1, std::min<int64_t>(n_windows, 2 * h->num_sms));
  CU(h->w_lsrc.ensure((size_t)(h->n_rows - row_lo) * sizeof(uint16_t)));
  CU(h->w_lscratch.ensure(layout_scratch_bytes(h->lay.W, ctas)));
  CU(launch_layout(h->d_codes, row_lo, h->n_rows, h->lay, h->d_rowid, (uint16_t *)h->w_lsrc.p, (uint4 *)h->w_lscratch.p, ctas, d_win,
                   n_windows, st));
  CU(cudaEventRecord(e1, st));
  CU(cudaEventSynchronize(e1));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  h->layout_ms += ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  h->w_lscratch.release();
  w_win.release();
  h->opt_n = h->n_rows;
  return VAQGPU_OK;
}

// Back to the arrival order (TI cluster ranges are defined on it): every row returns to the position its id names.
int restore_layout(vaqgpu_index *h, cudaStream_t st) {
  if (!h->d_rowid || h->opt_n == 0) return VAQGPU_OK;
  int rc = ensure_rowid(h, st);
  if (rc) return rc;
  const int64_t tiles = (h->n_rows + kTileRows - 1) / kTileRows;
  const size_t bytes = (size_t)tiles * kTileRows * h->lay.W * sizeof(uint4);
  uint4 *tmp = nullptr;
  CU(cudaMalloc(&tmp, bytes));
  cudaError_t e = cudaMemcpyAsync(tmp, h->d_codes, bytes, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess) e = launch_layout_restore(tmp, h->d_codes, h->lay.W, h->d_rowid, h->n_rows, st);
  if (e == cudaSuccess) e = launch_iota_u32(h->d_rowid, 0, h->n_rows, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(tmp);
  if (e != cudaSuccess) return fail(VAQGPU_ECUDA, "restore_layout: %s", cudaGetErrorString(e));
  h->opt_n = 0;
  h->cluster_windows = false;
  clear_scan_order(h);
  return VAQGPU_OK;
}

// Appending rows (or replacing the clusters) ends the current TI state: rows go back to their cluster-grouped order first.
int drop_clusters(vaqgpu_index *h) {
  if (!h->C) return VAQGPU_OK;
  int rc = restore_layout(h, h->stream);
  clear_clusters(h);
  return rc;
}

// The whole device-side search; exactly one of (d_labels,d_dists) / d_keys is the output.
//
// Scan kernel selection (the reference dispatches TI -> EA -> HEAP, VAQ.cpp:799-840):
//   EA / HEAP        -> adc_filter16_scan_kernel (fp16 lower-bound filter + exact scoring of the survivors), or
//                       adc_filter_scan_kernel when eight queries' fp16 tables do not fit in shared memory
//   TI / visit       -> the same fp16 filter kernel over the visited clusters (ti_plan.cu) when the cluster ranges partition
//                       the rows in ascending order and the tables fit; otherwise adc_scan_kernel over per-query row ranges
//   scan order       -> EA / HEAP searches of an index grouped by coarse cluster (build_scan_order) re-group the batch
//                       into query tiles by nearest cluster and start each tile's scan there
//   VAQGPU_SCAN_V1   -> force adc_scan_kernel (lane-per-row; exhaustive with HEAP, warp-uniform abandoning with EA)
int search_device_impl(vaqgpu_index *h, const float *d_queries, int nq, int k, uint32_t flags, int32_t *d_labels,
                       float *d_dists, uint64_t *d_keys, cudaStream_t st, bool record) {
  if (nq < 0 || k <= 0) return fail(VAQGPU_EINVAL, "nq=%d k=%d", nq, k);
  if (k > 2048) return fail(VAQGPU_EINVAL, "k=%d > 2048 unsupported", k);
  if (nq == 0) return VAQGPU_OK;
  const bool projected = flags & VAQGPU_PROJECTED;
  if (!projected && !h->d_eig) return fail(VAQGPU_ESTATE, "raw queries need eig_real in the model (or pass VAQGPU_PROJECTED)");
  const bool ti = (flags & VAQGPU_TI) != 0;
  if (ti && !h->d_clusters) return fail(VAQGPU_ESTATE, "VAQGPU_TI needs vaqgpu_set_clusters");
  const bool ea = (flags & VAQGPU_EA) != 0 || ti || !(flags & VAQGPU_HEAP);
  // HEAP (exhaustive) and EA return the same k best rows (VAQ.cpp:1718 vs :1750: same strict test, EA only skips
  // rows that cannot pass it), so both run the filter kernels; VAQGPU_SCAN_V1 keeps the literal exhaustive loop.
  // TI / visit runs on the fp16 filter kernel too (row tiles of unvisited clusters are skipped) when the cluster
  // ranges partition the rows in ascending order (what clusterTI produces) and the tables fit; otherwise, and under
  // VAQGPU_SCAN_V1, it runs on the lane-per-row kernel over per-query row ranges.
  bool filter = !(flags & VAQGPU_SCAN_V1) && h->n_rows > 0 && (!ti || (h->d_tile_cl && !(flags & VAQGPU_SCAN_F32)));
  const bool want_sqrt = (flags & VAQGPU_SQRT) != 0;
  int launches = 0;

  if (record) CU(cudaEventRecord(h->ev[0], st));
  const float *d_qproj = d_queries;
  if (!projected) {
    CU(h->w_qproj.ensure((size_t)nq * h->D * sizeof(float)));
    CU(launch_project(d_queries, nq, h->D, h->d_eig, (float *)h->w_qproj.p, st));
    d_qproj = (const float *)h->w_qproj.p;
    launches++;
  }
  if (record) CU(cudaEventRecord(h->ev[1], st));

  const int64_t n_tiles = (h->n_rows + kTileRows - 1) / kTileRows;
  ScanLayout lay;
  LutPlan plan;
  int32_t res_floats = 0, spill_floats = 0;

  // ---- fp16 lower-bound tables, query tiles of 8 (default when eight queries' tables fit) -------------
  bool filter16 = filter && !(flags & VAQGPU_SCAN_F32) && nq >= tune_knob("min16", 1);
  if (filter16) {
    apply_residency(h, (size_t)1 << 30, 8, lay, plan, res_floats, spill_floats);      // everything resident
    if (adc_filter16_smem_bytes(plan.row_stride, k, 1024, ti ? h->C : 0) > kSmemCap) filter16 = false;
  }
  if (ti && !filter16) filter = false;          // TI without the fp16 kernel: lane-per-row kernel over row ranges
  {
    int rc = filter ? ensure_layout(h, st) : ensure_rowid(h, st);
    if (rc) return rc;
  }

  // Per-query bound array of this search (filter kernels).  With an exported array (row-sharded deployment) the
  // search uses one half and resets the other for the next search: peers only write a half between this shard's
  // reset of it and the end of the search that uses it, because every search ends in a collective (the all-gather
  // of the key lists) that no shard passes before all shards finished scanning.
  uint32_t *thr_all = nullptr;
  PeerBounds peers{};
  if (filter) {
    if (h->d_bounds && nq <= h->bounds_cap) {
      thr_all = h->d_bounds + (size_t)h->bounds_parity * h->bounds_cap;
      uint32_t *next = h->d_bounds + (size_t)(h->bounds_parity ^ 1) * h->bounds_cap;
      for (int i = 0; i < h->peers.n; i++) peers.p[i] = h->peers.p[i] + (size_t)h->bounds_parity * h->bounds_cap;
      peers.n = h->peers.n;
      CU(launch_fill_u32(next, h->bounds_cap, 0xFFFFFFFFu, st));
      h->bounds_parity ^= 1;
    } else {
      const bool fresh = h->w_thr.p == nullptr;
      CU(h->w_thr.ensure((size_t)nq * sizeof(uint32_t)));
      thr_all = (uint32_t *)h->w_thr.p;
      // development (VAQGPU_TUNE=keepthr=1): repeat searches of the same batch start from the previous search's final bounds
      if (fresh || !tune_knob("keepthr", 0)) CU(launch_fill_u32(thr_all, nq, 0xFFFFFFFFu, st));
    }
    launches++;
  }

  if (filter16) {
    // 32 warps x one row per lane (measured faster than 16 warps x two rows per lane at every chunk length: the scan
    // is bound by shared-memory wavefronts and issue slots, which more resident warps fill better; the 512-thread
    // variant stays selectable with VAQGPU_TUNE=threads=512)
    const int T = 8, threads = tune_knob("threads", 1024);
    const size_t smem = adc_filter16_smem_bytes(plan.row_stride, k, ti ? 1024 : threads, ti ? h->C : 0);
    const int nwarps = threads / 32;
    const size_t bytes_per_q = (size_t)plan.row_stride * 4;
    int qb_max = (int)std::max<size_t>(T, std::min<size_t>((size_t)nq, kLutWorkspaceBytes / bytes_per_q));
    qb_max = (qb_max + T - 1) / T * T;
    const int qtiles_first = (std::min(nq, qb_max) + T - 1) / T;
    int64_t n_chunks = tune_knob("chunks", (int)choose_chunks(n_tiles, qtiles_first, nwarps, h->num_sms));
    const int64_t chunk_tiles = (n_tiles + n_chunks - 1) / n_chunks;
    n_chunks = (n_tiles + chunk_tiles - 1) / chunk_tiles;
    if (n_chunks > 65535) return fail(VAQGPU_EINVAL, "index too large for one launch (%lld chunks)", (long long)n_chunks);
    const int out_slots = (int)n_chunks;

    CU(h->w_lut.ensure((size_t)qb_max * bytes_per_q));
    CU(h->w_lut16.ensure((size_t)qb_max * bytes_per_q / 2));
    CU(h->w_scale.ensure((size_t)qb_max * sizeof(float)));
    CU(h->w_keys.ensure((size_t)qb_max * out_slots * k * sizeof(uint64_t)));
    if (out_slots > 16) CU(h->w_scratch.ensure((size_t)2 * qb_max * ((out_slots + 15) / 16) * k * sizeof(uint64_t)));
    // queries are re-grouped into tiles by nearest cluster: TI clusters (TI search) or the scan-order clusters
    const bool ordered = !ti && h->oc_C > 0 && h->C == 0 && tune_knob("rot", 1) != 0;
    const int plan_C = ti ? h->C : ordered ? h->oc_C : 0;
    if (plan_C > 0) {
      CU(h->w_vis.ensure((size_t)qb_max * plan_C));
      CU(h->w_near.ensure((size_t)qb_max * sizeof(int32_t)));
      CU(h->w_perm.ensure((size_t)qb_max * sizeof(int32_t)));
      CU(h->w_qperm.ensure((size_t)qb_max * h->D * sizeof(float)));
      CU(h->w_tmask.ensure((size_t)(qb_max / T) * plan_C));
      CU(h->w_rot.ensure((size_t)(qb_max / T) * sizeof(int32_t)));
    }

    for (int q0 = 0; q0 < nq; q0 += qb_max) {
      const int qb = std::min(qb_max, nq - q0);
      const int qb_pad = (qb + T - 1) / T * T;
      const float *qp = d_qproj + (size_t)q0 * h->D;
      if (ti) {
        // which clusters each query visits, queries grouped into tiles by nearest cluster, per-(tile, cluster) masks
        CU(launch_ti_plan(qp, qb, h->D, h->d_clusters_t, h->C, h->segdims, h->d_cl_rule ? h->d_cl_rule : h->d_cl_size, h->visit, k,
                          (uint8_t *)h->w_vis.p, (int32_t *)h->w_near.p, (int32_t *)h->w_perm.p, (float *)h->w_qperm.p,
                          (uint8_t *)h->w_tmask.p, st));
        qp = (const float *)h->w_qperm.p;          // the batch in tile order
        launches += 3;
        if (tune_knob("rot", 1) != 0) {
          CU(launch_rot_tiles((const int32_t *)h->w_near.p, (const int32_t *)h->w_perm.p, qb, h->d_cl_start, (int32_t *)h->w_rot.p, st));
          launches++;
        }
      } else if (ordered) {
        // nearest scan-order cluster of each query, queries grouped into tiles by it, start tile of each query tile
        CU(launch_ti_plan(qp, qb, h->D, h->d_oc_centres_t, h->oc_C, h->oc_dims, h->d_oc_size, 1.f, k, (uint8_t *)h->w_vis.p,
                          (int32_t *)h->w_near.p, (int32_t *)h->w_perm.p, (float *)h->w_qperm.p, (uint8_t *)h->w_tmask.p, st));
        CU(launch_rot_tiles((const int32_t *)h->w_near.p, (const int32_t *)h->w_perm.p, qb, h->d_oc_start, (int32_t *)h->w_rot.p, st));
        qp = (const float *)h->w_qperm.p;
        launches += 4;
      }
      CU(launch_lut_build(qp, qb, qb_pad, h->D, h->d_centroids, h->d_cent_rmax, plan, (float *)h->w_lut.p, h->w_lut16.p,
                          (float *)h->w_scale.p, st));
      launches += 2;
      if (record && q0 == 0) CU(cudaEventRecord(h->ev[2], st));
      AdcFilter16Args a{};
      a.codes = h->d_codes; a.n_rows = h->n_rows;
      a.lut16 = h->w_lut16.p; a.lut32 = (const float *)h->w_lut.p; a.scale = (const float *)h->w_scale.p;
      a.lut_stride = plan.row_stride;
      a.nq = qb; a.k = k; a.out_slots = out_slots; a.slot_base = 0;
      a.tile_lo = 0; a.tile_hi = n_tiles; a.chunk_tiles = (int32_t)chunk_tiles;
      a.out_keys = (uint64_t *)h->w_keys.p;
      a.thr_global = thr_all + q0;
      a.peers = peers;
      for (int i = 0; i < peers.n; i++) a.peers.p[i] += q0;
      a.seed = tune_knob("seed", 1);
      a.q3_cap = std::max(1, std::min(32, tune_knob("q3cap", 32)));
      a.seed_rows = std::max(1, std::min(16, tune_knob("spl", 3)));
      a.rowid = h->d_rowid;
      if (ti) {
        a.tile_cl = h->d_tile_cl; a.cl_start = h->d_cl_start; a.tmask = (const uint8_t *)h->w_tmask.p; a.C = h->C;
        a.qmap = (const int32_t *)h->w_perm.p;
        if (tune_knob("rot", 1) != 0) a.rot_tile = (const int32_t *)h->w_rot.p;
      } else if (ordered) {
        a.qmap = (const int32_t *)h->w_perm.p;
        a.rot_tile = (const int32_t *)h->w_rot.p;
      }
      a.chunks_fast = tune_knob("chunks_fast", 0);
      // a code matrix far larger than L2 streams from HBM: keep more of it in flight than the register prefetch holds
      a.l2_prefetch = tune_knob("l2pf", (size_t)h->n_rows * lay.W * 16 > (size_t)(96u << 20) ? 4 : 0);
      a.lay = lay;
      const bool dbg = tune_knob("dbg", 0) != 0;
      const size_t n_cta = (size_t)((qb + T - 1) / T) * n_chunks;
      if (dbg) {
        CU(h->w_dbg.ensure(n_cta * kDbgSlots * sizeof(long long)));
        CU(cudaMemsetAsync(h->w_dbg.p, 0, n_cta * kDbgSlots * sizeof(long long), st));
        a.dbg = (long long *)h->w_dbg.p;
      }
      CU(launch_adc_filter16_scan(a, threads, smem, st));
      if (dbg) {        // development only: per-CTA phase durations in SM clocks
        std::vector<long long> hd(n_cta * kDbgSlots);
        CU(cudaStreamSynchronize(st));
        CU(cudaMemcpy(hd.data(), h->w_dbg.p, hd.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        double ph[5] = {0, 0, 0, 0, 0};
        for (size_t c = 0; c < n_cta; c++)
          for (int i = 0; i < 5; i++) ph[i] += (double)(hd[c * kDbgSlots + i + 1] - hd[c * kDbgSlots + i]);
        fprintf(stderr, "[vaqgpu dbg] per-CTA clocks: table staging %.0f, seeding %.0f, scan (warp 0) %.0f, warp-0 tail %.0f, wait for other warps %.0f\n",
                ph[0] / n_cta, ph[1] / n_cta, ph[2] / n_cta, ph[3] / n_cta, ph[4] / n_cta);
        double sc[16] = {0};
        for (size_t c = 0; c < n_cta; c++)
          for (int i = 0; i < 16; i++) sc[i] += (double)hd[c * kDbgSlots + 8 + i];
        if (sc[0] > 0)        // -DVAQGPU_STATS build
          fprintf(stderr, "[vaqgpu stats] per CTA: stage-1 survivors %.0f (live pairs %.0f), level-1 passes %.0f -> rows %.0f, level-2 passes %.0f -> rows %.0f, "
                          "exact passes %.0f, pairs %.0f in %.0f rounds, insert candidates %.0f; warp-0 clocks level 1/2/3: %.0f %.0f %.0f\n",
                  sc[0] / n_cta, sc[9] / n_cta, sc[1] / n_cta, sc[2] / n_cta, sc[3] / n_cta, sc[4] / n_cta, sc[5] / n_cta, sc[6] / n_cta, sc[7] / n_cta,
                  sc[8] / n_cta, sc[13] / n_cta, sc[14] / n_cta, sc[15] / n_cta);
      }
      launches++;
      if (record && q0 + qb >= nq) CU(cudaEventRecord(h->ev[3], st));
      // TI: list q belongs to the query in slot q of the tile order -> written to that query's row; labels through id_map
      CU(launch_merge_keys((const uint64_t *)h->w_keys.p, k, (int64_t)out_slots * k, out_slots, qb, k, want_sqrt ? 1 : 0, 0,
                           d_labels ? d_labels + (size_t)q0 * k : nullptr, d_dists ? (void *)(d_dists + (size_t)q0 * k) : nullptr,
                           d_keys ? d_keys + (size_t)q0 * k : nullptr, ti ? h->d_id_map : nullptr, h->id_base,
                           (uint64_t *)h->w_scratch.p, st, (ti || ordered) ? (const int32_t *)h->w_perm.p : nullptr));
      launches += out_slots > 16 ? 2 : 1;
    }
    if (record) { CU(cudaEventRecord(h->ev[4], st)); h->timed = true; }
    h->cfg[0] = threads; h->cfg[1] = (int32_t)n_chunks; h->cfg[2] = res_floats; h->cfg[3] = 0;
    h->cfg[4] = (int32_t)smem; h->cfg[5] = lay.W; h->cfg[6] = launches; h->cfg[7] = qb_max;
    h->cfg[8] = T; h->cfg[9] = 3; h->cfg[10] = h->opt_n > 0; h->cfg[11] = (int32_t)(h->layout_ms * 1000.f);
    return VAQGPU_OK;
  }

  if (filter && !ti) {
    // ---- query-tile width T and residency -------------------------------------------------------
    const int threads = tune_knob("threads", 1024);
    int T = nq >= 8 ? 8 : (nq >= 3 ? 4 : nq);
    T = std::min(T, tune_knob("T", 8));
    // Tables that do not fit for T queries.  T > 1 only with every table resident: with tables spilled, a wider tile
    // keeps fewer of the leading tables in shared memory, and stage-1 gathers from L2 cost far more than the extra
    // passes over the rows (measured, GIST1M shape 1 K queries: T = 1 / 2 / 4 / 8 -> 15.3 / 23.8 / 33.1 / 35.6 ms;
    // the reference's min2/max13 SIFT1M setting, 10 K queries: 86 / 103 / 113 / 172 ms).  VAQGPU_TUNE=spillT=n overrides.
    const int spill_T = std::min(T, tune_knob("spillT", 1));
    for (;; T >>= 1) {
      const size_t fixed = adc_filter_smem_bytes(0, T, k, threads) + 1024;
      if (fixed >= kSmemCap) {
        if (T == 1) return fail(VAQGPU_EINVAL, "k=%d does not fit the scan's shared memory", k);
        continue;
      }
      const size_t budget = (kSmemCap - fixed) / (4 * (size_t)T);
      if ((size_t)h->total_entries + 4 <= budget || T <= spill_T) {
        apply_residency(h, budget, T, lay, plan, res_floats, spill_floats);
        break;
      }
    }
    const size_t smem = adc_filter_smem_bytes(res_floats, T, k, threads);
    if (smem > kSmemCap) return fail(VAQGPU_EINVAL, "k=%d needs %zu B shared memory (> %zu)", k, smem, kSmemCap);
    const int nwarps = threads / 32;
    const size_t bytes_per_q = (size_t)plan.row_stride * 4;
    int qb_max = (int)std::max<size_t>(T, std::min<size_t>((size_t)nq, kLutWorkspaceBytes / bytes_per_q));
    qb_max = (qb_max + T - 1) / T * T;
    // ---- row chunks: enough CTAs to fill the machine, chunks small enough to stay L2-resident while
    // every query tile sweeps them (query tiles are the fast grid dimension)
    const int qtiles_first = (std::min(nq, qb_max) + T - 1) / T;
    int64_t n_chunks = tune_knob("chunks", (int)choose_chunks(n_tiles, qtiles_first, nwarps, h->num_sms));
    int64_t chunk_tiles = (n_tiles + n_chunks - 1) / n_chunks;
    n_chunks = (n_tiles + chunk_tiles - 1) / chunk_tiles;
    if (n_chunks > 65535) return fail(VAQGPU_EINVAL, "index too large for one launch (%lld chunks)", (long long)n_chunks);

    const int out_slots = (int)n_chunks;

    CU(h->w_lut.ensure((size_t)qb_max * bytes_per_q));
    CU(h->w_keys.ensure((size_t)qb_max * out_slots * k * sizeof(uint64_t)));
    if (out_slots > 16) CU(h->w_scratch.ensure((size_t)2 * qb_max * ((out_slots + 15) / 16) * k * sizeof(uint64_t)));

    for (int q0 = 0; q0 < nq; q0 += qb_max) {
      const int qb = std::min(qb_max, nq - q0);
      const int qb_pad = (qb + T - 1) / T * T;
      const float *qp = d_qproj + (size_t)q0 * h->D;
      CU(launch_lut_build(qp, qb, qb_pad, h->D, h->d_centroids, h->d_cent_rmax, plan, (float *)h->w_lut.p, nullptr, nullptr, st));
      launches += 1;
      if (record && q0 == 0) CU(cudaEventRecord(h->ev[2], st));
      AdcFilterArgs a{};
      a.codes = h->d_codes; a.n_rows = h->n_rows;
      a.lut = (const float *)h->w_lut.p; a.lut_stride = plan.row_stride; a.smem_lut_floats = res_floats;
      a.nq = qb; a.k = k; a.out_slots = out_slots;
      a.out_keys = (uint64_t *)h->w_keys.p;
      a.thr_global = thr_all + q0;
      a.peers = peers;
      for (int i = 0; i < peers.n; i++) a.peers.p[i] += q0;
      a.lay = lay;
      a.rowid = h->d_rowid;
      a.seed = tune_knob("seed", 1);
      a.tile_lo = 0; a.tile_hi = n_tiles; a.chunk_tiles = (int32_t)chunk_tiles; a.slot_base = 0;
      CU(launch_adc_filter_scan(a, T, threads, smem, st));
      launches++;
      if (record && q0 + qb >= nq) CU(cudaEventRecord(h->ev[3], st));
      CU(launch_merge_keys((const uint64_t *)h->w_keys.p, k, (int64_t)out_slots * k, out_slots, qb, k, want_sqrt ? 1 : 0, 0,
                           d_labels ? d_labels + (size_t)q0 * k : nullptr, d_dists ? (void *)(d_dists + (size_t)q0 * k) : nullptr,
                           d_keys ? d_keys + (size_t)q0 * k : nullptr, nullptr, h->id_base, (uint64_t *)h->w_scratch.p, st));
      launches += out_slots > 16 ? 2 : 1;
    }
    if (record) { CU(cudaEventRecord(h->ev[4], st)); h->timed = true; }
    h->cfg[0] = threads; h->cfg[1] = (int32_t)n_chunks; h->cfg[2] = res_floats; h->cfg[3] = spill_floats;
    h->cfg[4] = (int32_t)smem; h->cfg[5] = lay.W; h->cfg[6] = launches; h->cfg[7] = qb_max;
    h->cfg[8] = T; h->cfg[9] = 2; h->cfg[10] = h->opt_n > 0; h->cfg[11] = (int32_t)(h->layout_ms * 1000.f);
    return VAQGPU_OK;
  }

  // ---- lane-per-row kernel (HEAP / TI / forced) ---------------------------------------------------
  const int threads = kScanThreads;
  {
    const size_t fixed = scan_smem_bytes(0, k, threads) + 1024;
    if (fixed >= kSmemCap) return fail(VAQGPU_EINVAL, "k=%d needs more shared memory than an SM has", k);
    apply_residency(h, (kSmemCap - fixed) / 4, 1, lay, plan, res_floats, spill_floats);
  }
  const size_t smem = scan_smem_bytes(res_floats, k, threads);
  int ctas_per_sm = 1;
  CU(adc_scan_occupancy(lay.W, threads, smem, &ctas_per_sm));
  if (ctas_per_sm < 1) return fail(VAQGPU_ECUDA, "ADC scan kernel does not fit an SM with %zu B shared memory", smem);
  const int nwarps = threads / 32;
  // grid.y of adc_scan_kernel is the query: at most 65535 per launch
  const int qb_max = (int)std::max<size_t>(1, std::min<size_t>({(size_t)nq, kLutWorkspaceBytes / ((size_t)plan.row_stride * 4), (size_t)65535}));
  // CTAs per query: fill ~2 waves of the machine when there are few queries, but leave each
  // warp at least 8 tiles so the per-CTA LUT staging stays amortised.
  const int64_t target = (int64_t)h->num_sms * ctas_per_sm * 2;
  int splits = (int)std::max<int64_t>(1, (target + std::min(nq, qb_max) - 1) / std::min(nq, qb_max));
  const int64_t max_splits = std::max<int64_t>(1, n_tiles / (nwarps * 8));
  splits = (int)std::min<int64_t>(splits, max_splits);
  if (ti) splits = std::min(splits, 4);

  CU(h->w_lut.ensure((size_t)qb_max * plan.row_stride * sizeof(float)));
  CU(h->w_keys.ensure((size_t)qb_max * splits * k * sizeof(uint64_t)));
  if (splits > 16) CU(h->w_scratch.ensure((size_t)2 * qb_max * ((splits + 15) / 16) * k * sizeof(uint64_t)));
  if (ti) {
    CU(h->w_ranges.ensure((size_t)qb_max * h->C * sizeof(int2)));
    CU(h->w_nranges.ensure((size_t)qb_max * sizeof(int32_t)));
  }

  for (int q0 = 0; q0 < nq; q0 += qb_max) {
    const int qb = std::min(qb_max, nq - q0);
    const float *qp = d_qproj + (size_t)q0 * h->D;
    if (ti) {
      CU(launch_rank_clusters(qp, qb, h->D, h->d_clusters, h->C, h->segdims, h->d_cl_start, h->d_cl_size, h->d_cl_rule ? h->d_cl_rule : h->d_cl_size, h->visit, k,
                              (int2 *)h->w_ranges.p, (int32_t *)h->w_nranges.p, st));
      launches++;
    }
    CU(launch_lut_build(qp, qb, qb, h->D, h->d_centroids, h->d_cent_rmax, plan, (float *)h->w_lut.p, nullptr, nullptr, st));
    launches++;
    if (record && q0 == 0) CU(cudaEventRecord(h->ev[2], st));
    AdcScanArgs a{};
    a.codes = h->d_codes; a.n_rows = h->n_rows;
    a.lut = (const float *)h->w_lut.p; a.lut_stride = plan.row_stride;
    a.smem_lut_floats = res_floats;
    a.nq = qb; a.k = k; a.splits = splits;
    a.early_abandon = ea ? 1 : 0;
    a.use_tma = 1;
    a.out_keys = (uint64_t *)h->w_keys.p;
    a.rowid = h->d_rowid;
    if (ti) { a.ranges = (const int2 *)h->w_ranges.p; a.n_ranges = (const int32_t *)h->w_nranges.p; a.max_ranges = h->C; }
    a.lay = lay;
    CU(launch_adc_scan(a, threads, smem, st));
    launches++;
    if (record && q0 + qb >= nq) CU(cudaEventRecord(h->ev[3], st));
    CU(launch_merge_keys((const uint64_t *)h->w_keys.p, k, (int64_t)splits * k, splits, qb, k, want_sqrt ? 1 : 0, 0,
                         d_labels ? d_labels + (size_t)q0 * k : nullptr, d_dists ? (void *)(d_dists + (size_t)q0 * k) : nullptr,
                         d_keys ? d_keys + (size_t)q0 * k : nullptr, ti ? h->d_id_map : nullptr, h->id_base,
                         (uint64_t *)h->w_scratch.p, st));
    launches += splits > 16 ? 2 : 1;
  }
  if (record) { CU(cudaEventRecord(h->ev[4], st)); h->timed = true; }
  h->cfg[0] = threads; h->cfg[1] = splits; h->cfg[2] = res_floats; h->cfg[3] = spill_floats;
  h->cfg[4] = (int32_t)smem; h->cfg[5] = lay.W; h->cfg[6] = launches; h->cfg[7] = qb_max;
  h->cfg[8] = 1; h->cfg[9] = 1; h->cfg[10] = h->opt_n > 0; h->cfg[11] = (int32_t)(h->layout_ms * 1000.f);
  return VAQGPU_OK;
}

}  // namespace

extern "C" {

const char *vaqgpu_last_error(void) { return g_err; }

int vaqgpu_device_count(int *count) {
  if (!count) return fail(VAQGPU_EINVAL, "count is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { *count = 0; return fail(VAQGPU_ECUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
  *count = n;
  return VAQGPU_OK;
}

int vaqgpu_create(const vaqgpu_model_desc *m, int device, vaqgpu_t **out) {
  if (!m || !out) return fail(VAQGPU_EINVAL, "model/out is NULL");
  *out = nullptr;
  if (m->M < 1 || m->M > kMaxSubspaces) return fail(VAQGPU_EINVAL, "M=%d out of range [1,%d]", m->M, kMaxSubspaces);
  if (m->L < 1 || m->L > 256) return fail(VAQGPU_EINVAL, "L=%d out of range", m->L);
  if (m->D != m->M * m->L) return fail(VAQGPU_EINVAL, "D=%d != M*L=%d", m->D, m->M * m->L);
  if (!m->bits || !m->centroids) return fail(VAQGPU_EINVAL, "bits/centroids is NULL");
  int64_t sum = 0;
  for (int s = 0; s < m->M; s++) {
    if (m->bits[s] < 1 || m->bits[s] > 15) return fail(VAQGPU_EINVAL, "bits[%d]=%d out of range [1,15]", s, m->bits[s]);
    sum += m->bits[s];
  }
  if (sum > 1024) return fail(VAQGPU_EINVAL, "sum(bits)=%lld > 1024", (long long)sum);
  int rc = check_device(device);
  if (rc) return rc;
  DeviceGuard g(device);
  vaqgpu_index *h = new (std::nothrow) vaqgpu_index();
  if (!h) return fail(VAQGPU_ENOMEM, "host allocation failed");
  h->device = device;
  h->D = m->D; h->M = m->M; h->L = m->L;
  h->bits.assign(m->bits, m->bits + m->M);
  cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
  rc = plan_model(h);
  if (rc) { delete h; return rc; }
#define CUX(call)                                                                                        \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      vaqgpu_destroy(h);                                                                                 \
      return fail(VAQGPU_ECUDA, "%s: %s", #call, cudaGetErrorString(e_));                                \
    }                                                                                                    \
  } while (0)
  size_t cent_floats = 0;
  for (int s = 0; s < m->M; s++) cent_floats += ((size_t)1 << m->bits[s]) * m->L;
  CUX(cudaMalloc(&h->d_centroids, cent_floats * sizeof(float)));
  CUX(cudaMemcpy(h->d_centroids, m->centroids, cent_floats * sizeof(float), cudaMemcpyHostToDevice));
  {
    std::vector<float> rmax(m->M, 0.f);
    size_t off = 0;
    for (int s = 0; s < m->M; s++) {
      const size_t K = (size_t)1 << m->bits[s];
      for (size_t c = 0; c < K; c++) {
        double n2 = 0;
        for (int j = 0; j < m->L; j++) { const double v = m->centroids[off + c * m->L + j]; n2 += v * v; }
        rmax[s] = std::max(rmax[s], (float)(sqrt(n2) * 1.0001));
      }
      off += K * m->L;
    }
    CUX(cudaMalloc(&h->d_cent_rmax, m->M * sizeof(float)));
    CUX(cudaMemcpy(h->d_cent_rmax, rmax.data(), m->M * sizeof(float), cudaMemcpyHostToDevice));
  }
  if (m->eig_real) {
    CUX(cudaMalloc(&h->d_eig, (size_t)m->D * m->D * sizeof(float)));
    CUX(cudaMemcpy(h->d_eig, m->eig_real, (size_t)m->D * m->D * sizeof(float), cudaMemcpyHostToDevice));
  }
  CUX(cudaMalloc(&h->d_bits, m->M * sizeof(int32_t)));
  CUX(cudaMemcpy(h->d_bits, m->bits, m->M * sizeof(int32_t), cudaMemcpyHostToDevice));
  CUX(cudaMalloc(&h->d_ent_off, (m->M + 1) * sizeof(int32_t)));
  CUX(cudaMemcpy(h->d_ent_off, h->plan.ent_off, (m->M + 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
  CUX(cudaMalloc(&h->d_cent_off, m->M * sizeof(int32_t)));
  CUX(cudaMemcpy(h->d_cent_off, h->plan.cent_off, m->M * sizeof(int32_t), cudaMemcpyHostToDevice));
  CUX(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  for (auto &e : h->ev) CUX(cudaEventCreate(&e));
#undef CUX
  *out = h;
  return VAQGPU_OK;
}

void vaqgpu_destroy(vaqgpu_t *h) {
  if (!h) return;
  DeviceGuard g(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->d_centroids); cudaFree(h->d_cent_rmax); cudaFree(h->d_eig); cudaFree(h->d_bits); cudaFree(h->d_ent_off); cudaFree(h->d_cent_off);
  cudaFree(h->d_codes); cudaFree(h->d_clusters); cudaFree(h->d_cl_start); cudaFree(h->d_cl_size); cudaFree(h->d_cl_rule); cudaFree(h->d_tile_cl); cudaFree(h->d_clusters_t);
  cudaFree(h->d_id_map); cudaFree(h->d_raw);
  for (void *p : h->ipc_opened) cudaIpcCloseMemHandle(p);
  cudaFree(h->d_bounds);
  cudaFree(h->d_rowid);
  cudaFree(h->d_oc_centres_t); cudaFree(h->d_oc_start); cudaFree(h->d_oc_size);
  for (DevBuf *b : {&h->w_vis, &h->w_near, &h->w_perm, &h->w_qperm, &h->w_tmask, &h->w_rot, &h->w_lsrc, &h->w_lscratch, &h->w_dbg, &h->w_lut16, &h->w_scale, &h->w_thr, &h->w_q, &h->w_qproj, &h->w_lut, &h->w_keys, &h->w_scratch, &h->w_ranges, &h->w_nranges, &h->w_stage,
                    &h->w_labels, &h->w_dists, &h->w_outkeys, &h->w_cdf, &h->w_x})
    b->release();
  for (auto &e : h->ev) if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int vaqgpu_set_id_base(vaqgpu_t *h, int64_t id_base) {
  if (!h) return fail(VAQGPU_EINVAL, "handle is NULL");
  h->id_base = id_base;
  return VAQGPU_OK;
}

int vaqgpu_reserve(vaqgpu_t *h, int64_t n_total) {
  if (!h) return fail(VAQGPU_EINVAL, "handle is NULL");
  DeviceGuard g(h->device);
  if (n_total > h->cap_rows) {
    // exact reservation (no head-room) by temporarily pretending the capacity is full
    int64_t cap = h->cap_rows;
    int64_t want = (n_total + kTileRows - 1) / kTileRows * kTileRows;
    uint4 *nw = nullptr;
    CU(cudaMalloc(&nw, (size_t)want * h->lay.W * sizeof(uint4)));
    const int64_t used_rows = (h->n_rows + kTileRows - 1) / kTileRows * kTileRows;
    const size_t used = (size_t)used_rows * h->lay.W * sizeof(uint4);
    if (h->d_codes && used) CU(cudaMemcpyAsync(nw, h->d_codes, used, cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaMemsetAsync(reinterpret_cast<unsigned char *>(nw) + used, 0, (size_t)want * h->lay.W * sizeof(uint4) - used, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (h->d_codes) cudaFree(h->d_codes);
    h->d_codes = nw;
    h->cap_rows = want;
    (void)cap;
  }
  return VAQGPU_OK;
}

int vaqgpu_add_codes_u16(vaqgpu_t *h, const uint16_t *codes, int64_t n) {
  if (!h || (!codes && n > 0)) return fail(VAQGPU_EINVAL, "handle/codes is NULL");
  if (n < 0) return fail(VAQGPU_EINVAL, "n=%lld", (long long)n);
  if (n == 0) return VAQGPU_OK;
  if (h->n_rows + n > 0x7FFFFFFFll) return fail(VAQGPU_EINVAL, "more than 2^31-1 rows per index (labels are int32, utils/Types.hpp:100)");
  DeviceGuard g(h->device);
  {
    int rc_ = drop_clusters(h);
    if (rc_) return rc_;
  }
  CU(grow_codes(&h->d_codes, &h->cap_rows, h->n_rows, h->lay.W, h->n_rows + n, h->stream));
  const int64_t chunk = std::min<int64_t>(n, kStageRows);
  CU(h->w_stage.ensure((size_t)chunk * h->M * sizeof(uint16_t)));
  for (int64_t r = 0; r < n; r += chunk) {
    const int64_t c = std::min(chunk, n - r);
    CU(cudaMemcpyAsync(h->w_stage.p, codes + (size_t)r * h->M, (size_t)c * h->M * sizeof(uint16_t), cudaMemcpyHostToDevice, h->stream));
    CU(launch_pack_codes((const uint16_t *)h->w_stage.p, c, h->n_rows + r, h->lay, h->d_codes, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  h->n_rows += n;
  return VAQGPU_OK;
}

int vaqgpu_encode_add(vaqgpu_t *h, const float *x_proj, int64_t n) {
  if (!h || (!x_proj && n > 0)) return fail(VAQGPU_EINVAL, "handle/x_proj is NULL");
  if (n < 0) return fail(VAQGPU_EINVAL, "n=%lld", (long long)n);
  if (n == 0) return VAQGPU_OK;
  if (h->n_rows + n > 0x7FFFFFFFll) return fail(VAQGPU_EINVAL, "more than 2^31-1 rows per index");
  DeviceGuard g(h->device);
  {
    int rc_ = drop_clusters(h);
    if (rc_) return rc_;
  }
  CU(grow_codes(&h->d_codes, &h->cap_rows, h->n_rows, h->lay.W, h->n_rows + n, h->stream));
  const int64_t chunk = std::min<int64_t>(n, std::max<int64_t>(1024, (int64_t)(256ull << 20) / (h->D * 4)));
  CU(h->w_stage.ensure((size_t)chunk * h->M * sizeof(uint16_t)));
  CU(h->w_x.ensure((size_t)chunk * h->D * sizeof(float)));
  for (int64_t r = 0; r < n; r += chunk) {
    const int64_t c = std::min(chunk, n - r);
    CU(cudaMemcpyAsync(h->w_x.p, x_proj + (size_t)r * h->D, (size_t)c * h->D * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CU(launch_encode((const float *)h->w_x.p, c, h->d_centroids, h->plan, (uint16_t *)h->w_stage.p, h->stream));
    CU(launch_pack_codes((const uint16_t *)h->w_stage.p, c, h->n_rows + r, h->lay, h->d_codes, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  h->n_rows += n;
  return VAQGPU_OK;
}

int vaqgpu_add_codes_synthetic(vaqgpu_t *h, int64_t n, uint64_t seed, const float *cdf) {
  if (!h) return fail(VAQGPU_EINVAL, "handle is NULL");
  if (n < 0) return fail(VAQGPU_EINVAL, "n=%lld", (long long)n);
  if (n == 0) return VAQGPU_OK;
  if (h->n_rows + n > 0x7FFFFFFFll) return fail(VAQGPU_EINVAL, "more than 2^31-1 rows per index");
  DeviceGuard g(h->device);
  {
    int rc_ = drop_clusters(h);
    if (rc_) return rc_;
  }
  CU(grow_codes(&h->d_codes, &h->cap_rows, h->n_rows, h->lay.W, h->n_rows + n, h->stream));
  const float *d_cdf = nullptr;
  if (cdf) {
    CU(h->w_cdf.ensure((size_t)h->total_entries * sizeof(float)));
    CU(cudaMemcpyAsync(h->w_cdf.p, cdf, (size_t)h->total_entries * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    d_cdf = (const float *)h->w_cdf.p;
  }
  const int64_t chunk = std::min<int64_t>(n, kStageRows);
  CU(h->w_stage.ensure((size_t)chunk * h->M * sizeof(uint16_t)));
  for (int64_t r = 0; r < n; r += chunk) {
    const int64_t c = std::min(chunk, n - r);
    CU(launch_synth_codes((uint16_t *)h->w_stage.p, c, h->id_base + h->n_rows + r, h->M, h->d_bits, d_cdf, h->d_ent_off, seed, h->stream));
    CU(launch_pack_codes((const uint16_t *)h->w_stage.p, c, h->n_rows + r, h->lay, h->d_codes, h->stream));
  }
  CU(cudaStreamSynchronize(h->stream));
  h->n_rows += n;
  return VAQGPU_OK;
}

int vaqgpu_num_rows(const vaqgpu_t *h, int64_t *n) {
  if (!h || !n) return fail(VAQGPU_EINVAL, "handle/n is NULL");
  *n = h->n_rows;
  return VAQGPU_OK;
}

int vaqgpu_row_bytes(const vaqgpu_t *h, int32_t *bytes) {
  if (!h || !bytes) return fail(VAQGPU_EINVAL, "handle/bytes is NULL");
  *bytes = h->lay.W * 16;
  return VAQGPU_OK;
}

int vaqgpu_get_codes_u16(vaqgpu_t *h, int64_t row0, int64_t n, uint16_t *out) {
  if (!h || (!out && n > 0)) return fail(VAQGPU_EINVAL, "handle/out is NULL");
  if (row0 < 0 || n < 0 || row0 + n > h->n_rows) return fail(VAQGPU_EINVAL, "rows [%lld,%lld) outside [0,%lld)", (long long)row0, (long long)(row0 + n), (long long)h->n_rows);
  if (n == 0) return VAQGPU_OK;
  DeviceGuard g(h->device);
  const int64_t chunk = std::min<int64_t>(n, kStageRows);
  CU(h->w_stage.ensure((size_t)chunk * h->M * sizeof(uint16_t)));
  {
    int rc = ensure_rowid(h, h->stream);
    if (rc) return rc;
  }
  for (int64_t r = 0; r < n; r += chunk) {
    const int64_t c = std::min(chunk, n - r);
    // storage rows that can hold the original rows [row0 + r, row0 + r + c): the windows they fall into
    // (windows that follow TI clusters are not aligned: every stored row is a candidate then)
    const int64_t slo = !h->d_rowid ? row0 + r : h->cluster_windows ? 0 : ((row0 + r) / kLayoutWin) * kLayoutWin;
    const int64_t shi = !h->d_rowid ? row0 + r + c
                        : h->cluster_windows ? h->n_rows : std::min<int64_t>(h->n_rows, ((row0 + r + c + kLayoutWin - 1) / kLayoutWin) * kLayoutWin);
    CU(launch_unpack_codes(h->d_codes, row0 + r, c, h->lay, (uint16_t *)h->w_stage.p, h->d_rowid, slo, shi, h->stream));
    CU(cudaMemcpyAsync(out + (size_t)r * h->M, h->w_stage.p, (size_t)c * h->M * sizeof(uint16_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  return VAQGPU_OK;
}

int vaqgpu_get_row_order(vaqgpu_t *h, int64_t srow0, int64_t n, uint32_t *out) {
  if (!h || (!out && n > 0)) return fail(VAQGPU_EINVAL, "handle/out is NULL");
  if (srow0 < 0 || n < 0 || srow0 + n > h->n_rows) return fail(VAQGPU_EINVAL, "rows [%lld,%lld) outside [0,%lld)", (long long)srow0, (long long)(srow0 + n), (long long)h->n_rows);
  if (n == 0) return VAQGPU_OK;
  if (!h->d_rowid) {
    for (int64_t i = 0; i < n; i++) out[i] = (uint32_t)(srow0 + i);
    return VAQGPU_OK;
  }
  DeviceGuard g(h->device);
  int rc = ensure_rowid(h, h->stream);
  if (rc) return rc;
  CU(cudaMemcpyAsync(out, h->d_rowid + srow0, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return VAQGPU_OK;
}

int vaqgpu_build_lut(vaqgpu_t *h, const float *q_proj, int32_t nq, float *lut_out) {
  if (!h || !q_proj || !lut_out) return fail(VAQGPU_EINVAL, "handle/q_proj/lut_out is NULL");
  if (nq <= 0) return nq == 0 ? VAQGPU_OK : fail(VAQGPU_EINVAL, "nq=%d", nq);
  DeviceGuard g(h->device);
  // compact plan: table s at ent_off[s]
  LutPlan p = h->plan;
  p.T = 1;
  for (int s = 0; s < h->M; s++) p.pos[s] = p.ent_off[s];
  p.row_stride = p.total_entries;
  CU(h->w_q.ensure((size_t)nq * h->D * sizeof(float)));
  CU(h->w_lut.ensure((size_t)nq * p.row_stride * sizeof(float)));
  CU(cudaMemcpyAsync(h->w_q.p, q_proj, (size_t)nq * h->D * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  CU(launch_lut_build((const float *)h->w_q.p, nq, nq, h->D, h->d_centroids, h->d_cent_rmax, p, (float *)h->w_lut.p, nullptr, nullptr, h->stream));
  CU(cudaMemcpyAsync(lut_out, h->w_lut.p, (size_t)nq * p.row_stride * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return VAQGPU_OK;
}

int vaqgpu_search_device(vaqgpu_t *h, const float *d_queries, int32_t nq, int32_t k, uint32_t flags, int32_t *d_labels,
                         float *d_dists, void *stream) {
  if (!h || (nq > 0 && (!d_queries || !d_labels || !d_dists))) return fail(VAQGPU_EINVAL, "NULL argument");
  DeviceGuard g(h->device);
  return search_device_impl(h, d_queries, nq, k, flags, d_labels, d_dists, nullptr, (cudaStream_t)stream, true);
}

int vaqgpu_search_keys_device(vaqgpu_t *h, const float *d_queries, int32_t nq, int32_t k, uint32_t flags, uint64_t *d_keys,
                              void *stream) {
  if (!h || (nq > 0 && (!d_queries || !d_keys))) return fail(VAQGPU_EINVAL, "NULL argument");
  DeviceGuard g(h->device);
  return search_device_impl(h, d_queries, nq, k, flags & ~VAQGPU_SQRT, nullptr, nullptr, d_keys, (cudaStream_t)stream, true);
}

int vaqgpu_merge_keys_device(const uint64_t *d_keys_in, int32_t G, int32_t nq, int32_t k, uint32_t flags, int32_t *d_labels,
                             float *d_dists, void *stream) {
  if (!d_keys_in || !d_labels || !d_dists) return fail(VAQGPU_EINVAL, "NULL argument");
  if (G < 1 || G > 16 || nq < 0 || k <= 0) return fail(VAQGPU_EINVAL, "G=%d (1..16) nq=%d k=%d", G, nq, k);
  CU(launch_merge_keys(d_keys_in, (int64_t)nq * k, k, G, nq, k, (flags & VAQGPU_SQRT) ? 1 : 0, 0, d_labels, d_dists, nullptr,
                       nullptr, 0, nullptr, (cudaStream_t)stream));
  return VAQGPU_OK;
}

int vaqgpu_search(vaqgpu_t *h, const float *queries, int32_t nq, int32_t k, uint32_t flags, int32_t *labels, float *dists) {
  if (!h || (nq > 0 && (!queries || !labels || !dists))) return fail(VAQGPU_EINVAL, "NULL argument");
  if (nq < 0 || k <= 0) return fail(VAQGPU_EINVAL, "nq=%d k=%d", nq, k);
  if (nq == 0) return VAQGPU_OK;
  DeviceGuard g(h->device);
  CU(h->w_q.ensure((size_t)nq * h->D * sizeof(float)));
  CU(h->w_labels.ensure((size_t)nq * k * sizeof(int32_t)));
  CU(h->w_dists.ensure((size_t)nq * k * sizeof(float)));
  CU(cudaMemcpyAsync(h->w_q.p, queries, (size_t)nq * h->D * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  int rc = search_device_impl(h, (const float *)h->w_q.p, nq, k, flags, (int32_t *)h->w_labels.p, (float *)h->w_dists.p, nullptr,
                              h->stream, true);
  if (rc) return rc;
  CU(cudaMemcpyAsync(labels, h->w_labels.p, (size_t)nq * k * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(dists, h->w_dists.p, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return VAQGPU_OK;
}

int vaqgpu_set_clusters(vaqgpu_t *h, const float *clusters, int32_t C, int32_t segdims, const int64_t *start,
                        const int64_t *size, const int32_t *id_map) {
  if (!h || !clusters || !start || !size) return fail(VAQGPU_EINVAL, "NULL argument");
  if (C < 1 || segdims < 1 || segdims > h->D) return fail(VAQGPU_EINVAL, "C=%d segdims=%d", C, segdims);
  for (int c = 0; c < C; c++)
    if (start[c] < 0 || size[c] < 0 || start[c] + size[c] > h->n_rows)
      return fail(VAQGPU_EINVAL, "cluster %d range [%lld,+%lld) outside the index (%lld rows)", c, (long long)start[c], (long long)size[c], (long long)h->n_rows);
  DeviceGuard g(h->device);
  clear_clusters(h);
  {
    int rc = restore_layout(h, h->stream);        // cluster ranges are row ranges of the arrival order
    if (rc) return rc;
  }
  auto up = [&](void **dst, const void *src, size_t bytes) -> cudaError_t {
    cudaError_t e = cudaMalloc(dst, bytes);
    return e != cudaSuccess ? e : cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
  };
  cudaError_t e = up((void **)&h->d_clusters, clusters, (size_t)C * segdims * sizeof(float));
  if (e == cudaSuccess) e = up((void **)&h->d_cl_start, start, (size_t)C * sizeof(int64_t));
  if (e == cudaSuccess) e = up((void **)&h->d_cl_size, size, (size_t)C * sizeof(int64_t));
  if (e == cudaSuccess && id_map) e = up((void **)&h->d_id_map, id_map, (size_t)h->n_rows * sizeof(int32_t));
  // filter-kernel TI needs the cluster of every row tile: possible when the ranges ascend without overlap (what
  // clusterTI produces: start[c] = start[c-1] + size[c-1]); anything else keeps the lane-per-row kernel
  bool ascending = C <= 0xFFFE && start[0] == 0 && start[C - 1] + size[C - 1] == h->n_rows;      // an exact partition
  for (int c = 1; c < C && ascending; c++) ascending = start[c] == start[c - 1] + size[c - 1];
  if (e == cudaSuccess && ascending) {
    const int64_t n_tiles = (h->n_rows + kTileRows - 1) / kTileRows;
    e = cudaMalloc(&h->d_tile_cl, (size_t)std::max<int64_t>(1, n_tiles) * sizeof(uint16_t));
    if (e == cudaSuccess) e = launch_tile_clusters(h->d_cl_start, C, h->n_rows, h->d_tile_cl, h->stream);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_clusters_t, (size_t)C * segdims * sizeof(float));
    if (e == cudaSuccess) e = launch_transpose(h->d_clusters, C, segdims, h->d_clusters_t, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  }
  if (e != cudaSuccess) {
    clear_clusters(h);
    return fail(e == cudaErrorMemoryAllocation ? VAQGPU_ENOMEM : VAQGPU_ECUDA, "vaqgpu_set_clusters: %s", cudaGetErrorString(e));
  }
  h->C = C; h->segdims = segdims;
  h->h_cl_start.assign(start, start + C);
  h->h_cl_size.assign(size, size + C);
  return VAQGPU_OK;
}

__global__ void compose_ids_kernel(const int32_t *__restrict__ older, int32_t *__restrict__ ids, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ids[i] = older[ids[i]];
}

int vaqgpu_cluster_ti(vaqgpu_t *h, int32_t C, int32_t n_segments, int32_t iters) {
  if (!h) return fail(VAQGPU_EINVAL, "handle is NULL");
  if (h->n_rows <= 0) return fail(VAQGPU_ESTATE, "vaqgpu_cluster_ti on an empty index");
  if (C < 1 || C > 0xFFFE || C > h->n_rows) return fail(VAQGPU_EINVAL, "C=%d (1..min(65534, rows))", C);
  if (iters < 0 || iters > 1000) return fail(VAQGPU_EINVAL, "iters=%d", iters);
  const int seg = (n_segments <= 0 || n_segments > h->M) ? h->M : n_segments;
  DeviceGuard g(h->device);
  cudaStream_t st = h->stream;
  int rc = restore_layout(h, st);          // ids of the regrouped rows are expressed in the arrival order
  if (rc) return rc;
  // a second clustering of an already regrouped index composes the id maps
  int32_t *old_map = h->d_id_map;
  h->d_id_map = nullptr;
  clear_clusters(h);
  const int64_t n = h->n_rows;
  const size_t tbl = cluster_ti_table_floats(h->plan, seg, C);
  const int64_t tiles = (n + kTileRows - 1) / kTileRows;
  float *centres = nullptr, *T = nullptr;
  int32_t *hist = nullptr, *assign = nullptr, *sizes = nullptr, *bh = nullptr, *id_map = nullptr;
  int64_t *start = nullptr, *size64 = nullptr;
  uint4 *ncodes = nullptr;
  uint16_t *tile_cl = nullptr;
  auto cleanup = [&]() {
    cudaFree(centres); cudaFree(T); cudaFree(hist); cudaFree(assign); cudaFree(sizes); cudaFree(bh); cudaFree(id_map);
    cudaFree(start); cudaFree(size64); cudaFree(ncodes); cudaFree(tile_cl); cudaFree(old_map);
  };
#define CT(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      cleanup();                                                                                         \
      return fail(e_ == cudaErrorMemoryAllocation ? VAQGPU_ENOMEM : VAQGPU_ECUDA, "vaqgpu_cluster_ti: %s: %s", #call, cudaGetErrorString(e_)); \
    }                                                                                                    \
  } while (0)
  CT(cudaMalloc(&centres, (size_t)C * seg * h->L * sizeof(float)));
  CT(cudaMalloc(&T, tbl * sizeof(float)));
  CT(cudaMalloc(&hist, tbl * sizeof(int32_t)));
  CT(cudaMalloc(&assign, (size_t)n * sizeof(int32_t)));
  CT(cudaMalloc(&sizes, (size_t)C * sizeof(int32_t)));
  CT(cudaMalloc(&bh, regroup_hist_ints(n, C) * sizeof(int32_t)));
  CT(cudaMalloc(&id_map, (size_t)n * sizeof(int32_t)));
  CT(cudaMalloc(&start, (size_t)C * sizeof(int64_t)));
  CT(cudaMalloc(&size64, (size_t)C * sizeof(int64_t)));
  CT(cudaMalloc(&ncodes, (size_t)tiles * kTileRows * h->lay.W * sizeof(uint4)));
  CT(cudaMalloc(&tile_cl, (size_t)tiles * sizeof(uint16_t)));
  CT(cudaMemsetAsync(ncodes, 0, (size_t)tiles * kTileRows * h->lay.W * sizeof(uint4), st));
  CT(launch_cluster_ti_kmeans(h->d_codes, h->lay, n, h->plan, seg, h->d_centroids, h->d_cent_off, h->d_ent_off, C, iters, centres, T, hist,
                              assign, sizes, st));
  CT(launch_regroup(assign, sizes, n, C, h->d_codes, ncodes, h->lay.W, id_map, start, size64, bh, st));
  if (old_map) {
    compose_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(old_map, id_map, n);
    CT(cudaGetLastError());
  }
  CT(launch_tile_clusters(start, C, n, tile_cl, st));
  float *centres_t = nullptr;
  CT(cudaMalloc(&centres_t, (size_t)C * seg * h->L * sizeof(float)));
  h->d_clusters_t = centres_t;          // owned by the handle from here on (clear_clusters frees it)
  CT(launch_transpose(centres, C, seg * h->L, centres_t, st));
  CT(cudaStreamSynchronize(st));
#undef CT
  cudaFree(h->d_codes);
  h->d_codes = ncodes; ncodes = nullptr;
  h->cap_rows = tiles * kTileRows;
  h->d_clusters = centres; centres = nullptr;
  h->d_cl_start = start; start = nullptr;
  h->d_cl_size = size64; size64 = nullptr;
  h->d_id_map = id_map; id_map = nullptr;
  h->d_tile_cl = tile_cl; tile_cl = nullptr;
  h->C = C; h->segdims = seg * h->L;
  h->opt_n = 0;
  h->h_cl_start.resize(C); h->h_cl_size.resize(C);
  cudaMemcpy(h->h_cl_start.data(), h->d_cl_start, (size_t)C * sizeof(int64_t), cudaMemcpyDeviceToHost);
  cudaMemcpy(h->h_cl_size.data(), h->d_cl_size, (size_t)C * sizeof(int64_t), cudaMemcpyDeviceToHost);
  cleanup();
  return VAQGPU_OK;
}

int vaqgpu_get_clusters(vaqgpu_t *h, int32_t *C, int32_t *segdims, float *clusters, int64_t *start, int64_t *size, int32_t *id_map) {
  if (!h) return fail(VAQGPU_EINVAL, "handle is NULL");
  if (C) *C = h->C;
  if (segdims) *segdims = h->segdims;
  if (!h->C) return (clusters || start || size || id_map) ? fail(VAQGPU_ESTATE, "no clusters set") : VAQGPU_OK;
  DeviceGuard g(h->device);
  if (clusters) CU(cudaMemcpy(clusters, h->d_clusters, (size_t)h->C * h->segdims * sizeof(float), cudaMemcpyDeviceToHost));
  if (start) CU(cudaMemcpy(start, h->d_cl_start, (size_t)h->C * sizeof(int64_t), cudaMemcpyDeviceToHost));
  if (size) CU(cudaMemcpy(size, h->d_cl_size, (size_t)h->C * sizeof(int64_t), cudaMemcpyDeviceToHost));
  if (id_map) {
    if (!h->d_id_map) return fail(VAQGPU_ESTATE, "clusters were set without an id map");
    CU(cudaMemcpy(id_map, h->d_id_map, (size_t)h->n_rows * sizeof(int32_t), cudaMemcpyDeviceToHost));
  }
  return VAQGPU_OK;
}

int vaqgpu_set_cluster_rule_sizes(vaqgpu_t *h, const int64_t *sizes) {
  if (!h || !sizes) return fail(VAQGPU_EINVAL, "NULL argument");
  if (!h->C) return fail(VAQGPU_ESTATE, "vaqgpu_set_clusters first");
  DeviceGuard g(h->device);
  if (!h->d_cl_rule) CU(cudaMalloc(&h->d_cl_rule, (size_t)h->C * sizeof(int64_t)));
  CU(cudaMemcpy(h->d_cl_rule, sizes, (size_t)h->C * sizeof(int64_t), cudaMemcpyHostToDevice));
  return VAQGPU_OK;
}

int vaqgpu_set_visit(vaqgpu_t *h, float visit) {
  if (!h) return fail(VAQGPU_EINVAL, "handle is NULL");
  if (!(visit > 0.f)) return fail(VAQGPU_EINVAL, "visit=%f must be > 0", visit);
  h->visit = visit;
  return VAQGPU_OK;
}

int vaqgpu_set_raw_vectors(vaqgpu_t *h, const float *xtrain, int64_t n, int32_t D0) {
  if (!h || !xtrain) return fail(VAQGPU_EINVAL, "NULL argument");
  if (n <= 0 || D0 <= 0) return fail(VAQGPU_EINVAL, "n=%lld D0=%d", (long long)n, D0);
  DeviceGuard g(h->device);
  cudaFree(h->d_raw); h->d_raw = nullptr;
  CU(cudaMalloc(&h->d_raw, (size_t)n * D0 * sizeof(float)));
  CU(cudaMemcpy(h->d_raw, xtrain, (size_t)n * D0 * sizeof(float), cudaMemcpyHostToDevice));
  h->raw_n = n; h->raw_D = D0;
  return VAQGPU_OK;
}

int vaqgpu_refine(vaqgpu_t *h, const float *queries, int32_t nq, const int32_t *in_labels, int32_t refine_num, int32_t k,
                  int32_t *labels, float *dists) {
  if (!h || !queries || !in_labels || !labels || !dists) return fail(VAQGPU_EINVAL, "NULL argument");
  if (!h->d_raw) return fail(VAQGPU_ESTATE, "vaqgpu_refine needs vaqgpu_set_raw_vectors");
  if (nq <= 0 || refine_num <= 0 || k <= 0 || k > refine_num) return fail(VAQGPU_EINVAL, "nq=%d refine_num=%d k=%d", nq, refine_num, k);
  if ((size_t)refine_num * 8 > kSmemCap) return fail(VAQGPU_EINVAL, "refine_num=%d too large", refine_num);
  DeviceGuard g(h->device);
  const int D0 = h->raw_D;
  CU(h->w_q.ensure((size_t)nq * D0 * sizeof(float)));
  CU(h->w_stage.ensure((size_t)nq * refine_num * sizeof(int32_t)));
  CU(h->w_labels.ensure((size_t)nq * k * sizeof(int32_t)));
  CU(h->w_dists.ensure((size_t)nq * k * sizeof(float)));
  CU(cudaMemcpyAsync(h->w_q.p, queries, (size_t)nq * D0 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->w_stage.p, in_labels, (size_t)nq * refine_num * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  CU(launch_refine(h->d_raw, h->raw_n, D0, (const float *)h->w_q.p, nq, (const int32_t *)h->w_stage.p, refine_num, k,
                   (int32_t *)h->w_labels.p, (float *)h->w_dists.p, h->stream));
  CU(cudaMemcpyAsync(labels, h->w_labels.p, (size_t)nq * k * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(dists, h->w_dists.p, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return VAQGPU_OK;
}

/* ---- cross-shard bound exchange -------------------------------------------------------------- */

int vaqgpu_bounds_export(vaqgpu_t *h, int32_t max_queries, unsigned char ipc_handle[64], void **d_ptr) {
  if (!h) return fail(VAQGPU_EINVAL, "handle is NULL");
  if (max_queries < 1) return fail(VAQGPU_EINVAL, "max_queries=%d", max_queries);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  DeviceGuard g(h->device);
  for (void *p : h->ipc_opened) cudaIpcCloseMemHandle(p);
  h->ipc_opened.clear();
  h->peers.n = 0;
  cudaFree(h->d_bounds);
  h->d_bounds = nullptr; h->bounds_cap = 0; h->bounds_parity = 0;
  const int32_t cap = (max_queries + 63) & ~63;
  CU(cudaMalloc(&h->d_bounds, (size_t)2 * cap * sizeof(uint32_t)));
  CU(cudaMemset(h->d_bounds, 0xFF, (size_t)2 * cap * sizeof(uint32_t)));
  CU(cudaDeviceSynchronize());
  h->bounds_cap = cap;
  if (ipc_handle) {
    cudaIpcMemHandle_t ih;
    CU(cudaIpcGetMemHandle(&ih, h->d_bounds));
    memcpy(ipc_handle, &ih, 64);
  }
  if (d_ptr) *d_ptr = h->d_bounds;
  return VAQGPU_OK;
}

int vaqgpu_bounds_attach_ipc(vaqgpu_t *h, int32_t n_peers, const unsigned char *ipc_handles) {
  if (!h || (n_peers > 0 && !ipc_handles)) return fail(VAQGPU_EINVAL, "NULL argument");
  if (!h->d_bounds) return fail(VAQGPU_ESTATE, "vaqgpu_bounds_export first");
  if (n_peers < 0 || n_peers > kMaxPeers) return fail(VAQGPU_EINVAL, "n_peers=%d (0..%d)", n_peers, kMaxPeers);
  DeviceGuard g(h->device);
  for (void *p : h->ipc_opened) cudaIpcCloseMemHandle(p);
  h->ipc_opened.clear();
  h->peers.n = 0;
  for (int i = 0; i < n_peers; i++) {
    cudaIpcMemHandle_t ih;
    memcpy(&ih, ipc_handles + (size_t)i * 64, 64);
    void *p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
    h->ipc_opened.push_back(p);
    h->peers.p[h->peers.n++] = (uint32_t *)p;
  }
  return VAQGPU_OK;
}

int vaqgpu_bounds_attach_ptr(vaqgpu_t *h, int32_t n_peers, void *const *peer_ptrs) {
  if (!h || (n_peers > 0 && !peer_ptrs)) return fail(VAQGPU_EINVAL, "NULL argument");
  if (!h->d_bounds) return fail(VAQGPU_ESTATE, "vaqgpu_bounds_export first");
  if (n_peers < 0 || n_peers > kMaxPeers) return fail(VAQGPU_EINVAL, "n_peers=%d (0..%d)", n_peers, kMaxPeers);
  h->peers.n = 0;
  for (int i = 0; i < n_peers; i++) h->peers.p[h->peers.n++] = (uint32_t *)peer_ptrs[i];
  return VAQGPU_OK;
}

int vaqgpu_last_timings(const vaqgpu_t *h, float ms[4]) {
  if (!h || !ms) return fail(VAQGPU_EINVAL, "NULL argument");
  if (!h->timed) return fail(VAQGPU_ESTATE, "no search has run on this handle");
  DeviceGuard g(h->device);
  CU(cudaEventSynchronize(h->ev[4]));
  for (int i = 0; i < 4; i++) CU(cudaEventElapsedTime(&ms[i], h->ev[i], h->ev[i + 1]));
  return VAQGPU_OK;
}

int vaqgpu_last_config(const vaqgpu_t *h, int32_t cfg[12]) {
  if (!h || !cfg) return fail(VAQGPU_EINVAL, "NULL argument");
  memcpy(cfg, h->cfg, sizeof(h->cfg));
  return VAQGPU_OK;
}

/* ---------------------------------------------------------------- Hamming ---- */

int hamgpu_create(int32_t nbits, int device, hamgpu_t **out) {
  if (!out) return fail(VAQGPU_EINVAL, "out is NULL");
  *out = nullptr;
  if (nbits < 1 || nbits > 1024) return fail(VAQGPU_EINVAL, "nbits=%d out of range [1,1024]", nbits);
  const int W = (nbits + 127) / 128;
  if (W == 5 || W == 7) return fail(VAQGPU_EINVAL, "nbits=%d (%d x 128-bit words) unsupported: use 1,2,3,4,6 or 8 words", nbits, W);
  int rc = check_device(device);
  if (rc) return rc;
  DeviceGuard g(device);
  hamgpu_index *h = new (std::nothrow) hamgpu_index();
  if (!h) return fail(VAQGPU_ENOMEM, "host allocation failed");
  h->device = device; h->nbits = nbits; h->w64 = (nbits + 63) / 64; h->W = W;
  cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
  rc = ensure_stream(&h->stream, h->ev, 3);
  if (rc) { hamgpu_destroy(h); return rc; }
  *out = h;
  return VAQGPU_OK;
}

void hamgpu_destroy(hamgpu_t *h) {
  if (!h) return;
  DeviceGuard g(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->d_codes);
  for (DevBuf *b : {&h->w_q, &h->w_qpad, &h->w_keys, &h->w_scratch, &h->w_stage, &h->w_idx, &h->w_dist}) b->release();
  for (auto &e : h->ev) if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int hamgpu_set_id_base(hamgpu_t *h, int64_t id_base) {
  if (!h) return fail(VAQGPU_EINVAL, "handle is NULL");
  h->id_base = id_base;
  return VAQGPU_OK;
}

int hamgpu_add(hamgpu_t *h, const uint64_t *words, int64_t n) {
  if (!h || (!words && n > 0)) return fail(VAQGPU_EINVAL, "handle/words is NULL");
  if (n < 0) return fail(VAQGPU_EINVAL, "n=%lld", (long long)n);
  if (n == 0) return VAQGPU_OK;
  if (h->n_rows + n > 0x7FFFFFFFll) return fail(VAQGPU_EINVAL, "more than 2^31-1 rows per index");
  DeviceGuard g(h->device);
  CU(grow_codes(&h->d_codes, &h->cap_rows, h->n_rows, h->W, h->n_rows + n, h->stream));
  const int64_t chunk = std::min<int64_t>(n, kStageRows);
  CU(h->w_stage.ensure((size_t)chunk * h->w64 * sizeof(uint64_t)));
  for (int64_t r = 0; r < n; r += chunk) {
    const int64_t c = std::min(chunk, n - r);
    CU(cudaMemcpyAsync(h->w_stage.p, words + (size_t)r * h->w64, (size_t)c * h->w64 * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
    CU(launch_ham_pack((const uint64_t *)h->w_stage.p, c, h->n_rows + r, h->w64, h->W, h->d_codes, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  h->n_rows += n;
  return VAQGPU_OK;
}

int hamgpu_add_synthetic(hamgpu_t *h, int64_t n, uint64_t seed) {
  if (!h) return fail(VAQGPU_EINVAL, "handle is NULL");
  if (n < 0) return fail(VAQGPU_EINVAL, "n=%lld", (long long)n);
  if (n == 0) return VAQGPU_OK;
  if (h->n_rows + n > 0x7FFFFFFFll) return fail(VAQGPU_EINVAL, "more than 2^31-1 rows per index");
  DeviceGuard g(h->device);
  CU(grow_codes(&h->d_codes, &h->cap_rows, h->n_rows, h->W, h->n_rows + n, h->stream));
  CU(launch_ham_synth(h->d_codes, n, h->n_rows, h->id_base + h->n_rows, h->nbits, h->W, seed, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->n_rows += n;
  return VAQGPU_OK;
}

int hamgpu_num_rows(const hamgpu_t *h, int64_t *n) {
  if (!h || !n) return fail(VAQGPU_EINVAL, "handle/n is NULL");
  *n = h->n_rows;
  return VAQGPU_OK;
}

}  // extern "C"

namespace {

int ham_query_impl(hamgpu_index *h, const uint64_t *d_queries, int nq, int k, int32_t *d_idx, uint32_t *d_dist,
                   uint64_t *d_keys, cudaStream_t st) {
  if (nq < 0 || k <= 0) return fail(VAQGPU_EINVAL, "nq=%d k=%d", nq, k);
  if (k > 2048) return fail(VAQGPU_EINVAL, "k=%d > 2048 unsupported", k);
  if (nq == 0) return VAQGPU_OK;
  const int W = h->W;
  const uint4 *dq;
  int launches = 0;
  if (h->w64 == 2 * W) {
    dq = reinterpret_cast<const uint4 *>(d_queries);
    if ((reinterpret_cast<uintptr_t>(d_queries) & 15) != 0) return fail(VAQGPU_EINVAL, "device queries must be 16-byte aligned");
  } else {
    CU(h->w_qpad.ensure((size_t)nq * W * sizeof(uint4)));
    const int total = nq * W;
    ham_pad_queries_kernel<<<(total + 255) / 256, 256, 0, st>>>(d_queries, nq, h->w64, W, (uint4 *)h->w_qpad.p);
    CU(cudaGetLastError());
    dq = (const uint4 *)h->w_qpad.p;
    launches++;
  }
  const int threads = kScanThreads, nwarps = threads / 32;
  // queries per CTA: as many as the per-(warp,query) lists allow in ~64 KB
  int qt = 8;
  while (qt > 1 && ((size_t)nwarps * qt * k * 8 > 64 * 1024 || qt > nq)) qt >>= 1;
  const size_t smem = (size_t)qt * W * 16 + ((size_t)nwarps * qt * k + k + qt) * sizeof(uint64_t);
  if (smem > kSmemCap) return fail(VAQGPU_EINVAL, "k=%d needs %zu B shared memory", k, smem);
  const int64_t n_tiles = (h->n_rows + kTileRows - 1) / kTileRows;
  const int qb_max = std::min(nq, 32768);                       // grid.y = query groups: bounded per launch
  const int qgroups = (qb_max + qt - 1) / qt;
  // two waves of resident CTAs (ham_scan_kernel's launch bounds: 6 per SM for one query per pass on rows of <= 256 bits — more
  // warps, hence more tiles in flight, where the pass is bound by load latency: 0.36 -> 0.335 ms on 64 M rows; with two
  // queries both 6 (40 registers, spills) and 5 (47 registers) measured slower: 0.57 / 0.49 vs 0.46 ms —, 4 otherwise, 2 for wider rows)
  const int occ = W <= 2 ? (qt == 1 ? 6 : 4) : 2;
  const int64_t target = (int64_t)h->num_sms * occ * 2;
  int splits = (int)std::max<int64_t>(1, (target + qgroups - 1) / qgroups);
  splits = (int)std::min<int64_t>(splits, std::max<int64_t>(1, n_tiles / (nwarps * 8)));
  CU(h->w_keys.ensure((size_t)qb_max * splits * k * sizeof(uint64_t)));
  if (splits > 16) CU(h->w_scratch.ensure((size_t)2 * qb_max * ((splits + 15) / 16) * k * sizeof(uint64_t)));
  CU(cudaEventRecord(h->ev[0], st));
  for (int q0 = 0; q0 < nq; q0 += qb_max) {
    const int qb = std::min(qb_max, nq - q0);
    HamScanArgs a{};
    a.codes = h->d_codes; a.n_rows = h->n_rows; a.W = W; a.queries = dq + (size_t)q0 * W;
    a.nq = qb; a.k = k; a.splits = splits; a.qt = qt;
    a.out_keys = (uint64_t *)h->w_keys.p;
    CU(launch_ham_scan(a, threads, smem, st));
    launches++;
    if (q0 + qb >= nq) CU(cudaEventRecord(h->ev[1], st));
    CU(launch_merge_keys((const uint64_t *)h->w_keys.p, k, (int64_t)splits * k, splits, qb, k, 0, 1, d_idx ? d_idx + (size_t)q0 * k : nullptr,
                         d_dist ? d_dist + (size_t)q0 * k : nullptr, d_keys ? d_keys + (size_t)q0 * k : nullptr, nullptr,
                         h->id_base, (uint64_t *)h->w_scratch.p, st));
    launches += splits > 16 ? 2 : 1;
  }
  CU(cudaEventRecord(h->ev[2], st));
  h->timed = true;
  h->cfg[0] = threads; h->cfg[1] = splits; h->cfg[2] = qt; h->cfg[3] = 0; h->cfg[4] = (int32_t)smem; h->cfg[5] = W;
  h->cfg[6] = launches; h->cfg[7] = nq;
  return VAQGPU_OK;
}

}  // namespace

extern "C" {

int hamgpu_query_device(hamgpu_t *h, const uint64_t *d_queries, int32_t nq, int32_t k, int32_t *d_idx, uint32_t *d_dist,
                        void *stream) {
  if (!h || (nq > 0 && (!d_queries || !d_idx || !d_dist))) return fail(VAQGPU_EINVAL, "NULL argument");
  DeviceGuard g(h->device);
  return ham_query_impl(h, d_queries, nq, k, d_idx, d_dist, nullptr, (cudaStream_t)stream);
}

int hamgpu_query_keys_device(hamgpu_t *h, const uint64_t *d_queries, int32_t nq, int32_t k, uint64_t *d_keys, void *stream) {
  if (!h || (nq > 0 && (!d_queries || !d_keys))) return fail(VAQGPU_EINVAL, "NULL argument");
  DeviceGuard g(h->device);
  return ham_query_impl(h, d_queries, nq, k, nullptr, nullptr, d_keys, (cudaStream_t)stream);
}

int hamgpu_merge_keys_device(const uint64_t *d_keys_in, int32_t G, int32_t nq, int32_t k, int32_t *d_idx, uint32_t *d_dist,
                             void *stream) {
  if (!d_keys_in || !d_idx || !d_dist) return fail(VAQGPU_EINVAL, "NULL argument");
  if (G < 1 || G > 16 || nq < 0 || k <= 0) return fail(VAQGPU_EINVAL, "G=%d (1..16) nq=%d k=%d", G, nq, k);
  CU(launch_merge_keys(d_keys_in, (int64_t)nq * k, k, G, nq, k, 0, 1, d_idx, d_dist, nullptr, nullptr, 0, nullptr,
                       (cudaStream_t)stream));
  return VAQGPU_OK;
}

int hamgpu_query(hamgpu_t *h, const uint64_t *queries, int32_t nq, int32_t k, int32_t *idx, uint32_t *dist) {
  if (!h || (nq > 0 && (!queries || !idx || !dist))) return fail(VAQGPU_EINVAL, "NULL argument");
  if (nq < 0 || k <= 0) return fail(VAQGPU_EINVAL, "nq=%d k=%d", nq, k);
  if (nq == 0) return VAQGPU_OK;
  DeviceGuard g(h->device);
  CU(h->w_q.ensure((size_t)nq * h->w64 * sizeof(uint64_t)));
  CU(h->w_idx.ensure((size_t)nq * k * sizeof(int32_t)));
  CU(h->w_dist.ensure((size_t)nq * k * sizeof(uint32_t)));
  CU(cudaMemcpyAsync(h->w_q.p, queries, (size_t)nq * h->w64 * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
  int rc = ham_query_impl(h, (const uint64_t *)h->w_q.p, nq, k, (int32_t *)h->w_idx.p, (uint32_t *)h->w_dist.p, nullptr, h->stream);
  if (rc) return rc;
  CU(cudaMemcpyAsync(idx, h->w_idx.p, (size_t)nq * k * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(dist, h->w_dist.p, (size_t)nq * k * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return VAQGPU_OK;
}

int hamgpu_last_timings(const hamgpu_t *h, float ms[2]) {
  if (!h || !ms) return fail(VAQGPU_EINVAL, "NULL argument");
  if (!h->timed) return fail(VAQGPU_ESTATE, "no query has run on this handle");
  DeviceGuard g(h->device);
  CU(cudaEventSynchronize(h->ev[2]));
  for (int i = 0; i < 2; i++) CU(cudaEventElapsedTime(&ms[i], h->ev[i], h->ev[i + 1]));
  return VAQGPU_OK;
}

int hamgpu_last_config(const hamgpu_t *h, int32_t cfg[8]) {
  if (!h || !cfg) return fail(VAQGPU_EINVAL, "NULL argument");
  memcpy(cfg, h->cfg, sizeof(h->cfg));
  return VAQGPU_OK;
}

}  // extern "C"
