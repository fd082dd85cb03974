// Row-sharded indexes over the GPUs of one box, driven from ONE host process: contiguous row blocks
// per GPU, the model and the query batch replicated, every GPU emits its shard-local top-k as sortable 64-bit keys,
// one ncclAllGather of the key lists (ncclCommInitAll communicator, one stream per device, grouped launch) and a
// device-side merge.  The reference's precedent for merging partial answers is concatenate + sort + resize(k),
// BitVecEngine.cpp:1599-1611; the merged answer is the k smallest (distance, id) keys overall, hence independent of
// the number of shards.  While the shards scan they exchange their running k-th-best bounds through NVLink peer
// memory (vaqgpu_bounds_*), so each GPU prunes as if it saw all rows.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): a process that already loaded a copy (e.g. through PyTorch)
// shares it, and single-GPU users of libvaqgpu.so need no NCCL at all.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "../../include/vaqgpu.h"

namespace vaqgpu {
int set_error(int code, const char *fmt, ...);      // vaqgpu_host.cu: formats into the thread-local message
}

namespace {

using vaqgpu::set_error;

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (lib) return true;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) return false;
#define SYM(field, sym) field = reinterpret_cast<decltype(field)>(dlsym(lib, sym))
    SYM(CommInitAll, "ncclCommInitAll");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllGather, "ncclAllGather");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return CommInitAll && CommDestroy && AllGather && GroupStart && GroupEnd && GetErrorString;
  }
};
NcclApi g_nccl;

#define CU(call)                                                                                              \
  do {                                                                                                        \
    cudaError_t e_ = (call);                                                                                  \
    if (e_ != cudaSuccess)                                                                                    \
      return set_error(e_ == cudaErrorMemoryAllocation ? VAQGPU_ENOMEM : VAQGPU_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, \
                       cudaGetErrorString(e_));                                                               \
  } while (0)
#define NC(call)                                                                                              \
  do {                                                                                                        \
    ncclResult_t r_ = (call);                                                                                 \
    if (r_ != ncclSuccess) return set_error(VAQGPU_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
  } while (0)
#define RC(call)                   \
  do {                             \
    int rc_ = (call);              \
    if (rc_ != VAQGPU_OK) return rc_; \
  } while (0)

// the multi-device entry points switch devices; put the caller's device back on exit
struct RestoreDevice {
  int prev = -1;
  RestoreDevice() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
  ~RestoreDevice() { if (prev >= 0) cudaSetDevice(prev); }
};

struct DevBufS {
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
};

// state shared by the VAQ and Hamming sharded handles: devices, streams, communicator, per-device buffers
struct ShardSet {
  int G = 0;
  std::vector<int> dev;
  std::vector<cudaStream_t> st;
  std::vector<ncclComm_t> comm;
  std::vector<DevBufS> q, keys, all;     // per device: queries, local key lists, gathered key lists
  DevBufS out_a, out_b;                  // device 0: labels / distances
  int64_t n_total = 0, per = 0, n_added = 0;

  int init(int n_gpus, const int *dev_ids, int64_t n_rows_total) {
    if (n_gpus < 1 || n_gpus > 16) return set_error(VAQGPU_EINVAL, "n_gpus=%d (1..16)", n_gpus);
    if (n_rows_total < 0) return set_error(VAQGPU_EINVAL, "n_rows_total=%lld", (long long)n_rows_total);
    G = n_gpus;
    dev.resize(G);
    for (int i = 0; i < G; i++) dev[i] = dev_ids ? dev_ids[i] : i;
    for (int i = 0; i < G; i++)
      for (int j = 0; j < i; j++)
        if (dev[i] == dev[j]) return set_error(VAQGPU_EINVAL, "device %d listed twice", dev[i]);
    n_total = n_rows_total;
    per = (n_total + G - 1) / G;
    st.assign(G, nullptr); comm.assign(G, nullptr);
    q.resize(G); keys.resize(G); all.resize(G);
    for (int i = 0; i < G; i++) {
      CU(cudaSetDevice(dev[i]));
      CU(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
    }
    if (G > 1) {
      if (!g_nccl.load()) return set_error(VAQGPU_ESTATE, "libnccl.so.2 not found (%s): a multi-GPU handle needs NCCL", dlerror());
      NC(g_nccl.CommInitAll(comm.data(), G, dev.data()));
      for (int i = 0; i < G; i++) {          // peer access for the bound exchange (NVLink / NVSwitch)
        CU(cudaSetDevice(dev[i]));
        for (int j = 0; j < G; j++) {
          if (i == j) continue;
          cudaError_t e = cudaDeviceEnablePeerAccess(dev[j], 0);
          if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
          else if (e != cudaSuccess) return set_error(VAQGPU_ECUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", dev[i], dev[j], cudaGetErrorString(e));
        }
      }
    }
    return VAQGPU_OK;
  }
  void destroy() {
    for (int i = 0; i < G; i++) {
      cudaSetDevice(dev[i]);
      if (st[i]) cudaStreamSynchronize(st[i]);
      if (comm[i]) g_nccl.CommDestroy(comm[i]);
      cudaFree(q[i].p); cudaFree(keys[i].p); cudaFree(all[i].p);
      if (st[i]) cudaStreamDestroy(st[i]);
    }
    if (G) { cudaSetDevice(dev[0]); cudaFree(out_a.p); cudaFree(out_b.p); }
    G = 0;
  }
  int64_t lo(int r) const { return std::min<int64_t>((int64_t)r * per, n_total); }
  // all-gather of [nq x k] key lists, every device receives [G][nq][k]
  int allgather_keys(size_t count) {
    if (G == 1) {
      CU(cudaSetDevice(dev[0]));
      CU(cudaMemcpyAsync(all[0].p, keys[0].p, count * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st[0]));
      return VAQGPU_OK;
    }
    NC(g_nccl.GroupStart());
    for (int i = 0; i < G; i++) NC(g_nccl.AllGather(keys[i].p, all[i].p, count, ncclUint64, comm[i], st[i]));
    NC(g_nccl.GroupEnd());
    return VAQGPU_OK;
  }
};

}  // namespace

struct vaqgpu_sharded {
  ShardSet s;
  std::vector<vaqgpu_t *> shard;
  int32_t D = 0, M = 0;
  int32_t bounds_cap = 0;
};

struct hamgpu_sharded {
  ShardSet s;
  std::vector<hamgpu_t *> shard;
  int32_t w64 = 0;
};

namespace {

int exchange_bounds(vaqgpu_sharded *h, int32_t nq) {
  if (h->s.G == 1 || nq <= h->bounds_cap) return VAQGPU_OK;
  const int G = h->s.G;
  for (int i = 0; i < G; i++) { CU(cudaSetDevice(h->s.dev[i])); CU(cudaDeviceSynchronize()); }
  const int32_t cap = std::max(nq, 16384);
  std::vector<void *> ptr(G, nullptr);
  for (int i = 0; i < G; i++) RC(vaqgpu_bounds_export(h->shard[i], cap, nullptr, &ptr[i]));
  for (int i = 0; i < G; i++) {
    std::vector<void *> peers;
    for (int j = 0; j < G; j++)
      if (j != i) peers.push_back(ptr[j]);
    RC(vaqgpu_bounds_attach_ptr(h->shard[i], (int32_t)peers.size(), peers.data()));
  }
  h->bounds_cap = cap;
  return VAQGPU_OK;
}

// split global rows [g0, g0+n) at the shard boundaries; f(shard, first row of the piece relative to g0, rows)
template <typename F>
int for_each_piece(const ShardSet &s, int64_t g0, int64_t n, F f) {
  if (g0 + n > s.n_total) return set_error(VAQGPU_EINVAL, "rows [%lld,%lld) exceed n_rows_total=%lld", (long long)g0, (long long)(g0 + n), (long long)s.n_total);
  for (int r = 0; r < s.G; r++) {
    const int64_t a = std::max(g0, s.lo(r)), b = std::min(g0 + n, s.lo(r + 1));
    if (b > a) RC(f(r, a - g0, b - a));
  }
  return VAQGPU_OK;
}

}  // namespace

extern "C" {

int vaqgpu_sharded_create(const vaqgpu_model_desc *model, int32_t n_gpus, const int *dev_ids, int64_t n_rows_total,
                          vaqgpu_sharded_t **out) {
  RestoreDevice restore;
  if (!model || !out) return set_error(VAQGPU_EINVAL, "model/out is NULL");
  *out = nullptr;
  vaqgpu_sharded *h = new (std::nothrow) vaqgpu_sharded();
  if (!h) return set_error(VAQGPU_ENOMEM, "host allocation failed");
  int rc = h->s.init(n_gpus, dev_ids, n_rows_total);
  for (int i = 0; rc == VAQGPU_OK && i < h->s.G; i++) {
    vaqgpu_t *x = nullptr;
    rc = vaqgpu_create(model, h->s.dev[i], &x);
    if (rc == VAQGPU_OK) {
      h->shard.push_back(x);
      rc = vaqgpu_set_id_base(x, h->s.lo(i));
      if (rc == VAQGPU_OK && h->s.lo(i + 1) > h->s.lo(i)) rc = vaqgpu_reserve(x, h->s.lo(i + 1) - h->s.lo(i));
    }
  }
  if (rc != VAQGPU_OK) { vaqgpu_sharded_destroy(h); return rc; }
  h->D = model->D; h->M = model->M;
  *out = h;
  return VAQGPU_OK;
}

void vaqgpu_sharded_destroy(vaqgpu_sharded_t *h) {
  RestoreDevice restore;
  if (!h) return;
  for (vaqgpu_t *x : h->shard) vaqgpu_destroy(x);
  h->s.destroy();
  delete h;
}

int vaqgpu_sharded_add_codes_u16(vaqgpu_sharded_t *h, const uint16_t *codes, int64_t n) {
  if (!h || (!codes && n > 0)) return set_error(VAQGPU_EINVAL, "handle/codes is NULL");
  RC(for_each_piece(h->s, h->s.n_added, n, [&](int r, int64_t off, int64_t cnt) {
    return vaqgpu_add_codes_u16(h->shard[r], codes + (size_t)off * h->M, cnt);
  }));
  h->s.n_added += n;
  return VAQGPU_OK;
}

int vaqgpu_sharded_encode_add(vaqgpu_sharded_t *h, const float *x_proj, int64_t n) {
  if (!h || (!x_proj && n > 0)) return set_error(VAQGPU_EINVAL, "handle/x_proj is NULL");
  RC(for_each_piece(h->s, h->s.n_added, n, [&](int r, int64_t off, int64_t cnt) {
    return vaqgpu_encode_add(h->shard[r], x_proj + (size_t)off * h->D, cnt);
  }));
  h->s.n_added += n;
  return VAQGPU_OK;
}

int vaqgpu_sharded_add_codes_synthetic(vaqgpu_sharded_t *h, int64_t n, uint64_t seed, const float *cdf) {
  if (!h) return set_error(VAQGPU_EINVAL, "handle is NULL");
  RC(for_each_piece(h->s, h->s.n_added, n, [&](int r, int64_t, int64_t cnt) {
    return vaqgpu_add_codes_synthetic(h->shard[r], cnt, seed, cdf);      // rows are a function of their global id
  }));
  h->s.n_added += n;
  return VAQGPU_OK;
}

int vaqgpu_sharded_num_shards(const vaqgpu_sharded_t *h, int32_t *n) {
  if (!h || !n) return set_error(VAQGPU_EINVAL, "NULL argument");
  *n = h->s.G;
  return VAQGPU_OK;
}

int vaqgpu_sharded_shard(vaqgpu_sharded_t *h, int32_t r, vaqgpu_t **shard) {
  if (!h || !shard || r < 0 || r >= h->s.G) return set_error(VAQGPU_EINVAL, "bad shard index");
  *shard = h->shard[r];
  return VAQGPU_OK;
}

int vaqgpu_sharded_search(vaqgpu_sharded_t *h, const float *queries, int32_t nq, int32_t k, uint32_t flags, int32_t *labels,
                          float *dists) {
  RestoreDevice restore;
  if (!h || (nq > 0 && (!queries || !labels || !dists))) return set_error(VAQGPU_EINVAL, "NULL argument");
  if (nq < 0 || k <= 0) return set_error(VAQGPU_EINVAL, "nq=%d k=%d", nq, k);
  if (nq == 0) return VAQGPU_OK;
  ShardSet &s = h->s;
  const int G = s.G;
  RC(exchange_bounds(h, nq));
  const size_t count = (size_t)nq * k;
  for (int i = 0; i < G; i++) {
    CU(cudaSetDevice(s.dev[i]));
    CU(s.q[i].ensure((size_t)nq * h->D * sizeof(float)));
    CU(s.keys[i].ensure(count * sizeof(uint64_t)));
    CU(s.all[i].ensure(count * G * sizeof(uint64_t)));
    CU(cudaMemcpyAsync(s.q[i].p, queries, (size_t)nq * h->D * sizeof(float), cudaMemcpyHostToDevice, s.st[i]));
    RC(vaqgpu_search_keys_device(h->shard[i], (const float *)s.q[i].p, nq, k, flags, (uint64_t *)s.keys[i].p, s.st[i]));
  }
  RC(s.allgather_keys(count));
  CU(cudaSetDevice(s.dev[0]));
  CU(s.out_a.ensure(count * sizeof(int32_t)));
  CU(s.out_b.ensure(count * sizeof(float)));
  RC(vaqgpu_merge_keys_device((const uint64_t *)s.all[0].p, G, nq, k, flags, (int32_t *)s.out_a.p, (float *)s.out_b.p, s.st[0]));
  CU(cudaMemcpyAsync(labels, s.out_a.p, count * sizeof(int32_t), cudaMemcpyDeviceToHost, s.st[0]));
  CU(cudaMemcpyAsync(dists, s.out_b.p, count * sizeof(float), cudaMemcpyDeviceToHost, s.st[0]));
  for (int i = 0; i < G; i++) { CU(cudaSetDevice(s.dev[i])); CU(cudaStreamSynchronize(s.st[i])); }
  return VAQGPU_OK;
}

/* ---------------------------------------------------------------- Hamming ---- */

int hamgpu_sharded_create(int32_t nbits, int32_t n_gpus, const int *dev_ids, int64_t n_rows_total, hamgpu_sharded_t **out) {
  RestoreDevice restore;
  if (!out) return set_error(VAQGPU_EINVAL, "out is NULL");
  *out = nullptr;
  hamgpu_sharded *h = new (std::nothrow) hamgpu_sharded();
  if (!h) return set_error(VAQGPU_ENOMEM, "host allocation failed");
  int rc = h->s.init(n_gpus, dev_ids, n_rows_total);
  for (int i = 0; rc == VAQGPU_OK && i < h->s.G; i++) {
    hamgpu_t *x = nullptr;
    rc = hamgpu_create(nbits, h->s.dev[i], &x);
    if (rc == VAQGPU_OK) {
      h->shard.push_back(x);
      rc = hamgpu_set_id_base(x, h->s.lo(i));
    }
  }
  if (rc != VAQGPU_OK) { hamgpu_sharded_destroy(h); return rc; }
  h->w64 = (nbits + 63) / 64;
  *out = h;
  return VAQGPU_OK;
}

void hamgpu_sharded_destroy(hamgpu_sharded_t *h) {
  RestoreDevice restore;
  if (!h) return;
  for (hamgpu_t *x : h->shard) hamgpu_destroy(x);
  h->s.destroy();
  delete h;
}

int hamgpu_sharded_add(hamgpu_sharded_t *h, const uint64_t *words, int64_t n) {
  if (!h || (!words && n > 0)) return set_error(VAQGPU_EINVAL, "handle/words is NULL");
  RC(for_each_piece(h->s, h->s.n_added, n, [&](int r, int64_t off, int64_t cnt) {
    return hamgpu_add(h->shard[r], words + (size_t)off * h->w64, cnt);
  }));
  h->s.n_added += n;
  return VAQGPU_OK;
}

int hamgpu_sharded_add_synthetic(hamgpu_sharded_t *h, int64_t n, uint64_t seed) {
  if (!h) return set_error(VAQGPU_EINVAL, "handle is NULL");
  RC(for_each_piece(h->s, h->s.n_added, n, [&](int r, int64_t, int64_t cnt) { return hamgpu_add_synthetic(h->shard[r], cnt, seed); }));
  h->s.n_added += n;
  return VAQGPU_OK;
}

int hamgpu_sharded_query(hamgpu_sharded_t *h, const uint64_t *queries, int32_t nq, int32_t k, int32_t *idx, uint32_t *dist) {
  RestoreDevice restore;
  if (!h || (nq > 0 && (!queries || !idx || !dist))) return set_error(VAQGPU_EINVAL, "NULL argument");
  if (nq < 0 || k <= 0) return set_error(VAQGPU_EINVAL, "nq=%d k=%d", nq, k);
  if (nq == 0) return VAQGPU_OK;
  ShardSet &s = h->s;
  const int G = s.G;
  const size_t count = (size_t)nq * k;
  for (int i = 0; i < G; i++) {
    CU(cudaSetDevice(s.dev[i]));
    CU(s.q[i].ensure((size_t)nq * h->w64 * sizeof(uint64_t)));
    CU(s.keys[i].ensure(count * sizeof(uint64_t)));
    CU(s.all[i].ensure(count * G * sizeof(uint64_t)));
    CU(cudaMemcpyAsync(s.q[i].p, queries, (size_t)nq * h->w64 * sizeof(uint64_t), cudaMemcpyHostToDevice, s.st[i]));
    RC(hamgpu_query_keys_device(h->shard[i], (const uint64_t *)s.q[i].p, nq, k, (uint64_t *)s.keys[i].p, s.st[i]));
  }
  RC(s.allgather_keys(count));
  CU(cudaSetDevice(s.dev[0]));
  CU(s.out_a.ensure(count * sizeof(int32_t)));
  CU(s.out_b.ensure(count * sizeof(uint32_t)));
  RC(hamgpu_merge_keys_device((const uint64_t *)s.all[0].p, G, nq, k, (int32_t *)s.out_a.p, (uint32_t *)s.out_b.p, s.st[0]));
  CU(cudaMemcpyAsync(idx, s.out_a.p, count * sizeof(int32_t), cudaMemcpyDeviceToHost, s.st[0]));
  CU(cudaMemcpyAsync(dist, s.out_b.p, count * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.st[0]));
  for (int i = 0; i < G; i++) { CU(cudaSetDevice(s.dev[i])); CU(cudaStreamSynchronize(s.st[i])); }
  return VAQGPU_OK;
}

}  // extern "C"
