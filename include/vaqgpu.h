/* vaqgpu.h — C ABI of the B200-native VAQ / BitVecEngine query-time search path.
 *
 * The reference (TheDatumOrg/VAQ) has no FFI layer: its boundary is the public
 * C++ surface of two classes.  Each entry point below names the reference
 * interface it replaces (file:line under the reference tree); INTEGRATION.md
 * shows the C++ shim that keeps `VAQ::search` / `BitVecEngine::query`
 * source-compatible on top of these calls.
 *
 * Conventions
 *  - plain pointers and sizes only; opaque handles; every function returns 0 on
 *    success and a negative VAQGPU_E* code on failure, with a human-readable
 *    message available from vaqgpu_last_error() (thread-local).  Nothing here
 *    exits the process (the reference prints and calls exit(0)/assert(false),
 *    VAQ.cpp:64-78, 453-456, 1263-1266).
 *  - there is NO CPU fallback: every entry point needs a CUDA device (sm_100a).
 *  - "host" entry points take host buffers and include the H2D/D2H copies;
 *    "_device" entry points take device pointers and run asynchronously on the
 *    caller's stream (a cudaStream_t passed as void*; NULL = default stream).
 *  - one handle may be used from one host thread at a time.
 *  - result order: ascending (distance, id) lexicographically — identical to the
 *    reference wherever distances are distinct (see DESIGN.md "tie rule");
 *    unfilled slots (fewer than k rows) are id = -1, dist = FLT_MAX (ADC) or
 *    0xFFFFFFFF (Hamming), as utils/Heap.hpp:230-233,344-347 leaves them.
 */
#ifndef VAQGPU_H_
#define VAQGPU_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAQGPU_OK 0
#define VAQGPU_EINVAL (-1)   /* bad argument */
#define VAQGPU_ECUDA (-2)    /* CUDA runtime error (message has the CUDA string) */
#define VAQGPU_ENOMEM (-3)   /* host or device allocation failed */
#define VAQGPU_ESTATE (-4)   /* call not valid in the handle's current state */

/* search flags.  The low byte mirrors VAQ::NNMethod (VAQ.hpp:38-49). */
#define VAQGPU_EA 0x02u         /* early abandon (searchEarlyAbandon, VAQ.cpp:1694) */
#define VAQGPU_TI 0x04u         /* cluster-ordered scan restricted by `visit` (VAQ.cpp:1540) */
#define VAQGPU_HEAP 0x80u       /* exhaustive scan (searchHeap, VAQ.cpp:1729) */
#define VAQGPU_PROJECTED 0x100u /* queries are already in PCA space: skip (X*V).real(), VAQ.cpp:777 */
#define VAQGPU_SQRT 0x200u      /* return sqrt(ADC) as the TI mode does (VAQ.cpp:1585) */
#define VAQGPU_SCAN_V1 0x1000u  /* diagnostics: force the lane-per-row scan kernel for EA searches */
#define VAQGPU_SCAN_F32 0x2000u /* diagnostics: force the fp32-table filter kernel (skip the fp16 lower-bound tables) */

typedef struct vaqgpu_index vaqgpu_t;
typedef struct hamgpu_index hamgpu_t;

/* Trained model, i.e. the public members a trained reference `VAQ` exposes
 * (VAQ.hpp:57-73): mSubsLen, mHighestSubs, mBitsAlloc, mCentroidsPerSubs,
 * real(mEigenVectors). */
typedef struct {
  int32_t D;               /* padded dimensionality = M * L */
  int32_t M;               /* mHighestSubs: number of subspaces scanned (1..128) */
  int32_t L;               /* mSubsLen: dims per subspace */
  const int32_t *bits;     /* [M] mBitsAlloc, 1..15 each, sum <= 1024 */
  const float *centroids;  /* concatenation over s of row-major [2^bits[s] x L] (mCentroidsPerSubs[s]) */
  const float *eig_real;   /* [D x D] row-major real(mEigenVectors), or NULL (queries always pre-projected) */
} vaqgpu_model_desc;

const char *vaqgpu_last_error(void);
int vaqgpu_device_count(int *count);

/* ---- VAQ index ---------------------------------------------------------- */

/* replaces: constructing a `VAQ` and filling it by train() (VAQ.hpp:36-114). */
int vaqgpu_create(const vaqgpu_model_desc *model, int device, vaqgpu_t **out);
void vaqgpu_destroy(vaqgpu_t *h);

/* Global id of this index's first row (row-sharded deployments: shard r holds rows
 * [id_base, id_base + n)).  Returned labels are id_base + local row. */
int vaqgpu_set_id_base(vaqgpu_t *h, int64_t id_base);

/* Append n encoded rows.  `codes` is row-major [n x M] uint16 == mCodebook
 * (VAQ.hpp:72, utils/Types.hpp:31), i.e. exactly what VAQ::encode (VAQ.cpp:663)
 * produced on the host.  The rows are bit-packed on the device into the scan layout. */
int vaqgpu_add_codes_u16(vaqgpu_t *h, const uint16_t *codes, int64_t n);

/* Encode n already-projected rows [n x D] on the device and append them
 * (replaces VAQ::encodeImpl, VAQ.cpp:728-748: nearest centroid, lowest code on ties). */
int vaqgpu_encode_add(vaqgpu_t *h, const float *x_proj, int64_t n);

/* Append n rows generated on the device: code[i][s] = inverse-CDF(hash(seed, id_base+i, s)).
 * `cdf` is the concatenation over s of 2^bits[s] cumulative probabilities (last = 1), or NULL
 * for uniform codes.  Used for the 100M / 1B-row shapes the host cannot hold; the same
 * generator is restated in numpy (vaq_b200/synth.py) for CPU parity on row slices. */
int vaqgpu_add_codes_synthetic(vaqgpu_t *h, int64_t n, uint64_t seed, const float *cdf);

/* Pre-size the packed code matrix for n_total rows (avoids regrowth copies on 100M+ row shards). */
int vaqgpu_reserve(vaqgpu_t *h, int64_t n_total);

int vaqgpu_num_rows(const vaqgpu_t *h, int64_t *n);
/* bytes per packed row in HBM (16 * ceil(sum(bits)/128)) */
int vaqgpu_row_bytes(const vaqgpu_t *h, int32_t *bytes);

/* Unpack rows [row0, row0+n) back to [n x M] uint16 (round-trip check against mCodebook). */
int vaqgpu_get_codes_u16(vaqgpu_t *h, int64_t row0, int64_t n, uint16_t *out);

/* Storage order of the packed matrix: out[i] = original (arrival) index of the row stored at position srow0 + i.
 * Before the first search the library groups the rows of an index of 32 K .. 8 M rows by a coarse clustering of their
 * leading subspaces (scan order: a query tile starts at the rows nearest to it, vaqgpu_host.cu build_scan_order) and
 * re-orders rows inside windows of at most 4096 so that the eight rows a quarter-warp gathers for hit different
 * shared-memory banks (csrc/layout.cu); ids, codes and results are always expressed in the original order — this
 * call exists for diagnostics and tests. */
int vaqgpu_get_row_order(vaqgpu_t *h, int64_t srow0, int64_t n, uint32_t *out);

/* replaces VAQ::CreateLUT (VAQ.hpp:128-167).  q_proj: host [nq x D] projected queries;
 * lut_out: host [nq x sum_s 2^bits[s]], table s at offset sum_{t<s} 2^bits[t]. */
int vaqgpu_build_lut(vaqgpu_t *h, const float *q_proj, int32_t nq, float *lut_out);

/* replaces VAQ::search (VAQ.cpp:776-847).  queries: [nq x D] (raw unless VAQGPU_PROJECTED);
 * labels/dists: [nq x k].  Squared ADC distances unless VAQGPU_SQRT. */
int vaqgpu_search(vaqgpu_t *h, const float *queries, int32_t nq, int32_t k, uint32_t flags,
                  int32_t *labels, float *dists);
int vaqgpu_search_device(vaqgpu_t *h, const float *d_queries, int32_t nq, int32_t k, uint32_t flags,
                         int32_t *d_labels, float *d_dists, void *stream);

/* Shard-local top-k as sortable 64-bit keys (float bits of the distance in the high word,
 * id in the low word) [nq x k]; combine the lists of G shards with vaqgpu_merge_keys_device.
 * The key lists replace the per-batch partial answers the reference concatenates and
 * re-sorts in BitVecEngine.cpp:1599-1611. */
int vaqgpu_search_keys_device(vaqgpu_t *h, const float *d_queries, int32_t nq, int32_t k, uint32_t flags,
                              uint64_t *d_keys, void *stream);
/* d_keys_in: [G x nq x k] (each [nq x k] block sorted ascending); writes the k smallest per query. */
int vaqgpu_merge_keys_device(const uint64_t *d_keys_in, int32_t G, int32_t nq, int32_t k, uint32_t flags,
                             int32_t *d_labels, float *d_dists, void *stream);

/* ---- cross-shard bound exchange (row-sharded deployments) ------------------------
 * Each shard owns a per-query array of "best k-th distance proven so far"; its scan publishes every new bound into
 * its own array and — through NVLink peer memory — into the arrays of the other shards, so all GPUs prune with the
 * tightest bound found anywhere (the reference's single running bsfK, VAQ.cpp:1700-1721, spans all rows; without
 * the exchange a shard only knows the k-th best of its own N/G rows).  Results are unchanged (exact pruning).
 *  - vaqgpu_bounds_export allocates the array (searches with nq <= max_queries use it), returns its device pointer
 *    and a 64-byte CUDA IPC handle for shards living in other processes (either out-pointer may be NULL);
 *  - vaqgpu_bounds_attach_ipc / _ptr hand this shard the arrays of its n_peers (<= 15) peers;
 *  - from then on vaqgpu_search*_device must be called collectively: every shard runs the same sequence of searches
 *    (same queries), each followed by a collective that waits for all shards (the all-gather of the key lists). */
int vaqgpu_bounds_export(vaqgpu_t *h, int32_t max_queries, unsigned char ipc_handle[64], void **d_ptr);
int vaqgpu_bounds_attach_ipc(vaqgpu_t *h, int32_t n_peers, const unsigned char *ipc_handles /* n_peers x 64 */);
int vaqgpu_bounds_attach_ptr(vaqgpu_t *h, int32_t n_peers, void *const *peer_ptrs);

/* TI / visit mode (replaces VAQ::clusterTI's outputs + searchTriangleInequality, VAQ.cpp:878-999,
 * 1540-1692).  The index rows must already be in cluster-grouped order (as clusterTI leaves
 * mCodebook).  clusters: [C x segdims]; start/size: [C] row ranges; id_map: [n] original id of each
 * grouped row (mTIClustersMember flattened).  A search with VAQGPU_TI ranks clusters by
 * ||q[0:segdims] - cc|| and scans the nearest floor(visit*C) (all when visit >= 1), extending
 * until at least k rows were scanned. */
int vaqgpu_set_clusters(vaqgpu_t *h, const float *clusters, int32_t C, int32_t segdims,
                        const int64_t *start, const int64_t *size, const int32_t *id_map);
/* replaces VAQ::clusterTI itself (VAQ.cpp:878-999) on the device: k-means (`iters` Lloyd iterations from evenly
 * spaced rows; 0 = assign to the seed rows only, as useKMeans = false) over the rows decoded in their first
 * n_segments subspaces (<= 0: all, mTISegmentNum = -1), rows regrouped by cluster in place (stable), cluster ranges and
 * the original id of every regrouped row kept on the device.  Equivalent to vaqgpu_set_clusters with the result.
 * Clustering parity with the reference is unpinned (it calls Armadillo's kmeans); searches over the clusters are exact.
 * vaqgpu_get_clusters reads the clustering back (any output may be NULL; clusters [C x segdims], start/size [C],
 * id_map [rows]). */
int vaqgpu_cluster_ti(vaqgpu_t *h, int32_t C, int32_t n_segments, int32_t iters);
int vaqgpu_get_clusters(vaqgpu_t *h, int32_t *C, int32_t *segdims, float *clusters, int64_t *start, int64_t *size,
                        int32_t *id_map);
int vaqgpu_set_visit(vaqgpu_t *h, float visit);
/* Row-sharded TI: a shard's vaqgpu_set_clusters holds the part of each cluster that falls into its rows; the visiting
 * rule ("continue while fewer than k rows were covered", VAQ.cpp:1555,1616-1618) must count the clusters' sizes in the
 * whole index.  sizes: [C]. */
int vaqgpu_set_cluster_rule_sizes(vaqgpu_t *h, const int64_t *sizes);

/* replaces VAQ::refine (VAQ.cpp:849-876): exact squared-L2 re-rank of `refine_num` candidate
 * labels per query against raw vectors.  xtrain: host [n x D0] raw rows (uploaded once and
 * cached by pointer+size); queries raw [nq x D0]. */
int vaqgpu_set_raw_vectors(vaqgpu_t *h, const float *xtrain, int64_t n, int32_t D0);
int vaqgpu_refine(vaqgpu_t *h, const float *queries, int32_t nq, const int32_t *in_labels,
                  int32_t refine_num, int32_t k, int32_t *labels, float *dists);

/* Timing of the kernels of the last search on this handle, in milliseconds (CUDA events on
 * the launching stream): [0]=projection, [1]=LUT build, [2]=ADC scan, [3]=merge/output. */
int vaqgpu_last_timings(const vaqgpu_t *h, float ms[4]);
/* Scan configuration chosen for the last search: [0]=threads/CTA, [1]=row chunks (CTAs per query tile),
 * [2]=LUT entries per query resident in shared memory, [3]=LUT entries spilled to L2, [4]=dynamic smem
 * bytes, [5]=uint4 words per row, [6]=kernel launches of the last search, [7]=queries per launch,
 * [8]=queries per CTA (tile width T), [9]=scan kernel (1 = lane-per-row, 2 = filter-and-refine on fp32
 * tables, 3 = filter-and-refine on fp16 lower-bound tables), [10]=1 when the rows are in the conflict-aware order
 * (csrc/layout.cu), [11]=microseconds spent re-ordering so far. */
int vaqgpu_last_config(const vaqgpu_t *h, int32_t cfg[12]);

/* ---- BitVecEngine Hamming index ---------------------------------------- */

/* replaces BitVecEngine(int N) + loadBitV / appendBitV (BitVecEngine.hpp:105-106, .cpp:12,1630).
 * words: row-major [n x ceil(nbits/64)] uint64 == bitvectors (BitVector.hpp:13-19). */
int hamgpu_create(int32_t nbits, int device, hamgpu_t **out);
void hamgpu_destroy(hamgpu_t *h);
int hamgpu_set_id_base(hamgpu_t *h, int64_t id_base);
int hamgpu_add(hamgpu_t *h, const uint64_t *words, int64_t n);
int hamgpu_add_synthetic(hamgpu_t *h, int64_t n, uint64_t seed);
int hamgpu_num_rows(const hamgpu_t *h, int64_t *n);

/* replaces BitVecEngine::query / queryParallel (BitVecEngine.cpp:509-519, 1264-1304):
 * k nearest rows by Hamming distance (utils/DistanceFunctions.hpp:164-172), ascending
 * (distance, id).  idx/dist: [nq x k] ({int idx; uint32_t dist}, utils/Types.hpp:42-51). */
int hamgpu_query(hamgpu_t *h, const uint64_t *queries, int32_t nq, int32_t k, int32_t *idx, uint32_t *dist);
int hamgpu_query_device(hamgpu_t *h, const uint64_t *d_queries, int32_t nq, int32_t k,
                        int32_t *d_idx, uint32_t *d_dist, void *stream);
int hamgpu_query_keys_device(hamgpu_t *h, const uint64_t *d_queries, int32_t nq, int32_t k,
                             uint64_t *d_keys, void *stream);
int hamgpu_merge_keys_device(const uint64_t *d_keys_in, int32_t G, int32_t nq, int32_t k,
                             int32_t *d_idx, uint32_t *d_dist, void *stream);
int hamgpu_last_timings(const hamgpu_t *h, float ms[2]); /* [0]=scan, [1]=merge/output */
int hamgpu_last_config(const hamgpu_t *h, int32_t cfg[8]);

/* ---- row-sharded handles: one host process, n_gpus devices -------------------
 * Shard r holds global rows [r * ceil(N/G), (r+1) * ceil(N/G)) of an index of n_rows_total rows; rows are appended in
 * global order and split at the shard boundaries.  A search copies the query batch to every GPU, runs the shard-local
 * scans concurrently (one stream per device; running k-th-best bounds exchanged through NVLink peer memory), gathers
 * the shard-local top-k key lists with one ncclAllGather (ncclCommInitAll communicator; NCCL is loaded at run time)
 * and merges them on the first device: same answer as one unsharded index, bit for bit.  dev_ids == NULL: 0..n_gpus-1.
 * The merge rule's precedent in the reference: BitVecEngine.cpp:1599-1611. */
typedef struct vaqgpu_sharded vaqgpu_sharded_t;
typedef struct hamgpu_sharded hamgpu_sharded_t;
int vaqgpu_sharded_create(const vaqgpu_model_desc *model, int32_t n_gpus, const int *dev_ids, int64_t n_rows_total,
                          vaqgpu_sharded_t **out);
void vaqgpu_sharded_destroy(vaqgpu_sharded_t *h);
int vaqgpu_sharded_add_codes_u16(vaqgpu_sharded_t *h, const uint16_t *codes, int64_t n);
int vaqgpu_sharded_encode_add(vaqgpu_sharded_t *h, const float *x_proj, int64_t n);
int vaqgpu_sharded_add_codes_synthetic(vaqgpu_sharded_t *h, int64_t n, uint64_t seed, const float *cdf);
int vaqgpu_sharded_num_shards(const vaqgpu_sharded_t *h, int32_t *n);
int vaqgpu_sharded_shard(vaqgpu_sharded_t *h, int32_t r, vaqgpu_t **shard);   /* borrowed: per-shard introspection */
/* replaces VAQ::search on the sharded index; host buffers, EA / HEAP flags (see vaqgpu_search) */
int vaqgpu_sharded_search(vaqgpu_sharded_t *h, const float *queries, int32_t nq, int32_t k, uint32_t flags,
                          int32_t *labels, float *dists);
int hamgpu_sharded_create(int32_t nbits, int32_t n_gpus, const int *dev_ids, int64_t n_rows_total, hamgpu_sharded_t **out);
void hamgpu_sharded_destroy(hamgpu_sharded_t *h);
int hamgpu_sharded_add(hamgpu_sharded_t *h, const uint64_t *words, int64_t n);
int hamgpu_sharded_add_synthetic(hamgpu_sharded_t *h, int64_t n, uint64_t seed);
/* replaces BitVecEngine::query / queryParallel on the sharded bit-vector matrix */
int hamgpu_sharded_query(hamgpu_sharded_t *h, const uint64_t *queries, int32_t nq, int32_t k, int32_t *idx, uint32_t *dist);

#ifdef __cplusplus
}
#endif
#endif /* VAQGPU_H_ */
