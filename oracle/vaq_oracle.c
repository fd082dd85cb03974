/* TEST INFRASTRUCTURE — CPU restatement ("oracle") of the reference's query-time
 * hot path, in plain C.  It is the checker the CUDA path is compared against;
 * it is never shipped, never imported by vaq_b200/, and never the thing measured
 * (except as bench.py's cpu_baseline "port" leg when oracle/_ref is absent).
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function here
 * against oracle/_ref/libvaq_ref.so (the unmodified reference compiled from
 * /root/reference, see oracle/Makefile) on seeded inputs, and against the
 * reference's own known-answer tests (test/test-distancefunction.cpp:11-63,
 * 118-132; test/test-bitvecengine.cpp:64-79, 165-179, 246-260); the outputs of
 * the reference are committed as tests/golden/ fixtures so the pin also holds on
 * machines without /root/reference.
 *
 * Every function cites the reference file:line it follows.  Layout difference
 * (values identical): the reference LUT is col-major [2^maxBits x M] with unused
 * tail entries zero (VAQ.hpp:130,137); here it is compact, table s starting at
 * lut_off[s] = sum_{t<s} 2^bits[t].
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ LUT */

/* utils/Math.hpp:147-171 fvec_L2sqr_ny with ElementOpL2 (:130-145): used by
 * CreateLUT for subspaces with fewer than 8 centroids (VAQ.hpp:161-165) and for
 * the TI query->centroid distances (VAQ.cpp:802).  The summation trees below are
 * those of the SSE specialisations for d in {1,2,4,8,12} (:38-128: packed mul,
 * then haddps pairs; for d=8/12 the 4-lane partial sums are accumulated with the
 * contraction GCC applies under the reference flags, -O3 -ffast-math -mfma:
 * accu = fma(t,t,accu)); other d fall to the scalar loop (:8-19) whose order
 * under -ffast-math is compiler-chosen => compared with a tolerance only. */
static float l2sqr_small(const float *x, const float *y, int d) {
  float t[12];
  switch (d) {
    case 1: {
      float a = x[0] - y[0];
      return a * a;
    }
    case 2: {
      float a = x[0] - y[0], b = x[1] - y[1];
      return a * a + b * b;
    }
    case 4: {
      for (int i = 0; i < 4; i++) { float a = x[i] - y[i]; t[i] = a * a; }
      return (t[0] + t[1]) + (t[2] + t[3]);
    }
    case 8: {
      for (int i = 0; i < 4; i++) {
        float a = x[i] - y[i], b = x[i + 4] - y[i + 4];
        t[i] = fmaf(b, b, a * a);
      }
      return (t[0] + t[1]) + (t[2] + t[3]);
    }
    case 12: {
      for (int i = 0; i < 4; i++) {
        float a = x[i] - y[i], b = x[i + 4] - y[i + 4], c = x[i + 8] - y[i + 8];
        t[i] = fmaf(c, c, fmaf(b, b, a * a));
      }
      return (t[0] + t[1]) + (t[2] + t[3]);
    }
    default: {
      float res = 0.f;
      for (int i = 0; i < d; i++) { float a = x[i] - y[i]; res += a * a; }
      return res;
    }
  }
}

void orc_l2sqr_ny(float *dis, const float *x, const float *y, int d, long ny) {
  for (long i = 0; i < ny; i++) dis[i] = l2sqr_small(x, y + (size_t)i * d, d);
}

/* VAQ::CreateLUT, AVX2 variant, VAQ.hpp:128-167.
 *  K_s >= 8 (:134-160): per centroid an accumulator starts at 0 and receives
 *  acc = fma(q_j - c_j, q_j - c_j, acc) for j = 0..L-1 in order (vfmadd231ps,
 *  utils/AVXUtils.hpp:11-15) -> one fused rounding per dimension.
 *  K_s < 8 (:161-165): fvec_L2sqr_ny over the row-major centroids.
 * centroids: concatenated row-major [K_s x L] blocks (mCentroidsPerSubs). */
void orc_create_lut(int M, int L, const int *bits, const float *centroids, const float *q, float *lut) {
  const float *cs = centroids;
  float *out = lut;
  for (int s = 0; s < M; s++) {
    const int K = 1 << bits[s];
    const float *qs = q + (size_t)s * L;
    if (K >= 8) {
      for (int c = 0; c < K; c++) {
        float acc = 0.f;
        for (int j = 0; j < L; j++) {
          float d = qs[j] - cs[(size_t)c * L + j];
          acc = fmaf(d, d, acc);
        }
        out[c] = acc;
      }
    } else {
      orc_l2sqr_ny(out, qs, cs, L, K);
    }
    cs += (size_t)K * L;
    out += K;
  }
}

/* ------------------------------------------------------------------ heap */
/* utils/Heap.hpp:73-88 CMax<float,int>::cmp(a,b) = a > b; neutral = FLT_MAX. */

/* Heap.hpp:115-144 heap_pop */
static void heap_pop(size_t k, float *bh_val, int *bh_ids) {
  bh_val--; bh_ids--;
  float val = bh_val[k];
  size_t i = 1, i1, i2;
  while (1) {
    i1 = i << 1; i2 = i1 + 1;
    if (i1 > k) break;
    if (i2 == k + 1 || bh_val[i1] > bh_val[i2]) {
      if (val > bh_val[i1]) break;
      bh_val[i] = bh_val[i1]; bh_ids[i] = bh_ids[i1]; i = i1;
    } else {
      if (val > bh_val[i2]) break;
      bh_val[i] = bh_val[i2]; bh_ids[i] = bh_ids[i2]; i = i2;
    }
  }
  bh_val[i] = bh_val[k]; bh_ids[i] = bh_ids[k];
}

/* Heap.hpp:151-169 heap_push */
static void heap_push(size_t k, float *bh_val, int *bh_ids, float val, int id) {
  bh_val--; bh_ids--;
  size_t i = k, i_father;
  while (i > 1) {
    i_father = i >> 1;
    if (!(val > bh_val[i_father])) break;
    bh_val[i] = bh_val[i_father]; bh_ids[i] = bh_ids[i_father]; i = i_father;
  }
  bh_val[i] = val; bh_ids[i] = id;
}

/* Heap.hpp:212-235 heap_heapify with k0 = 0: all slots neutral / -1 */
static void heap_heapify0(size_t k, float *bh_val, int *bh_ids) {
  for (size_t i = 0; i < k; i++) { bh_val[i] = FLT_MAX; bh_ids[i] = -1; }
}

/* Heap.hpp:322-349 heap_reorder: ascending order, unfilled slots (FLT_MAX,-1) last */
static size_t heap_reorder(size_t k, float *bh_val, int *bh_ids) {
  size_t i, ii;
  for (i = 0, ii = 0; i < k; i++) {
    float val = bh_val[0]; int id = bh_ids[0];
    heap_pop(k - i, bh_val, bh_ids);
    bh_val[k - ii - 1] = val; bh_ids[k - ii - 1] = id;
    if (id != -1) ii++;
  }
  size_t nel = ii;
  memmove(bh_val, bh_val + k - ii, ii * sizeof(*bh_val));
  memmove(bh_ids, bh_ids + k - ii, ii * sizeof(*bh_ids));
  for (; ii < k; ii++) { bh_val[ii] = FLT_MAX; bh_ids[ii] = -1; }
  return nel;
}

/* ------------------------------------------------------------------ ADC scans */

/* VAQ::searchHeap, VAQ.cpp:1729-1758.  codes: row-major [N x M] uint16
 * (mCodebook).  M must be a multiple of 4.  Sum grouping:
 * dism = ((l0+l1)+l2)+l3; dist += dism  (:1741-1748). */
void orc_search_heap(int M, const int *lut_off, const float *lut, const uint16_t *codes, long N, int k,
                     int *ids, float *dis) {
  heap_heapify0(k, dis, ids);
  const uint16_t *c = codes;
  for (long i = 0; i < N; i++) {
    float dist = 0.f;
    for (int col = 0; col < M; col += 4) {
      float dism = lut[lut_off[col] + c[0]];
      dism += lut[lut_off[col + 1] + c[1]];
      dism += lut[lut_off[col + 2] + c[2]];
      dism += lut[lut_off[col + 3] + c[3]];
      dist += dism;
      c += 4;
    }
    if (dis[0] > dist) {
      heap_pop(k, dis, ids);
      heap_push(k, dis, ids, dist, (int)i);
    }
  }
  heap_reorder(k, dis, ids);
}

/* VAQ::searchEarlyAbandon, VAQ.cpp:1694-1727: identical but the 4-subspace loop
 * stops once dist >= bsfK (:1708), bsfK = heap top after each insertion (:1721). */
void orc_search_ea(int M, const int *lut_off, const float *lut, const uint16_t *codes, long N, int k,
                   int *ids, float *dis) {
  heap_heapify0(k, dis, ids);
  float bsfK = FLT_MAX;
  const uint16_t *c = codes;
  for (long i = 0; i < N; i++) {
    float dist = 0.f;
    int col;
    for (col = 0; col < M && dist < bsfK; col += 4) {
      float dism = lut[lut_off[col] + c[col]];
      dism += lut[lut_off[col + 1] + c[col + 1]];
      dism += lut[lut_off[col + 2] + c[col + 2]];
      dism += lut[lut_off[col + 3] + c[col + 3]];
      dist += dism;
    }
    c += M;
    if (dis[0] > dist) {
      heap_pop(k, dis, ids);
      heap_push(k, dis, ids, dist, (int)i);
      bsfK = dis[0];
    }
  }
  heap_reorder(k, dis, ids);
}

typedef struct { float d; int i; } fi_pair;
static int cmp_fi(const void *a, const void *b) {
  const fi_pair *x = (const fi_pair *)a, *y = (const fi_pair *)b;
  if (x->d < y->d) return -1;
  if (x->d > y->d) return 1;
  return (x->i > y->i) - (x->i < y->i);
}

/* VAQ::searchTriangleInequality, VAQ.cpp:1540-1692, with the per-query prologue
 * of VAQ::search :799-827.
 *  - qToCC[c] = sqrt(fvec_L2sqr_ny(q[0:segdims], cluster c))            (:802-805)
 *  - clusters visited in ascending qToCC (std::sort, ties unspecified; here
 *    broken by cluster index)                                           (:815-820)
 *  - maxClusterVisit = floor(C * visit) if visit < 1 else C             (:1548-1551)
 *  - loop continues past maxClusterVisit while fewer than k rows seen   (:1555)
 *  - inside a cluster rows are stored far->near; once k rows are held, the
 *    member loop breaks when bsfK <= qToCC - codeToCC[id]               (:1566-1569)
 *  - distances are sqrt(ADC), ids are original ids from the member list (:1585-1607)
 * codes are the REGROUPED codebook (VAQ.cpp:984-996); members lists the original
 * ids in regrouped row order; code_to_cc is indexed by original id. */
void orc_search_ti(int M, const int *lut_off, const float *lut, const uint16_t *codes_grouped,
                   int C, int segdims, const float *clusters, const int *start_idx, const int *sizes,
                   const int *members, const float *code_to_cc, const float *q, float visit, int use_ea,
                   int k, int *ids, float *dis, long *pruned_out) {
  fi_pair *order = (fi_pair *)malloc(sizeof(fi_pair) * (size_t)C);
  float *qToCC = (float *)malloc(sizeof(float) * (size_t)C);
  orc_l2sqr_ny(qToCC, q, clusters, segdims, C);
  for (int c = 0; c < C; c++) { qToCC[c] = sqrtf(qToCC[c]); order[c].d = qToCC[c]; order[c].i = c; }
  qsort(order, (size_t)C, sizeof(fi_pair), cmp_fi);
  /* member list offsets in regrouped order == start_idx */
  heap_heapify0(k, dis, ids);
  float bsfK = 0.f, bsfK2 = 0.f;
  int counter = 0;
  long pruned = 0;
  int maxVisit = C;
  if (visit < 1.f) maxVisit = (int)((float)C * visit);
  int enough = 0;
  for (int cc = 0; (cc < maxVisit) || (!enough && cc < C); cc++) {
    const int cl = order[cc].i;
    const int st = start_idx[cl];
    if (sizes[cl] == 0) continue;
    const uint16_t *c = codes_grouped + (size_t)M * st;
    int inter = 0;
    for (int mi = 0; mi < sizes[cl]; mi++) {
      const int id = members[st + mi];
      if (counter >= k) {
        if (bsfK <= (qToCC[cl] - code_to_cc[id])) { pruned += sizes[cl] - inter; break; }
        float dist = 0.f;
        int col;
        for (col = 0; col < M && (!use_ea || dist < bsfK2); col += 4) {
          float dism = lut[lut_off[col] + c[col]];
          dism += lut[lut_off[col + 1] + c[col + 1]];
          dism += lut[lut_off[col + 2] + c[col + 2]];
          dism += lut[lut_off[col + 3] + c[col + 3]];
          dist += dism;
        }
        c += M;
        if (dist < bsfK2) {
          dist = sqrtf(dist);
          heap_pop(k, dis, ids);
          heap_push(k, dis, ids, dist, id);
          bsfK = dis[0];
          bsfK2 = bsfK * bsfK;
        }
      } else {
        float dist = 0.f;
        for (int col = 0; col < M; col += 4) {
          float dism = lut[lut_off[col] + c[col]];
          dism += lut[lut_off[col + 1] + c[col + 1]];
          dism += lut[lut_off[col + 2] + c[col + 2]];
          dism += lut[lut_off[col + 3] + c[col + 3]];
          dist += dism;
        }
        c += M;
        dist = sqrtf(dist);
        heap_pop(k, dis, ids);
        heap_push(k, dis, ids, dist, id);
        if (dist > bsfK) { bsfK = dist; bsfK2 = bsfK * bsfK; }
        counter++;
      }
      inter++;
    }
    if (counter >= k) enough = 1;
  }
  for (int i = maxVisit; i < C; i++) pruned += sizes[order[i].i];
  heap_reorder(k, dis, ids);
  if (pruned_out) *pruned_out = pruned;
  free(order); free(qToCC);
}

/* VAQ::search, VAQ.cpp:776-847 for already-projected queries (the projection
 * :777 is an Eigen GEMM, fed identically to both sides by the tests).
 * mode: 0 = HEAP (:831), 1 = EA (:829).  Queries are independent; nthreads>1
 * slices them (the reference loop :786 is serial). */
void orc_search(int M, int L, const int *bits, const float *centroids, const uint16_t *codes, long N,
                const float *q_proj, int nq, int k, int mode, int nthreads, int *labels, float *dists) {
  int *lut_off = (int *)malloc(sizeof(int) * (size_t)(M + 1));
  lut_off[0] = 0;
  for (int s = 0; s < M; s++) lut_off[s + 1] = lut_off[s] + (1 << bits[s]);
  const int D = M * L;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
  {
    float *lut = (float *)malloc(sizeof(float) * (size_t)lut_off[M]);
#pragma omp for schedule(dynamic, 1)
    for (int qi = 0; qi < nq; qi++) {
      orc_create_lut(M, L, bits, centroids, q_proj + (size_t)qi * D, lut);
      if (mode == 1)
        orc_search_ea(M, lut_off, lut, codes, N, k, labels + (size_t)qi * k, dists + (size_t)qi * k);
      else
        orc_search_heap(M, lut_off, lut, codes, N, k, labels + (size_t)qi * k, dists + (size_t)qi * k);
    }
    free(lut);
  }
  free(lut_off);
}

/* VAQ::refine, VAQ.cpp:849-876: exact squared L2 between the raw query and the
 * raw rows named by in_labels, k smallest through the same heap.  (Eigen's
 * squaredNorm reduction order is library-chosen => distances tolerance-only.) */
void orc_refine(const float *queries, int nq, int D, const int *in_labels, int refine_num,
                const float *xtrain, int k, int *labels, float *dists) {
  for (int qi = 0; qi < nq; qi++) {
    int *ids = labels + (size_t)qi * k;
    float *dis = dists + (size_t)qi * k;
    heap_heapify0(k, dis, ids);
    for (int i = 0; i < refine_num; i++) {
      const int id = in_labels[(size_t)qi * refine_num + i];
      const float *x = xtrain + (size_t)id * D, *q = queries + (size_t)qi * D;
      float dist = 0.f;
      for (int j = 0; j < D; j++) { float d = q[j] - x[j]; dist += d * d; }
      if (dis[0] > dist) { heap_pop(k, dis, ids); heap_push(k, dis, ids, dist, id); }
    }
    heap_reorder(k, dis, ids);
  }
}

/* VAQ::encodeImpl, VAQ.cpp:728-748: per (row, subspace) the centroid with the
 * smallest squared L2, strict '<' => lowest code wins ties (:739-742).  Distance
 * via Eigen squaredNorm (order library-chosen); restated as a sequential sum.
 * margin_out (optional) receives second-best minus best distance, so tests can
 * tell a genuine mismatch from a float near-tie. */
void orc_encode(int M, int L, const int *bits, const float *centroids, const float *x_proj, long N,
                uint16_t *codes, float *margin_out) {
  const int D = M * L;
  const float *cs = centroids;
  for (int s = 0; s < M; s++) {
    const int K = 1 << bits[s];
#pragma omp parallel for schedule(static)
    for (long r = 0; r < N; r++) {
      const float *x = x_proj + (size_t)r * D + (size_t)s * L;
      uint16_t best = 0;
      float bsf = FLT_MAX, second = FLT_MAX;
      for (int c = 0; c < K; c++) {
        float dist = 0.f;
        for (int j = 0; j < L; j++) { float d = x[j] - cs[(size_t)c * L + j]; dist += d * d; }
        if (dist < bsf) { second = bsf; best = (uint16_t)c; bsf = dist; }
        else if (dist < second) second = dist;
      }
      codes[(size_t)r * M + s] = best;
      if (margin_out) margin_out[(size_t)r * M + s] = second - bsf;
    }
    cs += (size_t)K * L;
  }
}

/* ------------------------------------------------------------------ Hamming */

/* utils/DistanceFunctions.hpp:164-172 hammingDist */
uint32_t orc_hamming_dist(const uint64_t *a, const uint64_t *b, int w) {
  uint32_t s = 0;
  for (int i = 0; i < w; i++) s += (uint32_t)__builtin_popcountll(a[i] ^ b[i]);
  return s;
}
/* DistanceFunctions.hpp:174-182 hammingDistEarlyAbandon */
static uint32_t hamming_ea(const uint64_t *a, const uint64_t *b, int w, uint32_t bsf) {
  uint32_t s = 0;
  for (int i = 0; i < w && s < bsf; i++) s += (uint32_t)__builtin_popcountll(a[i] ^ b[i]);
  return s;
}

typedef struct { int idx; uint32_t dist; } id_pair;   /* utils/Types.hpp:42-51 IdxDistPair */

/* libstdc++ <bits/stl_heap.h> (GCC 13, the toolchain the reference is built with
 * here): __push_heap / __adjust_heap / pop_heap / sort_heap with the comparator
 * a.dist < b.dist of BitVecEngine.cpp:133-135 — restated so equal-distance ties
 * resolve exactly as in query_heap / query_heap_early_abandon / queryParallel. */
static void std_push_heap(id_pair *first, long hole, long top, id_pair value) {
  long parent = (hole - 1) / 2;
  while (hole > top && first[parent].dist < value.dist) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}
static void std_adjust_heap(id_pair *first, long hole, long len, id_pair value) {
  const long top = hole;
  long child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (first[child].dist < first[child - 1].dist) child--;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  std_push_heap(first, hole, top, value);
}
static void std_pop_heap(id_pair *first, long len) {   /* [first, first+len) */
  if (len > 1) {
    id_pair value = first[len - 1];
    first[len - 1] = first[0];
    std_adjust_heap(first, 0, len - 1, value);
  }
}
static void std_sort_heap(id_pair *first, long len) {
  while (len > 1) { std_pop_heap(first, len); len--; }
}

/* BitVecEngine::maybeInsertNeighbor, BitVecEngine.hpp:110-127 */
static void maybe_insert(id_pair *nb, int len, id_pair nw) {
  int i = len - 1;
  if (nw.dist < nb[i].dist) nb[i] = nw;
  while (i > 0 && nb[i - 1].dist > nw.dist) {
    id_pair t = nb[i - 1]; nb[i - 1] = nb[i]; nb[i] = t; i--;
  }
}

/* BitVecEngine::query, BitVecEngine.cpp:509-519, method numbering of
 * BitVecEngine.hpp:82-84 {Heap=0, Sort=1, HeapEarlyAbandon=2, SortEarlyAbandon=3}.
 *  Sort   : query_sort :61-78 + KNNFromDists BitVecEngine.hpp:152-168 (std::sort of
 *           the first k pairs restated as a stable insertion sort — exact for
 *           k <= 16, libstdc++'s small-range path; above that equal-distance order
 *           among the first k rows is unspecified)
 *  Heap   : query_heap :131-156 (also queryParallel :1264-1304)
 *  HeapEA : query_heap_early_abandon :158-197
 *  SortEA : query_sort_early_abandon :80-129
 * data/queries: row-major [n x w] uint64 words.  Requires n >= k. */
void orc_bve_query(const uint64_t *data, long n, int w, const uint64_t *queries, int nq, int k, int method,
                   int *idx, uint32_t *dist) {
  id_pair *pairs = (id_pair *)malloc(sizeof(id_pair) * (size_t)(k + 2));
  for (int qi = 0; qi < nq; qi++) {
    const uint64_t *q = queries + (size_t)qi * w;
    long np = 0;
    if (method == 1) {
      for (int i = 0; i < k; i++) {           /* stable insertion sort of the first k */
        id_pair v = { i, orc_hamming_dist(q, data + (size_t)i * w, w) };
        long j = i;
        while (j > 0 && pairs[j - 1].dist > v.dist) { pairs[j] = pairs[j - 1]; j--; }
        pairs[j] = v;
      }
      for (long i = k; i < n; i++) {
        id_pair v = { (int)i, orc_hamming_dist(q, data + (size_t)i * w, w) };
        maybe_insert(pairs, k, v);
      }
      np = k;
    } else if (method == 0) {
      for (long i = 0; i < n; i++) {
        id_pair v = { (int)i, orc_hamming_dist(q, data + (size_t)i * w, w) };
        pairs[np++] = v;
        std_push_heap(pairs, np - 1, 0, v);
        if (i + 1 > k) { std_pop_heap(pairs, np); np--; }
      }
      std_sort_heap(pairs, np);
    } else if (method == 2) {
      uint32_t bsfK = 0;
      for (long i = 0; i < n; i++) {
        if (i < k) {
          id_pair v = { (int)i, orc_hamming_dist(q, data + (size_t)i * w, w) };
          pairs[np++] = v;
          std_push_heap(pairs, np - 1, 0, v);
          if (v.dist > bsfK) bsfK = v.dist;
        } else {
          uint32_t d = hamming_ea(q, data + (size_t)i * w, w, bsfK);
          if (d < bsfK) {
            id_pair v = { (int)i, d };
            pairs[np++] = v;
            std_push_heap(pairs, np - 1, 0, v);
            std_pop_heap(pairs, np); np--;
            bsfK = pairs[0].dist;
          }
        }
      }
      std_sort_heap(pairs, np);
    } else {
      /* insertionSort lambda :89-101: scan from idxStart-1 down to 1, stop at the first
       * position whose left neighbour is strictly smaller; insert there */
      uint32_t bsfK = 0;
      long i = 0;
      for (; i < k; i++) {
        uint32_t d = orc_hamming_dist(q, data + (size_t)i * w, w);
        if (d > bsfK) bsfK = d;
        long pos = 0;
        if (np > 0) { pos = i - 1; for (; pos > 0; pos--) if (d > pairs[pos - 1].dist) break; if (pos < 0) pos = 0; }
        memmove(pairs + pos + 1, pairs + pos, sizeof(id_pair) * (size_t)(np - pos));
        pairs[pos].idx = (int)i; pairs[pos].dist = d; np++;
      }
      for (; i < n; i++) {
        uint32_t d = hamming_ea(q, data + (size_t)i * w, w, bsfK);
        if (d < bsfK) {
          long pos = k - 1;
          for (; pos > 0; pos--) if (d > pairs[pos - 1].dist) break;
          memmove(pairs + pos + 1, pairs + pos, sizeof(id_pair) * (size_t)(np - pos));
          pairs[pos].idx = (int)i; pairs[pos].dist = d; np++;
          np--;                                  /* pairs.pop_back() */
          bsfK = pairs[k - 1].dist;
        }
      }
    }
    for (int j = 0; j < k; j++) {
      idx[(size_t)qi * k + j] = j < np ? pairs[j].idx : -1;
      dist[(size_t)qi * k + j] = j < np ? pairs[j].dist : 0xFFFFFFFFu;
    }
  }
  free(pairs);
}

/* ------------------------------------------------------------------ canonical order */
/* The selection rule the CUDA path implements: the k lexicographically smallest
 * (distance, id) pairs.  For float distances it returns the same set as
 * searchHeap whenever distances are distinct; rows tied with the k-th distance
 * are where the reference is heap-shape dependent and this rule
 * keeps the lowest ids.  Used by the tests to canonicalise both sides. */
void orc_topk_lex_f32(const float *d, long n, int id_base, int k, int *ids, float *dis) {
  fi_pair *best = (fi_pair *)malloc(sizeof(fi_pair) * (size_t)(k + 1));
  int nb = 0;
  for (long i = 0; i < n; i++) {
    fi_pair v = { d[i], (int)i + id_base };
    if (nb == k && cmp_fi(&v, &best[k - 1]) >= 0) continue;
    int j = nb < k ? nb++ : k - 1;
    while (j > 0 && cmp_fi(&v, &best[j - 1]) < 0) { best[j] = best[j - 1]; j--; }
    best[j] = v;
  }
  for (int j = 0; j < k; j++) {
    ids[j] = j < nb ? best[j].i : -1;
    dis[j] = j < nb ? best[j].d : FLT_MAX;
  }
  free(best);
}

/* full ADC distance of every row (no top-k): dist_i as in searchHeap :1741-1748 */
void orc_adc_all(int M, const int *lut_off, const float *lut, const uint16_t *codes, long N, float *out) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < N; i++) {
    const uint16_t *c = codes + (size_t)i * M;
    float dist = 0.f;
    for (int col = 0; col < M; col += 4) {
      float dism = lut[lut_off[col] + c[col]];
      dism += lut[lut_off[col + 1] + c[col + 1]];
      dism += lut[lut_off[col + 2] + c[col + 2]];
      dism += lut[lut_off[col + 3] + c[col + 3]];
      dist += dism;
    }
    out[i] = dist;
  }
}

int orc_nproc(void) {
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}
