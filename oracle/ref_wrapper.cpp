/* TEST INFRASTRUCTURE — C interface over the UNMODIFIED reference classes.
 *
 * This translation unit is compiled together with the reference's own sources
 * (where they lie under /root/reference) into oracle/_ref/libvaq_ref.so by
 * oracle/Makefile.  It contains no algorithm: it fills the public data members
 * of `VAQ` (reference bitvecengine/VAQ.hpp:51-91) / `BitVecEngine`
 * (BitVecEngine.hpp:86-106) from plain C arrays and calls the reference's own
 * entry points: VAQ::encode (VAQ.cpp:663), VAQ::search (VAQ.cpp:776),
 * VAQ::refine (VAQ.cpp:849), VAQ::clusterTI (VAQ.cpp:878), VAQ::CreateLUT
 * (VAQ.hpp:128, private — reached with the `#define private public` recipe of
 * SURVEY.md Appendix A), BitVecEngine::query / queryParallel
 * (BitVecEngine.cpp:509, 1264) and hammingDist (DistanceFunctions.hpp:164).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs load
 * the resulting library; the product (vaq_b200/) never does.
 */
#include <algorithm>
#include <array>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <vector>
#include <omp.h>
#include <Eigen/Eigenvalues>
#include <Eigen/StdVector>

#define private public
#include "VAQ.hpp"
#undef private
#include "BitVecEngine.hpp"
#include "utils/IO.hpp"

namespace {
/* The reference prints progress to std::cout (e.g. VAQ.cpp:910,976); silence it. */
struct CoutSilencer {
  std::streambuf *old;
  std::ostringstream sink;
  CoutSilencer() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~CoutSilencer() { std::cout.rdbuf(old); }
};
}  // namespace

extern "C" {

/* ---- VAQ ------------------------------------------------------------- */

/* centroids: concatenation over subspaces of row-major [K_s x L] blocks.
 * eig_real: D x D row-major real part of mEigenVectors, or NULL for identity
 * (then search() consumes already-projected queries exactly: x*1 + 0 terms). */
void *ref_vaq_create(int D, int M, int L, const int *bits, const float *centroids,
                     const float *eig_real, int max_bits, unsigned methods) {
  VAQ *v = new VAQ();
  v->mSubspaceNum = M;
  v->mHighestSubs = M;
  v->mSubsLen = L;
  v->mTotalDim = M * L;
  v->mPercentVarExplained = 1.0f;
  v->mMaxBitsPerSubs = max_bits;
  v->mMinBitsPerSubs = 1;
  v->mMethods = methods;
  v->mBitBudget = 0;
  v->mEigenVectors.resize(D, D);
  for (int i = 0; i < D; i++)
    for (int j = 0; j < D; j++)
      v->mEigenVectors(i, j) = Eigen::scomplex(eig_real ? eig_real[(size_t)i * D + j] : (i == j ? 1.f : 0.f), 0.f);
  v->mBitsAlloc.assign(bits, bits + M);
  v->mCentroidsNum.resize(M);
  v->mCentroidsPerSubs.resize(M);
  v->mCentroidsPerSubsCMajor.resize(M);
  const float *c = centroids;
  for (int s = 0; s < M; s++) {
    const int K = 1 << bits[s];
    v->mBitBudget += bits[s];
    v->mCentroidsNum[s] = K;
    v->mCentroidsPerSubs[s].resize(K, L);
    std::memcpy(v->mCentroidsPerSubs[s].data(), c, sizeof(float) * (size_t)K * L);
    /* same statement as reference VAQ.cpp:655-660 (col-major copy for the AVX LUT) */
    v->mCentroidsPerSubsCMajor[s] = v->mCentroidsPerSubs[s];
    c += (size_t)K * L;
  }
  return v;
}

void ref_vaq_destroy(void *h) { delete static_cast<VAQ *>(h); }

void ref_vaq_set_methods(void *h, unsigned methods) { static_cast<VAQ *>(h)->mMethods = methods; }
void ref_vaq_set_visit(void *h, float visit) { static_cast<VAQ *>(h)->mVisit = visit; }

/* copy N x M uint16 codes straight into mCodebook (VAQ.hpp:72) */
void ref_vaq_set_codes(void *h, const uint16_t *codes, long n) {
  VAQ *v = static_cast<VAQ *>(h);
  v->mCodebook.resize(n, v->mHighestSubs);
  std::memcpy(v->mCodebook.data(), codes, sizeof(uint16_t) * (size_t)n * v->mHighestSubs);
  v->mXTrainRows = (int)n;
  v->mXTrainCols = v->mHighestSubs * v->mSubsLen;
}

/* reference VAQ::encode on already-projected rows */
void ref_vaq_encode(void *h, const float *x_proj, long n, int D, int nthreads) {
  VAQ *v = static_cast<VAQ *>(h);
  Eigen::Map<const RowMatrixXf> X(x_proj, n, D);
  RowMatrixXf Xc = X;
  if (nthreads > 0) omp_set_num_threads(nthreads);
  v->encode(Xc);
}

long ref_vaq_num_codes(void *h) { return static_cast<VAQ *>(h)->mCodebook.rows(); }
void ref_vaq_get_codes(void *h, uint16_t *out) {
  VAQ *v = static_cast<VAQ *>(h);
  std::memcpy(out, v->mCodebook.data(), sizeof(uint16_t) * (size_t)v->mCodebook.rows() * v->mCodebook.cols());
}

/* reference CreateLUT (VAQ.hpp:128-167) with the same maxbit dispatch as
 * VAQ::search (VAQ.cpp:787-798); lut_out is col-major [2^max_bits x M]. */
void ref_vaq_create_lut(void *h, const float *q_proj, float *lut_out) {
  VAQ *v = static_cast<VAQ *>(h);
  const int D = v->mSubsLen * v->mHighestSubs;
  RowVectorXf q = Eigen::Map<const RowVectorXf>(q_proj, D);
  LUTType lut(1 << v->mMaxBitsPerSubs, v->mHighestSubs);
  switch (v->mMaxBitsPerSubs) {
    case 9: v->CreateLUT<9>(q, lut); break;
    case 10: v->CreateLUT<10>(q, lut); break;
    case 11: v->CreateLUT<11>(q, lut); break;
    case 12: v->CreateLUT<12>(q, lut); break;
    case 13: v->CreateLUT<13>(q, lut); break;
    case 14: v->CreateLUT<14>(q, lut); break;
    case 15: v->CreateLUT<15>(q, lut); break;
    default: v->CreateLUT(q, lut); break;
  }
  std::memcpy(lut_out, lut.data(), sizeof(float) * (size_t)lut.rows() * lut.cols());
}

/* reference VAQ::search (VAQ.cpp:776-847).  nthreads<=1: exactly as shipped
 * (serial).  nthreads>1: query slices run concurrently, each through the
 * unmodified search() (it only touches locals, VAQ.cpp:779-784). */
void ref_vaq_search(void *h, const float *queries, int nq, int D, int k, int nthreads,
                    int *labels, float *dists) {
  VAQ *v = static_cast<VAQ *>(h);
  if (nthreads <= 1) {
    RowMatrixXf X = Eigen::Map<const RowMatrixXf>(queries, nq, D);
    LabelDistVecF r = v->search(X, k, false);
    std::memcpy(labels, r.labels.data(), sizeof(int) * (size_t)nq * k);
    std::memcpy(dists, r.distances.data(), sizeof(float) * (size_t)nq * k);
    return;
  }
  const int chunk = (nq + nthreads - 1) / nthreads;
#pragma omp parallel for num_threads(nthreads) schedule(static, 1)
  for (int t = 0; t < nthreads; t++) {
    const int q0 = t * chunk, q1 = std::min(nq, q0 + chunk);
    if (q0 >= q1) continue;
    RowMatrixXf X = Eigen::Map<const RowMatrixXf>(queries + (size_t)q0 * D, q1 - q0, D);
    LabelDistVecF r = v->search(X, k, false);
    std::memcpy(labels + (size_t)q0 * k, r.labels.data(), sizeof(int) * (size_t)(q1 - q0) * k);
    std::memcpy(dists + (size_t)q0 * k, r.distances.data(), sizeof(float) * (size_t)(q1 - q0) * k);
  }
}

/* reference VAQ::refine (VAQ.cpp:849-876) */
void ref_vaq_refine(void *h, const float *queries, int nq, int D, const int *in_labels, int refine_num,
                    const float *xtrain, long n, int k, int *labels, float *dists) {
  VAQ *v = static_cast<VAQ *>(h);
  RowMatrixXf X = Eigen::Map<const RowMatrixXf>(queries, nq, D);
  RowMatrixXf T = Eigen::Map<const RowMatrixXf>(xtrain, n, D);
  LabelDistVecF in;
  in.labels.assign(in_labels, in_labels + (size_t)nq * refine_num);
  in.distances.assign((size_t)nq * refine_num, 0.f);
  LabelDistVecF r = v->refine(X, in, T, k);
  std::memcpy(labels, r.labels.data(), sizeof(int) * (size_t)nq * k);
  std::memcpy(dists, r.distances.data(), sizeof(float) * (size_t)nq * k);
}

/* reference VAQ::clusterTI (VAQ.cpp:878-999): regroups mCodebook by cluster. */
void ref_vaq_cluster_ti(void *h, int n_clusters, int n_segments, int use_kmeans, unsigned seed) {
  VAQ *v = static_cast<VAQ *>(h);
  CoutSilencer quiet;
  v->mTIClusterNum = n_clusters;
  v->mTISegmentNum = n_segments;
  v->mTIVariance = 1.f;
  v->mTIClustersMember.clear();
  srand(seed);
  v->clusterTI(use_kmeans != 0, false);
}

/* Hand the reference a clustering made elsewhere — its own public TI members (VAQ.hpp:77-84) filled the way
 * clusterTI leaves them: centres [C x segdims], start idx / sizes [C], member ids in regrouped row order (each cluster's
 * members sorted far -> near its centre, VAQ.cpp:968-982), codeToCC by ORIGINAL id.  mCodebook must be set to the
 * regrouped rows (ref_vaq_set_codes).  Lets bench.py run the reference's searchTriangleInequality on the clusters the
 * device built. */
void ref_vaq_set_ti(void *h, const float *clusters, int C, int segdims, const int *start_idx, const int *sizes, const int *members,
                    const float *code_to_cc, long n) {
  VAQ *v = static_cast<VAQ *>(h);
  v->mTIClusterNum = C;
  v->mTISegmentNum = segdims / v->mSubsLen;
  v->mTIVariance = 1.f;
  v->mTIClusters = Eigen::Map<const RowMatrixXf>(clusters, C, segdims);
  v->mClusterMembersStartIdx.assign(start_idx, start_idx + C);
  v->mTIClustersMember.assign((size_t)C, std::vector<int>());
  long pos = 0;
  for (int c = 0; c < C; c++) {
    v->mTIClustersMember[(size_t)c].assign(members + pos, members + pos + sizes[c]);
    pos += sizes[c];
  }
  v->mCodeToCCDist.assign(code_to_cc, code_to_cc + n);
}

int ref_vaq_ti_segdims(void *h) {
  VAQ *v = static_cast<VAQ *>(h);
  return v->mTISegmentNum * v->mSubsLen;
}
/* export the TI state: centroids [C x segdims], start idx [C], sizes [C],
 * member ids in regrouped row order [N], codeToCC indexed by ORIGINAL id [N] */
void ref_vaq_get_ti(void *h, float *clusters, int *start_idx, int *sizes, int *members, float *code_to_cc) {
  VAQ *v = static_cast<VAQ *>(h);
  const int C = (int)v->mTIClusters.rows();
  std::memcpy(clusters, v->mTIClusters.data(), sizeof(float) * (size_t)C * v->mTIClusters.cols());
  long pos = 0;
  for (int c = 0; c < C; c++) {
    start_idx[c] = v->mClusterMembersStartIdx[c];
    sizes[c] = (int)v->mTIClustersMember[c].size();
    for (int id : v->mTIClustersMember[c]) members[pos++] = id;
  }
  std::memcpy(code_to_cc, v->mCodeToCCDist.data(), sizeof(float) * v->mCodeToCCDist.size());
}

/* ---- BitVecEngine ------------------------------------------------------ */

static bitvectors to_bitvectors(const uint64_t *words, long n, int w) {
  bitvectors out((size_t)n);
  for (long i = 0; i < n; i++) out[(size_t)i].assign(words + (size_t)i * w, words + (size_t)(i + 1) * w);
  return out;
}

void *ref_bve_create(int nbits, const uint64_t *words, long n) {
  BitVecEngine *e = new BitVecEngine(nbits);
  e->loadBitV(to_bitvectors(words, n, e->actBitVLen));
  return e;
}
void ref_bve_destroy(void *h) { delete static_cast<BitVecEngine *>(h); }

/* method: BitVecEngine::QueryMethod {Heap=0, Sort=1, HeapEarlyAbandon=2, SortEarlyAbandon=3}
 * (BitVecEngine.hpp:82-84); threads>0 selects queryParallel (BitVecEngine.cpp:1264). */
void ref_bve_query(void *h, const uint64_t *qwords, int nq, int k, int method, int threads,
                   int *idx, uint32_t *dist) {
  BitVecEngine *e = static_cast<BitVecEngine *>(h);
  bitvectors q = to_bitvectors(qwords, nq, e->actBitVLen);
  std::vector<std::vector<IdxDistPair>> r =
      threads > 0 ? e->queryParallel(q, k, threads) : e->query(q, k, method);
  for (int i = 0; i < nq; i++)
    for (int j = 0; j < k; j++) {
      const bool ok = j < (int)r[(size_t)i].size();
      idx[(size_t)i * k + j] = ok ? r[(size_t)i][(size_t)j].idx : -1;
      dist[(size_t)i * k + j] = ok ? r[(size_t)i][(size_t)j].dist : 0xFFFFFFFFu;
    }
}

uint32_t ref_hamming_dist(const uint64_t *a, const uint64_t *b, int w) {
  bitv va(a, a + w), vb(b, b + w);
  return hammingDist(va, vb);
}
uint32_t ref_hamming_dist_sub(const uint64_t *a, const uint64_t *b, int w, int sublen, int subidx) {
  bitv va(a, a + w), vb(b, b + w);
  return hammingDistSub(va, vb, sublen, subidx);
}

/* BitVecEngine::generateDummyBitVectors (BitVecEngine.hpp:1433) — glibc rand() pinned by
 * test/test-bitvecengine.cpp:132-134, 213-215 */
void ref_generate_dummy(int nbits, int size, int seed, uint64_t *out) {
  bitvectors bv;
  BitVecEngine::generateDummyBitVectors(nbits, bv, size, seed);
  const int w = actualBitVLen(nbits);
  for (int i = 0; i < size; i++) std::memcpy(out + (size_t)i * w, bv[(size_t)i].data(), sizeof(uint64_t) * w);
}

/* createBitV(N, raw) (BitVector.hpp:46-61) */
void ref_create_bitv(int nbits, uint64_t raw, uint64_t *out) {
  bitv v = createBitV(nbits, raw);
  std::memcpy(out, v.data(), sizeof(uint64_t) * v.size());
}

/* ---- on-disk formats (utils/IO.hpp) — lets the tests check vaq_b200/io.py against files the reference
 * itself writes and reads: saveCentroids/loadCentroids (:736, :522), saveCodebook/loadCodebook (:757, :552),
 * readFVecsFromExternal (:126), readIVecsFromExternal (:334) */
void ref_save_codebook(const char *path, const uint16_t *codes, long n, int M) {
  CodebookType cb = Eigen::Map<const CodebookType>(codes, n, M);
  saveCodebook<CodebookType>(cb, path);
}
long ref_load_codebook(const char *path, uint16_t *out, long cap, int *M) {
  CodebookType cb = loadCodebook<CodebookType>(path);
  *M = (int)cb.cols();
  if ((long)cb.size() <= cap) std::memcpy(out, cb.data(), sizeof(uint16_t) * cb.size());
  return cb.rows();
}
void ref_save_centroids(const char *path, int M, int L, const int *bits, const float *centroids) {
  CentroidsPerSubsType c((size_t)M);
  size_t off = 0;
  for (int s = 0; s < M; s++) {
    const long K = 1L << bits[s];
    c[(size_t)s] = Eigen::Map<const CentroidsMatType>(centroids + off, K, L);
    off += (size_t)K * L;
  }
  saveCentroids(c, path);
}
long ref_load_centroids(const char *path, float *out, long cap, int *M, int *L) {
  CentroidsPerSubsType c = loadCentroids(path);
  *M = (int)c.size();
  *L = c.empty() ? 0 : (int)c[0].cols();
  long n = 0;
  for (const CentroidsMatType &m : c) {
    if (n + (long)m.size() <= cap) std::memcpy(out + n, m.data(), sizeof(float) * m.size());
    n += (long)m.size();
  }
  return n;
}
long ref_read_fvecs(const char *path, int dim, float *out, long max_rows) {
  RowMatrixXf d = RowMatrixXf::Zero(max_rows, dim);
  CoutSilencer quiet;
  readFVecsFromExternal(path, d, dim, (int)max_rows);
  std::memcpy(out, d.data(), sizeof(float) * (size_t)max_rows * dim);
  return max_rows;
}
long ref_read_ivecs(const char *path, int dim, int *out, long cap_rows) {
  std::vector<std::vector<int>> d;
  readIVecsFromExternal(path, d, dim);
  for (size_t i = 0; i < d.size() && (long)i < cap_rows; i++) std::memcpy(out + i * dim, d[i].data(), sizeof(int) * dim);
  return (long)d.size();
}

/* bit-vector CSV: readFromExternal(filepath, bitvectors&, cols, delim) (utils/IO.hpp:363-397) and
 * writeToExternal(filepath, const bitvectors&, N) (:681-704) */
long ref_read_bitv_csv(const char *path, int cols, uint64_t *out, long cap_rows) {
  bitvectors bv;
  readFromExternal(std::string(path), bv, cols, ',');
  const int w = actualBitVLen(cols);
  for (size_t i = 0; i < bv.size() && (long)i < cap_rows; i++) std::memcpy(out + i * w, bv[i].data(), sizeof(uint64_t) * w);
  return (long)bv.size();
}
void ref_write_bitv_csv(const char *path, const uint64_t *words, long n, int nbits) {
  const int w = actualBitVLen(nbits);
  bitvectors bv((size_t)n, bitv((size_t)w, 0));
  for (long i = 0; i < n; i++) std::memcpy(bv[(size_t)i].data(), words + (size_t)i * w, sizeof(uint64_t) * w);
  writeToExternal(std::string(path), bv, nbits);
}

int ref_nproc(void) { return omp_get_num_procs(); }

}  /* extern "C" */
