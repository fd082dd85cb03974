// vaq_gpu.hpp — header-only C++ host shim over the C ABI (vaqgpu.h) that mirrors the public surface of
// the reference's two index classes for the query-time path:
//
//   class VAQ           bitvecengine/VAQ.hpp:36-114   (parseMethodString, encode, search, refine, mVisit ...)
//   class BitVecEngine  bitvecengine/BitVecEngine.hpp:86-106, 1026, 1121, 1218  (loadBitV, appendBitV, query, queryParallel)
//
// Same names, argument meaning and result types (LabelDistVecF, IdxDistPair: utils/Types.hpp:42-51,
// 98-104), so a caller such as examples/demo_vaq.cpp:339-345 keeps its code and only changes the type
// it instantiates.  Training stays on the host in the reference (VAQ::train); a trained reference `VAQ`
// hands its public members to loadModel().  Matrices are passed as row-major float pointers — exactly
// RowMatrixXf::data() (utils/Types.hpp:14-18) — so the shim itself does not need Eigen.
//
// With Eigen included first (as every reference caller does, examples/demo_vaq.cpp:11-14) the classes also carry the
// reference's exact Eigen-typed signatures (VAQ.hpp:93-114) —
//     void encode(const RowMatrixXf &XTrain);
//     LabelDistVecF search(const RowMatrixXf &XTest, const int k, bool verbose = false);
//     LabelDistVecF refine(const RowMatrixXf &XTest, const LabelDistVecF &answersIn, const RowMatrixXf &XTrain, const int k);
// — and `adopt(const RefVAQ &)`, which takes the trained state straight out of a reference `VAQ` object (all its
// members are public), so the query phase of examples/demo_vaq.cpp:337-345 compiles unchanged against this header
// (tests/cpp/demo_sequence_check.cpp does exactly that).  If the reference's utils/Types.hpp was included, its
// LabelDistVecF / IdxDistPair are used as the result types, so results flow into the reference's own helpers.
//
// Error behaviour: the reference prints and calls exit(0)/assert(false) (VAQ.cpp:64-78, 1263-1266); the
// shim throws std::runtime_error carrying vaqgpu_last_error().  There is no CPU fallback.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "vaqgpu.h"

#if defined(EIGEN_WORLD_VERSION) && !defined(VAQGPU_NO_EIGEN)
#define VAQGPU_HAVE_EIGEN 1
#endif

namespace vaqgpu {

#ifdef TYPES_HPP_
// the reference's own result types (utils/Types.hpp:42-51, 98-104)
using LabelDistVecF = ::LabelDistVecF;
using IdxDistPair = ::IdxDistPair;
#else
// utils/Types.hpp:98-104
struct LabelDistVecF {
  std::vector<int> labels;
  std::vector<float> distances;
};
// utils/Types.hpp:42-51
struct IdxDistPair {
  int idx;
  uint32_t dist;
};
#endif
using bitv = std::vector<uint64_t>;        // BitVector.hpp:13
using bitvectors = std::vector<bitv>;      // BitVector.hpp:19
#ifdef VAQGPU_HAVE_EIGEN
using RowMatrixXf = Eigen::Matrix<float, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor>;      // utils/Types.hpp:14-16
#endif

// ---- on-disk formats of the reference (bitvecengine/utils/IO.hpp), so a C++ caller can load what the reference wrote ----
namespace io {

inline std::vector<unsigned char> slurp(const std::string &path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path);
  return std::vector<unsigned char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

// fvecs / ivecs / bvecs: per record an int32 dimension followed by that many float32 / int32 / uint8
// (readFVecsFromExternal IO.hpp:126-161, readIVecsFromExternal :334-361, readBVecsFromExternal :198-233 — the byte
// reader fills a float matrix, as here).  Returns row-major [rows x dim].
template <class T, class Out>
inline std::vector<Out> readVecs(const std::string &path, int &dim, long &rows, long max_rows = -1) {
  const std::vector<unsigned char> raw = slurp(path);
  if (raw.size() < 4) { dim = 0; rows = 0; return {}; }
  int32_t d; std::memcpy(&d, raw.data(), 4);
  const size_t rec = 4 + (size_t)d * sizeof(T);
  if (d <= 0 || raw.size() % rec) throw std::runtime_error(path + ": not a well-formed vecs file");
  dim = d; rows = (long)(raw.size() / rec);
  if (max_rows >= 0 && rows > max_rows) rows = max_rows;
  std::vector<Out> out((size_t)rows * d);
  for (long r = 0; r < rows; r++) {
    const unsigned char *p = raw.data() + (size_t)r * rec + 4;
    for (int j = 0; j < d; j++) { T v; std::memcpy(&v, p + (size_t)j * sizeof(T), sizeof(T)); out[(size_t)r * d + j] = (Out)v; }
  }
  return out;
}
inline std::vector<float> readFVecs(const std::string &path, int &dim, long &rows, long max_rows = -1) { return readVecs<float, float>(path, dim, rows, max_rows); }
inline std::vector<int> readIVecs(const std::string &path, int &dim, long &rows, long max_rows = -1) { return readVecs<int32_t, int>(path, dim, rows, max_rows); }
inline std::vector<float> readBVecs(const std::string &path, int &dim, long &rows, long max_rows = -1) { return readVecs<uint8_t, float>(path, dim, rows, max_rows); }

// saveCentroids / loadCentroids (IO.hpp:736-754, 522-549): size_t n; per subspace size_t rows, cols, float32[rows*cols].
// Returns the concatenated row-major blocks; K[s] = rows of block s, L = cols.
inline std::vector<float> loadCentroids(const std::string &path, std::vector<int> &K, int &L) {
  const std::vector<unsigned char> raw = slurp(path);
  size_t off = 0;
  auto rd64 = [&]() { if (off + 8 > raw.size()) throw std::runtime_error(path + ": truncated"); uint64_t v; std::memcpy(&v, raw.data() + off, 8); off += 8; return v; };
  const uint64_t n = rd64();
  std::vector<float> out;
  K.clear(); L = 0;
  for (uint64_t s = 0; s < n; s++) {
    const uint64_t rows = rd64(), cols = rd64();
    if (off + rows * cols * 4 > raw.size()) throw std::runtime_error(path + ": truncated");
    const size_t at = out.size();
    out.resize(at + rows * cols);
    std::memcpy(out.data() + at, raw.data() + off, rows * cols * 4);
    off += rows * cols * 4;
    K.push_back((int)rows); L = (int)cols;
  }
  return out;
}
inline void saveCentroids(const std::string &path, const float *cent, const std::vector<int> &K, int L) {
  std::ofstream f(path, std::ios::binary);
  const uint64_t n = K.size();
  f.write((const char *)&n, 8);
  size_t off = 0;
  for (int k : K) {
    const uint64_t rows = (uint64_t)k, cols = (uint64_t)L;
    f.write((const char *)&rows, 8); f.write((const char *)&cols, 8);
    f.write((const char *)(cent + off), (std::streamsize)(rows * cols * 4));
    off += rows * cols;
  }
}
// saveCodebook / loadCodebook (IO.hpp:757-772, 552-571): size_t rows, cols, uint16[rows*cols] row-major (mCodebook)
inline std::vector<uint16_t> loadCodebook(const std::string &path, size_t &rows, size_t &cols) {
  const std::vector<unsigned char> raw = slurp(path);
  if (raw.size() < 16) throw std::runtime_error(path + ": truncated");
  uint64_t r, c; std::memcpy(&r, raw.data(), 8); std::memcpy(&c, raw.data() + 8, 8);
  if (raw.size() < 16 + r * c * 2) throw std::runtime_error(path + ": truncated codebook");
  std::vector<uint16_t> out((size_t)(r * c));
  std::memcpy(out.data(), raw.data() + 16, (size_t)(r * c * 2));
  rows = (size_t)r; cols = (size_t)c;
  return out;
}
inline void saveCodebook(const std::string &path, const uint16_t *codes, size_t rows, size_t cols) {
  std::ofstream f(path, std::ios::binary);
  const uint64_t r = rows, c = cols;
  f.write((const char *)&r, 8); f.write((const char *)&c, 8);
  f.write((const char *)codes, (std::streamsize)(rows * cols * 2));
}

// actualBitVLen / createBitV (BitVector.hpp:36-76).  The scalar overload shifts a 64-bit value by multiples of 64
// for N > 64 (undefined in C++; x86-64 takes the count modulo 64): restated as the compiled reference behaves.
inline int actualBitVLen(int N) { return (N + 63) / 64; }
inline bitv createBitV(int N, uint64_t raw) {
  bitv v((size_t)actualBitVLen(N), 0);
  const int len = (int)v.size();
  if (N <= 64) { v[0] = raw; return v; }
  for (int i = 0; i < len - 1; i++) v[(size_t)i] = raw;
  const int rest = N - (len - 1) * 64;
  v[(size_t)len - 1] = rest >= 64 ? raw : (raw & ((1ull << rest) - 1ull));
  return v;
}
// bit-vector CSV: readFromExternal(filepath, bitvectors&, cols, delim) (IO.hpp:363-397) / writeToExternal (:681-704).
// Column c is bit 63 - (c % 64) of word c / 64.  With cols % 64 != 0 the reference shifts the last, partial word once
// too often (its first column falls out of the register); reproduced, as in vaq_b200/io.py.
inline void readBitVectorsCSV(const std::string &path, bitvectors &bv, int cols, char delim = ',') {
  std::ifstream in(path);
  std::string line, bit;
  while (std::getline(in, line)) {
    if (line.empty()) break;
    bitv v((size_t)actualBitVLen(cols), 0);
    std::stringstream ss(line);
    int counter = 0;
    uint64_t acc = 0;
    while (std::getline(ss, bit, delim) && counter < cols) {
      acc |= (uint64_t)std::stoi(bit);
      counter++;
      if (counter % 64 == 0) { v[(size_t)(counter / 64) - 1] = acc; acc = 0; }
      else acc <<= 1;
    }
    if (counter % 64 != 0) v[(size_t)(counter / 64)] = acc << (64 - (counter % 64));
    bv.push_back(v);
  }
}
inline void writeBitVectorsCSV(const std::string &path, const bitvectors &bv, int N) {
  std::ofstream out(path);
  for (const bitv &row : bv) {
    for (int i = 0; i < (int)row.size(); i++) {
      const int maxBin = ((i + 1) * 64 <= N) ? 64 : (N % 64);
      for (int b = 0; b < maxBin; b++) {
        out << (int)((row[(size_t)i] >> (63 - b)) & 1u);
        if (b != maxBin - 1) out << ',';
      }
      if (i != (int)row.size() - 1) out << ',';
    }
    out << '\n';
  }
}

}  // namespace io

inline void check(int rc) {
  if (rc != VAQGPU_OK) throw std::runtime_error(std::string("vaqgpu: ") + vaqgpu_last_error());
}

class VAQ {
 public:
  // VAQ::NNMethod, VAQ.hpp:38-49 (only the exact modes are implemented on the device; FAST* quantise the
  // LUT to uint8 and SORT/FAST2/FAST3 write their results through a defective helper)
  enum NNMethod { Sort = 1 << 0, EA = 1 << 1, TI = 1 << 2, Heap = 1 << 7 };

  // knobs with the reference's names (VAQ.hpp:51-55, 84)
  int mBitBudget = 0, mSubspaceNum = 0, mMinBitsPerSubs = 0, mMaxBitsPerSubs = 0;
  float mPercentVarExplained = 1.f;
  int mMethods = Heap;
  float mVisit = 1.f;
  int mTIClusterNum = 0;
  int mSubsLen = 0, mHighestSubs = 0;

  explicit VAQ(int device = 0) : device_(device) {}
  ~VAQ() { vaqgpu_destroy(h_); }
  VAQ(const VAQ &) = delete;
  VAQ &operator=(const VAQ &) = delete;

  // VAQ::parseMethodString, VAQ.cpp:1189-1267: "VAQ<budget>m<M>min<a>max<b>var<v>,<MODE>[_TI<c>]"
  void parseMethodString(const std::string &s) {
    float var = 1.f;
    if (std::sscanf(s.c_str(), "VAQ%dm%dmin%dmax%dvar%f", &mBitBudget, &mSubspaceNum, &mMinBitsPerSubs, &mMaxBitsPerSubs, &var) < 4)
      throw std::runtime_error("parseMethodString: expected VAQ<budget>m<M>min<a>max<b>var<v>,<MODE>");
    mPercentVarExplained = var;
    mMethods = 0;
    const size_t comma = s.find(',');
    const std::string mode = comma == std::string::npos ? "" : s.substr(comma + 1);
    if (mode.find("FAST") != std::string::npos) throw std::runtime_error("FAST* modes are not provided (lossy uint8 LUT)");
    if (mode.find("SORT") != std::string::npos) mMethods |= Sort;
    if (mode.find("HEAP") != std::string::npos) mMethods |= Heap;
    if (mode.find("EA") != std::string::npos) mMethods |= EA;
    const size_t ti = mode.find("TI");
    if (ti != std::string::npos) {
      mMethods |= TI;
      std::sscanf(mode.c_str() + ti, "TI%d", &mTIClusterNum);
    }
    if (!mMethods) mMethods = Heap;
  }

  // Hand over what VAQ::train produced (public members VAQ.hpp:57-73): mSubsLen, mHighestSubs, mBitsAlloc,
  // mCentroidsPerSubs (concatenated row-major [2^bits[s] x L] blocks), real(mEigenVectors) [D x D] or nullptr.
  void loadModel(int subsLen, int highestSubs, const int *bitsAlloc, const float *centroidsPerSubs, const float *eigReal) {
    vaqgpu_destroy(h_);
    h_ = nullptr;
    mSubsLen = subsLen;
    mHighestSubs = highestSubs;
    vaqgpu_model_desc d;
    d.D = subsLen * highestSubs; d.M = highestSubs; d.L = subsLen;
    d.bits = bitsAlloc; d.centroids = centroidsPerSubs; d.eig_real = eigReal;
    check(vaqgpu_create(&d, device_, &h_));
    has_eig_ = eigReal != nullptr;
  }

  // An index the reference saved with saveCentroids + saveCodebook (examples/demo_vaq.cpp:127-140, utils/IO.hpp:736-772):
  // bits follow from the centroid counts; eigReal [D x D] when queries arrive raw.
  void loadIndexFiles(const std::string &centroidsPath, const std::string &codebookPath, const float *eigReal = nullptr) {
    std::vector<int> K;
    int L = 0;
    const std::vector<float> cent = io::loadCentroids(centroidsPath, K, L);
    std::vector<int> bits;
    for (int k : K) {
      int b = 0;
      while ((1 << b) < k) b++;
      if ((1 << b) != k) throw std::runtime_error("loadIndexFiles: centroid count is not a power of two");
      bits.push_back(b);
    }
    loadModel(L, (int)K.size(), bits.data(), cent.data(), eigReal);
    size_t rows = 0, cols = 0;
    const std::vector<uint16_t> codes = io::loadCodebook(codebookPath, rows, cols);
    if ((int)cols != mHighestSubs) throw std::runtime_error("loadIndexFiles: codebook columns != subspaces");
    setCodebook(codes.data(), (int64_t)rows);
  }

  // mCodebook (VAQ.hpp:72): row-major [n x mHighestSubs] uint16, as VAQ::encode left it on the host
  void setCodebook(const uint16_t *codes, int64_t n) { check(vaqgpu_add_codes_u16(need(), codes, n)); }

  // VAQ::encode (VAQ.cpp:663): rows must already be projected (train() projects in place)
  void encode(const float *XTrainProjected, int64_t n) { check(vaqgpu_encode_add(need(), XTrainProjected, n)); }

  // VAQ::clusterTI's outputs (VAQ.hpp:77-84) for the TI / visit mode
  void setClusters(const float *clusters, int C, int segdims, const int64_t *start, const int64_t *size, const int32_t *members) {
    check(vaqgpu_set_clusters(need(), clusters, C, segdims, start, size, members));
  }

  // VAQ::search (VAQ.cpp:776-847): XTest is raw [nq x D] when the model has eigenvectors, else projected
  LabelDistVecF search(const float *XTest, int nq, int k, bool /*verbose*/ = false) {
    LabelDistVecF ret;
    ret.labels.resize((size_t)nq * k);
    ret.distances.resize((size_t)nq * k);
    uint32_t flags = has_eig_ ? 0u : VAQGPU_PROJECTED;
    if (mMethods & TI) {
      check(vaqgpu_set_visit(need(), mVisit));
      flags |= VAQGPU_TI | VAQGPU_EA | VAQGPU_SQRT;        // TI returns sqrt distances and original ids (VAQ.cpp:1585-1607)
    } else if (mMethods & EA) {
      flags |= VAQGPU_EA;
    } else {
      flags |= VAQGPU_HEAP;
    }
    check(vaqgpu_search(need(), XTest, nq, k, flags, ret.labels.data(), ret.distances.data()));
    return ret;
  }

  // VAQ::refine (VAQ.cpp:849-876): XTrain raw rows [n x D0]; answersIn.labels holds refineNum candidates per query
  LabelDistVecF refine(const float *XTest, int nq, const LabelDistVecF &answersIn, const float *XTrain, int64_t n, int D0, int k) {
    if (XTrain != raw_ || n != raw_n_) {
      check(vaqgpu_set_raw_vectors(need(), XTrain, n, D0));
      raw_ = XTrain; raw_n_ = n;
    }
    const int refineNum = (int)(answersIn.labels.size() / (size_t)nq);
    LabelDistVecF ret;
    ret.labels.resize((size_t)nq * k);
    ret.distances.resize((size_t)nq * k);
    check(vaqgpu_refine(h_, XTest, nq, answersIn.labels.data(), refineNum, k, ret.labels.data(), ret.distances.data()));
    return ret;
  }

  vaqgpu_t *handle() { return need(); }

#ifdef VAQGPU_HAVE_EIGEN
  // ---- the reference's Eigen-typed signatures (VAQ.hpp:93-114) ----------------------------------------------------
  // VAQ::encode (VAQ.cpp:663): rows already projected (train() projects its argument in place, VAQ.cpp:294)
  void encode(const RowMatrixXf &XTrain) { encode(XTrain.data(), (int64_t)XTrain.rows()); }

  // VAQ::search (VAQ.cpp:776-847)
  LabelDistVecF search(const RowMatrixXf &XTest, const int k, bool verbose = false) {
    const int D = mSubsLen * mHighestSubs;
    if ((int)XTest.cols() == D) return search(XTest.data(), (int)XTest.rows(), k, verbose);
    if ((int)XTest.cols() > D) throw std::runtime_error("vaqgpu::VAQ::search: XTest has more columns than the model's padded dimensionality");
    RowMatrixXf padded = RowMatrixXf::Zero(XTest.rows(), D);          // the demo pads its matrices the same way (demo_vaq.cpp:74-76,281)
    padded.leftCols(XTest.cols()) = XTest;
    return search(padded.data(), (int)padded.rows(), k, verbose);
  }

  // VAQ::refine (VAQ.cpp:849-876)
  LabelDistVecF refine(const RowMatrixXf &XTest, const LabelDistVecF &answersIn, const RowMatrixXf &XTrain, const int k) {
    if (XTest.cols() != XTrain.cols()) throw std::runtime_error("vaqgpu::VAQ::refine: XTest / XTrain column mismatch");
    return refine(XTest.data(), (int)XTest.rows(), answersIn, XTrain.data(), (int64_t)XTrain.rows(), (int)XTrain.cols(), k);
  }

  // Take over a trained (and, optionally, encoded / clustered) reference `VAQ` — any type with the reference's public
  // members (VAQ.hpp:51-84): mSubsLen, mHighestSubs, mBitsAlloc, mCentroidsPerSubs, mEigenVectors, the method knobs,
  // mCodebook when encode() already ran on the host, and clusterTI()'s outputs when the method string asks for TI.
  template <class RefVAQ>
  void adopt(const RefVAQ &v, bool with_codebook = true) {
    std::vector<float> cent;
    for (int s = 0; s < v.mHighestSubs; s++) {
      const RowMatrixXf c = v.mCentroidsPerSubs[(size_t)s];           // row-major [K_s x L]
      cent.insert(cent.end(), c.data(), c.data() + c.size());
    }
    const RowMatrixXf eig = v.mEigenVectors.real();                   // VAQ.hpp:57, consumed by search() at VAQ.cpp:777
    const int D = v.mSubsLen * v.mHighestSubs;
    loadModel(v.mSubsLen, v.mHighestSubs, v.mBitsAlloc.data(), cent.data(),
              (eig.rows() == D && eig.cols() == D) ? eig.data() : nullptr);
    mBitBudget = v.mBitBudget; mSubspaceNum = v.mSubspaceNum; mMinBitsPerSubs = v.mMinBitsPerSubs; mMaxBitsPerSubs = v.mMaxBitsPerSubs;
    mPercentVarExplained = v.mPercentVarExplained;
    mMethods = (int)v.mMethods; mVisit = v.mVisit; mTIClusterNum = v.mTIClusterNum;
    if (with_codebook && v.mCodebook.rows() > 0) {
      if ((int)v.mCodebook.cols() != v.mHighestSubs) throw std::runtime_error("vaqgpu::VAQ::adopt: mCodebook has the wrong number of columns");
      setCodebook(v.mCodebook.data(), (int64_t)v.mCodebook.rows());  // RowMatrix<uint16_t>, utils/Types.hpp:31
      if ((mMethods & TI) && v.mTIClusters.rows() > 0) {
        const int C = (int)v.mTIClusters.rows();
        std::vector<int64_t> start((size_t)C), size((size_t)C);
        std::vector<int32_t> members;
        for (int c = 0; c < C; c++) {
          start[(size_t)c] = v.mClusterMembersStartIdx[(size_t)c];
          size[(size_t)c] = (int64_t)v.mTIClustersMember[(size_t)c].size();
          members.insert(members.end(), v.mTIClustersMember[(size_t)c].begin(), v.mTIClustersMember[(size_t)c].end());
        }
        const RowMatrixXf cl = v.mTIClusters;
        setClusters(cl.data(), C, (int)cl.cols(), start.data(), size.data(), members.data());
      }
    }
  }
#endif  // VAQGPU_HAVE_EIGEN

 private:
  vaqgpu_t *need() {
    if (!h_) throw std::runtime_error("vaqgpu::VAQ: loadModel() first");
    return h_;
  }
  int device_;
  vaqgpu_t *h_ = nullptr;
  bool has_eig_ = false;
  const float *raw_ = nullptr;
  int64_t raw_n_ = 0;
};

class BitVecEngine {
 public:
  // BitVecEngine::QueryMethod, BitVecEngine.hpp:82-84.  Every method returns the k nearest rows; on the
  // device they all run the same scan and equal-distance rows are ordered by ascending index.
  enum QueryMethod { Heap = 0, Sort = 1, HeapEarlyAbandon = 2, SortEarlyAbandon = 3 };
  const int N;
  const int actBitVLen;

  explicit BitVecEngine(int _N, int device = 0) : N(_N), actBitVLen((_N + 63) / 64) { check(hamgpu_create(_N, device, &h_)); }
  ~BitVecEngine() { hamgpu_destroy(h_); }
  BitVecEngine(const BitVecEngine &) = delete;
  BitVecEngine &operator=(const BitVecEngine &) = delete;

  // BitVecEngine::loadBitV / appendBitV (BitVecEngine.cpp:12, 1630)
  void appendBitV(const bitvectors &bv) {
    std::vector<uint64_t> flat = flatten(bv);
    check(hamgpu_add(h_, flat.data(), (int64_t)bv.size()));
  }
  void loadBitV(const bitvectors &bv) {
    int64_t n = 0;
    check(hamgpu_num_rows(h_, &n));
    if (n != 0) throw std::runtime_error("loadBitV on a non-empty engine: create a new engine (rows are append-only on the device)");
    appendBitV(bv);
  }
  int64_t size() const {
    int64_t n = 0;
    check(hamgpu_num_rows(h_, &n));
    return n;
  }

  // BitVecEngine::query (BitVecEngine.cpp:509-519) / queryParallel (:1264-1304)
  std::vector<std::vector<IdxDistPair>> query(const bitvectors &queries, int k, int /*method*/ = Sort) const {
    const int nq = (int)queries.size();
    std::vector<uint64_t> flat = flatten(queries);
    std::vector<int32_t> idx((size_t)nq * k);
    std::vector<uint32_t> dist((size_t)nq * k);
    check(hamgpu_query(h_, flat.data(), nq, k, idx.data(), dist.data()));
    std::vector<std::vector<IdxDistPair>> out((size_t)nq);
    for (int q = 0; q < nq; q++) {
      for (int j = 0; j < k; j++) {
        if (idx[(size_t)q * k + j] < 0) break;      // fewer than k rows
        out[(size_t)q].push_back(IdxDistPair{idx[(size_t)q * k + j], dist[(size_t)q * k + j]});
      }
    }
    return out;
  }
  std::vector<std::vector<IdxDistPair>> queryParallel(const bitvectors &queries, int k, int /*thread*/) const { return query(queries, k); }

 private:
  std::vector<uint64_t> flatten(const bitvectors &bv) const {
    std::vector<uint64_t> flat(bv.size() * (size_t)actBitVLen, 0);
    for (size_t i = 0; i < bv.size(); i++) {
      if ((int)bv[i].size() != actBitVLen) throw std::runtime_error("bit vector length does not match the engine");
      for (int w = 0; w < actBitVLen; w++) flat[i * (size_t)actBitVLen + (size_t)w] = bv[i][(size_t)w];
    }
    return flat;
  }
  hamgpu_t *h_ = nullptr;
};

// ---- the same two classes over the GPUs of one box (one host process, vaqgpu_sharded_* / hamgpu_sharded_*) ----------
// Rows are appended in order and land in contiguous blocks of ceil(n_rows_total / n_gpus) per device; every
// device scans its block for all queries, the shards exchange their running k-th-best bounds through peer memory, and
// one ncclAllGather + device merge of the shard-local key lists gives the answer of a single index, bit for bit (the
// reference's precedent for merging partial answers: BitVecEngine.cpp:1599-1611).  n_rows_total fixes the block size and
// must be known before the first row arrives.
class ShardedVAQ {
 public:
  int mMethods = VAQ::Heap;                  // VAQ::EA or VAQ::Heap (TI: per-shard handles, vaqgpu_sharded_shard)
  int mSubsLen = 0, mHighestSubs = 0;

  ShardedVAQ(int n_gpus, int64_t n_rows_total, const int *dev_ids = nullptr)
      : n_gpus_(n_gpus), n_total_(n_rows_total), dev_ids_(dev_ids ? std::vector<int>(dev_ids, dev_ids + n_gpus) : std::vector<int>()) {}
  ~ShardedVAQ() { vaqgpu_sharded_destroy(h_); }
  ShardedVAQ(const ShardedVAQ &) = delete;
  ShardedVAQ &operator=(const ShardedVAQ &) = delete;

  // what VAQ::train produced, as in VAQ::loadModel above
  void loadModel(int subsLen, int highestSubs, const int *bitsAlloc, const float *centroidsPerSubs, const float *eigReal) {
    vaqgpu_sharded_destroy(h_);
    h_ = nullptr;
    mSubsLen = subsLen;
    mHighestSubs = highestSubs;
    vaqgpu_model_desc d;
    d.D = subsLen * highestSubs; d.M = highestSubs; d.L = subsLen;
    d.bits = bitsAlloc; d.centroids = centroidsPerSubs; d.eig_real = eigReal;
    check(vaqgpu_sharded_create(&d, n_gpus_, dev_ids_.empty() ? nullptr : dev_ids_.data(), n_total_, &h_));
    has_eig_ = eigReal != nullptr;
  }
  // mCodebook rows [n x mHighestSubs] uint16, appended in arrival order (may be called in pieces)
  void setCodebook(const uint16_t *codes, int64_t n) { check(vaqgpu_sharded_add_codes_u16(need(), codes, n)); }
  // VAQ::encode (VAQ.cpp:663) on the devices: projected rows, appended in arrival order
  void encode(const float *XTrainProjected, int64_t n) { check(vaqgpu_sharded_encode_add(need(), XTrainProjected, n)); }
  int numShards() const {
    int32_t g = 0;
    if (h_) check(vaqgpu_sharded_num_shards(h_, &g));
    return g;
  }
  // VAQ::search (VAQ.cpp:776-847): XTest raw [nq x D] when the model has eigenvectors, else projected
  LabelDistVecF search(const float *XTest, int nq, int k, bool /*verbose*/ = false) {
    LabelDistVecF ret;
    ret.labels.resize((size_t)nq * k);
    ret.distances.resize((size_t)nq * k);
    const uint32_t flags = (has_eig_ ? 0u : VAQGPU_PROJECTED) | ((mMethods & VAQ::EA) ? VAQGPU_EA : VAQGPU_HEAP);
    check(vaqgpu_sharded_search(need(), XTest, nq, k, flags, ret.labels.data(), ret.distances.data()));
    return ret;
  }
#ifdef VAQGPU_HAVE_EIGEN
  void encode(const RowMatrixXf &XTrain) { encode(XTrain.data(), (int64_t)XTrain.rows()); }
  LabelDistVecF search(const RowMatrixXf &XTest, const int k, bool verbose = false) {
    if ((int)XTest.cols() != mSubsLen * mHighestSubs) throw std::runtime_error("vaqgpu::ShardedVAQ::search: XTest must have M*L columns");
    return search(XTest.data(), (int)XTest.rows(), k, verbose);
  }
#endif

 private:
  vaqgpu_sharded_t *need() {
    if (!h_) throw std::runtime_error("vaqgpu::ShardedVAQ: loadModel first");
    return h_;
  }
  int n_gpus_;
  int64_t n_total_;
  std::vector<int> dev_ids_;
  vaqgpu_sharded_t *h_ = nullptr;
  bool has_eig_ = false;
};

class ShardedBitVecEngine {
 public:
  const int N;
  const int actBitVLen;
  ShardedBitVecEngine(int _N, int n_gpus, int64_t n_rows_total, const int *dev_ids = nullptr) : N(_N), actBitVLen((_N + 63) / 64) {
    check(hamgpu_sharded_create(_N, n_gpus, dev_ids, n_rows_total, &h_));
  }
  ~ShardedBitVecEngine() { hamgpu_sharded_destroy(h_); }
  ShardedBitVecEngine(const ShardedBitVecEngine &) = delete;
  ShardedBitVecEngine &operator=(const ShardedBitVecEngine &) = delete;

  // BitVecEngine::loadBitV / appendBitV: rows appended in arrival order
  void appendBitV(const bitvectors &bv) {
    std::vector<uint64_t> flat = flatten(bv);
    check(hamgpu_sharded_add(h_, flat.data(), (int64_t)bv.size()));
  }
  // BitVecEngine::query / queryParallel
  std::vector<std::vector<IdxDistPair>> query(const bitvectors &queries, int k, int /*method*/ = BitVecEngine::Sort) const {
    const int nq = (int)queries.size();
    std::vector<uint64_t> flat = flatten(queries);
    std::vector<int32_t> idx((size_t)nq * k);
    std::vector<uint32_t> dist((size_t)nq * k);
    check(hamgpu_sharded_query(h_, flat.data(), nq, k, idx.data(), dist.data()));
    std::vector<std::vector<IdxDistPair>> out((size_t)nq);
    for (int q = 0; q < nq; q++)
      for (int j = 0; j < k && idx[(size_t)q * k + j] >= 0; j++) out[(size_t)q].push_back(IdxDistPair{idx[(size_t)q * k + j], dist[(size_t)q * k + j]});
    return out;
  }

 private:
  std::vector<uint64_t> flatten(const bitvectors &bv) const {
    std::vector<uint64_t> flat(bv.size() * (size_t)actBitVLen, 0);
    for (size_t i = 0; i < bv.size(); i++) {
      if ((int)bv[i].size() != actBitVLen) throw std::runtime_error("bit vector length does not match the engine");
      for (int w = 0; w < actBitVLen; w++) flat[i * (size_t)actBitVLen + (size_t)w] = bv[i][(size_t)w];
    }
    return flat;
  }
  hamgpu_sharded_t *h_ = nullptr;
};

}  // namespace vaqgpu
