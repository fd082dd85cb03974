// Exercises include/vaq_gpu.hpp the way a reference caller would (examples/demo_vaq.cpp:339,
// test/test-bitvecengine.cpp:64-79,165-179,246-260).  Input/outputs are flat binary files written/read by
// tests/test_cpp_shim.py, which checks the results against the oracle.
//   usage: shim_check <in.bin> <out.bin>
// exit codes: 0 ok, 3 no usable GPU (std::runtime_error from the shim), 4 known-answer mismatch
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "vaq_gpu.hpp"

template <class T>
static std::vector<T> rd(FILE *f, size_t n) {
  std::vector<T> v(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
  return v;
}

// shim_check files <centroids.bin> <codebook.bin> <queries.fvecs> <k> <out.bin>: an index written in the reference's
// on-disk formats (utils/IO.hpp:736-772) + fvecs queries, loaded by the shim's own readers
static int files_mode(char **argv) {
  vaqgpu::VAQ vaq;
  vaq.parseMethodString("VAQ64m8min5max10var1,EA");
  vaq.loadIndexFiles(argv[2], argv[3]);
  int dim = 0; long nq = 0;
  std::vector<float> q = vaqgpu::io::readFVecs(argv[4], dim, nq);
  const int k = atoi(argv[5]);
  vaqgpu::LabelDistVecF ans = vaq.search(q.data(), (int)nq, k);
  // bit-vector CSV helpers round-trip
  vaqgpu::bitvectors bv = {vaqgpu::io::createBitV(128, 0x8000000000000001ull), {0x0123456789ABCDEFull, 0xFEDCBA9876543210ull}};
  const std::string csv = std::string(argv[6]) + ".csv";
  vaqgpu::io::writeBitVectorsCSV(csv, bv, 128);
  vaqgpu::bitvectors back;
  vaqgpu::io::readBitVectorsCSV(csv, back, 128);
  if (back != bv) { fprintf(stderr, "bit-vector CSV round trip failed\n"); return 4; }
  FILE *o = fopen(argv[6], "wb");
  fwrite(ans.labels.data(), sizeof(int), ans.labels.size(), o);
  fwrite(ans.distances.data(), sizeof(float), ans.distances.size(), o);
  fclose(o);
  printf("shim_check files ok\n");
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: shim_check in.bin out.bin | shim_check files centroids.bin codebook.bin queries.fvecs k out.bin\n"); return 2; }
  try {
    if (std::string(argv[1]) == "files") {
      if (argc < 7) return 2;
      return files_mode(argv);
    }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror("open"); return 2; }
    auto hdr = rd<int32_t>(f, 6);            // L, M, nq, k, has_eig, n
    const int L = hdr[0], M = hdr[1], nq = hdr[2], k = hdr[3], has_eig = hdr[4], n = hdr[5], D = L * M;
    auto bits = rd<int32_t>(f, M);
    size_t cent = 0;
    for (int s = 0; s < M; s++) cent += ((size_t)1 << bits[s]) * L;
    auto centroids = rd<float>(f, cent);
    auto eig = rd<float>(f, has_eig ? (size_t)D * D : 0);
    auto codes = rd<uint16_t>(f, (size_t)n * M);
    auto queries = rd<float>(f, (size_t)nq * D);
    fclose(f);

    vaqgpu::VAQ vaq;
    vaq.parseMethodString("VAQ64m8min5max10var1,EA");
    if (vaq.mBitBudget != 64 || vaq.mSubspaceNum != 8 || vaq.mMinBitsPerSubs != 5 || vaq.mMaxBitsPerSubs != 10 || !(vaq.mMethods & vaqgpu::VAQ::EA)) return 4;
    vaq.loadModel(L, M, bits.data(), centroids.data(), has_eig ? eig.data() : nullptr);
    vaq.setCodebook(codes.data(), n);
    vaqgpu::LabelDistVecF ea = vaq.search(queries.data(), nq, k);
    vaq.mMethods = vaqgpu::VAQ::Heap;
    vaqgpu::LabelDistVecF heap = vaq.search(queries.data(), nq, k);

    // BitVecEngine known answers (reference test/test-bitvecengine.cpp)
    int kat_ok = 1;
    {
      vaqgpu::BitVecEngine e(32);
      vaqgpu::bitvectors bv = {{0x6B8B4567ull}, {0x643C9869ull}, {0xFFFFFFF0ull}, {0xF0000000ull}, {0x0000000Full}};
      e.loadBitV(bv);
      auto r = e.query({bv[1]}, 3);
      kat_ok &= r.size() == 1 && r[0].size() == 3 && r[0][0].idx == 1 && r[0][1].idx == 3 && r[0][2].idx == 4;
      auto p1 = e.queryParallel({bv[0], bv[3]}, 2, 1), p2 = e.queryParallel({bv[0], bv[3]}, 2, 2);
      kat_ok &= p1[0][0].idx == p2[0][0].idx && p1[1][1].idx == p2[1][1].idx && p1[1][1].dist == p2[1][1].dist;
    }
    {
      vaqgpu::BitVecEngine e(64);
      vaqgpu::bitvectors bv = {{0x327B23C66B8B4567ull}, {0x19495CFF74B0DC51ull}, {0xFFFFFFF0FFFFFFFFull}, {0x00000000F0000000ull}, {0x000000000000000Full}};
      e.loadBitV(bv);
      auto r = e.query({bv[1]}, 3);
      kat_ok &= r[0][0].idx == 1 && r[0][1].idx == 3 && r[0][2].idx == 2;
    }
    {
      vaqgpu::BitVecEngine e(1);
      vaqgpu::bitvectors bv = {{1}, {1}};
      e.loadBitV(bv);
      e.appendBitV({{1}});
      e.appendBitV({{0}, {0}});
      auto r = e.query({bv[1]}, 3);
      kat_ok &= e.size() == 5 && r[0][0].idx == 0 && r[0][1].idx == 1 && r[0][2].idx == 2;
    }
    if (!kat_ok) { fprintf(stderr, "BitVecEngine known-answer mismatch\n"); return 4; }

    // the sharded classes over as many GPUs as the box has (at most 2 here): same answers as the single index, bit for bit
    {
      int n_dev = 0;
      vaqgpu::check(vaqgpu_device_count(&n_dev));
      const int G = n_dev >= 2 ? 2 : 1;
      vaqgpu::ShardedVAQ sv(G, n);
      sv.mMethods = vaqgpu::VAQ::EA;
      sv.loadModel(L, M, bits.data(), centroids.data(), has_eig ? eig.data() : nullptr);
      sv.setCodebook(codes.data(), n / 3);                                   // pieces that straddle the shard boundary
      sv.setCodebook(codes.data() + (size_t)(n / 3) * M, n - n / 3);
      vaqgpu::LabelDistVecF sa = sv.search(queries.data(), nq, k);
      if (sv.numShards() != G || sa.labels != ea.labels || sa.distances != ea.distances) { fprintf(stderr, "ShardedVAQ differs from VAQ\n"); return 4; }
      vaqgpu::ShardedBitVecEngine se(64, G, 5);
      vaqgpu::bitvectors bv = {{0x327B23C66B8B4567ull}, {0x19495CFF74B0DC51ull}, {0xFFFFFFF0FFFFFFFFull}, {0x00000000F0000000ull}, {0x000000000000000Full}};
      se.appendBitV({bv[0], bv[1]});
      se.appendBitV({bv[2], bv[3], bv[4]});
      auto r = se.query({bv[1]}, 3);
      if (!(r.size() == 1 && r[0].size() == 3 && r[0][0].idx == 1 && r[0][1].idx == 3 && r[0][2].idx == 2)) { fprintf(stderr, "ShardedBitVecEngine known-answer mismatch\n"); return 4; }
    }

    FILE *o = fopen(argv[2], "wb");
    fwrite(ea.labels.data(), sizeof(int), ea.labels.size(), o);
    fwrite(ea.distances.data(), sizeof(float), ea.distances.size(), o);
    fwrite(heap.labels.data(), sizeof(int), heap.labels.size(), o);
    fwrite(heap.distances.data(), sizeof(float), heap.distances.size(), o);
    fclose(o);
    printf("shim_check ok\n");
    return 0;
  } catch (const std::exception &e) {
    fprintf(stderr, "%s\n", e.what());
    return 3;
  }
}
