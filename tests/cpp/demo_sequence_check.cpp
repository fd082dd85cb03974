// The query phase of the reference's demo (examples/demo_vaq.cpp:306-345) compiled UNCHANGED against
// include/vaq_gpu.hpp: `vaq` is a vaqgpu::VAQ that adopted the trained state of a reference `VAQ` object, `queries` /
// `datasetrefine` are the reference's RowMatrixXf, `args` is the reference's ArgsParse (utils/Experiment.hpp:96) and
// the results land in the reference's LabelDistVecF (utils/Types.hpp:98-104).
// Built by tests/test_cpp_shim.py with -I<reference>/bitvecengine -I<reference>/external/eigen (skipped when the
// reference tree is not mounted); no reference source is copied here.
//   usage: demo_sequence_check <in.bin> <out.bin> --k K --refine R1,R2
// exit codes: 0 ok, 3 no usable GPU (std::runtime_error from the shim)
#include <iostream>
#include <fstream>
#include <sstream>
#include <vector>
#include <getopt.h>

#include <Eigen/Core>
#include <Eigen/Eigenvalues>

#include "VAQ.hpp"                 // the reference class: here only the holder of the trained state
#include "utils/Experiment.hpp"    // ArgsParse
#include "vaq_gpu.hpp"             // the replacement for the query phase

template <class T>
static std::vector<T> rd(FILE *f, size_t n) {
  std::vector<T> v(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
  return v;
}

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: demo_sequence_check in.bin out.bin [--k K] [--refine a,b]\n"); return 2; }
  const char *in_path = argv[1], *out_path = argv[2];
  std::vector<ArgsParse::opt> long_options{{"k", 'i', "100"}, {"refine", 's', ""}, {"method", 's', "VAQ64m8min5max10var1,EA"}};
  ArgsParse args = ArgsParse(argc - 2, argv + 2, long_options, "HELP");
  try {
    FILE *f = fopen(in_path, "rb");
    if (!f) { perror("open"); return 2; }
    auto hdr = rd<int32_t>(f, 6);            // L, M, nq, D0 (raw dims), n, has_eig
    const int L = hdr[0], M = hdr[1], nq = hdr[2], D0 = hdr[3], n = hdr[4], has_eig = hdr[5], D = L * M;
    auto bits = rd<int32_t>(f, M);
    // a reference VAQ object in the state train() + encode() leave it in (public members, VAQ.hpp:51-73)
    VAQ ref;
    ref.parseMethodString(args["method"]);
    ref.mSubsLen = L; ref.mHighestSubs = M;
    ref.mBitsAlloc.assign(bits.begin(), bits.end());
    ref.mCentroidsPerSubs.resize(M);
    for (int s = 0; s < M; s++) {
      const int K = 1 << bits[s];
      auto c = rd<float>(f, (size_t)K * L);
      ref.mCentroidsPerSubs[s] = Eigen::Map<RowMatrixXf>(c.data(), K, L);
    }
    if (has_eig) {
      auto e = rd<float>(f, (size_t)D * D);
      ref.mEigenVectors = Eigen::Map<RowMatrixXf>(e.data(), D, D).cast<Eigen::scomplex>();
    }
    auto codes = rd<uint16_t>(f, (size_t)n * M);
    ref.mCodebook = Eigen::Map<CodebookType>(codes.data(), n, M);
    auto qv = rd<float>(f, (size_t)nq * D0);
    auto xv = rd<float>(f, (size_t)n * D0);
    fclose(f);
    RowMatrixXf queries = Eigen::Map<RowMatrixXf>(qv.data(), nq, D0);
    RowMatrixXf datasetrefine = Eigen::Map<RowMatrixXf>(xv.data(), n, D0);

    vaqgpu::VAQ vaq;
    vaq.adopt(ref);

    std::vector<int> refines;
    if (args["refine"] != "") {
      std::stringstream ss(args["refine"]);
      while (ss.good())
      {
        std::string substr;
        getline(ss, substr, ',');
        refines.push_back(std::stoi(substr));
      }
    } else {
      refines.push_back(0);
    }

    FILE *o = fopen(out_path, "wb");
    // ---- examples/demo_vaq.cpp:337-345, unchanged -------------------------------------------------------------------
    for (const int refine: refines) {
      int searchK = refine >= args.at<int>("k") ? refine : args.at<int>("k");
      LabelDistVecF answers = vaq.search(queries, searchK, true);                                                                                                                


      if (refine >= args.at<int>("k")) {
        std::cout << "Refining the answer with Refine = " << refine << std::endl;
        answers = vaq.refine(queries, answers, datasetrefine, args.at<int>("k"));
      }
    // -------------------------------------------------------------------------------------------------------------------
      const int32_t cnt = (int32_t)answers.labels.size();
      fwrite(&cnt, sizeof(cnt), 1, o);
      fwrite(answers.labels.data(), sizeof(int), answers.labels.size(), o);
      fwrite(answers.distances.data(), sizeof(float), answers.distances.size(), o);
    }
    fclose(o);
    printf("demo_sequence_check ok\n");
    return 0;
  } catch (const std::exception &e) {
    fprintf(stderr, "%s\n", e.what());
    return 3;
  }
}
