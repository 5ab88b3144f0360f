// Test driver (tests/test_fuzz.py): the host half of the library (csrc/model_host.cc — dnn.bin parser and packer, blob validation,
// aligner, feature-file reader, Kaldi nnet1 importer) compiled with AddressSanitizer and UBSan, run over a directory of mutated
// files.  Every call must come back with a status code; any out-of-bounds access, overflow or leak of the sanitizers' kind ends the
// process with a report.  usage: host_fuzz_driver <dir> <count>   (files net_i.bin, feat_i.bin, nnet_i.txt, trans_i.txt)
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "../include/fdnn.h"
#include "../fast-dnn_b200/csrc/fdnn_internal.h"

using namespace fdnn;

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  const std::string dir = argv[1];
  const int n = std::atoi(argv[2]);
  std::mt19937 rng(7);
  long accepted[4] = {0, 0, 0, 0};
  for (int i = 0; i < n; ++i) {
    const std::string net = dir + "/net_" + std::to_string(i) + ".bin";
    std::vector<uint8_t> blob;
    if (pack_model(net.c_str(), 3.0f, blob) == FDNN_OK) {
      ++accepted[0];
      if (validate_blob(blob.data(), blob.size()) != FDNN_OK) {
        std::printf("a freshly packed blob fails validation: %s\n", net.c_str());
        return 1;
      }
      // what fdnn_load_blob does with bytes it did not pack itself: corrupt the index sections, truncate, validate
      for (int t = 0; t < 20; ++t) {
        std::vector<uint8_t> b2 = blob;
        const size_t lim = std::min<size_t>(b2.size(), 4096);
        for (int k = 0, e = 1 + int(rng() % 6); k < e; ++k) b2[rng() % lim] = uint8_t(rng());
        if (rng() % 4 == 0) b2.resize(rng() % b2.size());
        validate_blob(b2.data(), b2.size());
      }
    }
    if (align_dnn_bin(net.c_str(), (dir + "/aligned.bin").c_str(), 4, 16) == FDNN_OK) ++accepted[1];
    int frames = 0, dim = 0;
    std::vector<float> data;
    if (read_feature_bin((dir + "/feat_" + std::to_string(i) + ".bin").c_str(), &frames, &dim, data) == FDNN_OK) {
      ++accepted[2];
      if (size_t(frames) * size_t(dim) != data.size()) {
        std::printf("feature reader: %d x %d but %zu values\n", frames, dim, data.size());
        return 1;
      }
    }
    if (import_kaldi_nnet1((dir + "/nnet_" + std::to_string(i) + ".txt").c_str(), (dir + "/trans_" + std::to_string(i) + ".txt").c_str(),
                           (dir + "/imported.bin").c_str()) == FDNN_OK)
      ++accepted[3];
  }
  std::printf("accepted: pack %ld align %ld feat %ld kaldi %ld of %d\n", accepted[0], accepted[1], accepted[2], accepted[3], n);
  return 0;
}
