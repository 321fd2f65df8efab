// C entry point around the UNMODIFIED reference matcher (okvis_matcher/src/{DenseMatcher,MatchingAlgorithm,ThreadPool}.cpp,
// compiled where they lie under /root/reference by oracle/Makefile.ref into oracle/_ref/libokvis_matcher_ref.so).
// Test infrastructure only: tests/test_oracle_vs_reference_matcher.py checks the oracle's restatement of
// DenseMatcher::match against it.  The matching algorithm is the dense-matrix one of the reference's own test
// (okvis_matcher/test/testMatcher.cpp:46-67): distance(a, b) = D[a][b].
#include <cstdint>
#include <vector>

#include <okvis/DenseMatcher.hpp>

namespace {
class MatrixAlgorithm : public okvis::MatchingAlgorithm {
 public:
  const float* D; int nA, nB; const uint8_t *sA, *sB; float thr, ratio;
  std::vector<int>* out_a; std::vector<float>* out_d;
  size_t sizeA() const override { return (size_t)nA; }
  size_t sizeB() const override { return (size_t)nB; }
  float distanceThreshold() const override { return thr; }
  float distanceRatioThreshold() const override { return ratio; }
  bool skipA(size_t a) const override { return sA && sA[a]; }
  bool skipB(size_t b) const override { return sB && sB[b]; }
  float distance(size_t a, size_t b) const override { return D[a * (size_t)nB + b]; }
  void reserveMatches(size_t) override {}
  void setBestMatch(size_t a, size_t b, double d) override { (*out_a)[b] = (int)a; (*out_d)[b] = (float)d; }
};

// Hamming distance between descriptor lists, the way VioKeyframeWindowMatchingAlgorithm::specificDescriptorDistance does
// (popcount of the XOR over the 48 bytes; brisk::Hamming::PopcntofXORed(a, b, 3) in the reference, which is not in the tree)
class HammingAlgorithm : public okvis::MatchingAlgorithm {
 public:
  const uint8_t *A, *B; int nA, nB, bytes; float thr;
  std::vector<int>* out_a; std::vector<float>* out_d;
  size_t sizeA() const override { return (size_t)nA; }
  size_t sizeB() const override { return (size_t)nB; }
  float distanceThreshold() const override { return thr; }
  float distanceRatioThreshold() const override { return 3.0f; }
  float distance(size_t a, size_t b) const override {
    const uint64_t* pa = reinterpret_cast<const uint64_t*>(A + a * (size_t)bytes);
    const uint64_t* pb = reinterpret_cast<const uint64_t*>(B + b * (size_t)bytes);
    int d = 0;
    for (int i = 0; i < bytes / 8; ++i) d += __builtin_popcountll(pa[i] ^ pb[i]);
    return (float)d;
  }
  void reserveMatches(size_t) override {}
  void setBestMatch(size_t a, size_t b, double d) override { (*out_a)[b] = (int)a; (*out_d)[b] = (float)d; }
};
}  // namespace

// DenseMatcher::match over two descriptor lists (bytes a multiple of 8, rows 8-byte aligned) with `num_threads` matcher
// threads (the reference's frontend uses 4, Frontend.cpp:80).  Same outputs as okr_match.
extern "C" int okr_match_hamming(const uint8_t* A, int nA, const uint8_t* B, int nB, int bytes, float threshold, int num_best,
                                 int num_threads, int* out_a, float* out_d) {
  std::vector<int> a((size_t)nB, -1);
  std::vector<float> d((size_t)nB, 0.0f);
  HammingAlgorithm algo;
  algo.A = A; algo.B = B; algo.nA = nA; algo.nB = nB; algo.bytes = bytes; algo.thr = threshold; algo.out_a = &a; algo.out_d = &d;
  okvis::DenseMatcher matcher((unsigned char)num_threads, (unsigned char)num_best, false);
  matcher.match<HammingAlgorithm>(algo);
  int n = 0;
  for (int b = 0; b < nB; ++b) { out_a[b] = a[b]; out_d[b] = d[b]; n += a[b] >= 0; }
  return n;
}

// out_a[b] = index of the A element matched to b (-1: none), out_d[b] its distance.  Returns the number of matches.
extern "C" int okr_match(const float* D, int nA, int nB, const uint8_t* skipA, const uint8_t* skipB, float threshold, int num_best,
                         int use_ratio, float ratio_threshold, int num_threads, int* out_a, float* out_d) {
  std::vector<int> a((size_t)nB, -1);
  std::vector<float> d((size_t)nB, 0.0f);
  MatrixAlgorithm algo;
  algo.D = D; algo.nA = nA; algo.nB = nB; algo.sA = skipA; algo.sB = skipB; algo.thr = threshold; algo.ratio = ratio_threshold;
  algo.out_a = &a; algo.out_d = &d;
  okvis::DenseMatcher matcher((unsigned char)num_threads, (unsigned char)num_best, use_ratio != 0);
  matcher.match<MatrixAlgorithm>(algo);
  int n = 0;
  for (int b = 0; b < nB; ++b) { out_a[b] = a[b]; out_d[b] = d[b]; n += a[b] >= 0; }
  return n;
}
