// TEST INFRASTRUCTURE ONLY -- CPU restatement of okvis::DenseMatcher with SEQUENTIAL semantics
// (A ascending; the reference's 4-thread execution is order dependent under ties):
//   matchBody / doWorkLinearMatching / listBIteration  okvis_matcher/include/okvis/implementation/DenseMatcher.hpp:48-225
//   assignbest                                         okvis_matcher/src/DenseMatcher.cpp:69-110
// PINNED by the reference's two known-answer tests (okvis_matcher/test/testMatcher.cpp:69-155, tests/test_oracle_matcher.py)
// and by the reference ITSELF: okvis_matcher compiled unmodified from /root/reference (oracle/Makefile.ref ->
// oracle/_ref/libokvis_matcher_ref.so), bit-exact on random tie-heavy inputs (tests/test_oracle_vs_reference_matcher.py).
#pragma once
#include <cstdint>
#include <functional>
#include <limits>
#include <vector>

#include "../include/okvis_b200.h"

namespace oko {

struct Match { int a, b; float d; };

inline void assignbest(int indexA, std::vector<okb_pair>& vpairs, const std::vector<std::vector<okb_pair>>& best,
                       int numBest, int startidx) {
  const std::vector<okb_pair>& ai = best[indexA];
  for (int index = startidx; index < numBest && ai[index].index_a != -1; ++index) {
    const int b = ai[index].index_a;
    if (vpairs[b].index_a == -1) {
      vpairs[b].index_a = indexA;
      vpairs[b].distance = ai[index].distance;
      return;
    } else if (ai[index].distance < vpairs[b].distance) {
      const int old = vpairs[b].index_a;
      vpairs[b].index_a = indexA;
      vpairs[b].distance = ai[index].distance;
      assignbest(old, vpairs, best, numBest, 1);
      return;
    }
  }
}

// dist(a,b) is MatchingAlgorithm::distance.  topk (may be null) receives [nA][numBest]; pairs [nB].
inline void dense_match(int nA, int nB, const std::function<float(int, int)>& dist, const uint8_t* skipA,
                        const uint8_t* skipB, float threshold, int numBest, bool useRatio, float ratioThreshold,
                        okb_pair* topk, okb_pair* pairs, std::vector<Match>* matches) {
  std::vector<std::vector<okb_pair>> best(nA);
  std::vector<okb_pair> vpairs(nB, okb_pair{-1, std::numeric_limits<float>::max()});
  const float listThr = useRatio ? std::numeric_limits<float>::max() : threshold;
  for (int a = 0; a < nA; ++a) {
    if (skipA && skipA[a]) continue;
    std::vector<okb_pair>& ai = best[a];
    ai.assign(numBest, okb_pair{-1, listThr});
    for (int b = 0; b < nB; ++b) {
      if (skipB && skipB[b]) continue;
      const float t = dist(a, b);
      if (t < ai[numBest - 1].distance) {
        // std::lower_bound on distance: first element not less than t
        int lb = 0;
        while (lb < numBest && ai[lb].distance < t) ++lb;
        for (int k = numBest - 1; k > lb; --k) ai[k] = ai[k - 1];
        ai[lb] = okb_pair{b, t};
      }
    }
    assignbest(a, vpairs, best, numBest, 0);
  }
  if (topk)
    for (int a = 0; a < nA; ++a)
      for (int k = 0; k < numBest; ++k)
        topk[a * numBest + k] = best[a].empty() ? okb_pair{-1, listThr} : best[a][k];
  if (pairs)
    for (int b = 0; b < nB; ++b) pairs[b] = vpairs[b];
  if (matches) {
    matches->clear();
    for (int b = 0; b < nB; ++b) {
      if (useRatio && vpairs[b].distance < threshold) {
        const std::vector<okb_pair>& bl = best[vpairs[b].index_a];
        if (bl[1].index_a != -1) {
          const float d0 = bl[0].distance, d1 = bl[1].distance;
          if (d0 == 0 || d1 / d0 > ratioThreshold) matches->push_back({vpairs[b].index_a, b, vpairs[b].distance});
        } else {
          matches->push_back({vpairs[b].index_a, b, vpairs[b].distance});
        }
      } else if (vpairs[b].distance < threshold) {
        matches->push_back({vpairs[b].index_a, b, vpairs[b].distance});
      }
    }
  }
}

// brisk::Hamming::PopcntofXORed over desc_bytes (call site VioKeyframeWindowMatchingAlgorithm.hpp:253)
inline uint32_t hamming(const uint8_t* a, const uint8_t* b, int bytes) {
  uint32_t s = 0;
  for (int i = 0; i < bytes; ++i) s += __builtin_popcount((unsigned)(a[i] ^ b[i]));
  return s;
}

}  // namespace oko
