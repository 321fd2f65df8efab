// Host-side packing helpers of okb_window_upload (plain C++, no CUDA): the internal landmark order.
// Kept in a header so that tests/hostcheck can exercise it without a GPU.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace okb {

// Sort key of a landmark with frame-visibility mask m: first observing frame * 32 + last observing frame;
// landmarks without observations sort last.
inline uint32_t frame_range_key(uint32_t m) {
  if (!m) return 32u * 32u;
  return (uint32_t)__builtin_ctz(m) * 32u + (31u - (uint32_t)__builtin_clz(m));
}

// Internal landmark order: stable counting sort by frame_range_key.  vis[l] = frame-visibility mask of the caller's
// landmark l.  Outputs perm[j] = caller's index of internal landmark j, inv[l] = internal index of caller's l, and
// tile_range[t] = first | last << 8 over the observed landmarks of internal tile t (32 landmarks); first > last
// (value 1) marks a tile without observations.
inline void sort_landmarks_by_frame_range(const uint32_t* vis, int L, uint32_t* perm, uint32_t* inv, uint32_t* tile_range) {
  std::vector<uint32_t> count(32 * 32 + 2, 0u);
  for (int l = 0; l < L; ++l) ++count[frame_range_key(vis[l]) + 1];
  for (size_t k = 1; k < count.size(); ++k) count[k] += count[k - 1];
  for (int l = 0; l < L; ++l) {
    const uint32_t j = count[frame_range_key(vis[l])]++;
    perm[j] = (uint32_t)l;
    inv[l] = j;
  }
  const int n_tiles = (L + 31) / 32;
  for (int t = 0; t < n_tiles; ++t) tile_range[t] = 1u;
  for (int j = 0; j < L; ++j) {
    const uint32_t m = vis[perm[j]];
    if (!m) continue;
    const uint32_t fi = (uint32_t)__builtin_ctz(m), la = 31u - (uint32_t)__builtin_clz(m);
    uint32_t& tr = tile_range[j >> 5];
    const uint32_t a = tr & 0xffu, b = tr >> 8;
    tr = (a > b) ? (fi | (la << 8)) : (std::min(a, fi) | (std::max(b, la) << 8));
  }
}

}  // namespace okb
