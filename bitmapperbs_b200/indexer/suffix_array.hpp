// Suffix-array construction for small-alphabet texts (codes 0..2 or 0..3).
//
// Data-prep tooling (CPU): BitMapperBS builds its index on the CPU with an
// external suffix sorter (psascan, bwt.cpp:1013-1059) and that stays on the CPU
// (BASELINE.json north_star).  This is an independent sorter with the same
// contract -- SA of the text with an implicit end-of-text sentinel smaller
// than every symbol -- written for many-core hosts: suffixes are bucketed on
// their first 12 symbols, every bucket is sorted on a 32-symbol packed key,
// and ties are resolved by comparing further 32-symbol windows.  Suitable for
// genomes whose repeats are not megabase-long exact copies.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace bmbs {

class PackedText {
 public:
  // codes[i] in [0,3); stored as code+1 in 2 bits so that the zero padding
  // beyond the end sorts before every real symbol.
  PackedText(const uint8_t* codes, uint64_t n) : n_(n), w_((n + 31) / 32 + 3, 0) {
    for (uint64_t i = 0; i < n; ++i)
      w_[i >> 5] |= (uint64_t)(codes[i] + 1) << (62 - 2 * (i & 31));
  }
  // 32 symbols starting at position p (p may be >= n: zeros).
  inline uint64_t window(uint64_t p) const {
    if (p >= n_) return 0;
    uint64_t i = p >> 5, s = 2 * (p & 31);
    uint64_t a = w_[i];
    return s ? (a << s) | (w_[i + 1] >> (64 - s)) : a;
  }
  uint64_t size() const { return n_; }
 private:
  uint64_t n_;
  std::vector<uint64_t> w_;
};

// Returns SA[0..n): start positions of the suffixes in increasing order.
template <typename IdxT>
std::vector<IdxT> build_suffix_array(const uint8_t* codes, uint64_t n, int n_threads = 0) {
  if (n_threads <= 0) n_threads = std::max(1u, std::thread::hardware_concurrency());
  PackedText T(codes, n);
  constexpr int K = 12;                       // bucket prefix length (symbols)
  constexpr uint64_t NB = 1ull << (2 * K);
  std::vector<uint64_t> start(NB + 1, 0);
  auto key = [&](uint64_t p) { return T.window(p) >> (64 - 2 * K); };
  for (uint64_t p = 0; p < n; ++p) ++start[key(p) + 1];
  for (uint64_t b = 0; b < NB; ++b) start[b + 1] += start[b];
  std::vector<IdxT> sa(n);
  {
    std::vector<uint64_t> fill(start.begin(), start.end() - 1);
    for (uint64_t p = 0; p < n; ++p) sa[fill[key(p)]++] = (IdxT)p;
  }
  std::atomic<uint64_t> next{0};
  auto worker = [&]() {
    struct Item { uint64_t k; IdxT p; };
    std::vector<Item> items;
    constexpr uint64_t CHUNK = 256;
    for (;;) {
      uint64_t b0 = next.fetch_add(CHUNK);
      if (b0 >= NB) break;
      for (uint64_t b = b0; b < std::min(NB, b0 + CHUNK); ++b) {
        uint64_t lo = start[b], hi = start[b + 1];
        if (hi - lo < 2) continue;
        items.resize(hi - lo);
        for (uint64_t i = lo; i < hi; ++i) items[i - lo] = {T.window((uint64_t)sa[i] + K), sa[i]};
        std::sort(items.begin(), items.end(), [&](const Item& a, const Item& c) {
          if (a.k != c.k) return a.k < c.k;
          uint64_t pa = (uint64_t)a.p + K + 32, pc = (uint64_t)c.p + K + 32;
          for (;;) {
            // equal windows that ran past the end mean both suffixes ended:
            // impossible for distinct positions, so the loop terminates.
            uint64_t wa = T.window(pa), wc = T.window(pc);
            if (wa != wc) return wa < wc;
            if (pa >= n || pc >= n) return pa > pc;  // shorter suffix first
            pa += 32; pc += 32;
          }
        });
        for (uint64_t i = lo; i < hi; ++i) sa[i] = items[i - lo].p;
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t) th.emplace_back(worker);
  for (auto& t : th) t.join();
  return sa;
}

}  // namespace bmbs
