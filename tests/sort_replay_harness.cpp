// Test harness (CPU only): bmbs_sort_replay.h -- the device finishing's step-by-step replay of libstdc++'s std::sort on
// (vote << 16 | position) keys -- against std::sort itself on the reference's 32-byte vote records (Schema.cpp:27612,
// comparator :560-563): same permutation for random, periodic, sparse and ramp vote patterns, lists of 1..5000 windows.
#include <algorithm>
#include <cstdio>
#include <random>
#include <vector>
#include "../bitmapperbs_b200/csrc/bmbs_sort_replay.h"
struct H { uint64_t site, vote; uint32_t err; uint64_t end_site; };
int main() {
  std::mt19937_64 g(1);
  long bad = 0, fb = 0, tot = 0;
  for (int it = 0; it < 200000; ++it) {
    int n = 1 + g() % (it % 50 == 0 ? 5000 : 200);
    int vr = 1 + g() % 28;
    std::vector<H> h(n); std::vector<uint32_t> k(n);
    int mode = g() % 4;
    for (int i = 0; i < n; ++i) {
      uint64_t v = 1 + g() % vr;
      if (mode == 1) v = 1 + (i % vr); if (mode == 2) v = 1 + (g() % 10 == 0 ? g() % vr : 0); if (mode == 3) v = 1 + (uint64_t)(vr * (double)i / n);
      h[i] = {(uint64_t)i, v, 0, 0}; k[i] = (uint32_t)(v << 16 | i);
    }
    std::sort(h.begin(), h.end(), [](const H& a, const H& b) { return a.vote > b.vote; });
    bool ok = bmbs::sort_replay(k.data(), n);
    ++tot;
    if (!ok) { ++fb; continue; }
    for (int i = 0; i < n; ++i) if ((k[i] & 0xFFFF) != h[i].site) { ++bad; break; }
  }
  printf("tests %ld mismatching %ld fallback %ld\n", tot, bad, fb);
  return bad != 0;
}
