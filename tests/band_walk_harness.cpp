// Test harness (CPU only): bmbs_band_walk.h -- the run-to-run walk of the last band column that verify_windows uses on the
// device -- against the oracle's cell-by-cell restatement of BS_Reserve_Banded_BPM (Levenshtein_Cal.h:351-567).  The band
// states come from the column recurrence itself, in 32-bit (k <= 15) and 64-bit (k <= 31) words with the band mask as the
// kernel applies it, over windows with substitutions, indels, N, bisulfite conversions and unrelated sequence.
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "../oracle/oracle_core.hpp"
#include "../bitmapperbs_b200/csrc/bmbs_band_walk.h"

template <typename W>
static int product_band(const char* win, const char* read, int L, int k, uint32_t& err_out) {
  const int band = 2 * k + 1;
  const W mask = band >= (int)(8 * sizeof(W)) ? ~(W)0 : (W)(((W)1 << band) - 1);
  auto cls = [](char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4; };
  W VP = 0, VN = 0; int err = 0;
  for (int j = 0; j < L; ++j) {
    W eq = 0;                                               // window bits [j, j + word) of the read symbol's match plane
    const int c = cls(read[j]);
    for (int i = 0; i < (int)(8 * sizeof(W)) && j + i < L + 2 * k; ++i) {
      const int w = cls(win[j + i]);
      const bool m = c < 4 && w < 4 && (w == c || (c == 3 && w == 1));
      if (m) eq |= (W)1 << i;
    }
    const W X = (eq & mask) | VN;
    const W D0 = ((VP + (X & VP)) ^ VP) | X;
    const W HN = VP & D0, HP = VN | ~(VP | D0);
    const W X2 = D0 >> 1;
    VN = X2 & HP; VP = HN | ~(X2 | HP);
    err += !(D0 & 1);
  }
  int end; err_out = 0xFFFFFFFFu;
  if (err > 3 * k) return -1;
  bmbs::band_last_column<W>(VP, VN, err, k, L, end, err_out);
  return end;
}

int main() {
  std::mt19937_64 g(7);
  const char acgt[] = "ACGT";
  long tot = 0, bad = 0, hits = 0;
  for (int it = 0; it < 300000; ++it) {
    const int L = 1 + (int)(g() % (it % 7 == 0 ? 300 : 120));
    int k = (int)(g() % 32); if (it % 3) k = std::min(31, (int)(0.08 * L + g() % 3));
    const int plen = L + 2 * k;
    std::string win(plen + 8, 'A');
    for (auto& c : win) c = acgt[g() % 4];
    if (it % 29 == 0) for (int i = 0; i < 3; ++i) win[g() % plen] = 'N';
    if (it % 97 == 0) std::fill(win.begin(), win.end(), '\0');          // out-of-genome window
    std::string read;
    const int shift = (int)(g() % (2 * k + 1));                          // where in the band the alignment starts
    const int mode = (int)(g() % 10);
    for (int i = 0, w = shift; (int)read.size() < L; ++i) {
      char c = w < plen && mode != 0 ? win[w] : acgt[g() % 4];
      if (c == '\0') c = acgt[g() % 4];
      if (c == 'C' && g() % 50) c = 'T';
      const unsigned r = (unsigned)(g() % 1000);
      const unsigned rate = mode < 4 ? 10 : mode < 8 ? 50 : 150;
      if (r < rate) { const unsigned t = (unsigned)(g() % 6); if (t < 4) { read.push_back("ACGTN"[g() % 5]); ++w; } else if (t == 4) { ++w; } else read.push_back(acgt[g() % 4]); }
      else { read.push_back(c); ++w; }
    }
    read.resize(L);
    uint32_t oe, pe; 
    const int oend = oracle::banded_bs_edit(win.data(), read.data(), L, (unsigned)k, oe);
    const int pend = k <= 15 && (it & 1) ? product_band<uint32_t>(win.data(), read.data(), L, k, pe) : product_band<uint64_t>(win.data(), read.data(), L, k, pe);
    ++tot; hits += oend >= 0;
    if (oend != pend || (oend >= 0 && oe != pe) || (oend < 0 && pe != 0xFFFFFFFFu && pend >= 0)) {
      if (++bad < 5) printf("L %d k %d: oracle (%d, %u) product (%d, %u)\n", L, k, oend, oe, pend, pe);
    }
  }
  printf("tests %ld hits %ld mismatching %ld\n", tot, hits, bad);
  return bad != 0;
}
