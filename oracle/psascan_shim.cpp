// oracle/ (test infrastructure): stand-in for the external `psascan` binary the
// reference shells out to during --index (bwt.cpp:1031-1042:
// `./psascan tmp_ref.tmp -m 8192` -> tmp_ref.tmp.sa5, 5-byte little-endian
// entries, uint40.h).  The vendored pSAscan/libdivsufsort need cmake + OpenMP
// builds (their own build systems), so the oracle build links this instead.
// A suffix array is unique for a text, so any correct sorter yields the same
// .sa5 file; the reference's own index code then runs unmodified on it.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "../bitmapperbs_b200/indexer/suffix_array.hpp"

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: psascan FILE [-m MB]\n"); return 2; }
  std::string in = argv[1], out = in + ".sa5";
  FILE* f = fopen(in.c_str(), "rb");
  if (!f) { perror(in.c_str()); return 1; }
  fseek(f, 0, SEEK_END); uint64_t n = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> t(n);
  if (fread(t.data(), 1, n, f) != n) { perror("read"); return 1; }
  fclose(f);
  for (uint64_t i = 0; i < n; ++i) if (t[i] > 2) { fprintf(stderr, "psascan shim: symbol %u > 2 at %llu\n", t[i], (unsigned long long)i); return 1; }
  FILE* g = fopen(out.c_str(), "wb");
  if (!g) { perror(out.c_str()); return 1; }
  std::vector<uint8_t> buf; buf.reserve(5u << 20);
  auto emit = [&](uint64_t v) { for (int b = 0; b < 5; ++b) buf.push_back((uint8_t)(v >> (8 * b))); if (buf.size() >= (5u << 20)) { fwrite(buf.data(), 1, buf.size(), g); buf.clear(); } };
  if (n < 0xFFFFFFFFull) { auto sa = bmbs::build_suffix_array<uint32_t>(t.data(), n); for (auto v : sa) emit(v); }
  else { auto sa = bmbs::build_suffix_array<uint64_t>(t.data(), n); for (auto v : sa) emit(v); }
  fwrite(buf.data(), 1, buf.size(), g);
  fclose(g);
  fprintf(stderr, "psascan shim: wrote %llu entries to %s\n", (unsigned long long)n, out.c_str());
  return 0;
}
