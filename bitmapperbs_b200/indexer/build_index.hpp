// bmbs-index: writes BitMapperBS's on-disk index (the files `bitmapperBS
// --index` produces) for a FASTA genome.  Data-prep tooling on the CPU, as the
// north_star keeps index construction on the CPU; the mapper (GPU) and the
// reference mapper both load these files unchanged.
//
// File formats follow the reference writer (cited per block below):
//   <fa>.index              Index.cpp:134-159   chromosome table + N
//   <fa>.index.bs.pac       Index.cpp:734-831   2-bit genome, first base in the top bits
//   <fa>.index.bs.index     bwt.cpp:1715-1729, :1830-1835   68-byte header
//   <fa>.index.bs.index.bwt bwt.cpp:1345-1531 (bit-plane BWT + 16-bit counters), :1863-2084 (3^16 table)
//   <fa>.index.bs.index.occ bwt.cpp:1435-1446   absolute counts every 65536 symbols
//   <fa>.index.bs.index.sa  bwt.cpp:1558-1696 (flag bit-vector with rank words), :1751-1816 (sampled SA)
// The text indexed is complement(genome) followed by reverse(genome), both with
// C->T, over the alphabet G<T<A (Index.cpp:591-692, bwt.cpp:1135-1140).
//
// Not written: <fa>.index.methy (only --methy_extract reads it; out of scope).
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#pragma once
#include "suffix_array.hpp"

namespace bmbs {
namespace indexer {

struct Chrom { std::string name; uint64_t len; };

[[noreturn]] void die(const std::string& m) { fprintf(stderr, "bmbs-index: %s\n", m.c_str()); exit(1); }

void wr(FILE* f, const void* p, size_t sz, size_t n) { if (n && fwrite(p, sz, n, f) != n) die("short write"); }

// FASTA: name ends at the first blank; bases upper-cased (Ref_Genome.cpp:32-113).
void read_fasta(const char* path, std::vector<Chrom>& chroms, std::vector<uint8_t>& g) {
  FILE* f = fopen(path, "rb");
  if (!f) die(std::string("cannot open ") + path);
  fseek(f, 0, SEEK_END); size_t sz = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<char> buf(sz);
  if (fread(buf.data(), 1, sz, f) != sz) die("short read");
  fclose(f);
  g.reserve(sz);
  size_t i = 0;
  uint64_t rng = 0x9E3779B97F4A7C15ull, replaced = 0;
  while (i < sz) {
    char c = buf[i];
    if (c == '>') {
      size_t e = i + 1;
      while (e < sz && buf[e] != '\n') ++e;
      size_t s = i + 1, t = s;
      while (t < e && buf[t] != ' ') ++t;
      if (!chroms.empty()) chroms.back().len = g.size() - chroms.back().len;
      chroms.push_back({std::string(buf.data() + s, t - s), g.size()});  // len holds start for now
      i = e + 1;
    } else {
      if (!isspace((unsigned char)c)) {
        char u = toupper(c);
        if (u != 'A' && u != 'C' && u != 'G' && u != 'T') {
          // the reference replaces every base outside ACGT by a random one seeded with the time of day (Index.cpp:696-729),
          // so two of its own builds differ there; here the replacement is a fixed-seed sequence (reproducible builds)
          rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
          u = "ACGT"[(rng >> 33) & 3]; ++replaced;
        }
        if (chroms.empty()) die("sequence before first FASTA header");
        g.push_back(u);
      }
      ++i;
    }
  }
  if (chroms.empty()) die("no FASTA records");
  chroms.back().len = g.size() - chroms.back().len;
  if (replaced) fprintf(stderr, "bmbs-index: %llu bases outside ACGT replaced by pseudo-random bases (fixed seed)\n", (unsigned long long)replaced);
}

// ---- writers shared by the CPU builder below and the device builder (gpu_index.cu): same bytes from the same arrays
inline void write_chrom_table(const std::string& fa, const std::vector<Chrom>& chroms, uint64_t N) {
  FILE* f = fopen((fa + ".index").c_str(), "wb"); if (!f) die("cannot write .index");
  uint64_t nc = chroms.size(); wr(f, &nc, 8, 1);
  for (auto& c : chroms) { uint64_t l = c.name.size(); wr(f, &l, 8, 1); wr(f, c.name.data(), 1, l); wr(f, &c.len, 8, 1); }
  wr(f, &N, 8, 1); fclose(f);
}
inline void write_pac(const std::string& fa, const std::vector<uint8_t>& g) {
  const uint64_t N = g.size();
  std::vector<uint8_t> pac((N + 3) / 4, 0);
  for (uint64_t i = 0; i < N; ++i) {
    uint8_t c = g[i] == 'A' ? 0 : g[i] == 'C' ? 1 : g[i] == 'G' ? 2 : 3;
    pac[i >> 2] |= c << (6 - 2 * (i & 3));
  }
  FILE* f = fopen((fa + ".index.bs.pac").c_str(), "wb"); if (!f) die("cannot write .pac");
  uint64_t nb = pac.size(); wr(f, &nb, 8, 1); wr(f, pac.data(), 1, nb); fclose(f);
}
// kc[h]: occurrences of 16-mer h in the text; tail: codes of the last min(15, n) symbols of the text
inline void write_bwt_files(const std::string& fa, uint64_t R, uint64_t shapline, uint64_t cnt0, uint64_t cnt1, uint64_t cnt2,
                            const std::vector<uint64_t>& bwt, uint64_t bwt_words, const std::vector<uint64_t>& high_occ,
                            const std::vector<uint32_t>& ssa, const std::vector<uint64_t>& flag, uint64_t flag_words,
                            const std::vector<uint32_t>& kc, const std::vector<uint8_t>& tail, uint64_t n) {
  const std::string p = fa + ".index.bs.index";
  {
    FILE* f = fopen((p + ".occ").c_str(), "wb"); if (!f) die("cannot write .occ");
    uint64_t l = high_occ.size(); wr(f, &l, 8, 1); wr(f, high_occ.data(), 8, l); fclose(f);
  }
  uint64_t nacgt[5] = {1, 1 + cnt0, 1 + cnt0 + cnt1, 1 + cnt0 + cnt1 + cnt2, 1 + cnt0 + cnt1 + cnt2};
  {
    FILE* f = fopen(p.c_str(), "wb"); if (!f) die("cannot write .bs.index");
    wr(f, &R, 8, 1); wr(f, &shapline, 8, 1); wr(f, nacgt, 8, 5);
    uint32_t prm[3] = {8, 64, 128}; wr(f, prm, 4, 3); fclose(f);
  }
  {
    FILE* f = fopen((p + ".sa").c_str(), "wb"); if (!f) die("cannot write .sa");
    uint64_t l = ssa.size(); wr(f, &l, 8, 1); wr(f, ssa.data(), 4, l);
    wr(f, &flag_words, 8, 1); wr(f, flag.data(), 8, flag_words); fclose(f);
  }
  // ---- 3^16 table: entry h = first row of 16-mer h; rows of suffixes shorter than 16
  // symbols that sit between two consecutive 16-mer intervals are recorded as a 4-bit
  // gap in the top bits of the following entry (bwt.cpp:1893-2007, query bwt.h:284-306).
  const uint64_t H = 43046721ull;  // 3^16
  std::vector<uint64_t> short_pad;  // padded keys of the suffixes shorter than 16 (excluding the empty one)
  for (uint64_t l = 1; l < 16 && l <= n; ++l) {
    uint64_t v = 0;
    for (uint64_t i = tail.size() - l; i < tail.size(); ++i) v = v * 3 + tail[i];
    for (uint64_t i = l; i < 16; ++i) v *= 3;
    short_pad.push_back(v);
  }
  std::sort(short_pad.begin(), short_pad.end());
  std::vector<uint32_t> hi(H + 1, 0); std::vector<uint8_t> lo(H + 1, 0);
  {
    uint64_t rows_before = 1;   // the empty suffix
    uint64_t chain = 1;         // what the table holds for entry h before it is visited
    size_t sp_i = 0;
    for (uint64_t h = 0; h < H; ++h) {
      while (sp_i < short_pad.size() && short_pad[sp_i] <= h) { ++rows_before; ++sp_i; }
      uint64_t top, bot; uint32_t diff = 0;
      if (kc[h] == 0) { top = bot = chain; }
      else { top = rows_before; bot = top + kc[h]; diff = (uint32_t)(top - chain) << 28; }
      hi[h] = (uint32_t)(top >> 8) | diff; lo[h] = top & 255;
      hi[h + 1] = (uint32_t)(bot >> 8); lo[h + 1] = bot & 255;
      chain = (((uint64_t)(hi[h + 1] & 0x0FFFFFFFu)) << 8) | lo[h + 1];
      rows_before += kc[h];
    }
  }
  {
    FILE* f = fopen((p + ".bwt").c_str(), "wb"); if (!f) die("cannot write .bwt");
    wr(f, &bwt_words, 8, 1); wr(f, bwt.data(), 8, bwt_words);
    uint64_t hn = H + 1; wr(f, &hn, 8, 1); wr(f, hi.data(), 4, hn); wr(f, lo.data(), 1, hn); fclose(f);
  }
  fprintf(stderr, "bmbs-index: wrote %s{,.bwt,.sa,.occ}\n", p.c_str());
}

}  // namespace indexer

// Builds every index file next to `fa`; returns 0 on success (fatal problems exit with a message).
inline int build_index(const std::string& fa, int threads = 0) {
  using namespace indexer;

  std::vector<Chrom> chroms; std::vector<uint8_t> g;
  read_fasta(fa.c_str(), chroms, g);
  const uint64_t N = g.size(), n = 2 * N;
  fprintf(stderr, "bmbs-index: %zu chromosomes, %llu bases\n", chroms.size(), (unsigned long long)N);

  write_chrom_table(fa, chroms, N);
  write_pac(fa, g);

  // 3-letter double-strand text, codes G=0 T=1 A=2
  std::vector<uint8_t> t(n);
  for (uint64_t i = 0; i < N; ++i) {
    uint8_t c = g[i];
    t[i] = c == 'A' ? 1 : c == 'C' ? 0 : c == 'G' ? 1 : 2;           // complement, then C->T
    t[n - 1 - i] = c == 'A' ? 2 : c == 'G' ? 0 : 1;                  // reversed, C->T
  }
  std::vector<uint8_t>().swap(g);

  const bool wide = n >= 0xFFFFFFFFull;
  std::vector<uint32_t> sa32; std::vector<uint64_t> sa64;
  if (wide) sa64 = bmbs::build_suffix_array<uint64_t>(t.data(), n, threads);
  else sa32 = bmbs::build_suffix_array<uint32_t>(t.data(), n, threads);
  auto SA = [&](uint64_t row) -> uint64_t { return row == 0 ? n : (wide ? sa64[row - 1] : (uint64_t)sa32[row - 1]); };
  fprintf(stderr, "bmbs-index: suffix array done\n");

  const uint64_t R = n + 1;  // rows including the empty suffix
  // ---- BWT bit-planes with interleaved 16-bit counters + 65536-symbol absolute table
  uint64_t S = n;            // BWT symbols (the row whose SA is 0 carries '$' and is dropped)
  uint64_t bwt_words = 1 + 2 * (S / 64) + (S / 128) + 2;
  std::vector<uint64_t> bwt(bwt_words + 8, 0), high_occ(2, 0);
  uint64_t cnt[3] = {0, 0, 0}, shapline = 0, j = 0;
  for (uint64_t row = 0; row < R; ++row) {
    uint64_t s = SA(row);
    if (s == 0) { shapline = row; continue; }
    uint8_t ch = t[s - 1];
    uint64_t w = (j >> 7) * 5 + 1 + ((j >> 6) & 1) * 2, bit = 63 - (j & 63);
    bwt[w] |= (uint64_t)(ch & 1) << bit;
    bwt[w + 1] |= (uint64_t)(ch >> 1) << bit;
    ++cnt[ch]; ++j;
    if ((j & 65535) == 0) { high_occ.push_back(cnt[1]); high_occ.push_back(cnt[2]); }
    if ((j & 63) == 0) {
      uint64_t hw = (j >> 7) * 5, base = (j >> 16) * 2;
      uint64_t c1 = cnt[1] - high_occ[base], c2 = cnt[2] - high_occ[base + 1];
      bwt[hw] |= (j & 64) ? (c1 << 16) | c2 : (c1 << 48) | (c2 << 32);
    }
  }
  // ---- flag bit-vector (rows whose SA is a multiple of 8) + sampled SA
  uint64_t flag_words = 1 + R / 64 + (R % 64 ? 1 : 0) + R / 256 + 1;
  std::vector<uint64_t> flag(flag_words + 8, 0);
  std::vector<uint32_t> ssa; ssa.reserve(n / 8 + 2);
  for (uint64_t row = 0; row < R; ++row) {
    if ((row & 255) == 0) flag[(row >> 8) * 5] = ssa.size();
    uint64_t s = SA(row);
    if ((s & 7) == 0) {
      flag[(row >> 8) * 5 + 1 + ((row & 255) >> 6)] |= 1ull << (63 - (row & 63));
      uint32_t ch = s ? t[s - 1] : 1;
      ssa.push_back((ch << 30) | (uint32_t)(s >> 3));
    }
  }
  if ((R & 255) == 0) flag[(R >> 8) * 5] = ssa.size();

  std::vector<uint32_t> kc(43046721ull, 0);
  if (n >= 16) {
    uint64_t key = 0; const uint64_t P15 = 14348907ull;
    for (int i = 0; i < 16; ++i) key = key * 3 + t[i];
    for (uint64_t q = 0;; ++q) {
      ++kc[key];
      if (q + 16 >= n) break;
      key = (key - t[q] * P15) * 3 + t[q + 16];
    }
  }
  std::vector<uint8_t> tail(t.end() - (n >= 15 ? 15 : n), t.end());
  write_bwt_files(fa, R, shapline, cnt[0], cnt[1], cnt[2], bwt, bwt_words, high_occ, ssa, flag, flag_words, kc, tail, n);
  return 0;
}

}  // namespace bmbs
