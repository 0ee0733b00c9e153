// SAM text records in the reference's exact layout (Schema.cpp:12596-12812 single
// end; :9453-9700, :10922-11170 paired end) and the --mapstats block
// (Bitmapper_main.cpp:266-308).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <string_view>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include "postprocess.hpp"

namespace bmbs {

// decimal text of v appended to out (std::to_string allocates a temporary per field)
inline void append_uint(std::string& out, uint64_t v) {
  char buf[24]; int n = 0;
  do { buf[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  const size_t at = out.size(); out.resize(at + (size_t)n);
  for (int i = 0; i < n; ++i) out[at + (size_t)i] = buf[n - 1 - i];
}
inline void append_int(std::string& out, long long v) { if (v < 0) { out += '-'; append_uint(out, (uint64_t)(-(v + 1)) + 1u); } else append_uint(out, (uint64_t)v); }
// dst[i] = src[n-1-i], sixteen bytes at a time (byte shuffle) where the CPU has SSSE3; complement: A<->T, C<->G, anything else stays
#if defined(__x86_64__)
__attribute__((target("ssse3"))) inline void reverse_bytes_ssse3(char* dst, const char* src, size_t n, bool complement) {
  const __m128i rev = _mm_set_epi8(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
  const __m128i cA = _mm_set1_epi8('A'), cC = _mm_set1_epi8('C'), cG = _mm_set1_epi8('G'), cT = _mm_set1_epi8('T');
  const __m128i xAT = _mm_set1_epi8('A' ^ 'T'), xCG = _mm_set1_epi8('C' ^ 'G');
  size_t i = 0;
  for (; i + 16 <= n; i += 16) {
    __m128i v = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(src + n - 16 - i)), rev);
    if (complement) {
      const __m128i at = _mm_or_si128(_mm_cmpeq_epi8(v, cA), _mm_cmpeq_epi8(v, cT)), cg = _mm_or_si128(_mm_cmpeq_epi8(v, cC), _mm_cmpeq_epi8(v, cG));
      v = _mm_xor_si128(v, _mm_or_si128(_mm_and_si128(at, xAT), _mm_and_si128(cg, xCG)));
    }
    _mm_storeu_si128((__m128i*)(dst + i), v);
  }
  for (; i < n; ++i) { const char c = src[n - 1 - i]; dst[i] = complement ? (c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : c) : c; }
}
inline bool cpu_has_ssse3() { static const bool v = __builtin_cpu_supports("ssse3"); return v; }
#endif
inline void append_reversed(std::string& out, std::string_view s) {
  const size_t at = out.size(), n = s.size(); out.resize(at + n);
#if defined(__x86_64__)
  if (cpu_has_ssse3()) { reverse_bytes_ssse3(&out[at], s.data(), n, false); return; }
#endif
  for (size_t i = 0; i < n; ++i) out[at + i] = s[n - 1 - i];
}
inline char complement_base(char c) { return c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : c; }
struct ComplementTable { char t[256]; ComplementTable() { for (int i = 0; i < 256; ++i) t[i] = complement_base((char)i); } };
inline const char* complement_table() { static const ComplementTable c; return c.t; }
inline void append_revcomp(std::string& out, std::string_view s) {
  const char* t = complement_table();
  const size_t at = out.size(), n = s.size(); out.resize(at + n);
#if defined(__x86_64__)
  if (cpu_has_ssse3()) { reverse_bytes_ssse3(&out[at], s.data(), n, true); return; }
#endif
  for (size_t i = 0; i < n; ++i) out[at + i] = t[(unsigned char)s[n - 1 - i]];
}

inline void sam_header(std::string& out, const ChromTable& ct, const std::string& cmdline) {
  out += "@HD\tVN:1.4\tSO:unsorted\n";
  for (size_t i = 0; i < ct.name.size(); ++i) out += "@SQ\tSN:" + ct.name[i] + "\tLN:" + std::to_string(ct.len[i]) + "\n";
  out += "@PG\tID:BitMapperBS\tVN:1.0.2.3\tCL:" + cmdline + "\n";
}

// Single-end record.  `seq`/`qual` as in the FASTQ; reverse-strand hits print
// the reverse complement and the reversed qualities.
// raw field writers for the hot single-end record: the record's text is written through a pointer into space reserved once
inline char* put_bytes(char* w, std::string_view s) { memcpy(w, s.data(), s.size()); return w + s.size(); }
inline char* put_uint(char* w, uint64_t v) {
  char buf[24]; int n = 0;
  do { buf[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (n) *w++ = buf[--n];
  return w;
}
inline void sam_record_se(std::string& out, std::string_view name, std::string_view seq, std::string_view qual,
                          const ChromTable& ct, const Placed& p, int mapq, std::string_view cigar, unsigned nm) {
  const std::string& cn = ct.name[p.chrom];
  const size_t at = out.size();
  out.resize(at + name.size() + cn.size() + cigar.size() + seq.size() + qual.size() + 96);
  char* const w0 = &out[at]; char* w = w0;
  w = put_bytes(w, name); *w++ = '\t';
  w = put_uint(w, (uint64_t)p.flag); *w++ = '\t';
  w = put_bytes(w, cn); *w++ = '\t';
  w = put_uint(w, p.pos); *w++ = '\t';
  if (mapq < 0) { *w++ = '-'; w = put_uint(w, (uint64_t)(-(long long)mapq)); } else w = put_uint(w, (uint64_t)mapq);
  *w++ = '\t';
  w = put_bytes(w, cigar); w = put_bytes(w, "\t*\t0\t0\t");
  if (p.flag == 0) { w = put_bytes(w, seq); *w++ = '\t'; w = put_bytes(w, qual); }
  else {
#if defined(__x86_64__)
    if (cpu_has_ssse3()) { reverse_bytes_ssse3(w, seq.data(), seq.size(), true); w += seq.size(); *w++ = '\t'; reverse_bytes_ssse3(w, qual.data(), qual.size(), false); w += qual.size(); }
    else
#endif
    {
      const char* t = complement_table();
      for (size_t i = 0, n = seq.size(); i < n; ++i) *w++ = t[(unsigned char)seq[n - 1 - i]];
      *w++ = '\t';
      for (size_t i = 0, n = qual.size(); i < n; ++i) *w++ = qual[n - 1 - i];
    }
  }
  w = put_bytes(w, "\tNM:i:"); w = put_uint(w, nm); *w++ = '\n';
  out.resize(at + (size_t)(w - w0));
}

// --pbat single-end record (output_sam_end_to_end_pbat_output_buffer, Schema.cpp:13215-13450): `seq` is what was aligned (the
// reverse complement of the FASTQ record `raw`), `qual` in FASTQ order
inline void sam_record_se_pbat(std::string& out, std::string_view name, std::string_view seq, std::string_view raw, std::string_view qual,
                               const ChromTable& ct, const Placed& p, int mapq, std::string_view cigar, unsigned nm) {
  out += name; out += '\t';
  append_uint(out, (uint64_t)p.flag); out += '\t';
  out += ct.name[p.chrom]; out += '\t';
  append_uint(out, p.pos); out += '\t';
  append_int(out, mapq); out += '\t';
  out += cigar; out += "\t*\t0\t0\t";
  if (p.flag == 0) { out += seq; out += '\t'; append_reversed(out, qual); }
  else { out += raw; out += '\t'; out += qual; }
  out += "\tNM:i:"; append_uint(out, nm); out += '\n';
}

// Paired-end record (Schema.cpp:9453-9700 mate 1, :10922-11170 mate 2).  `seq` is
// what was aligned, `rseq` its reverse complement (empty: computed here), `qual` the FASTQ-order
// qualities; for mate 2 `seq` is the reverse complement of the FASTQ record.
// strand_flag: 0 = aligned sequence lies on the forward strand, 16 = reverse.
inline void sam_record_pe(std::string& out, bool first, std::string_view name, std::string_view seq, std::string_view rseq,
                          std::string_view qual, const ChromTable& ct, int strand_flag, size_t chrom, uint64_t pos, int mapq,
                          const std::string& cigar, uint64_t mate_pos, long long tlen, unsigned nm) {
  const int flag = first ? (strand_flag == 0 ? 99 : 83) : (strand_flag == 0 ? 147 : 163);
  out += name; out += '\t'; append_uint(out, (uint64_t)flag); out += '\t'; out += ct.name[chrom]; out += '\t';
  append_uint(out, pos); out += '\t'; append_int(out, mapq); out += '\t'; out += cigar; out += "\t=\t";
  append_uint(out, mate_pos); out += '\t';
  if (mate_pos < pos || (mate_pos == pos && !first)) out += '-';
  append_int(out, (int)tlen); out += '\t';
  auto put_rseq = [&] { if (rseq.empty()) append_revcomp(out, seq); else out += rseq; };
  if (first) {
    if (flag & 32) { out += seq; out += '\t'; out += qual; } else { put_rseq(); out += '\t'; append_reversed(out, qual); }
  } else {
    if (flag & 16) { out += seq; out += '\t'; append_reversed(out, qual); } else { put_rseq(); out += '\t'; out += qual; }
  }
  out += "\tNM:i:"; append_uint(out, nm); out += '\n';
}

// --unmapped_out record (Schema.cpp:13041-13110; paired end :10392-10520 with flags 77 / 141): the read as it stands in the FASTQ
inline void sam_record_unmapped(std::string& out, std::string_view name, int flag, std::string_view seq, std::string_view qual) {
  out += name; out += '\t'; append_uint(out, (uint64_t)flag); out += "\t*\t0\t0\t*\t*\t0\t0\t"; out += seq; out += '\t'; out += qual; out += '\n';
}

struct MapStats { uint64_t reads = 0, unique = 0, ambiguous = 0, bases = 0, err_bases = 0; };

}  // namespace bmbs
