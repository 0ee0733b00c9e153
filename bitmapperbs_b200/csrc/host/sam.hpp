// SAM text records in the reference's exact layout (Schema.cpp:12596-12812 single
// end; :9453-9700, :10922-11170 paired end) and the --mapstats block
// (Bitmapper_main.cpp:266-308).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <string>
#include <string_view>
#include "postprocess.hpp"

namespace bmbs {

inline void sam_header(std::string& out, const ChromTable& ct, const std::string& cmdline) {
  out += "@HD\tVN:1.4\tSO:unsorted\n";
  for (size_t i = 0; i < ct.name.size(); ++i) out += "@SQ\tSN:" + ct.name[i] + "\tLN:" + std::to_string(ct.len[i]) + "\n";
  out += "@PG\tID:BitMapperBS\tVN:1.0.2.3\tCL:" + cmdline + "\n";
}

// Single-end record.  `seq`/`qual` as in the FASTQ; reverse-strand hits print
// the reverse complement and the reversed qualities.
inline void sam_record_se(std::string& out, std::string_view name, std::string_view seq, std::string_view qual,
                          const ChromTable& ct, const Placed& p, int mapq, const std::string& cigar, unsigned nm,
                          std::string_view rseq) {
  out += name; out += '\t';
  out += std::to_string(p.flag); out += '\t';
  out += ct.name[p.chrom]; out += '\t';
  out += std::to_string(p.pos); out += '\t';
  out += std::to_string(mapq); out += '\t';
  out += cigar; out += "\t*\t0\t0\t";
  if (p.flag == 0) { out += seq; out += '\t'; out += qual; }
  else { out += rseq; out += '\t'; out.append(qual.rbegin(), qual.rend()); }
  out += "\tNM:i:"; out += std::to_string(nm); out += '\n';
}

// Paired-end record (Schema.cpp:9453-9700 mate 1, :10922-11170 mate 2).  `seq` is
// what was aligned, `rseq` its reverse complement, `qual` the FASTQ-order
// qualities; for mate 2 `seq` is the reverse complement of the FASTQ record.
// strand_flag: 0 = aligned sequence lies on the forward strand, 16 = reverse.
inline void sam_record_pe(std::string& out, bool first, std::string_view name, std::string_view seq, std::string_view rseq,
                          std::string_view qual, const ChromTable& ct, int strand_flag, size_t chrom, uint64_t pos, int mapq,
                          const std::string& cigar, uint64_t mate_pos, long long tlen, unsigned nm) {
  const int flag = first ? (strand_flag == 0 ? 99 : 83) : (strand_flag == 0 ? 147 : 163);
  out += name; out += '\t'; out += std::to_string(flag); out += '\t'; out += ct.name[chrom]; out += '\t';
  out += std::to_string(pos); out += '\t'; out += std::to_string(mapq); out += '\t'; out += cigar; out += "\t=\t";
  out += std::to_string(mate_pos); out += '\t';
  if (mate_pos < pos || (mate_pos == pos && !first)) out += '-';
  out += std::to_string((int)tlen); out += '\t';
  if (first) {
    if (flag & 32) { out += seq; out += '\t'; out += qual; } else { out += rseq; out += '\t'; out.append(qual.rbegin(), qual.rend()); }
  } else {
    if (flag & 16) { out += seq; out += '\t'; out.append(qual.rbegin(), qual.rend()); } else { out += rseq; out += '\t'; out += qual; }
  }
  out += "\tNM:i:"; out += std::to_string(nm); out += '\n';
}

struct MapStats { uint64_t reads = 0, unique = 0, ambiguous = 0, bases = 0, err_bases = 0; };

}  // namespace bmbs
