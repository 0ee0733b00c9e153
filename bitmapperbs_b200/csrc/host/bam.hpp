// --bam: BAM output without htslib.  The reference turns every SAM line it has formatted into a BAM record
// (convert_string_to_bam, bam_prase.cpp:248-274: sam_parse1 + bam_write1 through its patched htslib) and writes the
// header from the chromosome table (init_bam_header, bam_prase.cpp:62-112).  This does the same on the finished SAM
// text of a sub-block: SAM line -> BAM record (SAM/BAM specification, section 4.2), records -> BGZF blocks (zlib raw
// deflate, one gzip member of at most 64 KiB of payload each).  Sub-blocks are compressed by the finishing threads and
// concatenated in input order; BGZF members concatenate into a valid file, which ends with the empty EOF member.
#pragma once
#include <zlib.h>
#include <cstdint>
#include <cstring>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>
#include "postprocess.hpp"

namespace bmbs {

struct BamWriter {
  std::unordered_map<std::string, int32_t> ref_id;

  static void put32(std::string& o, uint32_t v) { char b[4] = {(char)v, (char)(v >> 8), (char)(v >> 16), (char)(v >> 24)}; o.append(b, 4); }
  static void put16(std::string& o, uint32_t v) { char b[2] = {(char)v, (char)(v >> 8)}; o.append(b, 2); }

  // "BAM\1", header text, reference names and lengths
  void header(const ChromTable& ct, const std::string& sam_header_text, std::string& raw) {
    raw.append("BAM\1", 4);
    put32(raw, (uint32_t)sam_header_text.size()); raw += sam_header_text;
    put32(raw, (uint32_t)ct.name.size());
    for (size_t i = 0; i < ct.name.size(); ++i) {
      put32(raw, (uint32_t)ct.name[i].size() + 1); raw += ct.name[i]; raw += '\0';
      put32(raw, (uint32_t)ct.len[i]);
      ref_id[ct.name[i]] = (int32_t)i;
    }
  }

  // UCSC binning scheme (specification 5.3)
  static uint32_t reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return (uint32_t)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (uint32_t)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (uint32_t)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (uint32_t)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (uint32_t)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
  }

  static std::string_view field(const char*& p, const char* e) {
    const char* t = (const char*)memchr(p, '\t', (size_t)(e - p));
    const char* q = t ? t : e;
    std::string_view v(p, (size_t)(q - p));
    p = t ? t + 1 : e;
    return v;
  }
  static int64_t to_int(std::string_view v) { int64_t x = 0; bool neg = false; size_t i = 0; if (!v.empty() && v[0] == '-') { neg = true; i = 1; } for (; i < v.size(); ++i) x = x * 10 + (v[i] - '0'); return neg ? -x : x; }

  // one SAM line (without the newline) -> one BAM record appended to raw
  void record(std::string_view line, std::string& raw) const {
    const char* p = line.data(); const char* e = p + line.size();
    const std::string_view qname = field(p, e), flag_s = field(p, e), rname = field(p, e), pos_s = field(p, e), mapq_s = field(p, e),
                           cigar = field(p, e), rnext = field(p, e), pnext_s = field(p, e), tlen_s = field(p, e), seq = field(p, e), qual = field(p, e);
    auto ref_of = [&](std::string_view n) -> int32_t { if (n == "*") return -1; auto it = ref_id.find(std::string(n)); return it == ref_id.end() ? -1 : it->second; };
    const int32_t rid = ref_of(rname);
    const int32_t nrid = rnext == "=" ? rid : ref_of(rnext);
    const int64_t pos = to_int(pos_s) - 1, pnext = to_int(pnext_s) - 1;
    std::vector<uint32_t> ops; int64_t ref_len = 0;
    if (cigar != "*") {
      uint32_t len = 0;
      for (char c : cigar) {
        if (c >= '0' && c <= '9') { len = len * 10 + (uint32_t)(c - '0'); continue; }
        const char* codes = "MIDNSHP=X"; const char* w = strchr(codes, c);
        const uint32_t op = w ? (uint32_t)(w - codes) : 0;
        ops.push_back(len << 4 | op);
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_len += len;
        len = 0;
      }
    }
    const int64_t end = pos + (ref_len > 0 ? ref_len : 1);
    const uint32_t l_seq = seq == "*" ? 0u : (uint32_t)seq.size();
    // optional fields: the mapper only writes NM:i:<n>; integers take the smallest unsigned type, like sam_parse1
    std::string aux;
    while (p < e) {
      const std::string_view f = field(p, e);
      if (f.size() >= 5 && f[2] == ':' && f[3] == 'i' && f[4] == ':') {
        const int64_t v = to_int(f.substr(5));
        aux += f[0]; aux += f[1];
        if (v >= 0) { if (v <= 0xFF) { aux += 'C'; aux += (char)v; } else if (v <= 0xFFFF) { aux += 'S'; put16(aux, (uint32_t)v); } else { aux += 'I'; put32(aux, (uint32_t)v); } }
        else { if (v >= -128) { aux += 'c'; aux += (char)v; } else if (v >= -32768) { aux += 's'; put16(aux, (uint32_t)v); } else { aux += 'i'; put32(aux, (uint32_t)v); } }
      } else if (f.size() >= 5 && f[2] == ':' && f[3] == 'Z' && f[4] == ':') { aux += f[0]; aux += f[1]; aux += 'Z'; aux += f.substr(5); aux += '\0'; }
    }
    const uint32_t l_name = (uint32_t)qname.size() + 1;
    const uint32_t block = 32 + l_name + 4 * (uint32_t)ops.size() + (l_seq + 1) / 2 + l_seq + (uint32_t)aux.size();
    put32(raw, block);
    put32(raw, (uint32_t)rid); put32(raw, (uint32_t)(int32_t)pos);
    raw += (char)l_name; raw += (char)to_int(mapq_s); put16(raw, reg2bin(pos, end));
    put16(raw, (uint32_t)ops.size()); put16(raw, (uint32_t)to_int(flag_s));
    put32(raw, l_seq);
    put32(raw, (uint32_t)nrid); put32(raw, (uint32_t)(int32_t)pnext); put32(raw, (uint32_t)(int32_t)to_int(tlen_s));
    raw += qname; raw += '\0';
    for (uint32_t o : ops) put32(raw, o);
    static const struct Nt16 { uint8_t t[256]; Nt16() { memset(t, 15, 256); const char* s = "=ACMGRSVTWYHKDBN"; for (int i = 0; i < 16; ++i) { t[(uint8_t)s[i]] = (uint8_t)i; t[(uint8_t)(s[i] | 0x20)] = (uint8_t)i; } } } nt;
    for (uint32_t i = 0; i < l_seq; i += 2) raw += (char)(nt.t[(uint8_t)seq[i]] << 4 | (i + 1 < l_seq ? nt.t[(uint8_t)seq[i + 1]] : 0));
    if (qual == "*" || qual.size() != l_seq) raw.append(l_seq, (char)0xFF);
    else for (uint32_t i = 0; i < l_seq; ++i) raw += (char)(qual[i] - 33);
    raw += aux;
  }

  // every line of a sub-block's SAM text
  void records(std::string_view sam_text, std::string& raw) const {
    const char* p = sam_text.data(); const char* e = p + sam_text.size();
    while (p < e) {
      const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
      const char* q = nl ? nl : e;
      if (q > p && *p != '@') record(std::string_view(p, (size_t)(q - p)), raw);
      p = nl ? nl + 1 : e;
    }
  }

  // raw bytes -> BGZF members appended to out (specification 4.1)
  static bool bgzf(std::string_view raw, std::string& out, int level = 6) {
    const size_t CHUNK = 0xff00;
    for (size_t at = 0; at < raw.size(); at += CHUNK) {
      const size_t n = std::min(CHUNK, raw.size() - at);
      unsigned char buf[0x10000 + 64];
      z_stream z; memset(&z, 0, sizeof z);
      if (deflateInit2(&z, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
      z.next_in = (Bytef*)(raw.data() + at); z.avail_in = (uInt)n; z.next_out = buf; z.avail_out = sizeof buf;
      const int rc = deflate(&z, Z_FINISH);
      const size_t clen = z.total_out;
      deflateEnd(&z);
      if (rc != Z_STREAM_END || clen + 26 > 0x10000) return false;
      const unsigned char hdr[12] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0};
      out.append((const char*)hdr, 12);
      out += 'B'; out += 'C'; put16(out, 2); put16(out, (uint32_t)(clen + 25));
      out.append((const char*)buf, clen);
      put32(out, (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef*)(raw.data() + at), (uInt)n)); put32(out, (uint32_t)n);
    }
    return true;
  }
  static void eof_marker(std::string& out) {
    static const unsigned char m[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    out.append((const char*)m, 28);
  }
};

}  // namespace bmbs
