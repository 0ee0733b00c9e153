// Test harness for bitmapperbs_b200/csrc/host/bam.hpp: SAM file (header with @SQ lines + records) -> BAM file, through the
// same calls bmbs --bam makes (header, records of a block of text, BGZF members, EOF marker).  tests/test_host_bam.py
// decodes the result in Python and compares every field with the SAM text.  CPU only.
#include <cstdio>
#include <fstream>
#include <sstream>
#include "../bitmapperbs_b200/csrc/host/bam.hpp"

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::ifstream in(argv[1], std::ios::binary);
  std::stringstream ss; ss << in.rdbuf();
  const std::string text = ss.str();
  bmbs::ChromTable ct; std::string header, body;
  size_t at = 0;
  while (at < text.size()) {
    size_t nl = text.find('\n', at); if (nl == std::string::npos) nl = text.size();
    const std::string line = text.substr(at, nl - at);
    if (!line.empty() && line[0] == '@') {
      header += line + "\n";
      if (line.rfind("@SQ", 0) == 0) {
        const size_t sn = line.find("SN:"), ln = line.find("LN:");
        const size_t se = line.find('\t', sn);
        ct.name.push_back(line.substr(sn + 3, se - sn - 3)); ct.len.push_back(strtoull(line.c_str() + ln + 3, nullptr, 10));
      }
    } else if (!line.empty()) body += line + "\n";
    at = nl + 1;
  }
  bmbs::BamWriter bw; std::string raw, out;
  bw.header(ct, header, raw);
  if (!bmbs::BamWriter::bgzf(raw, out)) return 1;
  // records in blocks of a few hundred lines, like the sub-blocks of the mapper
  size_t p = 0; int lines = 0; size_t start = 0;
  while (p < body.size()) {
    const size_t nl = body.find('\n', p);
    p = nl + 1;
    if (++lines == 300 || p >= body.size()) {
      raw.clear(); bw.records(std::string_view(body.data() + start, p - start), raw);
      if (!bmbs::BamWriter::bgzf(raw, out)) return 1;
      start = p; lines = 0;
    }
  }
  bmbs::BamWriter::eof_marker(out);
  FILE* f = fopen(argv[2], "wb"); fwrite(out.data(), 1, out.size(), f); fclose(f);
  return 0;
}
