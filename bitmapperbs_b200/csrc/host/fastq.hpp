// FASTQ(.gz) record reader with the reference's normalisation
// (Process_Reads.cpp:62-90, :321-472): names lose the leading '@' and are cut at
// the first ' ' or '/', bases are upper-cased; mate 2 of a pair is additionally
// cut where the two names first differ.  Every record must end with '\n'.
#pragma once
#include <zlib.h>
#include <cctype>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace bmbs {

struct FastqRecord { std::string name, seq, qual; };

class FastqReader {
 public:
  bool open(const std::string& path) { gz_ = gzopen(path.c_str(), "rb"); if (gz_) gzbuffer(gz_, 1 << 20); return gz_ != nullptr; }
  ~FastqReader() { if (gz_) gzclose(gz_); }
  bool next(FastqRecord& r) {
    if (!line(r.name)) return false;
    std::string plus;
    if (!line(r.seq) || !line(plus) || !line(r.qual)) return false;
    if (!r.name.empty() && r.name[0] == '@') r.name.erase(0, 1);
    for (auto& c : r.seq) c = toupper((unsigned char)c);
    if (r.qual.size() > r.seq.size()) r.qual.resize(r.seq.size());
    return true;
  }
 private:
  bool line(std::string& s) {
    s.clear();
    char buf[4096];
    for (;;) {
      if (!gzgets(gz_, buf, sizeof buf)) return !s.empty();
      size_t n = strlen(buf);
      if (n && buf[n - 1] == '\n') { s.append(buf, n - 1); if (!s.empty() && s.back() == '\r') s.pop_back(); return true; }
      s.append(buf, n);
    }
  }
  gzFile gz_ = nullptr;
};

inline void cut_name_se(std::string& n) { size_t p = n.find_first_of(" /"); if (p != std::string::npos) n.resize(p); }
inline void cut_name_pe(std::string& a, std::string& b) {
  size_t j = 0;
  for (; j < a.size(); ++j) if (j >= b.size() || a[j] != b[j] || a[j] == ' ' || a[j] == '/') break;
  if (j < a.size()) { a.resize(j); if (b.size() > j) b.resize(j); }
}
inline std::string revcomp(const std::string& s) {
  std::string r(s.size(), 'N');
  for (size_t i = 0; i < s.size(); ++i) {
    char c = s[s.size() - 1 - i];
    r[i] = c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : c;
  }
  return r;
}

}  // namespace bmbs
