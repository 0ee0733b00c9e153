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
#include <string_view>
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

// Block reader for the mapper: hands out raw FASTQ text holding a whole number of 4-line records, so that parsing can run
// in worker threads.  Plain and gzip input (gzread passes plain files through).
class FastqBlockReader {
 public:
  bool open(const std::string& path) { gz_ = gzopen(path.c_str(), "rb"); if (gz_) gzbuffer(gz_, 1 << 22); return gz_ != nullptr; }
  ~FastqBlockReader() { if (gz_) gzclose(gz_); }
  // up to max_rec records -> out (ends with '\n'); returns the number of records, 0 at end of input
  size_t next(size_t max_rec, std::string& out) {
    out.clear();
    size_t lines = 0, scan = pos_;
    const size_t want = max_rec * 4;
    for (;;) {
      while (lines < want) {
        const char* nl = scan < buf_.size() ? (const char*)memchr(buf_.data() + scan, '\n', buf_.size() - scan) : nullptr;
        if (!nl) break;
        scan = (size_t)(nl - buf_.data()) + 1; ++lines;
      }
      if (lines == want || eof_) break;
      // need more input: drop what was handed out already, then append a chunk
      if (pos_) { buf_.erase(0, pos_); scan -= pos_; pos_ = 0; }
      const size_t old = buf_.size(), chunk = (size_t)16 << 20;
      buf_.resize(old + chunk);
      const int got = gzread(gz_, &buf_[old], (unsigned)chunk);
      buf_.resize(old + (got > 0 ? (size_t)got : 0));
      if (got <= 0) eof_ = true;
    }
    if (eof_ && lines < want && scan < buf_.size()) { buf_ += '\n'; scan = buf_.size(); ++lines; }   // last line without a newline
    const size_t rec = lines / 4;
    if (rec == 0) { pos_ = buf_.size(); return 0; }
    // hand out whole records only
    size_t end = pos_, l = 0;
    if (lines == rec * 4) end = scan;
    else for (; l < rec * 4; ++l) end = (size_t)((const char*)memchr(buf_.data() + end, '\n', buf_.size() - end) - buf_.data()) + 1;
    out.assign(buf_.data() + pos_, end - pos_);
    pos_ = end;
    return rec;
  }
 private:
  gzFile gz_ = nullptr; std::string buf_; size_t pos_ = 0; bool eof_ = false;
};

// one line of a raw block, without the line end ('\n' or "\r\n"); advances p
inline std::string_view next_line(const char*& p, const char* end) {
  const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
  const char* e = nl ? nl : end;
  std::string_view v(p, (size_t)(e - p));
  p = nl ? nl + 1 : end;
  if (!v.empty() && v.back() == '\r') v.remove_suffix(1);
  return v;
}
inline void cut_name_se(std::string_view& n) { size_t p = n.find_first_of(" /"); if (p != std::string_view::npos) n = n.substr(0, p); }
inline void cut_name_pe(std::string_view& a, std::string_view& b) {
  size_t j = 0;
  for (; j < a.size(); ++j) if (j >= b.size() || a[j] != b[j] || a[j] == ' ' || a[j] == '/') break;
  if (j < a.size()) { a = a.substr(0, j); if (b.size() > j) b = b.substr(0, j); }
}

inline void cut_name_se(std::string& n) { size_t p = n.find_first_of(" /"); if (p != std::string::npos) n.resize(p); }
inline void cut_name_pe(std::string& a, std::string& b) {
  size_t j = 0;
  for (; j < a.size(); ++j) if (j >= b.size() || a[j] != b[j] || a[j] == ' ' || a[j] == '/') break;
  if (j < a.size()) { a.resize(j); if (b.size() > j) b.resize(j); }
}
inline std::string revcomp(std::string_view s) {
  std::string r(s.size(), 'N');
  for (size_t i = 0; i < s.size(); ++i) {
    char c = s[s.size() - 1 - i];
    r[i] = c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : c;
  }
  return r;
}

}  // namespace bmbs
