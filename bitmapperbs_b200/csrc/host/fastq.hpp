// FASTQ(.gz) record reader with the reference's normalisation
// (Process_Reads.cpp:62-90, :321-472): names lose the leading '@' and are cut at
// the first ' ' or '/', bases are upper-cased; mate 2 of a pair is additionally
// cut where the two names first differ.  Every record must end with '\n'.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <algorithm>
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

// number of '\n' in [p, p+n): 16 bytes per step with SSE2 (part of x86-64), SWAR elsewhere
inline size_t count_newlines(const char* p, size_t n) {
  size_t c = 0, i = 0;
#if defined(__SSE2__)
  const __m128i nl = _mm_set1_epi8('\n');
  for (; i + 64 <= n; i += 64) {
    const unsigned m0 = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128((const __m128i*)(p + i)), nl));
    const unsigned m1 = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128((const __m128i*)(p + i + 16)), nl));
    const unsigned m2 = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128((const __m128i*)(p + i + 32)), nl));
    const unsigned m3 = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128((const __m128i*)(p + i + 48)), nl));
    c += (size_t)__builtin_popcountll((unsigned long long)m0 | ((unsigned long long)m1 << 16) | ((unsigned long long)m2 << 32) | ((unsigned long long)m3 << 48));
  }
#else
  for (; i + 8 <= n; i += 8) {
    unsigned long long x; memcpy(&x, p + i, 8); x ^= 0x0A0A0A0A0A0A0A0Aull;
    const unsigned long long t = ~(((x & 0x7F7F7F7F7F7F7F7Full) + 0x7F7F7F7F7F7F7F7Full) | x | 0x7F7F7F7F7F7F7F7Full);
    c += (size_t)__builtin_popcountll(t);
  }
#endif
  for (; i < n; ++i) c += p[i] == '\n';
  return c;
}

// Block reader for the mapper: hands out raw FASTQ text holding a whole number of 4-line records, so that parsing can run
// in worker threads.  A plain file is mapped into memory and handed out as views (no copy; the mapping is private and
// writable because parsing upper-cases bases in place -- only pages that really hold lower-case bases get copied); gzip
// input goes through gzread into an owned buffer.
class FastqBlockReader {
 public:
  bool open(const std::string& path) {
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    unsigned char magic[2] = {0, 0};
    const ssize_t got = ::pread(fd, magic, 2, 0);
    struct stat st;
    if (got == 2 && !(magic[0] == 0x1f && magic[1] == 0x8b) && fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
      void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ | PROT_WRITE, MAP_PRIVATE, fd, 0);
      if (m != MAP_FAILED) {
        map_ = (char*)m; map_size_ = (size_t)st.st_size;
        madvise(map_, map_size_, MADV_SEQUENTIAL);
        ::close(fd);
        return true;
      }
    }
    ::close(fd);
    gz_ = gzopen(path.c_str(), "rb"); if (gz_) gzbuffer(gz_, 1 << 22); return gz_ != nullptr;
  }
  ~FastqBlockReader() { if (gz_) gzclose(gz_); if (map_) munmap(map_, map_size_); }
  // up to max_rec records -> view (whole records; the last line of the input may lack its '\n'); `owned` backs the view for
  // gzip input.  Returns the number of records, 0 at end of input.
  size_t next(size_t max_rec, std::string_view& view, std::string& owned) {
    if (!map_) { const size_t n = next_copy(max_rec, owned); view = owned; return n; }
    const size_t want = max_rec * 4;
    size_t lines = 0, at = pos_;
    constexpr size_t BLK = 4096;
    while (lines < want && at < map_size_) {
      const size_t n = std::min(BLK, map_size_ - at);
      const size_t c = count_newlines(map_ + at, n);
      if (lines + c < want) { lines += c; at += n; continue; }
      while (lines < want) { at = (size_t)((const char*)memchr(map_ + at, '\n', map_size_ - at) - map_) + 1; ++lines; }   // the want-th newline is inside this block
    }
    if (lines < want && at == map_size_ && map_size_ > pos_ && map_[map_size_ - 1] != '\n') ++lines;   // last line without a newline
    const size_t rec = lines / 4;
    if (rec == 0) { pos_ = map_size_; return 0; }
    size_t end = at;
    if (lines != rec * 4) {      // trailing partial record: stop after the last whole one
      end = pos_;
      for (size_t l = 0; l < rec * 4; ++l) end = (size_t)((const char*)memchr(map_ + end, '\n', map_size_ - end) - map_) + 1;
    }
    view = std::string_view(map_ + pos_, end - pos_);
    pos_ = end;
    return rec;
  }
  bool mapped() const { return map_ != nullptr; }
  // Plain (mapped) input of 4-line records: the next ~target bytes as a view, cut at a record boundary that is found by looking
  // at the lines around the cut -- the text is not scanned here (the parse workers count the records of their block).  A record
  // starts at a line that begins with '@' and whose next line but one begins with '+'; a quality line may begin with '@' as
  // well, but two lines below it stands a sequence.  Returns the number of bytes, 0 at end of input.
  size_t next_bytes(size_t target, std::string_view& view) {
    if (pos_ >= map_size_) return 0;
    size_t end = pos_ + (target ? target : 1);
    end = end >= map_size_ ? map_size_ : record_start_after(end);
    view = std::string_view(map_ + pos_, end - pos_);
    pos_ = end;
    return view.size();
  }
  // bytes per record over the first (up to) 64 records of the input: sizes the blocks of next_bytes
  size_t bytes_per_record() const {
    size_t at = 0, lines = 0;
    while (lines < 256 && at < map_size_) { const char* nl = (const char*)memchr(map_ + at, '\n', map_size_ - at); at = nl ? (size_t)(nl - map_) + 1 : map_size_; ++lines; }
    return lines >= 4 ? at / (lines / 4) : (at ? at : 1);
  }
 private:
  size_t record_start_after(size_t at) const {
    size_t p = at, l[6];
    if (p > 0 && map_[p - 1] != '\n') { const char* nl = (const char*)memchr(map_ + p, '\n', map_size_ - p); if (!nl) return map_size_; p = (size_t)(nl - map_) + 1; }
    for (int i = 0; i < 6; ++i) {
      if (p >= map_size_) return map_size_;              // fewer than six lines left: the tail goes out with this block
      l[i] = p;
      const char* nl = (const char*)memchr(map_ + p, '\n', map_size_ - p);
      p = nl ? (size_t)(nl - map_) + 1 : map_size_;
    }
    for (int i = 0; i < 4; ++i) if (map_[l[i]] == '@' && map_[l[i + 2]] == '+') return l[i];
    return map_size_;                                    // not 4-line FASTQ: one block, the parser reports what it finds
  }
  // gzip path: up to max_rec records -> out (ends with '\n')
  size_t next_copy(size_t max_rec, std::string& out) {
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
  gzFile gz_ = nullptr; std::string buf_; size_t pos_ = 0; bool eof_ = false;
  char* map_ = nullptr; size_t map_size_ = 0;
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
  // the reference's loop also looks at mate 1's terminating NUL (Process_Reads.cpp:392-405): a mate-2 name that merely
  // continues mate 1's is cut to the same length
  if (j < a.size()) a = a.substr(0, j);
  if (b.size() > j) b = b.substr(0, j);
}

inline void cut_name_se(std::string& n) { size_t p = n.find_first_of(" /"); if (p != std::string::npos) n.resize(p); }
inline void cut_name_pe(std::string& a, std::string& b) {
  size_t j = 0;
  for (; j < a.size(); ++j) if (j >= b.size() || a[j] != b[j] || a[j] == ' ' || a[j] == '/') break;
  if (j < a.size()) a.resize(j);
  if (b.size() > j) b.resize(j);
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
