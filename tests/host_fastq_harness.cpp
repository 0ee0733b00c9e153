// Test harness for bitmapperbs_b200/csrc/host/fastq.hpp: prints, for every block the reader hands out, the record count and
// a checksum of the text, then the total; used by tests/test_host_fastq.py (CPU only).
#include <cstdio>
#include <cstdlib>
#include "../bitmapperbs_b200/csrc/host/fastq.hpp"

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  bmbs::FastqBlockReader r;
  if (!r.open(argv[1])) { printf("open failed\n"); return 1; }
  const size_t per = (size_t)atoll(argv[2]);
  size_t total = 0;
  std::string_view v; std::string own;
  if (argc > 3 && r.mapped()) {      // blocks cut by size (the single-end path of the mapper): argv[3] = bytes per block, records counted like the parse worker does
    const size_t target = (size_t)atoll(argv[3]);
    while (r.next_bytes(target, v)) {
      size_t lines = bmbs::count_newlines(v.data(), v.size());
      if (!v.empty() && v.back() != '\n') ++lines;
      const size_t n = lines / 4;
      total += n;
      const char* p = v.data(); const char* e = p + v.size();
      for (size_t i = 0; i < n; ++i) {
        std::string_view name = bmbs::next_line(p, e), seq = bmbs::next_line(p, e); bmbs::next_line(p, e); std::string_view q = bmbs::next_line(p, e);
        printf("%.*s\t%.*s\t%.*s\n", (int)name.size(), name.data(), (int)seq.size(), seq.data(), (int)q.size(), q.data());
      }
    }
    printf("TOTAL %zu\n", total);
    return 0;
  }
  for (;;) {
    const size_t n = r.next(per, v, own);
    if (!n) break;
    total += n;
    const char* p = v.data(); const char* e = p + v.size();
    for (size_t i = 0; i < n; ++i) {
      std::string_view name = bmbs::next_line(p, e), seq = bmbs::next_line(p, e); bmbs::next_line(p, e); std::string_view q = bmbs::next_line(p, e);
      printf("%.*s\t%.*s\t%.*s\n", (int)name.size(), name.data(), (int)seq.size(), seq.data(), (int)q.size(), q.data());
    }
    if (p != e) { printf("LEFTOVER %zu bytes in block\n", (size_t)(e - p)); return 1; }
  }
  printf("TOTAL %zu\n", total);
  return 0;
}
