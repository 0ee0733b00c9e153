// oracle/ (test infrastructure): exposes the REFERENCE's own FM-index functions (bwt.h / bwt.cpp compiled
// from a scratch copy of /root/reference that only has the missing `return`s added, see build_ref.sh)
// through a C ABI: index load, one LF step, the greedy seed, the exact count and single-row locate.
#include <cstdint>
#include <cstring>
#include "bwt.h"

extern "C" int ref_fm_load(const char* prefix_bs) { char buf[1000]; strncpy(buf, prefix_bs, 999); buf[999] = 0; load_index(buf); return bitmapper_index_params.SA_length ? 0 : -1; }
extern "C" uint64_t ref_fm_lf(uint64_t row, int c) { return find_occ_fm_index(row, c, bitmapper_index_params.bwt, bitmapper_index_params.high_occ_table); }
extern "C" void ref_fm_lf_pair(uint64_t sp, uint64_t ep, int c, uint64_t* nsp, uint64_t* nep) {
  bitmapper_bs_iter a, b; find_occ_fm_index_combine(sp, ep, &a, &b, c, bitmapper_index_params.bwt); *nsp = a; *nep = b;
}
extern "C" uint64_t ref_fm_seed(const char* pat, uint64_t len, uint64_t* sp, uint64_t* ep, uint64_t* mlen) {
  bitmapper_bs_iter a = *sp, b = *ep, a1 = 0, b1 = 0, m = 0;
  bitmapper_bs_iter h = count_backward_as_much_1_terminate((char*)pat, len, &a, &b, &a1, &b1, &m);
  *sp = a; *ep = b; *mlen = m; return h;
}
extern "C" uint64_t ref_fm_count(const char* pat, uint64_t len, uint64_t* sp, uint64_t* ep) {
  bitmapper_bs_iter a = *sp, b = *ep, a1 = 0, b1 = 0;
  bitmapper_bs_iter h = count_hash_table((char*)pat, len, &a, &b, &a1, &b1);
  *sp = a; *ep = b; return h;
}
extern "C" int ref_fm_locate(uint64_t row, uint64_t* sa) {
  bitmapper_bs_iter out[4] = {0, 0, 0, 0}, n = 0;
  locate_one_position(out, row, &n);
  *sa = out[0]; return (int)n;
}
extern "C" uint64_t ref_fm_locate_interval(const char* pat, uint64_t sp, uint64_t ep, uint64_t sp1, uint64_t ep1, uint64_t len, uint64_t* out) {
  bwt_locate_queue q; init_locate_queue_muti_thread(&q);
  bitmapper_bs_iter n = 0;
  locate_muti_thread((char*)pat, sp, ep, sp1, ep1, out, len, &n, &q);
  return n;
}
