// oracle/ (test infrastructure): exposes the REFERENCE's own verification kernels -- compiled from
// /root/reference/Levenshtein_Cal.h where it lies (header-only, unmodified) -- through a C ABI so
// tests can pin the restatement (oracle_core.hpp banded_bs_edit) and the CUDA kernel against them.
#include <cstring>
#include "Levenshtein_Cal.h"

extern "C" int ref_bpm_scalar(const char* win, int p_len, const char* read, int t_len, unsigned short k, unsigned int* err) {
  return BS_Reserve_Banded_BPM((char*)win, p_len, (char*)read, t_len, k, err);
}
// 8 x u32 lanes (k <= 15); wins: 8 windows of p_len bytes each, stride `stride`
extern "C" void ref_bpm_8(const char* wins, int stride, int p_len, const char* read, int t_len, unsigned short k, int* sites, unsigned int* errs) {
  __m256i Peq[256];
  for (int i = 0; i < 256; ++i) Peq[i] = _mm256_setzero_si256();
  char* w = (char*)wins;
  BS_Reserve_Banded_BPM_8_SSE(w, w + stride, w + 2 * stride, w + 3 * stride, w + 4 * stride, w + 5 * stride, w + 6 * stride, w + 7 * stride,
                              p_len, (char*)read, t_len, sites, errs, k, Peq);
}
// 4 x u64 lanes (k <= 31)
extern "C" void ref_bpm_4(const char* wins, int stride, int p_len, const char* read, int t_len, unsigned short k, int* sites, unsigned int* errs) {
  __m256i Peq[256];
  for (int i = 0; i < 256; ++i) Peq[i] = _mm256_setzero_si256();
  char* w = (char*)wins;
  BS_Reserve_Banded_BPM_4_SSE(w, w + stride, w + 2 * stride, w + 3 * stride, p_len, (char*)read, t_len, sites, errs, k, Peq);
}
