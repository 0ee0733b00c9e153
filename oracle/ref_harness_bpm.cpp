// oracle/ (test infrastructure): exposes the REFERENCE's own verification kernels -- compiled from
// /root/reference/Levenshtein_Cal.h where it lies (header-only, unmodified) -- through a C ABI so
// tests can pin the restatement (oracle_core.hpp banded_bs_edit) and the CUDA kernel against them.
#include <cstring>
#include <thread>   // before the reference header: it defines min/max macros
#include <vector>
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

// Threaded batch driver for the verification microbench (BASELINE.json config 5): n_reads reads, each against its own
// group of 8 windows (k <= 15: one 8-lane call; k > 15: two 4-lane calls), the way map_candidate_votes_mutiple_cut_end_to_end_8/_4
// (Schema.cpp:7707, :6740) feed them.  wins: [n_reads][8][stride] ASCII, reads: [n_reads][t_len].
extern "C" void ref_bpm_batch(const char* wins, size_t n_reads, int stride, int p_len, const char* reads, int t_len, unsigned short k,
                              int threads, int* sites, unsigned int* errs) {
  auto work = [&](int t) {
    __m256i Peq[256];
    for (int i = 0; i < 256; ++i) Peq[i] = _mm256_setzero_si256();
    std::vector<char> rd(t_len + 64, 0);
    for (size_t r = t; r < n_reads; r += threads) {
      char* w = (char*)wins + r * 8 * (size_t)stride;
      memcpy(rd.data(), reads + r * (size_t)t_len, t_len);
      if (k <= 15)
        BS_Reserve_Banded_BPM_8_SSE(w, w + stride, w + 2 * stride, w + 3 * stride, w + 4 * stride, w + 5 * stride, w + 6 * stride, w + 7 * stride,
                                    p_len, rd.data(), t_len, sites + r * 8, errs + r * 8, k, Peq);
      else
        for (int h = 0; h < 2; ++h) {
          char* v = w + 4 * h * (size_t)stride;
          BS_Reserve_Banded_BPM_4_SSE(v, v + stride, v + 2 * stride, v + 3 * stride, p_len, rd.data(), t_len, sites + r * 8 + 4 * h, errs + r * 8 + 4 * h, k, Peq);
        }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < threads; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
}
