// oracle/seam_cpu_shim.cpp -- TEST INFRASTRUCTURE.  The entry points of include/bmbs.h that the reference-side seam
// (integration/bmbs_seam.h) calls, served by the CPU oracle (oracle_capi.cpp) instead of libbmbs_gpu.so, so that the patched
// reference (oracle/_ref/bitmapperBS_seam_cpu) can be diffed against the stock one on a machine without a GPU.  What this
// tests is the splice in integration/patch_reference.py, not the product; the GPU test links the same patched sources
// against the real library (oracle/_ref/bitmapperBS_gpu).
#include <cstdint>
#include <cstring>
#include <string>
#include "../include/bmbs.h"

extern "C" {
void* orc_load(const char* prefix);
void orc_free(void* h);
int orc_map_se(void* h, const char* seqs, const uint64_t* offs, int n, double e_rate, int seed_len, bmbs_read_result* res, bmbs_cand* cand, size_t cap, size_t* used);
int orc_map_pe(void* h, const char* seqs, const uint64_t* offs, int n_pairs, double e_rate, int seed_len, int min_ins, int max_ins,
               bmbs_read_result* res, bmbs_cand* cand, size_t cap, size_t* used);

static thread_local std::string shim_err;
const char* bmbs_last_error(void) { return shim_err.c_str(); }
void bmbs_params_default(bmbs_params* p) { p->e_rate = 0.08; p->seed_len = 30; p->min_ins = 0; p->max_ins = 500; p->sensitive = 0; p->ambiguous_out = 0; }
int bmbs_index_load(const char* prefix, const int*, int, bmbs_index** out) {
  void* h = orc_load(prefix);
  if (!h) { shim_err = std::string("cannot load ") + prefix; return BMBS_ERR_IO; }
  *out = (bmbs_index*)h; return BMBS_OK;
}
void bmbs_index_free(bmbs_index* idx) { if (idx) orc_free(idx); }
int bmbs_map_batch_se(bmbs_index* idx, int, const char* seqs, const uint64_t* offsets, int n_reads, const bmbs_params* prm,
                      bmbs_read_result* res, bmbs_cand* cand, size_t cand_cap, size_t* cand_used) {
  const int rc = orc_map_se(idx, seqs, offsets, n_reads, prm->e_rate, prm->seed_len, res, cand, cand_cap, cand_used);
  if (rc) shim_err = "cand[] too small";
  return rc;
}
int bmbs_map_batch_pe(bmbs_index* idx, int, const char* seqs, const uint64_t* offsets, int n_pairs, const bmbs_params* prm,
                      bmbs_read_result* res, bmbs_cand* cand, size_t cand_cap, size_t* cand_used) {
  const int rc = orc_map_pe(idx, seqs, offsets, n_pairs, prm->e_rate, prm->seed_len, prm->min_ins, prm->max_ins, res, cand, cand_cap, cand_used);
  if (rc) shim_err = "cand[] too small";
  return rc;
}
}
