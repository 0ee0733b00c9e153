// bmbs_seam.h -- the reference-side binding of libbmbs_gpu.so: what a BitMapperBS maintainer adds to Schema.cpp so that the
// pthread workers hand each sub-block of reads to the GPU library instead of seeding and verifying read by read.
//
// Included by the patched Schema.cpp (integration/patch_reference.py splices the calls in; nothing of the reference is copied
// into this repository).  It only uses the reference's own types and globals:
//   Read                         Process_Reads.h:35-45   (name / seq / rseq / qual / length; mate 2's seq is already the reverse
//                                                          complement of the FASTQ record, Process_Reads.cpp:359-369)
//   seed_votes                   Schema.h:169-176        (site, vote, err, end_site)
//   fileName[1]                  Auxiliary.h:45          ("<genome.fa>.index", what Load_Index opens)
//   thread_e_f, over_all_seed_length, bs_available_seed_length, minDistance_pair, maxDistance_pair   Auxiliary.h:41-57
// and the C ABI in include/bmbs.h.  Seams replaced:
//   Map_Single_Seq_split   Schema.cpp:27113-27520 (seeding) and :27592-27674 (sort, votes, verification + reduction)
//   Map_Pair_Seq_split_fast Schema.cpp:21961-22180 (get_candidates_muti_thread x 2, filter_pairs, verify_candidate_locations)
// Everything after the seam -- CIGAR, MAPQ, SAM/BAM text, counters, output queue -- stays the reference's own code.
#ifndef BMBS_SEAM_H
#define BMBS_SEAM_H
#include <pthread.h>
#include <algorithm>
#include <vector>
#include "bmbs.h"

static bmbs_index* bmbs_seam_index = NULL;
static pthread_once_t bmbs_seam_once = PTHREAD_ONCE_INIT;

static void bmbs_seam_die(const char* what) {
  fprintf(stderr, "bmbs seam: %s: %s\n", what, bmbs_last_error());
  exit(1);
}

// once per process, from the first worker that needs it: the same "<genome>.index" prefix Load_Index was given
static void bmbs_seam_load_once() {
  if (bmbs_index_load(fileName[1], NULL, 0, &bmbs_seam_index)) bmbs_seam_die("index load");
}
static void bmbs_seam_load() { pthread_once(&bmbs_seam_once, bmbs_seam_load_once); }

static void bmbs_seam_params(bmbs_params* p, int sensitive) {
  bmbs_params_default(p);
  p->e_rate = thread_e_f;
  p->seed_len = bs_available_seed_length != -1 ? bs_available_seed_length : over_all_seed_length;
  p->min_ins = minDistance_pair;
  p->max_ins = maxDistance_pair;
  p->sensitive = sensitive;
}

// one worker's view of one sub-block: the reads flattened for the library, and what came back
struct bmbs_seam_block {
  std::vector<char> seqs;
  std::vector<uint64_t> offs;
  std::vector<bmbs_read_result> res;
  std::vector<bmbs_cand> cand;
  size_t used = 0;

  void add(const Read& r) { seqs.insert(seqs.end(), r.seq, r.seq + r.length); offs.push_back(seqs.size()); }

  // single end: reads r1[0..n); paired end: mates r1[i], r2[i] interleaved
  void map(Read* r1, Read* r2, int n, int sensitive) {
    bmbs_seam_load();
    seqs.clear(); offs.clear(); offs.push_back(0);
    for (int i = 0; i < n; ++i) { add(r1[i]); if (r2) add(r2[i]); }
    const int n_reads = r2 ? 2 * n : n;
    res.resize((size_t)n_reads + 1);
    if (cand.size() < (size_t)n_reads * 32 + 1024) cand.resize((size_t)n_reads * 32 + 1024);
    seqs.resize(seqs.size() + 64);                       // the library reads whole words: slack behind the last base
    bmbs_params p; bmbs_seam_params(&p, sensitive);
    for (;;) {
      const int rc = r2 ? bmbs_map_batch_pe(bmbs_seam_index, 0, seqs.data(), offs.data(), n, &p, res.data(), cand.data(), cand.size(), &used)
                        : bmbs_map_batch_se(bmbs_seam_index, 0, seqs.data(), offs.data(), n, &p, res.data(), cand.data(), cand.size(), &used);
      if (rc == BMBS_ERR_CAPACITY && used > cand.size()) { cand.resize(used + used / 4 + 1024); continue; }
      if (rc) bmbs_seam_die("map batch");
      return;
    }
  }

  // a read's verified windows as the reference's seed_votes (site order, as generate_candidate_votes_shift leaves them)
  bitmapper_bs_iter votes_of(int read, seed_votes* out) const {
    const bmbs_read_result& r = res[read];
    for (uint32_t j = 0; j < r.n_cand; ++j) {
      const bmbs_cand& c = cand[r.first_cand + j];
      out[j].site = c.site; out[j].vote = c.vote;
      out[j].err = c.err == 0xFFFF ? (unsigned int)-1 : c.err;
      out[j].end_site = (bitmapper_bs_iter)(long long)c.end_site;
    }
    return r.n_cand;
  }
};

static inline bool bmbs_seam_resolved(const bmbs_read_result& r) {
  return r.state == BMBS_EXACT_UNIQUE || r.state == BMBS_MULTI_EXACT || r.state == BMBS_ONE_MISMATCH;
}

// What is left of map_candidate_votes_mutiple{,_cut}_end_to_end_{8,4} (Schema.cpp:7707-8183 / :8202-8745) once every window
// carries its (end_site, err): the sequential reduction in vote order.  `stop_at_two_exact` is the non-cut variant's early
// return (:8359-8362), taken when the first seed was a full-length multi-hit.
static inline void bmbs_seam_reduce(seed_votes* v, bitmapper_bs_iter n, bool stop_at_two_exact,
                                    unsigned int* min_err, int* min_err_index, unsigned int* second_best_diff) {
  *min_err = ((unsigned int)-1) - 1; *min_err_index = -1;
  bitmapper_bs_iter min_err_site = (bitmapper_bs_iter)-1;
  for (bitmapper_bs_iter i = 0; i < n; ++i) {
    const bitmapper_bs_iter end_abs = v[i].site + v[i].end_site;
    if (v[i].err == *min_err && min_err_site != end_abs && (stop_at_two_exact || *min_err_index >= 0)) {
      *second_best_diff = 0;
      if (*min_err_index >= 0) *min_err_index = -2 - *min_err_index;
      if (stop_at_two_exact && *min_err == 0) break;
    } else if (v[i].err < *min_err) {
      *second_best_diff = *min_err - v[i].err;
      *min_err = v[i].err; *min_err_index = (int)i; min_err_site = end_abs;
    }
  }
  if (*min_err_index >= 0) { v[0].err = v[*min_err_index].err; v[0].end_site = v[*min_err_index].end_site; v[0].site = v[*min_err_index].site; }
}

// What is left of map_candidate_votes_mutiple_cut_end_to_end_*_for_paired_end (Schema.cpp:7334-7698, compaction :7502-7512):
// hits within k whose absolute end differs from that of the window verified right before them, moved to the front.
static inline int bmbs_seam_keep_hits(seed_votes* v, bitmapper_bs_iter n, bitmapper_bs_iter k) {
  int kept = 0; bitmapper_bs_iter prev = (bitmapper_bs_iter)-1;
  for (bitmapper_bs_iter i = 0; i < n; ++i) {
    const bitmapper_bs_iter end_abs = v[i].site + v[i].end_site;
    if (v[i].err <= k && prev != end_abs) { v[kept].site = v[i].site; v[kept].err = v[i].err; v[kept].end_site = v[i].end_site; ++kept; }
    prev = end_abs;
  }
  return kept;
}
#endif
