// oracle/oracle_pe_sensitive.hpp -- TEST INFRASTRUCTURE (see oracle_core.hpp).
// `--pe --sensitive` restatement: Map_Pair_Seq_split (Schema.cpp:22450-23953) and its helpers
// process_rest_seed_muti_thread (:17852), process_rest_seed_filter_muti_thread (:17244),
// reseed_filter_muti_thread (:16998), select_best_seeds (:16630), select_suit_candidates (:4775),
// generate_candidate_votes_shift_filter (:4884).
//
// Shape of the reference: the first seed of both mates; the mate with fewer candidates after it is the
// "primary" and is seeded, voted and verified without a filter; the other ("secondary") is seeded the same
// way but only windows within the insert-size range of a verified primary hit are kept (C3) and verified;
// a secondary left without a hit is re-seeded (two halves / gaps of the seeds used so far, then a seed every
// 8 bases, minimum multi-hit seed length 20) and filtered the same way; then the best pair is picked (V5).
// Seeding of a mate never depends on the other mate, so both are seeded up front here.
#pragma once
#include "oracle_pe.hpp"

namespace oracle {

// select_suit_candidates, Schema.cpp:4775-4824: is there a verified mate hit within [dmin, dmax] of `site`?
// `next` persists across the (ascending) sites of one read, as in the reference.
inline int has_mate_in_range(u64 site, const std::vector<Vote>& hits, int nh, int dmax, int dmin, int& next) {
  for (int i = next; i < nh; ++i) {
    if (hits[i].site > site) {
      const long long d = (long long)(hits[i].site - site);
      if (d > dmax) return 0;
      if (d <= dmax && d >= dmin) return 1;
    } else {
      const long long d = (long long)(site - hits[i].site);
      if (d > dmax) next = i + 1;
      else if (d >= dmin) return 1;
    }
  }
  return 0;
}

// C3: generate_candidate_votes_shift_filter, Schema.cpp:4884-4992
inline void votes_from_sorted_filtered(const std::vector<u64>& cand, u64 k, std::vector<Vote>& out,
                                       const std::vector<Vote>& hits, int nh, int dmax, int dmin) {
  out.clear();
  if (cand.empty()) return;
  int next = 0;
  u64 prev = cand[0], vote = 1; size_t i = 1;
  auto emit = [&](u64 site) { if (has_mate_in_range(site, hits, nh, dmax, dmin, next)) out.push_back({site, vote, 0, 0}); };
  while (prev < k && i < cand.size()) {
    if (cand[i] == prev) { ++vote; } else { emit(0); vote = 1; prev = cand[i]; }
    ++i;
  }
  while (i < cand.size()) {
    if (cand[i] == prev) { ++vote; } else { emit(prev - k); vote = 1; prev = cand[i]; }
    ++i;
  }
  emit(prev >= k ? prev - k : 0);
}

// tail of process_rest_seed{,_filter}_muti_thread (Schema.cpp:17990-18040, :17480-17530): candidates -> hits
inline int candidates_to_hits(const Index& ix, const char* read, int L, u64 k, const SeedTrace& t, std::vector<Vote>& v,
                              const std::vector<Vote>* mate_hits, int mate_n, int dmax, int dmin) {
  v.clear();
  if (!t.extra && (t.cand.size() == 1 || (t.cand.size() == 2 && t.cand[0] == t.cand[1]))) {
    v.push_back({t.cand[0], 0, 1, (u64)(L - 1)});
    return 1;
  }
  if (t.cand.empty()) return 0;
  std::vector<u64> c = t.cand; std::sort(c.begin(), c.end());
  if (mate_hits) votes_from_sorted_filtered(c, k, v, *mate_hits, mate_n, dmax, dmin);
  else votes_from_sorted(c, k, v);
  return verify_keep_hits(ix, read, L, k, v);
}

// select_best_seeds, Schema.cpp:16630-16670.  With no seed used so far the reference reads element [-1] of its two
// malloc'ed int arrays; with glibc that is the upper half of the chunk-size word, i.e. 0, which yields one seed
// covering the whole read -- restated as that.
inline int select_best_seeds(const std::vector<int>& start, const std::vector<int>& len, int L, int rs[3], int rl[3]) {
  const size_t n = start.size();
  int m = 0;
  if (n >= 2) {
    m = 2;
    rs[0] = start[0]; rl[0] = start[1] - start[0];
    rs[1] = start[n - 2] + len[n - 2]; rl[1] = L - rs[1];
  } else if (n == 1) {
    m = 2;
    rs[0] = start[0]; rl[0] = L / 2;
    rs[1] = rs[0] + rl[0]; rl[1] = L - rs[1];
  }
  const int last_end = n ? start[n - 1] + len[n - 1] : 0;
  if (last_end < L) { rs[m] = last_end; rl[m] = L - last_end; ++m; }
  return m;
}

// reseed_filter_muti_thread, Schema.cpp:16998-17240
inline int reseed_filtered(const Index& ix, const char* read, int L, u64 k, const SeedTrace& t, std::vector<Vote>& v,
                           const std::vector<Vote>& mate_hits, int mate_n, int dmax, int dmin) {
  std::string bs; int c_site; reverse_c_to_t(read, L, bs, c_site);
  u64 max_seeds = (u64)L / 10 - 1; if (max_seeds > 25) max_seeds = 25;
  const u64 min_multi_len = 20, max_hits = 1000, step = 8;
  std::vector<u64> cand; SeedTrace dummy;
  int rs[3], rl[3];
  const int nsel = select_best_seeds(t.seed_start, t.seed_len, L, rs, rl);
  u64 seed_id = 0, sp = 0, ep = 0;
  while (seed_id < (u64)nsel) {
    const u64 off = (u64)rs[seed_id], cur = (u64)L - off, mlen = (u64)(long long)rl[seed_id];
    const u64 hits = count_exact(ix, bs.data() + (cur - mlen), mlen, sp, ep);
    if (hits == 1) { u64 sa; if (ix.locate_row(sp, sa)) cand.push_back(ix.total_sa_length - sa - mlen - off); }
    else if (mlen >= min_multi_len && hits <= max_hits) { if (hits) locate_interval(ix, sp, ep, mlen, off, cand, dummy); }
    else if (cur == mlen) break;
    ++seed_id;
  }
  u64 off = t.seed_start.size() > 1 ? (u64)((t.seed_start[0] + t.seed_start[1]) / 2) : step / 2;
  while (seed_id < max_seeds && off < (u64)L) {
    const u64 cur = (u64)L - off;
    SeedHit h = seed_until_unique(ix, bs.data(), cur, sp, ep);
    if (h.hits == 1) { u64 sa; if (ix.locate_row(sp, sa)) cand.push_back(ix.total_sa_length - sa - h.mlen - off); }
    else if (h.mlen >= min_multi_len && h.hits <= max_hits) { if (h.hits) locate_interval(ix, sp, ep, h.mlen, off, cand, dummy); }
    else if (cur == h.mlen) break;
    off += step;
    ++seed_id;
  }
  v.clear();
  if (cand.empty()) return 0;
  std::sort(cand.begin(), cand.end());
  votes_from_sorted_filtered(cand, k, v, mate_hits, mate_n, dmax, dmin);
  return verify_keep_hits(ix, read, L, k, v);
}

// One mate's state after seeding, sensitive mode
struct SensMate { SeedTrace t; bool jump = false; int occ = 0; std::vector<Vote> v; size_t first_cands = 0; bool reseeded = false; };

inline void sens_seed(const Index& ix, const Params& P, const char* read, int L, SensMate& m) {
  m = SensMate();
  seed_read(ix, P, read, L, m.t, true, 10000);
  if (m.t.exact_unique) { m.v.push_back({m.t.cand[0], 0, 0, (u64)(L - 1)}); m.occ = 1; m.jump = true; }
  else if (m.t.multi_exact_noC) {
    for (u64 s : m.t.multi_sites) m.v.push_back({s, 0, 0, (u64)(L - 1)});
    m.occ = (int)m.v.size(); m.jump = true;
  } else m.first_cands = m.t.first_cands;
}

struct SensDebug { int primary = 0; SensMate a, b; };

inline bool run_pe_sensitive_pair(const Index& ix, const Params& P, const char* read1, const char* qual1, int L1,
                                  const char* read2, const char* qual2, int L2, PairOutcome& o, Stats& st, SensDebug* dbg) {
  o = PairOutcome(); ++st.reads;
  const u64 k1 = u64_k(P.e_rate, L1), k2 = u64_k(P.e_rate, L2), kl = k1 > k2 ? k1 : k2;
  const int maxlen = L1 > L2 ? L1 : L2;
  const int dmax = (int)((u64)P.max_ins + kl * 2), dmin = (int)((u64)P.min_ins - kl * 2 - (u64)maxlen);
  SensMate a, b;
  sens_seed(ix, P, read1, L1, a);
  sens_seed(ix, P, read2, L2, b);
  const bool first_is_1 = a.first_cands <= b.first_cands;        // Schema.cpp:23326
  SensMate& pri = first_is_1 ? a : b; SensMate& sec = first_is_1 ? b : a;
  const char* rp = first_is_1 ? read1 : read2; const char* rs = first_is_1 ? read2 : read1;
  const int Lp = first_is_1 ? L1 : L2, Ls = first_is_1 ? L2 : L1;
  const u64 kp = first_is_1 ? k1 : k2, ks = first_is_1 ? k2 : k1;
  auto done = [&]() { if (dbg) { dbg->primary = first_is_1 ? 0 : 1; dbg->a = a; dbg->b = b; } };
  if (!pri.jump) pri.occ = candidates_to_hits(ix, rp, Lp, kp, pri.t, pri.v, nullptr, 0, dmax, dmin);
  if (pri.occ == 0) { done(); return true; }
  if (!sec.jump) sec.occ = candidates_to_hits(ix, rs, Ls, ks, sec.t, sec.v, &pri.v, pri.occ, dmax, dmin);
  if (sec.occ == 0) { sec.occ = reseed_filtered(ix, rs, Ls, ks, sec.t, sec.v, pri.v, pri.occ, dmax, dmin); sec.reseeded = true; }
  o.pick = pick_pair(a.v, a.occ, b.v, b.occ, (int)kl, dmax, dmin);
  o.n_pairs = o.pick.n;
  if (o.n_pairs == 1) {
    finish_mate(ix, P, read1, qual1, L1, k1, a.v[o.pick.i1], false, o.m1);
    finish_mate(ix, P, read2, qual2, L2, k2, b.v[o.pick.i2], true, o.m2);
    o.tlen = tlen_of((long long)o.m1.pos, o.m1.span, (long long)o.m2.pos, o.m2.span);
    const int t = (int)o.tlen;
    if (t <= P.max_ins && t >= P.min_ins && o.m1.pos + o.m1.span <= ix.chroms.len[o.m1.chrom] + 1 &&
        o.m2.pos + o.m2.span <= ix.chroms.len[o.m2.chrom] + 1) {
      ++st.unique; st.bases += L1 + L2; st.err_bases += o.m1.err + o.m2.err;
      o.mapq = bmbs::mapq_from(o.pick.second_best_diff, (u32)(k1 + k2), o.m1.score + o.m2.score, P.sc);
      o.written = true;
    }
  } else if (o.n_pairs > 1) ++st.ambiguous;
  done();
  return true;
}

}  // namespace oracle
