// oracle/oracle_pe.hpp -- TEST INFRASTRUCTURE (see oracle_core.hpp).
// Paired-end restatement: Map_Pair_Seq_split_fast (Schema.cpp:21460-22444, `--pe`) and the
// helpers it calls.  `--pe --sensitive` (Map_Pair_Seq_split, :22450-23953) is in oracle_pe_sensitive.hpp.
#pragma once
#include "oracle_core.hpp"
#include "../bitmapperbs_b200/csrc/host/fastq.hpp"
#include "../bitmapperbs_b200/csrc/host/sam.hpp"

namespace oracle {

// One mate after seeding: get_candidates_muti_thread, Schema.cpp:19550-19947.
//   occ == -1 : `v` holds site-sorted candidate windows still to verify
//   occ ==  0 : nothing
//   occ  >  0 : `v[0..occ)` are finished hits (site, err, end_site)
struct MateCands { int occ = -1; std::vector<Vote> v; SeedTrace trace; };

inline void mate_candidates(const Index& ix, const Params& P, const char* read, int L, u64 k, MateCands& m) {
  m = MateCands();
  seed_read(ix, P, read, L, m.trace, true);
  SeedTrace& t = m.trace;
  if (t.exact_unique) { m.v.push_back({t.cand[0], 0, 0, (u64)(L - 1)}); m.occ = 1; return; }
  if (t.multi_exact_noC) {
    for (u64 s : t.multi_sites) m.v.push_back({s, 0, 0, (u64)(L - 1)});
    m.occ = (int)m.v.size();
    return;
  }
  if (!t.extra && (t.cand.size() == 1 || (t.cand.size() == 2 && t.cand[0] == t.cand[1]))) {
    m.v.push_back({t.cand[0], 0, 1, (u64)(L - 1)}); m.occ = 1; return;
  }
  if (!t.cand.empty()) { std::vector<u64> c = t.cand; std::sort(c.begin(), c.end()); votes_from_sorted(c, k, m.v); }
  else m.occ = 0;
}

// C5: Schema.cpp:16052-16180.  Keeps, in order and without repeats, the entries of both lists that have
// a partner within [dmin, dmax] (distances between window starts; int bounds, u64 sites).
inline void filter_pairs(std::vector<Vote>& a, std::vector<Vote>& b, int dmax, int dmin) {
  std::vector<Vote> ka, kb;
  long long first = 0;
  for (long long i = 0; i < (long long)a.size(); ++i) {
    for (long long j = first; j < (long long)b.size(); ++j) {
      bool in = false;
      if (a[i].site > b[j].site) {
        long long d = (long long)(a[i].site - b[j].site);
        if (d > dmax) first = j + 1; else if (d >= dmin) in = true;
      } else {
        long long d = (long long)(b[j].site - a[i].site);
        if (d > dmax) break;
        if (d >= dmin) in = true;
      }
      if (in) {
        if (ka.empty() || a[i].site > ka.back().site) ka.push_back(a[i]);
        if (kb.empty() || b[j].site > kb.back().site) kb.push_back(b[j]);
      }
    }
  }
  a.swap(ka); b.swap(kb);
}

// Schema.cpp:16186-16288: keep the entries of `b` that pair with one of the first `na` hits of `a`.
inline void filter_single_side(const std::vector<Vote>& a, int na, std::vector<Vote>& b, int dmax, int dmin) {
  long long first = 0; size_t kept = 0;
  for (long long i = 0; i < na; ++i) {
    for (long long j = first; j < (long long)b.size(); ++j) {
      bool in = false;
      if (a[i].site > b[j].site) {
        long long d = (long long)(a[i].site - b[j].site);
        if (d > dmax) first = j + 1; else if (d >= dmin) in = true;
      } else {
        long long d = (long long)(b[j].site - a[i].site);
        if (d > dmax) break;
        if (d >= dmin) in = true;
      }
      if (in) { b[kept].site = b[j].site; b[kept].err = b[j].err; b[kept].end_site = b[j].end_site; ++kept; first = j + 1; }
    }
  }
  if (na > 0) b.resize(kept); else b.clear();
}

// V4: verify every candidate in site order, keep hits (err <= k) whose absolute end differs from that of
// the candidate just before it (hit or not).  Schema.cpp:7334-7698.  Returns the number of hits, compacted
// to the front of `v`.
inline int verify_keep_hits(const Index& ix, const char* read, int L, u64 k, std::vector<Vote>& v) {
  std::vector<char> win; int kept = 0; u64 prev_end = (u64)-1;
  for (size_t i = 0; i < v.size(); ++i) {
    verify_one(ix, read, L, k, v[i], win);
    u64 end_abs = v[i].site + v[i].end_site;
    if (v[i].err <= k && prev_end != end_abs) { v[kept].site = v[i].site; v[kept].err = v[i].err; v[kept].end_site = v[i].end_site; ++kept; }
    prev_end = end_abs;
  }
  return kept;
}

// V5: Schema.cpp:15773-15959.
struct PairPick { int n = 0; long long i1 = 0, i2 = 0; u32 second_best_diff = 0; };
inline PairPick pick_pair(const std::vector<Vote>& a, int na, const std::vector<Vote>& b, int nb, int k_large, int dmax, int dmin) {
  PairPick r; int best = 4 * k_large + 2; long long second = (long long)best * 2; long long first = 0;
  if (na > 0 && nb > 0) {
    for (int i = 0; i < na; ++i) {
      for (int j = (int)first; j < nb; ++j) {
        bool in = false;
        if (a[i].site > b[j].site) {
          long long d = (long long)(a[i].site - b[j].site);
          if (d > dmax) first = j + 1; else if (d >= dmin) in = true;
        } else {
          long long d = (long long)(b[j].site - a[i].site);
          if (d > dmax) break;
          if (d >= dmin) in = true;
        }
        if (!in) continue;
        long long sum = (long long)a[i].err + b[j].err;
        if (sum < best) { second = best; best = (int)sum; r.i1 = i; r.i2 = j; r.n = 1; }
        else if (sum == best) { second = best; ++r.n; if (best == 0) { r.second_best_diff = 0; return r; } }
      }
    }
  }
  if (r.n != 0) r.second_best_diff = (u32)(second - best);
  return r;
}

struct MateResult { int flag = 0; size_t chrom = 0; u64 pos = 0; u64 origin = 0, end_site = 0; u32 err = 0; int score = 0; std::string cigar; int span = 0; };

// calculate_best_map_cigar_end_to_end_return + output_sam_end_to_end_return, Schema.cpp:14602-14697, :9188-9244
inline void finish_mate(const Index& ix, const Params& P, const char* read, const char* qual, int L, u64 k, const Vote& hit,
                        bool reverse_quality, MateResult& r) {
  r.origin = hit.site; r.end_site = hit.end_site; r.err = hit.err;
  int start;
  if (hit.err != 0) {
    const int plen = L + 2 * (int)k; std::vector<char> win(plen + 8);
    ix.genome.window(hit.site, plen, win.data());
    bmbs::Refined rf;
    bmbs::refine_alignment(win.data(), plen, read, L, (int)k, (int)hit.end_site, hit.err, hit.site < ix.N, qual, reverse_quality, P.sc, rf);
    r.end_site = rf.end_site; r.err = rf.err; r.score = rf.score; r.cigar = rf.cigar; start = rf.start_site;
  } else { r.score = 0; start = (int)(hit.end_site + 1 - L); r.cigar = std::to_string(L) + "M"; }
  bmbs::Placed p = bmbs::place(ix.chroms, r.origin, (u64)(long long)start, r.end_site);
  r.flag = p.flag; r.chrom = p.chrom; r.pos = p.pos;
  r.span = hit.err != 0 ? (int)(r.end_site - start + 1) : L;
}

inline long long tlen_of(long long p1, long long l1, long long p2, long long l2) {  // Schema.h:1587-1600
  long long lo = p1 < p2 ? p1 : p2, hi = p1 + l1 - 1;
  if (hi < p2 + l2 - 1) hi = p2 + l2 - 1;
  return hi - lo + 1;
}

struct PairOutcome { int n_pairs = 0; bool written = false; MateResult m1, m2; int mapq = 0; long long tlen = 0; MateCands c1, c2; PairPick pick; };

// One pair, fast mode.  read2 is the reverse complement of the second FASTQ record, qual2 its qualities as stored.
inline void map_pair_fast(const Index& ix, const Params& P, const char* read1, const char* qual1, int L1,
                          const char* read2, const char* qual2, int L2, PairOutcome& o, Stats& st) {
  o = PairOutcome(); ++st.reads;
  const u64 k1 = u64_k(P.e_rate, L1), k2 = u64_k(P.e_rate, L2), kl = k1 > k2 ? k1 : k2;
  const int maxlen = L1 > L2 ? L1 : L2;
  const int dmax = (int)((u64)P.max_ins + kl * 2), dmin = (int)((u64)P.min_ins - kl * 2 - (u64)maxlen);
  mate_candidates(ix, P, read1, L1, k1, o.c1);
  mate_candidates(ix, P, read2, L2, k2, o.c2);
  int occ1 = o.c1.occ, occ2 = o.c2.occ;
  std::vector<Vote>& v1 = o.c1.v; std::vector<Vote>& v2 = o.c2.v;
  if (occ1 > 0 && occ2 > 0) { occ1 = (int)v1.size(); occ2 = (int)v2.size(); }
  else {
    if (occ1 == 0 || occ2 == 0) return;
    filter_pairs(v1, v2, dmax, dmin);
    if (v1.empty() || v2.empty()) return;
    if (occ1 == -1 && occ2 == -1) {
      if (v1.size() <= v2.size()) {
        occ1 = verify_keep_hits(ix, read1, L1, k1, v1);
        if (occ1 == 0) return;
        filter_single_side(v1, occ1, v2, dmax, dmin);
        occ2 = verify_keep_hits(ix, read2, L2, k2, v2);
      } else {
        occ2 = verify_keep_hits(ix, read2, L2, k2, v2);
        if (occ2 == 0) return;
        filter_single_side(v2, occ2, v1, dmax, dmin);
        occ1 = verify_keep_hits(ix, read1, L1, k1, v1);
      }
    } else if (occ1 != -1) {
      if ((int)v1.size() < occ1) occ1 = (int)v1.size();
      if (occ2 == -1) occ2 = verify_keep_hits(ix, read2, L2, k2, v2);
    } else if (occ2 != -1) {
      if ((int)v2.size() < occ2) occ2 = (int)v2.size();
      if (occ1 == -1) occ1 = verify_keep_hits(ix, read1, L1, k1, v1);
    }
  }
  o.pick = pick_pair(v1, occ1, v2, occ2, (int)kl, dmax, dmin);
  o.n_pairs = o.pick.n;
  if (o.n_pairs == 1) {
    finish_mate(ix, P, read1, qual1, L1, k1, v1[o.pick.i1], false, o.m1);
    finish_mate(ix, P, read2, qual2, L2, k2, v2[o.pick.i2], true, o.m2);
    o.tlen = tlen_of((long long)o.m1.pos, o.m1.span, (long long)o.m2.pos, o.m2.span);
    int t = (int)o.tlen;
    if (t <= P.max_ins && t >= P.min_ins && o.m1.pos + o.m1.span <= ix.chroms.len[o.m1.chrom] + 1 &&
        o.m2.pos + o.m2.span <= ix.chroms.len[o.m2.chrom] + 1) {
      ++st.unique; st.bases += L1 + L2; st.err_bases += o.m1.err + o.m2.err;
      o.mapq = bmbs::mapq_from(o.pick.second_best_diff, (u32)(k1 + k2), o.m1.score + o.m2.score, P.sc);
      o.written = true;
    }
  } else if (o.n_pairs > 1) ++st.ambiguous;
}

struct SensDebug;
bool run_pe_sensitive_pair(const Index&, const Params&, const char*, const char*, int, const char*, const char*, int, PairOutcome&, Stats&, SensDebug* dbg = nullptr);

}  // namespace oracle

namespace oracle {
inline bool run_pe(const Index& ix, const Params& P, const char* f1, const char* f2, const char* outp, std::string& out, Stats& st) {
  bmbs::FastqReader q1, q2;
  if (!q1.open(f1) || !q2.open(f2)) { fprintf(stderr, "cannot open reads\n"); return false; }
  FILE* fo = fopen(outp, "w"); if (!fo) return false;
  bmbs::FastqRecord a, b; PairOutcome o;
  while (q1.next(a) && q2.next(b)) {
    bmbs::cut_name_pe(a.name, b.name);
    std::string seq2 = bmbs::revcomp(b.seq);
    if (P.sensitive) run_pe_sensitive_pair(ix, P, a.seq.c_str(), a.qual.c_str(), (int)a.seq.size(), seq2.c_str(), b.qual.c_str(), (int)seq2.size(), o, st);
    else map_pair_fast(ix, P, a.seq.c_str(), a.qual.c_str(), (int)a.seq.size(), seq2.c_str(), b.qual.c_str(), (int)seq2.size(), o, st);
    if (o.written) {
      bmbs::sam_record_pe(out, true, a.name, a.seq, bmbs::revcomp(a.seq), a.qual, ix.chroms, o.m1.flag, o.m1.chrom, o.m1.pos, o.mapq, o.m1.cigar,
                          o.m2.pos, o.tlen, o.m1.err);
      bmbs::sam_record_pe(out, false, b.name, seq2, b.seq, b.qual, ix.chroms, o.m2.flag, o.m2.chrom, o.m2.pos, o.mapq, o.m2.cigar,
                          o.m1.pos, o.tlen, o.m2.err);
    }
    if (out.size() > (1u << 20)) { fwrite(out.data(), 1, out.size(), fo); out.clear(); }
  }
  fwrite(out.data(), 1, out.size(), fo); fclose(fo);
  return true;
}
}  // namespace oracle
