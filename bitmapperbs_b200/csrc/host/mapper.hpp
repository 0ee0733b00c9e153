// Host side of the GPU mapper: takes the per-read records libbmbs_gpu.so returns
// (include/bmbs.h) and finishes each read the way the reference's worker does after
// its verification calls -- vote-ordered reduction, pairing, CIGAR, MAPQ, SAM text.
//
//   single end  Schema.cpp:27527-27751 (decision), :7847-8056 / :8325-8745 (reduction in std::sort order)
//   paired end  Schema.cpp:22014-22382 (Map_Pair_Seq_split_fast after get_candidates + filter_pairs),
//               :7502-7608 (hit compaction), :16186-16288 (single-side filter), :15773-15959 (pair pick)
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <string_view>
#include <vector>
#include "../../../include/bmbs.h"
#include "fastq.hpp"
#include "postprocess.hpp"
#include "sam.hpp"

namespace bmbs {

struct HostHit { uint64_t site; uint64_t vote; uint32_t err; uint64_t end_site; };  // 32 bytes, the reference's seed_votes shape

inline void unpack_cands(const bmbs_cand* c, uint32_t n, std::vector<HostHit>& v) {
  v.resize(n);
  for (uint32_t i = 0; i < n; ++i) {
    v[i].site = c[i].site; v[i].vote = c[i].vote;
    v[i].err = c[i].err == 0xFFFF ? 0xFFFFFFFFu : c[i].err;
    v[i].end_site = (uint64_t)(int64_t)c[i].end_site;
  }
}

struct ReadView { std::string_view name, seq, qual, raw; };   // seq: as aligned; raw: as it stands in the FASTQ (differs with --pbat)

struct HostContext {
  ChromTable chroms;
  Genome2bit genome;
  Scoring sc;
  bmbs_params prm;
  bool ambiguous_out = false;   // --ambiguous_out: report the first hit of an ambiguously mapped read (pair)
  // --pbat, single end (Map_Single_Seq_split_pbat, Schema.cpp:27879; reads stored by post_process_single_reads_pbat,
  // Process_Reads.cpp:476): the reverse complement of the FASTQ record is what gets aligned -- the treatment mate 2 of a pair
  // gets -- so qualities are taken in reverse for scoring and the record itself is printed for reverse-strand hits
  // (output_sam_end_to_end_pbat_output_buffer, :13215).  Paired end: the two input files change places (exchange_two_reads).
  bool pbat = false;
};

inline uint64_t threshold_k(double e_rate, size_t L) { uint64_t k = (uint64_t)(e_rate * L); return k >= 31 ? 31 : k; }

// ---- single end -------------------------------------------------------------------------------
inline void finish_single(const HostContext& hc, const ReadView& rd, const bmbs_read_result& res, const bmbs_cand* cand,
                          std::string& out, MapStats& st, std::vector<HostHit>& hits, std::vector<char>& win, DpQueue* dq = nullptr) {
  const std::string_view seq = rd.seq, qual = rd.qual;
  const int L = (int)seq.size();
  const uint64_t k = threshold_k(hc.prm.e_rate, L);
  ++st.reads;
  auto emit = [&](uint64_t site, uint64_t end_site, int start_site, unsigned nm, const std::string& cigar, int mapq) -> bool {
    Placed p = place(hc.chroms, site, (uint64_t)(int64_t)start_site, end_site);
    if (p.off_chrom) return false;
    if (!hc.pbat) sam_record_se(out, rd.name, seq, qual, hc.chroms, p, mapq, cigar, nm);
    else sam_record_se_pbat(out, rd.name, seq, rd.raw, qual, hc.chroms, p, mapq, cigar, nm);
    return true;
  };
  switch (res.state) {
    case BMBS_EXACT_UNIQUE:
      if (emit(res.site, L - 1, 0, 0, std::to_string(L) + "M", 42)) { ++st.unique; st.bases += L; }
      return;
    case BMBS_MULTI_EXACT:
      // several exact hits of a read without C.  --ambiguous_out: the first located row (suffix-array order, at most 1000)
      // that stays inside a chromosome, MAPQ 1; counted only when one was written (Schema.cpp:27216-27245, :26704-26758)
      if (!hc.ambiguous_out) { ++st.ambiguous; return; }
      for (uint32_t i = 0; i < res.n_cand; ++i)
        if (emit(cand[res.first_cand + i].site, L - 1, 0, 0, std::to_string(L) + "M", 1)) { ++st.ambiguous; return; }
      return;
    case BMBS_ONE_MISMATCH: {
      const int pos = res.one_mismatch_pos;
      int score = 0;
      if (seq[pos] == 'N') score -= hc.sc.n_pen; else score -= mismatch_penalty(hc.sc, qual[hc.pbat ? L - 1 - pos : pos]);   // :28703 pbat reads the quality from the other end
      const int mapq = mapq_from(0xFFFFFFFFu, (unsigned)k, score, hc.sc);
      if (emit(res.site, L - 1, 0, 1, std::to_string(L) + "M", mapq)) { ++st.unique; st.bases += L; st.err_bases += 1; }
      return;
    }
    case BMBS_VERIFY: break;
    default: return;
  }
  unpack_cands(cand + res.first_cand, res.n_cand, hits);
  // the reference reduces in the order its (unstable) std::sort by vote leaves; the same call on the same
  // site-ordered array reproduces that order (Schema.cpp:27612, comparator :560-563)
  std::sort(hits.begin(), hits.end(), [](const HostHit& a, const HostHit& b) { return a.vote > b.vote; });
  uint32_t min_err = 0xFFFFFFFEu, sbd = 0; int idx = -1; uint64_t best_end = ~0ull;
  const bool early = res.is_multiple_map != 0;
  for (size_t i = 0; i < hits.size(); ++i) {
    const uint32_t e = hits[i].err; const uint64_t end_abs = hits[i].site + hits[i].end_site;
    if (!early) {
      if (e == min_err && best_end != end_abs && idx >= 0) { sbd = 0; idx = -2 - idx; }
      else if (e < min_err) { sbd = min_err - e; min_err = e; idx = (int)i; best_end = end_abs; }
    } else {
      if (e == min_err && best_end != end_abs) { sbd = 0; if (idx >= 0) idx = -2 - idx; if (min_err == 0) break; }
      else if (e < min_err) { sbd = min_err - e; min_err = e; idx = (int)i; best_end = end_abs; }
    }
  }
  const bool ambiguous = idx <= -2;
  if (ambiguous) {                       // --ambiguous_out reports the first of the equally good hits (Schema.cpp:27707-27733)
    if (!hc.ambiguous_out) { ++st.ambiguous; return; }
    idx = -2 - idx;
  }
  if (idx < 0) return;
  const HostHit& b = hits[idx];
  Refined rf;
  if (b.err != 0) {
    const int plen = L + 2 * (int)k; win.resize(plen + 8);
    hc.genome.window(b.site, plen, win.data());
    refine_alignment(win.data(), plen, seq.data(), L, (int)k, (int)b.end_site, b.err, b.site < hc.chroms.N, qual.data(), hc.pbat, hc.sc, rf, b.site, dq);
  } else { rf.score = 0; rf.start_site = (int)b.end_site - L + 1; rf.end_site = b.end_site; rf.err = 0; rf.cigar = std::to_string(L) + "M"; }
  const int mapq = mapq_from(sbd, (unsigned)k, rf.score, hc.sc);
  if (emit(b.site, rf.end_site, rf.start_site, rf.err, rf.cigar, mapq)) {
    if (ambiguous) ++st.ambiguous; else { ++st.unique; st.bases += L; st.err_bases += rf.err; }
  }
}

// ---- single end, from the device's finished records (bmbs_batch_finish) --------------------------
// The reduction, the ungapped CIGAR check and the coordinates were done on the device; what is left here is the score of the
// returned mismatch positions (MismatchPenaltyByQuality needs the quality string, which stays on the host), MAPQ and the
// record text.  BMBS_FIN_DP reads take the banded DP (one bmbs_refine call per sub-block through `dq`), BMBS_FIN_HOST reads
// the host reduction above with the window list that came back.
inline void finish_single_final(const HostContext& hc, const ReadView& rd, const bmbs_final& f, const uint16_t* mism, const bmbs_cand* fb_cand,
                                std::string& out, MapStats& st, std::vector<HostHit>& hits, std::vector<char>& win, DpQueue* dq = nullptr) {
  const std::string_view seq = rd.seq, qual = rd.qual;
  const int L = (int)seq.size();
  const uint64_t k = threshold_k(hc.prm.e_rate, L);
  auto count = [&](unsigned nm) { if (f.flags & BMBS_FINF_AMBIGUOUS) ++st.ambiguous; else { ++st.unique; st.bases += L; st.err_bases += nm; } };
  auto record = [&](const Placed& p, int mapq, std::string_view cigar, unsigned nm) {
    if (!hc.pbat) sam_record_se(out, rd.name, seq, qual, hc.chroms, p, mapq, cigar, nm);
    else sam_record_se_pbat(out, rd.name, seq, rd.raw, qual, hc.chroms, p, mapq, cigar, nm);
  };
  switch (f.status) {
    case BMBS_FIN_UNIQUE: {
      ++st.reads;
      int score = 0;
      for (uint32_t j = 0; j < f.n_aux; ++j) {
        const int pos = mism[f.aux_first + j];
        if (seq[pos] == 'N') score -= hc.sc.n_pen; else score -= mismatch_penalty(hc.sc, qual[hc.pbat ? L - 1 - pos : pos]);
      }
      const int mapq = f.mapq_fixed ? f.mapq_fixed : mapq_from(f.sbd, (unsigned)k, score, hc.sc);
      Placed p; p.flag = (f.flags & BMBS_FINF_REVERSE) ? 16 : 0; p.chrom = (size_t)(f.chrom_pos >> 40); p.pos = f.chrom_pos & 0xFFFFFFFFFFull; p.off_chrom = false;
      char cg[16]; int cp = 16; cg[--cp] = 'M';
      { unsigned v = (unsigned)L; do { cg[--cp] = (char)('0' + v % 10); v /= 10; } while (v); }
      record(p, mapq, std::string_view(cg + cp, (size_t)(16 - cp)), f.nm);
      count(f.nm);
      return;
    }
    case BMBS_FIN_AMBIGUOUS: ++st.reads; ++st.ambiguous; return;
    case BMBS_FIN_DP: {
      ++st.reads;
      Refined rf;
      const bool forward = f.site < hc.chroms.N;
      if (dq && dq->mode == DpQueue::COLLECT) {        // the ungapped check already failed on the device: straight to the DP queue
        if (!hc.pbat) dq->request(f.site, seq.data(), qual.data(), L, (int)k);
        else { std::string rq(qual.rbegin(), qual.rend()); dq->request(f.site, seq.data(), rq.data(), L, (int)k); }
        return;
      }
      if (dq) dq->take(forward, rf);
      else {                                           // no device queue (tests): the whole refinement on the CPU
        const int plen = L + 2 * (int)k; win.resize(plen + 8);
        hc.genome.window(f.site, plen, win.data());
        refine_alignment(win.data(), plen, seq.data(), L, (int)k, (int)f.end_site, f.nm, forward, qual.data(), hc.pbat, hc.sc, rf, f.site, nullptr);
      }
      const int mapq = mapq_from(f.sbd, (unsigned)k, rf.score, hc.sc);
      const Placed p = place(hc.chroms, f.site, (uint64_t)(int64_t)rf.start_site, rf.end_site);
      if (p.off_chrom) return;
      record(p, mapq, rf.cigar, rf.err);
      count(rf.err);
      return;
    }
    case BMBS_FIN_HOST: {
      bmbs_read_result r{}; r.state = BMBS_VERIFY; r.first_cand = f.aux_first; r.n_cand = f.n_aux; r.is_multiple_map = (uint8_t)f.site;
      finish_single(hc, rd, r, fb_cand, out, st, hits, win, dq && dq->external ? nullptr : dq);   // (no queue to ask: the DP on the CPU)
      return;
    }
    default: ++st.reads; return;
  }
}

// ---- paired end -------------------------------------------------------------------------------
namespace pe {
// keep hits (err <= k) whose absolute end differs from the candidate right before them
inline int keep_hits(std::vector<HostHit>& v, uint64_t k) {
  int kept = 0; uint64_t prev = ~0ull;
  for (size_t i = 0; i < v.size(); ++i) {
    const uint64_t end_abs = v[i].site + v[i].end_site;
    if (v[i].err <= k && prev != end_abs) { v[kept].site = v[i].site; v[kept].err = v[i].err; v[kept].end_site = v[i].end_site; ++kept; }
    prev = end_abs;
  }
  return kept;
}
inline bool in_range(uint64_t a, uint64_t b, int dmax, int dmin, long long j, long long& first, bool& stop) {
  stop = false;
  if (a > b) { const long long d = (long long)(a - b); if (d > dmax) { first = j + 1; return false; } return d >= dmin; }
  const long long d = (long long)(b - a);
  if (d > dmax) { stop = true; return false; }
  return d >= dmin;
}
inline void single_side(const std::vector<HostHit>& a, int na, std::vector<HostHit>& b, int dmax, int dmin) {
  long long first = 0; size_t kept = 0;
  for (long long i = 0; i < na; ++i)
    for (long long j = first; j < (long long)b.size(); ++j) {
      bool stop; const bool in = in_range(a[i].site, b[j].site, dmax, dmin, j, first, stop);
      if (stop) break;
      if (in) { b[kept].site = b[j].site; b[kept].err = b[j].err; b[kept].end_site = b[j].end_site; ++kept; first = j + 1; }
    }
  b.resize(na > 0 ? kept : 0);
}
struct Pick { int n = 0; long long i1 = 0, i2 = 0; uint32_t sbd = 0; };
inline Pick pick(const std::vector<HostHit>& a, int na, const std::vector<HostHit>& b, int nb, int k_large, int dmax, int dmin) {
  Pick r; int best = 4 * k_large + 2; long long second = 2LL * best, first = 0;
  if (na > 0 && nb > 0)
    for (int i = 0; i < na; ++i)
      for (int j = (int)first; j < nb; ++j) {
        bool stop; const bool in = in_range(a[i].site, b[j].site, dmax, dmin, j, first, stop);
        if (stop) break;
        if (!in) continue;
        const long long sum = (long long)a[i].err + b[j].err;
        if (sum < best) { second = best; best = (int)sum; r.i1 = i; r.i2 = j; r.n = 1; }
        else if (sum == best) { second = best; ++r.n; if (best == 0) { r.sbd = 0; return r; } }
      }
  if (r.n) r.sbd = (uint32_t)(second - best);
  return r;
}
struct Mate { int flag = 0; size_t chrom = 0; uint64_t pos = 0; unsigned err = 0; int score = 0, span = 0; std::string cigar; };
inline void finish_mate(const HostContext& hc, std::string_view seq, std::string_view qual, uint64_t k, const HostHit& h,
                        bool reverse_quality, Mate& m, std::vector<char>& win, DpQueue* dq) {
  const int L = (int)seq.size();
  int start; uint64_t end = h.end_site;
  m.err = h.err;
  if (h.err != 0) {
    const int plen = L + 2 * (int)k; win.resize(plen + 8);
    hc.genome.window(h.site, plen, win.data());
    Refined rf;
    refine_alignment(win.data(), plen, seq.data(), L, (int)k, (int)h.end_site, h.err, h.site < hc.chroms.N, qual.data(), reverse_quality, hc.sc, rf, h.site, dq);
    end = rf.end_site; m.err = rf.err; m.score = rf.score; m.cigar = rf.cigar; start = rf.start_site;
    m.span = (int)(end - start + 1);
  } else { m.score = 0; start = (int)(h.end_site + 1 - L); m.cigar = std::to_string(L) + "M"; m.span = L; }
  const Placed p = place(hc.chroms, h.site, (uint64_t)(int64_t)start, end);
  m.flag = p.flag; m.chrom = p.chrom; m.pos = p.pos;
}
}  // namespace pe

// seq2 is mate 2 as aligned (reverse complement of the FASTQ record `raw2`), qual2 in FASTQ order.
inline void finish_pair(const HostContext& hc, std::string_view name1, std::string_view seq1, std::string_view qual1,
                        std::string_view name2, std::string_view seq2, std::string_view raw2, std::string_view qual2,
                        const bmbs_read_result& r1, const bmbs_read_result& r2, const bmbs_cand* cand,
                        std::string& out, MapStats& st, std::vector<HostHit>& v1, std::vector<HostHit>& v2, std::vector<char>& win, DpQueue* dq = nullptr) {
  ++st.reads;
  const int L1 = (int)seq1.size(), L2 = (int)seq2.size();
  const uint64_t k1 = threshold_k(hc.prm.e_rate, L1), k2 = threshold_k(hc.prm.e_rate, L2), kl = k1 > k2 ? k1 : k2;
  const int dmax = (int)((uint64_t)(long long)hc.prm.max_ins + kl * 2);
  const int dmin = (int)((uint64_t)(long long)hc.prm.min_ins - kl * 2 - (uint64_t)(L1 > L2 ? L1 : L2));
  auto resolved = [](const bmbs_read_result& r) { return r.state == BMBS_EXACT_UNIQUE || r.state == BMBS_MULTI_EXACT || r.state == BMBS_ONE_MISMATCH; };
  const bool res1 = resolved(r1), res2 = resolved(r2);
  if (r1.n_cand == 0 || r2.n_cand == 0) return;           // a mate without candidates, or nothing survived the distance filter
  unpack_cands(cand + r1.first_cand, r1.n_cand, v1);
  unpack_cands(cand + r2.first_cand, r2.n_cand, v2);
  int occ1, occ2;
  // --pe --sensitive (Map_Pair_Seq_split, Schema.cpp:22450): the library already returns each mate's final hits
  if (hc.prm.sensitive || (res1 && res2)) { occ1 = (int)v1.size(); occ2 = (int)v2.size(); }
  else if (!res1 && !res2) {
    if (v1.size() <= v2.size()) {
      occ1 = pe::keep_hits(v1, k1); if (occ1 == 0) return;
      pe::single_side(v1, occ1, v2, dmax, dmin);
      occ2 = pe::keep_hits(v2, k2);
    } else {
      occ2 = pe::keep_hits(v2, k2); if (occ2 == 0) return;
      pe::single_side(v2, occ2, v1, dmax, dmin);
      occ1 = pe::keep_hits(v1, k1);
    }
  } else if (res1) { occ1 = (int)v1.size(); occ2 = pe::keep_hits(v2, k2); }
  else { occ2 = (int)v2.size(); occ1 = pe::keep_hits(v1, k1); }
  const pe::Pick pk = pe::pick(v1, occ1, v2, occ2, (int)kl, dmax, dmin);
  if (pk.n > 1 && !hc.ambiguous_out) { ++st.ambiguous; return; }   // --ambiguous_out: the first best pair is reported (Schema.cpp:22218-22222)
  if (pk.n < 1) return;
  pe::Mate m1, m2;
  pe::finish_mate(hc, seq1, qual1, k1, v1[pk.i1], false, m1, win, dq);
  pe::finish_mate(hc, seq2, qual2, k2, v2[pk.i2], true, m2, win, dq);
  long long lo = (long long)std::min(m1.pos, m2.pos), hi = std::max((long long)m1.pos + m1.span - 1, (long long)m2.pos + m2.span - 1);
  const int tlen = (int)(hi - lo + 1);                     // calculate_TLEN, Schema.h:1587-1600
  if (!(tlen <= hc.prm.max_ins && tlen >= hc.prm.min_ins)) return;
  if (!(m1.pos + m1.span <= hc.chroms.len[m1.chrom] + 1 && m2.pos + m2.span <= hc.chroms.len[m2.chrom] + 1)) return;
  if (pk.n == 1) ++st.unique; else ++st.ambiguous;
  st.bases += L1 + L2; st.err_bases += m1.err + m2.err;
  const int mapq = mapq_from(pk.sbd, (unsigned)(k1 + k2), m1.score + m2.score, hc.sc);
  sam_record_pe(out, true, name1, seq1, std::string_view(), qual1, hc.chroms, m1.flag, m1.chrom, m1.pos, mapq, m1.cigar, m2.pos, tlen, m1.err);
  sam_record_pe(out, false, name2, seq2, raw2, qual2, hc.chroms, m2.flag, m2.chrom, m2.pos, mapq, m2.cigar, m1.pos, tlen, m2.err);
}

// ---- paired end, from the device's finished records (bmbs_batch_finish on a paired batch) -----------------------
// Hit compaction, the single-side filter and the pair pick ran on the device (finish_pe), and so did the ungapped CIGAR check
// and the coordinates of the two chosen hits; what is left here mirrors the tail of finish_pair: the score of the returned
// mismatch positions (mate 2's qualities are read from the other end: it was aligned as its reverse complement), the banded DP
// of BMBS_FIN_DP mates through `dq`, TLEN and its limits, the chromosome-end check, MAPQ, the two records.
inline void finish_pair_final(const HostContext& hc, std::string_view name1, std::string_view seq1, std::string_view qual1,
                              std::string_view name2, std::string_view seq2, std::string_view raw2, std::string_view qual2,
                              const bmbs_final& f1, const bmbs_final& f2, const uint16_t* mism,
                              std::string& out, MapStats& st, std::vector<char>& win, DpQueue* dq = nullptr) {
  ++st.reads;
  if (f1.status == BMBS_FIN_AMBIGUOUS) { ++st.ambiguous; return; }
  if (f1.status == BMBS_FIN_UNMAPPED) return;
  const int L1 = (int)seq1.size(), L2 = (int)seq2.size();
  const uint64_t k1 = threshold_k(hc.prm.e_rate, L1), k2 = threshold_k(hc.prm.e_rate, L2);
  if (dq && dq->mode == DpQueue::COLLECT) {          // the requests of this pair, mate 1 first; the records are written in the replay pass
    if (f1.status == BMBS_FIN_DP) dq->request(f1.site, seq1.data(), qual1.data(), L1, (int)k1);
    if (f2.status == BMBS_FIN_DP) { std::string rq(qual2.rbegin(), qual2.rend()); dq->request(f2.site, seq2.data(), rq.data(), L2, (int)k2); }
    return;
  }
  auto mate = [&](const bmbs_final& f, std::string_view seq, std::string_view qual, uint64_t k, bool reverse_quality, pe::Mate& m) {
    const int L = (int)seq.size();
    if (f.status == BMBS_FIN_UNIQUE) {
      int score = 0;
      for (uint32_t j = 0; j < f.n_aux; ++j) {
        const int pos = mism[f.aux_first + j];
        if (seq[pos] == 'N') score -= hc.sc.n_pen; else score -= mismatch_penalty(hc.sc, qual[reverse_quality ? L - 1 - pos : pos]);
      }
      m.err = f.nm; m.score = score; m.span = L;
      char cg[16]; int cp = 16; cg[--cp] = 'M';
      { unsigned v = (unsigned)L; do { cg[--cp] = (char)('0' + v % 10); v /= 10; } while (v); }
      m.cigar.assign(cg + cp, (size_t)(16 - cp));
      m.flag = (f.flags & BMBS_FINF_REVERSE) ? 16 : 0; m.chrom = (size_t)(f.chrom_pos >> 40); m.pos = f.chrom_pos & 0xFFFFFFFFFFull;
      return;
    }
    Refined rf;
    const bool forward = f.site < hc.chroms.N;
    if (dq) dq->take(forward, rf);
    else {                                             // no device queue (tests): the whole refinement on the CPU
      const int plen = L + 2 * (int)k; win.resize(plen + 8);
      hc.genome.window(f.site, plen, win.data());
      refine_alignment(win.data(), plen, seq.data(), L, (int)k, (int)f.end_site, f.nm, forward, qual.data(), reverse_quality, hc.sc, rf, f.site, nullptr);
    }
    m.err = rf.err; m.score = rf.score; m.cigar = rf.cigar; m.span = (int)(rf.end_site - rf.start_site + 1);
    const Placed p = place(hc.chroms, f.site, (uint64_t)(int64_t)rf.start_site, rf.end_site);
    m.flag = p.flag; m.chrom = p.chrom; m.pos = p.pos;
  };
  pe::Mate m1, m2;
  mate(f1, seq1, qual1, k1, false, m1);
  mate(f2, seq2, qual2, k2, true, m2);
  long long lo = (long long)std::min(m1.pos, m2.pos), hi = std::max((long long)m1.pos + m1.span - 1, (long long)m2.pos + m2.span - 1);
  const int tlen = (int)(hi - lo + 1);                     // calculate_TLEN, Schema.h:1587-1600
  if (!(tlen <= hc.prm.max_ins && tlen >= hc.prm.min_ins)) return;
  if (!(m1.pos + m1.span <= hc.chroms.len[m1.chrom] + 1 && m2.pos + m2.span <= hc.chroms.len[m2.chrom] + 1)) return;
  if (f1.flags & BMBS_FINF_AMBIGUOUS) ++st.ambiguous; else ++st.unique;
  st.bases += L1 + L2; st.err_bases += m1.err + m2.err;
  const unsigned sbd = f1.sbd == 255 ? 0xFFFFu : f1.sbd;
  const int mapq = mapq_from(sbd, (unsigned)(k1 + k2), m1.score + m2.score, hc.sc);
  sam_record_pe(out, true, name1, seq1, std::string_view(), qual1, hc.chroms, m1.flag, m1.chrom, m1.pos, mapq, m1.cigar, m2.pos, tlen, m1.err);
  sam_record_pe(out, false, name2, seq2, raw2, qual2, hc.chroms, m2.flag, m2.chrom, m2.pos, mapq, m2.cigar, m1.pos, tlen, m2.err);
}

}  // namespace bmbs
