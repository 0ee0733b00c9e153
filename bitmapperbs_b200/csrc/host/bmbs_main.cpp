// bmbs: BitMapperBS-compatible command line over the GPU seed-and-verify library.
//
// Same flags as the reference for the supported modes (Process_CommandLines.cpp:88-132):
//   --index <genome.fa>                          build the on-disk index (CPU; same files as the reference)
//   --search <genome.fa> --seq <r.fq[.gz]>       single-end mapping
//   --search <genome.fa> --seq1 <a> --seq2 <b> --pe   paired-end mapping (fast mode)
//   -o <out.sam>  -t <host threads>  -e <rate>  --seed <len>  --min/--max <insert>  --mapstats <file>
//   --unmapped_out  --ambiguous_out  --bam  --pbat
//   --mp_max/--mp_min/--np/--gap_open/--gap_extension  --phred33/--phred64   -g/--gpus <n>
// Records are written in input order (the reference's `-t 1` order).
// Pipeline: block splitter -> FASTQ parse workers -> GPU threads (three batches in flight per device: H2D, kernels,
// D2H through include/bmbs.h) -> host finishing workers (reduction, CIGAR, MAPQ, SAM text) -> ordered writer.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <sys/stat.h>
#include <unistd.h>
#include "../../../include/bmbs.h"
#include "../../indexer/build_index.hpp"
#include "bam.hpp"
#include "mapper.hpp"

using namespace bmbs;

namespace {

struct Options {
  std::string mode, genome, seq, seq1, seq2, out = "output", mapstats;
  bool pe = false, sensitive = false, unmapped_out = false, ambiguous_out = false, bam = false, pbat = false;
  int threads = 1, gpus = 1;
  size_t batch_reads = 1 << 15;                      // reads (pairs) per batch
  bmbs_params prm; Scoring sc;
};

struct RawBatch { size_t seq_no = 0; size_t n_rec = 0; std::string_view raw1, raw2; std::string own1, own2; };   // views into the mapped files (own*: gzip input)

// What one device launch returned.  Several sub-blocks go to the device together (a 32 k-read sub-block does not fill a B200);
// they share the launch's page-locked result buffers and each looks at its own range of reads.  Recycled through a pool:
// page-locking memory costs more than the copies into it.
struct GpuResult {
  bmbs_final* fin = nullptr; uint16_t* mism = nullptr; bmbs_read_result* res = nullptr; bmbs_cand* cand = nullptr;
  size_t fin_cap = 0, mism_cap = 0, res_cap = 0, cand_cap = 0;
  std::vector<bmbs_refine_result> dp_res; std::vector<uint32_t> dp_ops;   // the launch's banded DPs (BMBS_FIN_DP reads), in read order
  template <class T> static void grow(T*& p, size_t& cap, size_t need) {
    if (need <= cap) return;
    bmbs_pinned_free(p); cap = 2 * need + 1024; p = (T*)bmbs_pinned_alloc(cap * sizeof(T));
    if (!p) { fprintf(stderr, "bmbs: cannot allocate page-locked result buffers\n"); exit(1); }
  }
  ~GpuResult() { bmbs_pinned_free(fin); bmbs_pinned_free(mism); bmbs_pinned_free(res); bmbs_pinned_free(cand); }
};

struct Batch {
  size_t seq_no = 0; int n = 0;                      // n reads (SE) or mates (PE, even)
  std::string_view raw1, raw2; std::string own1, own2; // FASTQ text the views below point into (bases upper-cased in place)
  std::vector<std::string_view> name, qual, fq_seq;  // per read / mate; fq_seq: the sequence as it stands in the FASTQ record
  std::string flat; std::vector<uint64_t> offsets;   // sequences as aligned (mate 2 reverse-complemented), back to back
  // results of the launch this sub-block was part of: read u of the sub-block is read r0 + u of the launch; slices of cand[] and
  // mism[] are addressed by the records themselves.  final: single end, finished records of the device (cand: handed-back lists)
  std::shared_ptr<GpuResult> gr; size_t r0 = 0, dp_first = 0; bool final = false;   // dp_first: the sub-block's first entry of gr->dp_res
  std::string sam; MapStats st;
  std::string_view seq(int i) const { return std::string_view(flat.data() + offsets[i], (size_t)(offsets[i + 1] - offsets[i])); }
};

template <class T> class Channel {
 public:
  explicit Channel(size_t cap) : cap_(cap) {}
  void push(T v) { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [&] { return q_.size() < cap_; }); q_.push_back(std::move(v)); cv_.notify_all(); }
  bool pop(T& v) { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [&] { return !q_.empty() || closed_; }); if (q_.empty()) return false; v = std::move(q_.front()); q_.pop_front(); cv_.notify_all(); return true; }
  void close() { std::lock_guard<std::mutex> l(m_); closed_ = true; cv_.notify_all(); }
  // without waiting: used for the pool of written-out batches whose buffers the parsers take over
  bool try_push(T& v) { std::lock_guard<std::mutex> l(m_); if (q_.size() >= cap_) return false; q_.push_back(std::move(v)); return true; }
  bool try_pop(T& v) { std::lock_guard<std::mutex> l(m_); if (q_.empty()) return false; v = std::move(q_.front()); q_.pop_front(); cv_.notify_all(); return true; }
 private:
  std::mutex m_; std::condition_variable cv_; std::deque<T> q_; size_t cap_; bool closed_ = false;
};

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

[[noreturn]] void die(const std::string& m) { fprintf(stderr, "bmbs: %s\n", m.c_str()); exit(1); }

void parse(int argc, char** argv, Options& o) {
  bmbs_params_default(&o.prm);
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto val = [&]() -> std::string { if (i + 1 >= argc) die("missing value for " + a); return argv[++i]; };
    if (a == "--index" || a == "-i") { o.mode = "index"; o.genome = val(); }
    else if (a == "--search") { o.mode = "search"; o.genome = val(); }
    else if (a == "--seq") o.seq = val();
    else if (a == "--seq1") o.seq1 = val();
    else if (a == "--seq2") o.seq2 = val();
    else if (a == "--pe") o.pe = true;
    else if (a == "--sensitive") o.sensitive = true;
    else if (a == "--fast") o.sensitive = false;
    else if (a == "--unmapped_out") o.unmapped_out = true;
    else if (a == "--bam") o.bam = true;
    else if (a == "--pbat") o.pbat = true;
    else if (a == "--sam") o.bam = false;
    else if (a == "--ambiguous_out") o.ambiguous_out = true;
    else if (a == "-o") o.out = val();
    else if (a == "-t" || a == "--threads") o.threads = atoi(val().c_str());
    else if (a == "-e") o.prm.e_rate = atof(val().c_str());
    else if (a == "--seed") o.prm.seed_len = atoi(val().c_str());
    else if (a == "--min") o.prm.min_ins = atoi(val().c_str());
    else if (a == "--max") o.prm.max_ins = atoi(val().c_str());
    else if (a == "--mapstats") o.mapstats = val();
    else if (a == "--mp_max") o.sc.mp_max = atoi(val().c_str());
    else if (a == "--mp_min") o.sc.mp_min = atoi(val().c_str());
    else if (a == "--np") o.sc.n_pen = atoi(val().c_str());
    else if (a == "--gap_open") o.sc.gap_open = atoi(val().c_str());
    else if (a == "--gap_extension") o.sc.gap_ext = atoi(val().c_str());
    else if (a == "--phred33") o.sc.q_base = 33;
    else if (a == "--phred64") o.sc.q_base = 64;
    else if (a == "-g" || a == "--gpus") o.gpus = atoi(val().c_str());
    else if (a == "--batch") o.batch_reads = (size_t)atoll(val().c_str());
    else die("unknown or unsupported option " + a + " (supported: --index --search --seq --seq1 --seq2 --pe --fast --sensitive --unmapped_out --ambiguous_out --bam --sam --pbat -o -t -e --seed --min --max --mapstats scoring flags --phred33/64 --gpus --batch)");
  }
  if (!o.seq1.empty() && !o.seq2.empty()) o.pe = true;   // Process_CommandLines.cpp:314-317
  if (o.pbat && o.pe) std::swap(o.seq1, o.seq2);           // exchange_two_reads, Bitmapper_main.cpp:169-172
  if (o.threads < 1) o.threads = 1;
  unsigned hw = std::thread::hardware_concurrency();
  if (hw && (unsigned)o.threads > hw) o.threads = (int)hw;  // the reference caps -t at the online CPUs (:254-258)
  if (o.gpus < 1) o.gpus = 1;
  o.prm.sensitive = (o.sensitive && o.pe) ? 1 : 0;
  o.prm.ambiguous_out = o.ambiguous_out ? 1 : 0;      // --sensitive only selects the pair worker (Bitmapper_main.cpp)
}

// Bases are upper-cased in place (Process_Reads.cpp:321-472).  Lower-case letters have bit 5 set and A/C/G/T/N do not, so
// eight bytes at a time are skipped unless one of them could be lower case; nothing is written to clean input (the
// mapped file stays shared with the page cache).
inline void upper_in_place(std::string_view v) {
  char* p = const_cast<char*>(v.data());
  const size_t n = v.size(); size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    unsigned long long x; memcpy(&x, p + i, 8);
    if (!(x & 0x2020202020202020ull)) continue;
    for (size_t j = i; j < i + 8; ++j) { const char c = p[j]; if (c >= 'a' && c <= 'z') p[j] = (char)(c - 32); }
  }
  for (; i < n; ++i) { const char c = p[i]; if (c >= 'a' && c <= 'z') p[i] = (char)(c - 32); }
}

// FASTQ text -> batch (names cut, bases upper-cased, mate 2 reverse-complemented for alignment): Process_Reads.cpp:62-90, :321-472
void parse_batch(RawBatch& rb, bool pe, bool pbat_se, Batch& b) {
  b.seq_no = rb.seq_no;
  // moving a std::string keeps its heap buffer, so views into own* stay valid; short strings live inside the object: re-point
  const bool o1 = rb.raw1.data() == rb.own1.data(), o2 = rb.raw2.data() == rb.own2.data();
  b.own1 = std::move(rb.own1); b.own2 = std::move(rb.own2);
  b.raw1 = o1 ? std::string_view(b.own1) : rb.raw1; b.raw2 = o2 ? std::string_view(b.own2) : rb.raw2;
  if (rb.n_rec == SIZE_MAX) {      // a block cut by bytes (single end, mapped input): its records are counted here
    size_t lines = count_newlines(b.raw1.data(), b.raw1.size());
    if (!b.raw1.empty() && b.raw1.back() != '\n') ++lines;
    rb.n_rec = lines / 4;
  }
  const size_t n = rb.n_rec * (pe ? 2 : 1);
  b.n = (int)n;
  b.name.resize(n); b.qual.resize(n); b.fq_seq.resize(n);
  b.offsets.resize(n + 1); b.offsets[0] = 0;
  b.flat.clear(); b.flat.reserve((pe ? b.raw1.size() + b.raw2.size() : b.raw1.size()) / 2 + 64);
  const char* p1 = b.raw1.data(); const char* e1 = p1 + b.raw1.size();
  const char* p2 = b.raw2.data(); const char* e2 = p2 + b.raw2.size();
  auto record = [](const char*& p, const char* e, std::string_view& name, std::string_view& seq, std::string_view& qual) {
    name = next_line(p, e); seq = next_line(p, e); next_line(p, e); qual = next_line(p, e);
    if (!name.empty() && name[0] == '@') name.remove_prefix(1);
    upper_in_place(seq);
    if (qual.size() > seq.size()) qual = qual.substr(0, seq.size());
    // every later stage indexes the qualities by read position: a record whose quality line is shorter than its sequence
    // (truncated / malformed input) is refused rather than read past its end
    if (qual.size() < seq.size()) die("malformed FASTQ record '" + std::string(name) + "': quality line shorter than the sequence");
  };
  for (size_t u = 0; u < rb.n_rec; ++u) {
    if (!pe) {
      record(p1, e1, b.name[u], b.fq_seq[u], b.qual[u]);
      cut_name_se(b.name[u]);
      if (!pbat_se) b.flat.append(b.fq_seq[u]);
      else {      // --pbat: the reverse complement is aligned (post_process_single_reads_pbat, Process_Reads.cpp:476)
        const std::string_view s1 = b.fq_seq[u]; const size_t at = b.flat.size(); b.flat.resize(at + s1.size());
        const char* comp = complement_table();
        for (size_t t = 0; t < s1.size(); ++t) b.flat[at + t] = comp[(unsigned char)s1[s1.size() - 1 - t]];
      }
      b.offsets[u + 1] = b.flat.size();
    } else {
      const size_t i = 2 * u, k = i + 1;
      record(p1, e1, b.name[i], b.fq_seq[i], b.qual[i]);
      record(p2, e2, b.name[k], b.fq_seq[k], b.qual[k]);
      cut_name_pe(b.name[i], b.name[k]);
      b.flat.append(b.fq_seq[i]); b.offsets[i + 1] = b.flat.size();
      const std::string_view s2 = b.fq_seq[k];
      const size_t at = b.flat.size(); b.flat.resize(at + s2.size());
      const char* comp = complement_table();
      for (size_t t = 0; t < s2.size(); ++t) b.flat[at + t] = comp[(unsigned char)s2[s2.size() - 1 - t]];
      b.offsets[k + 1] = b.flat.size();
    }
  }
}

// Finishing of one sub-block.  The banded DPs (alignments with indels) of the whole sub-block run on the GPU in one
// bmbs_refine call: the first pass writes the SAM text of every unit that needs none and collects the requests of the
// others, the second pass redoes only those units with the results, and their text is spliced back in input order.
struct FinishScratch { std::vector<HostHit> v1, v2; std::vector<char> win; DpQueue dq; std::string side; std::vector<uint32_t> unit; std::vector<size_t> at, side_end; double refine_s = 0; };

void finish_batch(const HostContext& hc, Batch& b, bool pe, bool unmapped_out, bmbs_refiner* refiner, FinishScratch& fs, long long& n_dp) {
  const int units = pe ? b.n / 2 : b.n;
  b.sam.clear(); b.sam.reserve((size_t)units * (pe ? 900 : 400));
  DpQueue& dq = fs.dq; dq.clear();
  fs.unit.clear(); fs.at.clear();
  const GpuResult& gr = *b.gr;
  const bmbs_final* fin = b.final ? gr.fin + b.r0 : nullptr;
  const bmbs_read_result* res = b.final ? nullptr : gr.res + b.r0;
  auto one = [&](int u, std::string& out, MapStats& st) {
    if (!pe) {
      ReadView rv{b.name[u], b.seq(u), b.qual[u], b.fq_seq[u]};
      if (fin) finish_single_final(hc, rv, fin[u], gr.mism, gr.cand, out, st, fs.v1, fs.win, &dq);
      else finish_single(hc, rv, res[u], gr.cand, out, st, fs.v1, fs.win, &dq);
    } else if (fin) {
      finish_pair_final(hc, b.name[2 * u], b.seq(2 * u), b.qual[2 * u], b.name[2 * u + 1], b.seq(2 * u + 1), b.fq_seq[2 * u + 1], b.qual[2 * u + 1],
                        fin[2 * u], fin[2 * u + 1], gr.mism, out, st, fs.win, &dq);
    } else {
      finish_pair(hc, b.name[2 * u], b.seq(2 * u), b.qual[2 * u], b.name[2 * u + 1], b.seq(2 * u + 1), b.fq_seq[2 * u + 1], b.qual[2 * u + 1],
                  res[2 * u], res[2 * u + 1], gr.cand, out, st, fs.v1, fs.v2, fs.win, &dq);
    }
  };
  // --unmapped_out: a read (pair) that was counted neither as unique nor as ambiguous gets flag-4 (77 / 141) records where its
  // alignment would have stood (Schema.cpp:27087-27096 / :13041 single end, :21864-21880 / :10392 paired end; this includes
  // unique hits dropped because they run over a chromosome end)
  auto unmapped = [&](int u, std::string& out, const MapStats& t) {
    if (!unmapped_out || t.unique || t.ambiguous) return;
    if (!pe) sam_record_unmapped(out, b.name[u], 4, b.fq_seq[u], b.qual[u]);
    else { sam_record_unmapped(out, b.name[2 * u], 77, b.fq_seq[2 * u], b.qual[2 * u]); sam_record_unmapped(out, b.name[2 * u + 1], 141, b.fq_seq[2 * u + 1], b.qual[2 * u + 1]); }
  };
  auto add = [&](const MapStats& t) { b.st.reads += t.reads; b.st.unique += t.unique; b.st.ambiguous += t.ambiguous; b.st.bases += t.bases; b.st.err_bases += t.err_bases; };
  if (fin) {
    // behind the device finishing: the GPU thread already ran the banded DPs of the whole launch (it knows which reads need one
    // from their records), so the records are written once, in order, and nothing here waits for the device
    dq.replay_from(gr.dp_res.data(), gr.dp_ops.data(), b.dp_first);
    for (int u = 0; u < units; ++u) { MapStats t; one(u, b.sam, t); add(t); unmapped(u, b.sam, t); }
    return;
  }
  for (int u = 0; u < units; ++u) {
    const size_t mark = b.sam.size();
    MapStats t; dq.pending = false;
    one(u, b.sam, t);
    if (dq.pending) { b.sam.resize(mark); fs.unit.push_back((uint32_t)u); fs.at.push_back(mark); } else { add(t); unmapped(u, b.sam, t); }
  }
  if (fs.unit.empty()) return;
  n_dp += (long long)dq.items.size();
  dq.res.resize(dq.items.size()); dq.ops.resize(dq.ops_bound);
  const bmbs_scoring sc{hc.sc.mp_max, hc.sc.mp_min, hc.sc.n_pen, hc.sc.gap_open, hc.sc.gap_ext, hc.sc.q_base};
  size_t used = 0;
  const double tr = now();
  if (bmbs_refine(refiner, dq.seqs.data(), dq.quals.data(), dq.seqs.size(), dq.items.data(), dq.items.size(), &sc, dq.res.data(), dq.ops.data(), dq.ops.size(), &used))
    die(std::string("gpu CIGAR refinement failed: ") + bmbs_last_error());
  fs.refine_s += now() - tr;
  dq.mode = DpQueue::REPLAY; dq.next = 0;
  fs.side.clear(); fs.side_end.clear();
  for (uint32_t u : fs.unit) { MapStats t; one((int)u, fs.side, t); add(t); unmapped((int)u, fs.side, t); fs.side_end.push_back(fs.side.size()); }
  std::string merged; merged.reserve(b.sam.size() + fs.side.size());
  size_t from = 0, sfrom = 0;
  for (size_t i = 0; i < fs.unit.size(); ++i) {
    merged.append(b.sam, from, fs.at[i] - from); from = fs.at[i];
    merged.append(fs.side, sfrom, fs.side_end[i] - sfrom); sfrom = fs.side_end[i];
  }
  merged.append(b.sam, from, std::string::npos);
  b.sam.swap(merged);
}

void print_stats(FILE* f, const MapStats& st) {
  const long long n = (long long)st.reads, u = (long long)st.unique, a = (long long)st.ambiguous, un = n - u - a;
  fprintf(f, "%-48s%lld\n", "No. of Reads:", n);
  fprintf(f, "%-48s%lld (%0.2f%%)\n", "No. of Unique Mapped Reads:", u, ((double)u / (double)n) * 100);
  fprintf(f, "%-48s%lld (%0.2f%%)\n", "No. of Ambiguous Mapped Reads:", a, ((double)a / (double)n) * 100);
  fprintf(f, "%-48s%lld (%0.2f%%)\n", "No. of Unmapped Reads:", un, ((double)un / (double)n) * 100);
  fprintf(f, "%-47s %0.2f%%\n", "Mismatch and Indel Rate:", ((double)st.err_bases / (double)st.bases) * 100);
}

int search(const Options& o, const std::string& cmdline) {
  const double t0 = now();
  HostContext hc; hc.sc = o.sc; hc.prm = o.prm; hc.ambiguous_out = o.ambiguous_out; hc.pbat = o.pbat && !o.pe;
  const std::string prefix = o.genome + ".index";
  if (!hc.chroms.load(prefix)) die("cannot open " + prefix);
  if (!hc.genome.load(prefix + ".bs.pac", hc.chroms.N)) die("cannot open " + prefix + ".bs.pac");
  std::vector<int> devs; for (int g = 0; g < o.gpus; ++g) devs.push_back(g);
  bmbs_index* idx = nullptr;
  if (bmbs_index_load(prefix.c_str(), devs.data(), (int)devs.size(), &idx)) die(std::string("index load failed: ") + bmbs_last_error());
  const double t_load = now() - t0;

  FastqBlockReader q1, q2;
  const bool pe = o.pe;
  if (pe) { if (!q1.open(o.seq1) || !q2.open(o.seq2)) die("cannot open read files"); }
  else if (!q1.open(o.seq)) die("cannot open " + o.seq);
  FILE* fo = fopen(o.out.c_str(), "w");
  if (!fo) die("cannot write " + o.out);
  BamWriter bam;
  {
    std::string h; sam_header(h, hc.chroms, cmdline);
    if (o.bam) { std::string raw, z; bam.header(hc.chroms, h, raw); if (!BamWriter::bgzf(raw, z)) die("BGZF compression failed"); h.swap(z); }
    fwrite(h.data(), 1, h.size(), fo); fflush(fo);
  }
  // batches of SAM text go straight to the descriptor: no second copy through the stdio buffer
  auto write_all = [&](const char* p, size_t n) {
    while (n) { const ssize_t w = ::write(fileno(fo), p, n); if (w <= 0) die("write failed on " + o.out); p += w; n -= (size_t)w; }
  };

  // splitter -> parse workers -> GPU threads (three launches in flight per device) -> finish workers -> ordered writer
  const double t1 = now();
  // Launches in flight per device: while one launch computes, the next one's reads go up and the previous one's records come down
  // (BMBS_INFLIGHT to change).  A launch takes the sub-blocks that are waiting, up to BMBS_GROUP of them: a 32 k-read sub-block
  // does not fill a B200 (the device is 3 x faster per read on 256 k reads), but the host side sets the pace of the program and
  // larger groups only add latency and burstiness to it: measured at 10 M reads, groups of 1-2 map in 0.51-0.53 s, of 8 in 0.59 s.
  auto env_int = [](const char* k, int d) { const char* e = getenv(k); return e ? std::max(1, atoi(e)) : d; };
  const int inflight = env_int("BMBS_INFLIGHT", 3), group_max = env_int("BMBS_GROUP", 2);
  // host threads: -t counts the workers that do the per-read work; the splitter, the sequencer, the writer and the GPU threads
  // mostly wait.  Measured per 10 M reads of 150 bp: parsing 1.2 core-seconds, finishing (records already reduced on the
  // device, DPs replayed) 1.4, the GPU threads' own host work 1.3, writing 0.9.
  const int n_parse = env_int("BMBS_PARSE_THREADS", std::max(1, (3 * o.threads + 8) / 10)), n_finish = env_int("BMBS_FINISH_THREADS", std::max(1, o.threads / 2));
  const int n_gpu = inflight * (int)devs.size();
  std::atomic<long long> us_split(0), us_parse(0), us_gpu(0), us_finish(0), us_write(0), n_retry(0), n_batches(0), n_launches(0), us_dev(0), us_up(0), us_run(0), us_down(0), us_prep(0), us_refine(0);
  std::atomic<long long> us_stage[8] = {}; std::atomic<long long> n_dp_total(0);
  auto us = [](double a, double b) { return (long long)((b - a) * 1e6); };
  // result buffers of finished launches, handed back by the last sub-block that drops its reference (declared before the queues: it outlives every batch)
  struct ResultPool {
    std::mutex m; std::vector<GpuResult*> idle;
    ~ResultPool() { for (GpuResult* g : idle) delete g; }
    std::shared_ptr<GpuResult> take() {
      GpuResult* g = nullptr;
      { std::lock_guard<std::mutex> l(m); if (!idle.empty()) { g = idle.back(); idle.pop_back(); } }
      if (!g) g = new GpuResult();
      return std::shared_ptr<GpuResult>(g, [this](GpuResult* x) { std::lock_guard<std::mutex> l(m); idle.push_back(x); });
    }
  } result_pool;
  Channel<std::unique_ptr<RawBatch>> raw_q(2 * n_parse);
  Channel<std::unique_ptr<Batch>> gpu_q((size_t)(n_gpu * group_max)), fin_q((size_t)(2 * n_finish + n_gpu * group_max)), out_q(4 * n_finish);
  // batches that have been written out go back to the parsers with their buffers (a fresh 25 MB text buffer per sub-block costs
  // more in page faults than the text that goes into it)
  Channel<std::unique_ptr<Batch>> spare((size_t)(2 * n_gpu * group_max + 8 * n_finish + 2 * n_parse));
  // a plain single-end file is cut into blocks by size at record boundaries (no pass over the text in this one thread); pairs
  // and gzip input are cut by counting lines, the two files of a pair side by side (BMBS_SPLIT_SCAN=1: always by lines)
  const bool by_bytes = !pe && q1.mapped() && !getenv("BMBS_SPLIT_SCAN");
  const size_t block_bytes = by_bytes ? o.batch_reads * q1.bytes_per_record() : 0;
  std::thread splitter([&] {
    size_t seq_no = 0;
    for (;;) {
      std::unique_ptr<RawBatch> rb(new RawBatch()); rb->seq_no = seq_no++;
      const double ts = now();
      size_t n2 = 0;
      std::thread second;                              // the two files of a pair are split side by side
      if (pe) second = std::thread([&] { n2 = q2.next(o.batch_reads, rb->raw2, rb->own2); });
      if (by_bytes) rb->n_rec = q1.next_bytes(block_bytes, rb->raw1) ? SIZE_MAX : 0;      // records counted by the parse worker
      else rb->n_rec = q1.next(o.batch_reads, rb->raw1, rb->own1);
      if (pe) { second.join(); if (n2 < rb->n_rec) rb->n_rec = n2; }   // the shorter file ends the run, as in the reference's paired reader
      us_split += us(ts, now());
      if (rb->n_rec == 0) break;
      raw_q.push(std::move(rb));
    }
    raw_q.close();
  });
  std::atomic<int> live_parse(n_parse), live_gpu(n_gpu), live_finish(n_finish);
  std::vector<std::thread> pool;
  for (int t = 0; t < n_parse; ++t) pool.emplace_back([&] {
    std::unique_ptr<RawBatch> rb;
    while (raw_q.pop(rb)) {
      const double ts = now();
      std::unique_ptr<Batch> b;
      if (!spare.try_pop(b)) b.reset(new Batch());
      parse_batch(*rb, pe, o.pbat && !pe, *b);
      us_parse += us(ts, now()); gpu_q.push(std::move(b));
    }
    if (--live_parse == 0) gpu_q.close();
  });
  for (int g = 0; g < n_gpu; ++g) pool.emplace_back([&, g] {
    const int dev = devs[g % devs.size()];
    // one batch context and one page-locked staging area for the reads per GPU thread, grown on demand
    bmbs_batch* ctx = nullptr; size_t cap_reads = 0, cap_bases = 0, cap_cand = 0;
    char* h_seq = nullptr; uint64_t* h_off = nullptr; size_t h_seq_cap = 0, h_off_cap = 0;
    // the device finishes the reads (single end: reduction in the reference's order; pairs: hit compaction, single-side filter, pair
    // pick; both: ungapped CIGAR check, coordinates) and one 32-byte record per read comes back; BMBS_HOST_FINISH=1 keeps the
    // host reduction / pair pick over the full window lists
    const bool dev_finish = !getenv("BMBS_HOST_FINISH");
    // (re-creating a context or re-locking host memory costs milliseconds and serialises the GPU threads in the driver: the first
    // launch that outgrows the initial two sub-blocks gets room for a full group at once)
    const size_t full_reads = o.batch_reads * (pe ? 2 : 1) * (size_t)group_max;
    auto ensure = [&](size_t reads, size_t bases, size_t cands) {
      if (!ctx || reads > cap_reads || bases > cap_bases || cands > cap_cand) {
        if (ctx) {
          bmbs_batch_free(ctx);
          const size_t per_read = reads ? bases / reads + 1 : 160;
          reads = std::max(reads, full_reads + full_reads / 16); bases = std::max(bases, reads * per_read + 64); cands = std::max(cands, reads * 24 + (1u << 20));
        }
        cap_reads = std::max(cap_reads, reads); cap_bases = std::max(cap_bases, bases); cap_cand = std::max(cap_cand, cands);
        if (bmbs_batch_create(idx, dev, cap_reads, cap_bases, cap_cand, &ctx)) die(std::string("batch create: ") + bmbs_last_error());
      }
      if (cap_bases > h_seq_cap) { bmbs_pinned_free(h_seq); h_seq_cap = cap_bases; h_seq = (char*)bmbs_pinned_alloc(h_seq_cap); }
      if (cap_reads + 1 > h_off_cap) { bmbs_pinned_free(h_off); h_off_cap = cap_reads + 1; h_off = (uint64_t*)bmbs_pinned_alloc(h_off_cap * sizeof(uint64_t)); }
      if (!h_seq || !h_off) die("cannot allocate page-locked staging buffers");
    };
    {   // sized for two sub-blocks before the first one arrives (a short input never grows it)
      const size_t r0 = o.batch_reads * (pe ? 2 : 1) * (size_t)std::min(group_max, 2);
      ensure(r0 + r0 / 16, r0 * 160 + 64, r0 * 24 + (1u << 20));
    }
    // the banded DPs of a launch (reads whose ungapped check failed on the device) go to the device in one call from here
    bmbs_refiner* refiner = nullptr;
    if (dev_finish && bmbs_refiner_create(idx, dev, &refiner)) die(std::string("refiner create: ") + bmbs_last_error());
    DpQueue gdq; std::string rq; long long n_dp = 0;
    const bmbs_scoring bsc{hc.sc.mp_max, hc.sc.mp_min, hc.sc.n_pen, hc.sc.gap_open, hc.sc.gap_ext, hc.sc.q_base};
    std::vector<std::unique_ptr<Batch>> grp;
    std::unique_ptr<Batch> b;
    while (gpu_q.pop(b)) {
      const double ts = now();
      grp.clear(); grp.push_back(std::move(b));
      while ((int)grp.size() < group_max) { std::unique_ptr<Batch> x; if (!gpu_q.try_pop(x)) break; grp.push_back(std::move(x)); }
      n_batches += (long long)grp.size(); ++n_launches;
      size_t reads = 0, bases = 0;
      for (auto& x : grp) { reads += (size_t)x->n; bases += x->flat.size(); }
      size_t want_cand = std::max<size_t>(cap_cand, reads * 24 + (1u << 20));
      std::shared_ptr<GpuResult> gr = result_pool.take();
      for (;;) {
        const double t_a = now();
        ensure(reads, bases + 64, want_cand);
        {   // the sub-blocks' reads back to back in the staging area
          size_t r = 0, at = 0;
          for (auto& x : grp) {
            memcpy(h_seq + at, x->flat.data(), x->flat.size());
            for (int i = 0; i < x->n; ++i) h_off[r + (size_t)i] = at + x->offsets[(size_t)i];
            r += (size_t)x->n; at += x->flat.size();
          }
          h_off[r] = at;
        }
        const double t_b = now();
        int rc = bmbs_batch_upload(ctx, h_seq, h_off, (int)reads, pe ? 1 : 0);
        const double t_c = now();
        if (!rc) rc = bmbs_batch_run(ctx, &o.prm);
        if (!rc && dev_finish) rc = bmbs_batch_finish(ctx);
        const double t_d = now();
        size_t n_cand = 0, n_mism = 0, got_cand = 0, got_mism = 0;
        if (!rc) rc = bmbs_batch_output_sizes(ctx, &n_cand, &n_mism);
        if (!rc) {
          GpuResult::grow(gr->cand, gr->cand_cap, n_cand + 64);
          if (dev_finish) {
            GpuResult::grow(gr->fin, gr->fin_cap, reads); GpuResult::grow(gr->mism, gr->mism_cap, n_mism + 64);
            rc = bmbs_batch_download_final(ctx, gr->fin, gr->mism, gr->mism_cap, &got_mism, gr->cand, gr->cand_cap, &got_cand);
          } else {
            GpuResult::grow(gr->res, gr->res_cap, reads);
            rc = bmbs_batch_download(ctx, gr->res, gr->cand, gr->cand_cap, &got_cand);
          }
        }
        const double t_e = now();
        us_prep += us(t_a, t_b); us_up += us(t_b, t_c); us_run += us(t_c, t_d); us_down += us(t_d, t_e);
        if (!rc) { float ms[8]; if (!bmbs_batch_timings(ctx, ms)) { us_dev += (long long)(ms[0] * 1000); for (int q = 1; q < 8; ++q) us_stage[q] += (long long)(ms[q] * 1000); } }
        if (rc == BMBS_ERR_CAPACITY) { want_cand = std::max(cap_cand * 2, n_cand + (n_cand >> 2) + 1024); ++n_retry; continue; }
        if (rc) die(std::string("gpu batch failed: ") + bmbs_last_error());
        break;
      }
      if (dev_finish) {
        const double tr = now();
        gdq.clear();
        size_t r = 0;
        for (auto& x : grp) {
          x->dp_first = gdq.items.size();
          for (int u = 0; u < x->n; ++u) {
            const bmbs_final& f = gr->fin[r + (size_t)u];
            if (f.status != BMBS_FIN_DP) continue;
            const std::string_view sq = x->seq(u), ql = x->qual[(size_t)u];
            // the DP sees the qualities in the order of the aligned sequence: reversed for mate 2 and for --pbat single-end reads
            if (pe ? (u & 1) != 0 : hc.pbat) { rq.assign(ql.rbegin(), ql.rend()); gdq.request(f.site, sq.data(), rq.data(), (int)sq.size(), (int)f.k); }
            else gdq.request(f.site, sq.data(), ql.data(), (int)sq.size(), (int)f.k);
          }
          r += (size_t)x->n;
        }
        gr->dp_res.resize(gdq.items.size()); gr->dp_ops.resize(gdq.ops_bound);
        if (!gdq.items.empty()) {
          n_dp += (long long)gdq.items.size();
          size_t used = 0;
          if (bmbs_refine(refiner, gdq.seqs.data(), gdq.quals.data(), gdq.seqs.size(), gdq.items.data(), gdq.items.size(), &bsc, gr->dp_res.data(), gr->dp_ops.data(), gr->dp_ops.size(), &used))
            die(std::string("gpu CIGAR refinement failed: ") + bmbs_last_error());
        }
        us_refine += us(tr, now());
      }
      us_gpu += us(ts, now());
      size_t r = 0;
      for (auto& x : grp) { x->gr = gr; x->r0 = r; x->final = dev_finish; r += (size_t)x->n; fin_q.push(std::move(x)); }
    }
    n_dp_total += n_dp;
    if (refiner) bmbs_refiner_free(refiner);
    if (ctx) bmbs_batch_free(ctx);
    bmbs_pinned_free(h_seq); bmbs_pinned_free(h_off);
    if (--live_gpu == 0) fin_q.close();
  });
  for (int t = 0; t < n_finish; ++t) pool.emplace_back([&, t] {
    // host finishing (BMBS_HOST_FINISH=1): each finishing thread owns a refiner (stream + device buffers) on one of the GPUs and
    // sends its sub-block's DPs itself; behind the device finishing the GPU threads have done that already
    bmbs_refiner* refiner = nullptr;
    if (getenv("BMBS_HOST_FINISH") && bmbs_refiner_create(idx, devs[t % devs.size()], &refiner)) die(std::string("refiner create: ") + bmbs_last_error());
    FinishScratch fs; long long n_dp = 0;
    std::unique_ptr<Batch> b;
    std::string raw;
    while (fin_q.pop(b)) {
      const double ts = now();
      finish_batch(hc, *b, pe, o.unmapped_out, refiner, fs, n_dp);
      if (o.bam) {      // --bam: the sub-block's records as BGZF members (bam_prase.cpp:248-274 does this per line through htslib)
        raw.clear(); bam.records(b->sam, raw);
        std::string z; z.reserve(raw.size() / 3 + 64);
        if (!BamWriter::bgzf(raw, z)) die("BGZF compression failed");
        b->sam.swap(z);
      }
      us_finish += us(ts, now()); out_q.push(std::move(b));
    }
    if (refiner) bmbs_refiner_free(refiner);
    n_dp_total += n_dp; us_refine += (long long)(fs.refine_s * 1e6);
    if (--live_finish == 0) out_q.close();
  });

  // Output.  This thread puts the finished sub-blocks back in input order.  Into a regular file their text is then written by a
  // writer thread of its own, each block at the offset its predecessors' sizes give it (copying 350 bytes per read into the page
  // cache is the slowest stage of the program: 3.7 GB/s on the test boxes, where more writers, BMBS_WRITE_THREADS, only queue
  // on the file's lock); anything else (a pipe, /dev/null) is written here, in order.
  MapStats total;
  struct stat ost; const int ofd = fileno(fo);
  const bool positioned = fstat(ofd, &ost) == 0 && S_ISREG(ost.st_mode);
  const int n_write = positioned ? env_int("BMBS_WRITE_THREADS", 1) : 0;
  off_t out_off = positioned ? lseek(ofd, 0, SEEK_CUR) : 0;
  struct WriteJob { std::unique_ptr<Batch> b; off_t off = 0; };
  Channel<WriteJob> write_q((size_t)(4 * std::max(1, n_write)));
  auto recycle = [&](std::unique_ptr<Batch>& done) { done->sam.clear(); done->st = MapStats(); done->gr.reset(); spare.try_push(done); };
  std::vector<std::thread> writers;
  for (int t = 0; t < n_write; ++t) writers.emplace_back([&] {
    WriteJob j;
    while (write_q.pop(j)) {
      const double ts = now();
      const char* p = j.b->sam.data(); size_t n = j.b->sam.size(); off_t at = j.off;
      while (n) { const ssize_t w = ::pwrite(ofd, p, n, at); if (w <= 0) die("write failed on " + o.out); p += w; n -= (size_t)w; at += w; }
      us_write += us(ts, now());
      recycle(j.b);
    }
  });
  {
    std::map<size_t, std::unique_ptr<Batch>> pending; size_t next = 0;
    std::unique_ptr<Batch> b;
    while (out_q.pop(b)) {
      pending[b->seq_no] = std::move(b);
      while (!pending.empty() && pending.begin()->first == next) {
        std::unique_ptr<Batch> done = std::move(pending.begin()->second);
        pending.erase(pending.begin()); ++next;
        const Batch& x = *done;
        total.reads += x.st.reads; total.unique += x.st.unique; total.ambiguous += x.st.ambiguous; total.bases += x.st.bases; total.err_bases += x.st.err_bases;
        if (n_write) { WriteJob j; j.off = out_off; out_off += (off_t)x.sam.size(); j.b = std::move(done); write_q.push(std::move(j)); }
        else {
          const double ts = now();
          write_all(x.sam.data(), x.sam.size());
          us_write += us(ts, now());
          recycle(done);
        }
      }
    }
  }
  write_q.close(); for (auto& w : writers) w.join();
  if (o.bam) {
    std::string z; BamWriter::eof_marker(z);
    if (n_write) { if (::pwrite(ofd, z.data(), z.size(), out_off) != (ssize_t)z.size()) die("write failed on " + o.out); }
    else write_all(z.data(), z.size());
  }
  splitter.join(); for (auto& w : pool) w.join();
  if (getenv("BMBS_TIMING")) {
    fprintf(stderr, "[bmbs timing] gpu threads: prepare %.2f  upload %.2f  run(enqueue) %.2f  download(wait+copy) %.2f  | device %.3f s: pack %.3f seed %.3f locate %.3f votes %.3f pairfilter %.3f verify %.3f sensitive %.3f\n",
            us_prep / 1e6, us_up / 1e6, us_run / 1e6, us_down / 1e6, us_dev / 1e6, us_stage[1] / 1e6, us_stage[2] / 1e6, us_stage[3] / 1e6, us_stage[4] / 1e6, us_stage[5] / 1e6, us_stage[6] / 1e6, us_stage[7] / 1e6);
  }
  if (getenv("BMBS_TIMING"))
    fprintf(stderr, "[bmbs timing] busy seconds summed over threads: split %.2f  parse %.2f (%d thr)  gpu %.2f (%d thr, %lld sub-blocks in %lld launches, %lld capacity retries)  finish %.2f (%d thr)  banded DPs %.2f (%lld on the GPU)  write %.2f\n",
            us_split / 1e6, us_parse / 1e6, n_parse, us_gpu / 1e6, n_gpu, (long long)n_batches, (long long)n_launches, (long long)n_retry, us_finish / 1e6, n_finish, us_refine / 1e6, (long long)n_dp_total, us_write / 1e6);
  fclose(fo);
  const double t_map = now() - t1;
  bmbs_index_free(idx);
  fprintf(stderr, "-----------------------------------------------------------------------------------------------------------\n");
  fprintf(stderr, "%19s%16.2f%18.2f\n\n", "Total:", t_load, t_map);
  fprintf(stderr, "%-42s%10.2f\n", "Total Time:", t_load + t_map);
  print_stats(stderr, total);
  if (!o.mapstats.empty()) { FILE* f = fopen(o.mapstats.c_str(), "w"); if (f) { print_stats(f, total); fclose(f); } }
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  Options o; parse(argc, argv, o);
  std::string cmdline; for (int i = 0; i < argc; ++i) { cmdline += argv[i]; cmdline += ' '; }
  if (o.mode == "index") return bmbs::build_index(o.genome, o.threads > 1 ? o.threads : 0);
  if (o.mode == "search") return search(o, cmdline);
  fprintf(stderr, "usage: bmbs --index genome.fa | --search genome.fa (--seq r.fq | --seq1 a.fq --seq2 b.fq --pe) [-o out.sam] [-t N] [--gpus G]\n");
  return 2;
}
