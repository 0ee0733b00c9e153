// bmbs: BitMapperBS-compatible command line over the GPU seed-and-verify library.
//
// Same flags as the reference for the supported modes (Process_CommandLines.cpp:88-132):
//   --index <genome.fa>                          build the on-disk index (CPU; same files as the reference)
//   --search <genome.fa> --seq <r.fq[.gz]>       single-end mapping
//   --search <genome.fa> --seq1 <a> --seq2 <b> --pe   paired-end mapping (fast mode)
//   -o <out.sam>  -t <host threads>  -e <rate>  --seed <len>  --min/--max <insert>  --mapstats <file>
//   --mp_max/--mp_min/--np/--gap_open/--gap_extension  --phred33/--phred64   -g/--gpus <n>
// Records are written in input order (the reference's `-t 1` order).
// Pipeline: reader thread -> per-GPU worker (H2D, kernels, D2H through include/bmbs.h) ->
// host finishing threads (reduction, CIGAR, MAPQ, SAM text) -> ordered writer.
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
#include "../../../include/bmbs.h"
#include "../../indexer/build_index.hpp"
#include "mapper.hpp"

using namespace bmbs;

namespace {

struct Options {
  std::string mode, genome, seq, seq1, seq2, out = "output", mapstats;
  bool pe = false, sensitive = false;
  int threads = 1, gpus = 1;
  size_t batch_reads = 1 << 18;
  bmbs_params prm; Scoring sc;
};

struct Batch {
  size_t seq_no = 0; int n = 0;                      // n reads (SE) or mates (PE, even)
  std::vector<std::string> name, seq, qual, raw;     // raw: mate-2 FASTQ record (PE)
  std::string flat; std::vector<uint64_t> offsets;
  std::vector<bmbs_read_result> res; std::vector<bmbs_cand> cand;
  std::string sam; MapStats st;
};

template <class T> class Channel {
 public:
  explicit Channel(size_t cap) : cap_(cap) {}
  void push(T v) { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [&] { return q_.size() < cap_; }); q_.push_back(std::move(v)); cv_.notify_all(); }
  bool pop(T& v) { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [&] { return !q_.empty() || closed_; }); if (q_.empty()) return false; v = std::move(q_.front()); q_.pop_front(); cv_.notify_all(); return true; }
  void close() { std::lock_guard<std::mutex> l(m_); closed_ = true; cv_.notify_all(); }
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
    else die("unknown or unsupported option " + a + " (supported: --index --search --seq --seq1 --seq2 --pe --fast --sensitive -o -t -e --seed --min --max --mapstats scoring flags --phred33/64 --gpus --batch)");
  }
  if (!o.seq1.empty() && !o.seq2.empty()) o.pe = true;   // Process_CommandLines.cpp:314-317
  if (o.threads < 1) o.threads = 1;
  unsigned hw = std::thread::hardware_concurrency();
  if (hw && (unsigned)o.threads > hw) o.threads = (int)hw;  // the reference caps -t at the online CPUs (:254-258)
  if (o.gpus < 1) o.gpus = 1;
  o.prm.sensitive = (o.sensitive && o.pe) ? 1 : 0;      // --sensitive only selects the pair worker (Bitmapper_main.cpp)
}

void finish_batch(const HostContext& hc, Batch& b, bool pe, int threads) {
  const int units = pe ? b.n / 2 : b.n;
  const int T = std::max(1, std::min(threads, units / 256 + 1));
  std::vector<std::string> out(T); std::vector<MapStats> st(T);
  auto work = [&](int t) {
    std::vector<HostHit> v1, v2; std::vector<char> win;
    const int lo = (int)((long long)units * t / T), hi = (int)((long long)units * (t + 1) / T);
    out[t].reserve((size_t)(hi - lo) * (pe ? 900 : 400));
    for (int u = lo; u < hi; ++u) {
      if (!pe) {
        ReadView rv{&b.name[u], &b.seq[u], &b.qual[u]};
        finish_single(hc, rv, b.res[u], b.cand.data(), out[t], st[t], v1, win);
      } else {
        finish_pair(hc, b.name[2 * u], b.seq[2 * u], b.qual[2 * u], b.name[2 * u + 1], b.seq[2 * u + 1], b.raw[u], b.qual[2 * u + 1],
                    b.res[2 * u], b.res[2 * u + 1], b.cand.data(), out[t], st[t], v1, v2, win);
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < T; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
  size_t total = 0; for (auto& s : out) total += s.size();
  b.sam.clear(); b.sam.reserve(total);
  for (int t = 0; t < T; ++t) {
    b.sam += out[t];
    b.st.reads += st[t].reads; b.st.unique += st[t].unique; b.st.ambiguous += st[t].ambiguous; b.st.bases += st[t].bases; b.st.err_bases += st[t].err_bases;
  }
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
  HostContext hc; hc.sc = o.sc; hc.prm = o.prm;
  const std::string prefix = o.genome + ".index";
  if (!hc.chroms.load(prefix)) die("cannot open " + prefix);
  if (!hc.genome.load(prefix + ".bs.pac", hc.chroms.N)) die("cannot open " + prefix + ".bs.pac");
  std::vector<int> devs; for (int g = 0; g < o.gpus; ++g) devs.push_back(g);
  bmbs_index* idx = nullptr;
  if (bmbs_index_load(prefix.c_str(), devs.data(), (int)devs.size(), &idx)) die(std::string("index load failed: ") + bmbs_last_error());
  const double t_load = now() - t0;

  FastqReader q1, q2;
  const bool pe = o.pe;
  if (pe) { if (!q1.open(o.seq1) || !q2.open(o.seq2)) die("cannot open read files"); }
  else if (!q1.open(o.seq)) die("cannot open " + o.seq);
  FILE* fo = fopen(o.out.c_str(), "w");
  if (!fo) die("cannot write " + o.out);
  { std::string h; sam_header(h, hc.chroms, cmdline); fwrite(h.data(), 1, h.size(), fo); }

  const double t1 = now();
  Channel<std::unique_ptr<Batch>> to_gpu(2 * devs.size()), to_writer(4 * devs.size());
  std::thread reader([&] {
    size_t seq_no = 0;
    for (;;) {
      std::unique_ptr<Batch> b(new Batch()); b->seq_no = seq_no++;
      FastqRecord a, c;
      const size_t want = o.batch_reads;
      b->offsets.push_back(0);
      while ((size_t)(pe ? b->n / 2 : b->n) < want && q1.next(a)) {
        if (pe) {
          if (!q2.next(c)) break;
          cut_name_pe(a.name, c.name);
          std::string rc2 = revcomp(c.seq);
          b->flat += a.seq; b->offsets.push_back(b->flat.size());
          b->flat += rc2; b->offsets.push_back(b->flat.size());
          b->name.push_back(a.name); b->name.push_back(c.name);
          b->seq.push_back(std::move(a.seq)); b->seq.push_back(std::move(rc2));
          b->qual.push_back(std::move(a.qual)); b->qual.push_back(std::move(c.qual));
          b->raw.push_back(std::move(c.seq));
          b->n += 2;
        } else {
          cut_name_se(a.name);
          b->flat += a.seq; b->offsets.push_back(b->flat.size());
          b->name.push_back(std::move(a.name)); b->seq.push_back(std::move(a.seq)); b->qual.push_back(std::move(a.qual));
          b->n += 1;
        }
      }
      if (b->n == 0) break;
      to_gpu.push(std::move(b));
    }
    to_gpu.close();
  });

  std::atomic<int> live((int)devs.size());
  std::vector<std::thread> workers;
  const int finish_threads = std::max(1, o.threads / (int)devs.size());
  for (int dev : devs) workers.emplace_back([&, dev] {
    bmbs_batch* ctx = nullptr; size_t cap_reads = 0, cap_bases = 0, cap_cand = 0;
    std::unique_ptr<Batch> b;
    while (to_gpu.pop(b)) {
      const size_t bases = b->flat.size() + 64;
      size_t want_cand = std::max<size_t>(cap_cand, (size_t)b->n * 24 + (1u << 20));
      for (;;) {
        if (!ctx || (size_t)b->n > cap_reads || bases > cap_bases || want_cand > cap_cand) {
          if (ctx) bmbs_batch_free(ctx);
          cap_reads = std::max(cap_reads, (size_t)b->n); cap_bases = std::max(cap_bases, bases); cap_cand = want_cand;
          if (bmbs_batch_create(idx, dev, cap_reads, cap_bases, cap_cand, &ctx)) die(std::string("batch create: ") + bmbs_last_error());
        }
        b->res.resize(b->n); b->cand.resize(cap_cand);
        size_t used = 0;
        int rc = bmbs_batch_upload(ctx, b->flat.data(), b->offsets.data(), b->n, pe ? 1 : 0);
        if (!rc) rc = bmbs_batch_run(ctx, &o.prm);
        if (!rc) rc = bmbs_batch_download(ctx, b->res.data(), b->cand.data(), b->cand.size(), &used);
        if (rc == BMBS_ERR_CAPACITY) { want_cand = std::max(cap_cand * 2, used + (used >> 2) + 1024); continue; }
        if (rc) die(std::string("gpu batch failed: ") + bmbs_last_error());
        b->cand.resize(used);
        break;
      }
      finish_batch(hc, *b, pe, finish_threads);
      to_writer.push(std::move(b));
    }
    if (ctx) bmbs_batch_free(ctx);
    if (--live == 0) to_writer.close();
  });

  MapStats total;
  {
    std::map<size_t, std::unique_ptr<Batch>> pending; size_t next = 0;
    std::unique_ptr<Batch> b;
    while (to_writer.pop(b)) {
      pending[b->seq_no] = std::move(b);
      while (!pending.empty() && pending.begin()->first == next) {
        Batch& x = *pending.begin()->second;
        fwrite(x.sam.data(), 1, x.sam.size(), fo);
        total.reads += x.st.reads; total.unique += x.st.unique; total.ambiguous += x.st.ambiguous; total.bases += x.st.bases; total.err_bases += x.st.err_bases;
        pending.erase(pending.begin()); ++next;
      }
    }
  }
  reader.join(); for (auto& w : workers) w.join();
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
