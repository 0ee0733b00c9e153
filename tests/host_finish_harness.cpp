// Test harness (CPU only): the product's host finishing code (bitmapperbs_b200/csrc/host/mapper.hpp: vote-ordered reduction,
// pair pick, CIGAR refinement with the DP on the CPU, MAPQ, SAM text) fed with the per-read records of the ORACLE instead of
// the GPU library -- both produce the same bmbs_read_result / bmbs_cand arrays (tests/test_gpu_parity.py checks that on the
// GPU box).  The SAM it writes must equal the reference's golden SAM: that pins the host glue without a GPU.
// Mode sef: single end through the finished records (bmbs_final) -- the oracle's restatement of the device finishing
// (orc_finish_se) feeds mapper.hpp's finish_single_final, the consumer the mapper uses behind bmbs_batch_finish.
// Modes pef / pesf: pairs (fast / sensitive) through the finished records -- orc_finish_pe (the oracle's restatement of the pair
// logic the device runs in finish_pe) feeds finish_pair_final.
//   host_finish_harness se|sef|pe|pes|pef|pesf <genome.fa> <out.sam> <reads.fq> [<mates.fq>] [--unmapped_out] [--ambiguous_out] [--pbat]
// (--ambiguous_out: paired end only here -- the single-end multi-exact case needs the located rows in row order, which only the
// device path returns; --pbat as in bmbs_main.cpp: single end aligns the reverse complement, paired end swaps the files)
#include <cstdio>
#include <chrono>
#include <cstring>
#include <string>
#include <vector>
#include "../oracle/oracle_capi.cpp"
#include "../bitmapperbs_b200/csrc/host/mapper.hpp"

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  const std::string mode = argv[1], fa = argv[2], out_path = argv[3];
  const bool fin_mode = mode == "sef" || mode == "pef" || mode == "pesf";      // through finished records (bmbs_final)
  const bool pe = mode == "pe" || mode == "pes" || mode == "pef" || mode == "pesf", sens = mode == "pes" || mode == "pesf";
  bool unmapped_out = false, ambiguous_out = false, pbat = false;
  std::vector<std::string> files;
  for (int i = 4; i < argc; ++i) { if (!strcmp(argv[i], "--unmapped_out")) unmapped_out = true; else if (!strcmp(argv[i], "--ambiguous_out")) ambiguous_out = true; else if (!strcmp(argv[i], "--pbat")) pbat = true; else files.push_back(argv[i]); }
  if (pbat && pe) std::swap(files[0], files[1]);
  bmbs::HostContext hc;
  hc.prm.e_rate = 0.08; hc.prm.seed_len = 30; hc.prm.min_ins = 0; hc.prm.max_ins = 500; hc.prm.sensitive = sens ? 1 : 0; hc.prm.ambiguous_out = 0; hc.ambiguous_out = ambiguous_out; hc.pbat = pbat && !pe;
  const std::string prefix = fa + ".index";
  if (!hc.chroms.load(prefix) || !hc.genome.load(prefix + ".bs.pac", hc.chroms.N)) { fprintf(stderr, "cannot load %s\n", prefix.c_str()); return 1; }
  void* h = orc_load(prefix.c_str());
  if (!h) return 1;
  bmbs::FastqReader f1, f2;
  if (!f1.open(files[0]) || (pe && !f2.open(files[1]))) return 1;
  std::vector<bmbs::FastqRecord> recs; bmbs::FastqRecord r;
  std::string flat; std::vector<uint64_t> offs(1, 0); std::vector<std::string> raw2;
  std::vector<std::string> raw1;
  if (!pe) { while (f1.next(r)) { bmbs::cut_name_se(r.name); raw1.push_back(r.seq); flat += hc.pbat ? bmbs::revcomp(r.seq) : r.seq; offs.push_back(flat.size()); recs.push_back(r); } }
  else {
    bmbs::FastqRecord a, b;
    while (f1.next(a) && f2.next(b)) {
      bmbs::cut_name_pe(a.name, b.name);
      flat += a.seq; offs.push_back(flat.size());
      raw2.push_back(b.seq); b.seq = bmbs::revcomp(b.seq);
      flat += b.seq; offs.push_back(flat.size());
      recs.push_back(a); recs.push_back(b);
    }
  }
  const int n = (int)recs.size();
  std::vector<bmbs_read_result> res(n + 1); std::vector<bmbs_cand> cand((size_t)n * 64 + 1024); size_t used = 0;
  for (;;) {
    int rc = !pe ? orc_map_se(h, flat.data(), offs.data(), n, 0.08, 30, res.data(), cand.data(), cand.size(), &used)
             : !sens ? orc_map_pe(h, flat.data(), offs.data(), n / 2, 0.08, 30, 0, 500, res.data(), cand.data(), cand.size(), &used)
                     : orc_map_pe_sensitive(h, flat.data(), offs.data(), n / 2, 0.08, 30, 0, 500, res.data(), cand.data(), cand.size(), &used, nullptr);
    if (rc == 0) break;
    cand.resize(used + 1024);
  }
  std::vector<bmbs_final> fin; std::vector<uint16_t> mism;
  if (fin_mode) {
    fin.resize(n + 1); mism.resize((size_t)n * 32 + 64); size_t mused = 0;
    if (!pe) { if (orc_finish_se(h, flat.data(), offs.data(), n, 0.08, ambiguous_out ? 1 : 0, res.data(), cand.data(), fin.data(), mism.data(), mism.size(), &mused)) return 1; }
    else if (orc_finish_pe(h, flat.data(), offs.data(), n / 2, 0.08, 0, 500, sens ? 1 : 0, ambiguous_out ? 1 : 0, res.data(), cand.data(), fin.data(), mism.data(), mism.size(), &mused)) return 1;
  }
  std::string out; bmbs::sam_header(out, hc.chroms, "host_finish_harness");
  bmbs::MapStats st; std::vector<bmbs::HostHit> v1, v2; std::vector<char> win;
  auto seq = [&](int i) { return std::string_view(flat.data() + offs[i], (size_t)(offs[i + 1] - offs[i])); };
  const int units = pe ? n / 2 : n;
  const auto t_begin = std::chrono::steady_clock::now();
  for (int u = 0; u < units; ++u) {
    bmbs::MapStats t;
    if (fin_mode && pe) bmbs::finish_pair_final(hc, recs[2 * u].name, seq(2 * u), recs[2 * u].qual, recs[2 * u + 1].name, seq(2 * u + 1), raw2[u], recs[2 * u + 1].qual,
                                                fin[2 * u], fin[2 * u + 1], mism.data(), out, t, win);
    else if (fin_mode) { bmbs::ReadView rv{recs[u].name, seq(u), recs[u].qual, raw1[u]}; bmbs::finish_single_final(hc, rv, fin[u], mism.data(), cand.data(), out, t, v1, win); }
    else if (!pe) { bmbs::ReadView rv{recs[u].name, seq(u), recs[u].qual, raw1[u]}; bmbs::finish_single(hc, rv, res[u], cand.data(), out, t, v1, win); }
    else bmbs::finish_pair(hc, recs[2 * u].name, seq(2 * u), recs[2 * u].qual, recs[2 * u + 1].name, seq(2 * u + 1), raw2[u], recs[2 * u + 1].qual,
                           res[2 * u], res[2 * u + 1], cand.data(), out, t, v1, v2, win);
    if (unmapped_out && !t.unique && !t.ambiguous) {
      if (!pe) bmbs::sam_record_unmapped(out, recs[u].name, 4, raw1[u], recs[u].qual);
      else { bmbs::sam_record_unmapped(out, recs[2 * u].name, 77, seq(2 * u), recs[2 * u].qual); bmbs::sam_record_unmapped(out, recs[2 * u + 1].name, 141, raw2[u], recs[2 * u + 1].qual); }
    }
    st.reads += t.reads; st.unique += t.unique; st.ambiguous += t.ambiguous; st.bases += t.bases; st.err_bases += t.err_bases;
  }
  fprintf(stderr, "finish: %.3f s for %d units\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count(), units);
  FILE* fo = fopen(out_path.c_str(), "w"); fwrite(out.data(), 1, out.size(), fo); fclose(fo);
  const long long nn = (long long)st.reads, uq = (long long)st.unique, am = (long long)st.ambiguous;
  printf("%lld %lld %lld %.2f\n", nn, uq, am, st.bases ? (double)st.err_bases / (double)st.bases * 100 : 0.0);
  return 0;
}
