// oracle/oracle_cli.cpp -- TEST INFRASTRUCTURE: single-threaded CPU mapper built from
// the restatement in oracle_core.hpp.  Prints SAM in input order (the reference's
// `-t 1` order) so it can be diffed against oracle/_ref/bitmapperBS and the GPU path.
//   oracle_cli se  <genome.fa> <reads.fq> <out.sam> [e_rate]
//   oracle_cli pe  <genome.fa> <r1.fq> <r2.fq> <out.sam> [sensitive(0/1)] [e_rate] [min] [max]
#include <cstdio>
#include <cstdlib>
#include <string>
#include "oracle_core.hpp"
#include "oracle_pe.hpp"
#include "oracle_pe_sensitive.hpp"
#include "../bitmapperbs_b200/csrc/host/fastq.hpp"
#include "../bitmapperbs_b200/csrc/host/sam.hpp"

static void print_stats(const oracle::Stats& st) {
  unsigned long long un = st.reads - st.unique - st.ambiguous;
  fprintf(stderr, "No. of Reads: %llu\nUnique: %llu\nAmbiguous: %llu\nUnmapped: %llu\nErrBases/Bases: %llu/%llu\n",
          (unsigned long long)st.reads, (unsigned long long)st.unique, (unsigned long long)st.ambiguous, un,
          (unsigned long long)st.err_bases, (unsigned long long)st.bases);
}

int main(int argc, char** argv) {
  if (argc < 5) { fprintf(stderr, "usage: oracle_cli se|pe ...\n"); return 2; }
  std::string mode = argv[1], fa = argv[2];
  oracle::Index ix;
  if (!ix.load(fa + ".index")) { fprintf(stderr, "oracle: cannot load index %s.index*\n", fa.c_str()); return 1; }
  oracle::Params P; oracle::Stats st;
  std::string out; bmbs::sam_header(out, ix.chroms, "oracle");
  if (mode == "se") {
    if (argc > 5) P.e_rate = atof(argv[5]);
    bmbs::FastqReader fq; if (!fq.open(argv[3])) { fprintf(stderr, "cannot open %s\n", argv[3]); return 1; }
    bmbs::FastqRecord r; oracle::SeOutcome o; oracle::u32 carry = 0;
    FILE* fo = fopen(argv[4], "w");
    while (fq.next(r)) {
      bmbs::cut_name_se(r.name);
      oracle::map_single(ix, P, r.seq.c_str(), r.qual.c_str(), (int)r.seq.size(), o, st, carry);
      if (o.kind == oracle::SeOutcome::UNIQUE && !o.dropped_off_chrom)
        bmbs::sam_record_se(out, r.name, r.seq, r.qual, ix.chroms, o.placed, o.mapq, o.cigar, o.err);
      if (out.size() > (1u << 20)) { fwrite(out.data(), 1, out.size(), fo); out.clear(); }
    }
    fwrite(out.data(), 1, out.size(), fo); fclose(fo);
  } else {
    if (argc < 6) return 2;
    if (argc > 6) P.sensitive = atoi(argv[6]) != 0;
    if (argc > 7) P.e_rate = atof(argv[7]);
    if (argc > 8) P.min_ins = atoi(argv[8]);
    if (argc > 9) P.max_ins = atoi(argv[9]);
    return oracle::run_pe(ix, P, argv[3], argv[4], argv[5], out, st) ? (print_stats(st), 0) : 1;
  }
  print_stats(st);
  return 0;
}
