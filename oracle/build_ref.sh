#!/usr/bin/env bash
# oracle/ (test infrastructure, never on the product path): compile the REAL
# reference mapper from the sources where they lie under /root/reference.
#
#   oracle/_ref/bitmapperBS  the reference CLI (SE / PE / PE --sensitive, SAM text)
#   oracle/_ref/bitmapperBS_bam + libref_hts.so   the same with its BAM writer (the vendored patched htslib, compiled with gcc)
#   oracle/_ref/bitmapperBS_gpu, bitmapperBS_seam_cpu   the reference with its workers calling include/bmbs.h per sub-block
#   oracle/_ref/psascan      suffix-sorter stand-in the reference shells out to in --index
#   oracle/_ref/libref_bpm.so, libref_fm.so, libref_ksw.so   C entry points into the reference's own BPM / FM-index / CIGAR functions
#
# Nothing from /root/reference is copied into the repo: sources are copied to a
# scratch dir under /tmp, the 43 missing-`return` sites (UB that crashes an -O3
# build with g++ 13) get `return 0;`, everything is compiled by
# one g++ command mirroring the reference Makefile:23,47 (the reference's own
# build system is not run), and only the binaries land in oracle/_ref/
# (git-ignored; travels to the GPU box with the snapshot).  In the main binary htslib
# (BAM only) is replaced by abort() stubs, pSAscan/libdivsufsort (cmake/OpenMP
# builds) by psascan_shim.cpp.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${BMBS_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "build_ref: $REF absent (GPU box?) - keeping prebuilt $OUT" >&2; exit 0; }
mkdir -p "$OUT"
if [ -x "$OUT/bitmapperBS" ] && [ -x "$OUT/bitmapperBS_bam" ] && [ -x "$OUT/bitmapperBS_gpu" ] && [ -x "$OUT/bitmapperBS_seam_cpu" ] && [ "$OUT/bitmapperBS_gpu" -nt "$HERE/../integration/bmbs_seam.h" ] && [ "$OUT/bitmapperBS_gpu" -nt "$HERE/../integration/patch_reference.py" ] && [ -x "$OUT/psascan" ] && [ -f "$OUT/libref_bpm.so" ] && [ -f "$OUT/libref_fm.so" ] && [ -f "$OUT/libref_ksw.so" ] && [ "${1:-}" != "--force" ]; then
  echo "build_ref: $OUT up to date"; exit 0
fi
TMP="$(mktemp -d /tmp/bmbs_refbuild.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
cp "$REF"/*.cpp "$REF"/*.h "$TMP"/
chmod u+w "$TMP"/*
python3 "$HERE/patch_returns.py" "$TMP" "$REF/htslib"
gcc -c -O1 "$HERE/hts_stub.c" -o "$TMP/hts_stub.o"
( cd "$TMP" && g++ -w -O3 -mavx2 -mpopcnt -fomit-frame-pointer -D__AVX2__ -I "$REF/htslib" \
    saca-k.cpp bwt.cpp Bitmapper_main.cpp Process_CommandLines.cpp Auxiliary.cpp Index.cpp Schema.cpp \
    Process_sam_out.cpp Process_Reads.cpp Ref_Genome.cpp Levenshtein_Cal.cpp SAM_queue.cpp bam_prase.cpp ksw.cpp \
    hts_stub.o -o bitmapperBS -lm -lz -lpthread )
cp "$TMP/bitmapperBS" "$OUT/bitmapperBS"
# ---- the same reference with the GPU seam spliced into its workers (integration/patch_reference.py, integration/bmbs_seam.h):
#   bitmapperBS_gpu       linked against the product library libbmbs_gpu.so (GPU tests diff its SAM with the stock binary's)
#   bitmapperBS_seam_cpu  linked against the CPU oracle behind the same C ABI (seam_cpu_shim.cpp): tests the splice without a GPU
ROOT="$(cd "$HERE/.." && pwd)"
cp "$TMP/Schema.cpp" "$TMP/Schema_seam.cpp"
python3 "$ROOT/integration/patch_reference.py" "$TMP/Schema_seam.cpp"
( cd "$TMP" && for f in saca-k bwt Bitmapper_main Process_CommandLines Auxiliary Index Process_sam_out Process_Reads Ref_Genome Levenshtein_Cal SAM_queue bam_prase ksw; do
    g++ -w -O3 -mavx2 -mpopcnt -fomit-frame-pointer -D__AVX2__ -I "$REF/htslib" -c $f.cpp -o $f.o & done; wait
  g++ -w -O3 -mavx2 -mpopcnt -fomit-frame-pointer -D__AVX2__ -I "$REF/htslib" -I "$ROOT/include" -I "$ROOT/integration" -c Schema_seam.cpp -o Schema_seam.o )
OBJS="saca-k.o bwt.o Bitmapper_main.o Process_CommandLines.o Auxiliary.o Index.o Schema_seam.o Process_sam_out.o Process_Reads.o Ref_Genome.o Levenshtein_Cal.o SAM_queue.o bam_prase.o ksw.o hts_stub.o"
( cd "$TMP" && g++ -O2 -std=c++17 -w -c "$HERE/seam_cpu_shim.cpp" -o seam_cpu_shim.o && g++ -O2 -std=c++17 -w -Wno-sign-compare -c "$HERE/oracle_capi.cpp" -o oracle_capi.o \
  && g++ $OBJS seam_cpu_shim.o oracle_capi.o -o bitmapperBS_seam_cpu -lm -lz -lpthread )
cp "$TMP/bitmapperBS_seam_cpu" "$OUT/bitmapperBS_seam_cpu"
if [ -f "$ROOT/bitmapperbs_b200/libbmbs_gpu.so" ]; then
  ( cd "$TMP" && g++ $OBJS -o bitmapperBS_gpu -L"$ROOT/bitmapperbs_b200" -lbmbs_gpu -Wl,-rpath,'$ORIGIN/../../bitmapperbs_b200' -lm -lz -lpthread )
  cp "$TMP/bitmapperBS_gpu" "$OUT/bitmapperBS_gpu"
else
  echo "build_ref: libbmbs_gpu.so not built yet -- bitmapperBS_gpu skipped (run bitmapperbs_b200/build.py first)" >&2
fi
# ---- the reference with its own BAM writer (--bam): the vendored, patched htslib 1.9 compiled from its sources with plain gcc
# (its Makefile / version.sh are not run: config.h and version.h are written here, bz2 / lzma / curl left out) into
# oracle/_ref/libref_hts.so, and the stock reference linked against it -> oracle/_ref/bitmapperBS_bam.  Dynamic linking is needed:
# bam_prase.cpp and bgzf.c both define queueMutex_bam.  Only the BAM parity test uses it (tests/test_host_bam.py).
if [ ! -x "$OUT/bitmapperBS_bam" ] || [ ! -f "$OUT/libref_hts.so" ] || [ "${1:-}" = "--force" ]; then
  H="$TMP/hts"; mkdir -p "$H"; cp -r "$REF"/htslib/* "$H"/; chmod -R u+w "$H"
  printf '#define HAVE_FSEEKO 1\n#define HAVE_DRAND48 1\n' > "$H/config.h"
  printf '#define HTS_VERSION "1.9"\n' > "$H/version.h"
  HOBJ="kfunc knetfile kstring bcf_sr_sort bgzf errmod faidx hfile hfile_net hts hts_os md5 multipart probaln realn regidx sam synced_bcf_reader vcf_sweep tbx textutils thread_pool vcf vcfutils cram/cram_codecs cram/cram_decode cram/cram_encode cram/cram_external cram/cram_index cram/cram_io cram/cram_samtools cram/cram_stats cram/files cram/mFILE cram/open_trace_file cram/pooled_alloc cram/rANS_static cram/sam_header cram/string_alloc"
  ( cd "$H" && for f in $HOBJ; do gcc -w -O2 -fPIC -I. -c $f.c -o $f.o & done; wait; gcc -shared -Wl,-soname,libref_hts.so -o "$OUT/libref_hts.so" *.o cram/*.o -lz -lm -lpthread )
  ( cd "$TMP" && g++ -w -O3 -mavx2 -mpopcnt -fomit-frame-pointer -D__AVX2__ -I "$H" \
      saca-k.cpp bwt.cpp Bitmapper_main.cpp Process_CommandLines.cpp Auxiliary.cpp Index.cpp Schema.cpp \
      Process_sam_out.cpp Process_Reads.cpp Ref_Genome.cpp Levenshtein_Cal.cpp SAM_queue.cpp bam_prase.cpp ksw.cpp \
      -o bitmapperBS_bam -L"$OUT" -l:libref_hts.so -Wl,-rpath,'$ORIGIN' -lm -lz -lpthread )
  cp "$TMP/bitmapperBS_bam" "$OUT/bitmapperBS_bam"
fi
g++ -O2 -std=c++17 -pthread "$HERE/psascan_shim.cpp" -o "$OUT/psascan"
# function-level harnesses over the reference's own headers / sources
g++ -w -O3 -mavx2 -mpopcnt -D__AVX2__ -shared -fPIC -pthread -I "$REF" "$HERE/ref_harness_bpm.cpp" -o "$OUT/libref_bpm.so"
( cd "$TMP" && g++ -w -O2 -mpopcnt -shared -fPIC -I "$TMP" "$HERE/ref_harness_fm.cpp" bwt.cpp saca-k.cpp -o "$OUT/libref_fm.so" )
( cd "$TMP" && g++ -w -O2 -msse4.1 -mpopcnt -shared -fPIC -I "$TMP" "$HERE/ref_harness_ksw.cpp" ksw.cpp -o "$OUT/libref_ksw.so" )
echo "build_ref: built $OUT/{bitmapperBS,psascan,libref_bpm.so,libref_fm.so,libref_ksw.so}"
