#!/usr/bin/env bash
# oracle/ (test infrastructure, never on the product path): compile the REAL
# reference mapper from the sources where they lie under /root/reference.
#
#   oracle/_ref/bitmapperBS  the reference CLI (SE / PE / PE --sensitive, SAM text)
#   oracle/_ref/psascan      suffix-sorter stand-in the reference shells out to in --index
#   oracle/_ref/libref_bpm.so, libref_fm.so, libref_ksw.so   C entry points into the reference's own BPM / FM-index / CIGAR functions
#
# Nothing from /root/reference is copied into the repo: sources are copied to a
# scratch dir under /tmp, the 43 missing-`return` sites (UB that crashes an -O3
# build with g++ 13) get `return 0;`, everything is compiled by
# one g++ command mirroring the reference Makefile:23,47 (the reference's own
# build system is not run), and only the binaries land in oracle/_ref/
# (git-ignored; travels to the GPU box with the snapshot).  htslib (BAM only) is
# replaced by abort() stubs, pSAscan/libdivsufsort (cmake/OpenMP builds) by
# psascan_shim.cpp.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${BMBS_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "build_ref: $REF absent (GPU box?) - keeping prebuilt $OUT" >&2; exit 0; }
mkdir -p "$OUT"
if [ -x "$OUT/bitmapperBS" ] && [ -x "$OUT/psascan" ] && [ -f "$OUT/libref_bpm.so" ] && [ -f "$OUT/libref_fm.so" ] && [ -f "$OUT/libref_ksw.so" ] && [ "${1:-}" != "--force" ]; then
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
g++ -O2 -std=c++17 -pthread "$HERE/psascan_shim.cpp" -o "$OUT/psascan"
# function-level harnesses over the reference's own headers / sources
g++ -w -O3 -mavx2 -mpopcnt -D__AVX2__ -shared -fPIC -pthread -I "$REF" "$HERE/ref_harness_bpm.cpp" -o "$OUT/libref_bpm.so"
( cd "$TMP" && g++ -w -O2 -mpopcnt -shared -fPIC -I "$TMP" "$HERE/ref_harness_fm.cpp" bwt.cpp saca-k.cpp -o "$OUT/libref_fm.so" )
( cd "$TMP" && g++ -w -O2 -msse4.1 -mpopcnt -shared -fPIC -I "$TMP" "$HERE/ref_harness_ksw.cpp" ksw.cpp -o "$OUT/libref_ksw.so" )
echo "build_ref: built $OUT/{bitmapperBS,psascan,libref_bpm.so,libref_fm.so,libref_ksw.so}"
