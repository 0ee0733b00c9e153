// Device-side view of the index and the primitive FM-index / genome operations.
//
// HBM layout (built once at load from the reference's on-disk files, which stay
// untouched):
//   occ    : one 32-byte block per 64 BWT symbols = exactly one DRAM sector per
//            occ lookup: { u64 plane_lo, u64 plane_hi, u64 count(T), u64 count(A) }
//            with ABSOLUTE counts (the reference's 16-bit relative counters +
//            65536-row table, bwt.h:1007-1058, are folded together at load).
//   flag   : one 16-byte block per 64 rows: { u64 sampled-row flags, u64 rank
//            before the block } (reference: 5 words per 256 rows, bwt.h:2449-2560).
//   hash   : u64 per 16-mer: 36-bit first row | 4-bit gap << 60 (bwt.h:284-306).
//   ssa    : u32 sampled suffix array, unchanged (top two bits masked on use); replaced at load by the
//            dense array dsa (u32 per row, +u8 above 2^32 rows) when HBM allows: see locate_row.
//   planes : the 2N-base double-strand sequence  G ++ revcomp(G)  as interleaved
//            bit-planes {lo, hi} per 32 bases, LSB = first base, so that any
//            window is two funnel shifts away (reference: 2-bit bytes decoded
//            through 256-entry LUTs, Schema.cpp:4998-5115).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bmbs {

typedef unsigned long long u64;
typedef unsigned int u32;

struct DevIndex {
  const ulonglong2* occ;
  const ulonglong2* flag;
  const u64* hash;
  const u32* ssa;
 const uint2* planes;
  const u64* ktab;            // deep seed table over K-mers, K = 16 + kdepth (null: only the 16-mer table); see kmer_entry
  u32 kdepth, kpow;           // K - 16 (1..4), 3^(K-16)
  const u32* dsa_lo;          // dense suffix array (one entry per row), low 32 bits; null = walk to a sampled row
  const unsigned char* dsa_hi; // bits 32..39 when the text is longer than 2^32
  u64 C[3];        // first row of symbols G(0), T(1), A(2)   (nacgt[c], bwt.cpp:1715-1729)
  u64 shapline;    // row whose BWT symbol is '$' (omitted from the planes)
  u64 n_rows;      // text length + 1
  u64 N;           // genome length; text length = 2N
  const u64* chrom_start;  // first forward-strand coordinate of every chromosome, chrom_start[n_chrom] = N (finishing: place_hit)
  u32 n_chrom;
};

// nibble codes of a read base: A0 C1 G2 T3 other 4; FM alphabet G0 T1 A2 with C->T, other 3
__device__ __forceinline__ int fm_code(int c) { return (0x31012 >> (4 * c)) & 0xF; }

__device__ __forceinline__ int read_code(const u32* __restrict__ w, int i) {
  return (__ldg(w + (i >> 3)) >> (4 * (i & 7))) & 0xF;
}

struct OccBlock { ulonglong2 planes, cnt; u64 blk; };

// one 32-byte occ block = one 256-bit load (LDG.E.256 on sm_100) = one L1 wavefront and one DRAM sector
__device__ __forceinline__ OccBlock load_occ(const DevIndex& ix, u64 adj_row) {
  OccBlock b;
  b.blk = adj_row >> 6;
  const ulonglong2* p = ix.occ + b.blk * 2;
  asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];"
               : "=l"(b.planes.x), "=l"(b.planes.y), "=l"(b.cnt.x), "=l"(b.cnt.y) : "l"(p));
  return b;
}

// C[c] by selects: a run-time index into the by-value kernel parameter would make the compiler copy the whole struct to local memory
__device__ __forceinline__ u64 first_row_of(const DevIndex& ix, int c) { return c == 0 ? ix.C[0] : c == 1 ? ix.C[1] : ix.C[2]; }

// C[c] + occ(c, row) for a row already adjusted for the '$' row.  bwt.h:1373-1465.
__device__ __forceinline__ u64 rank_in(const DevIndex& ix, const OccBlock& b, u64 adj_row, int c) {
  const unsigned part = (unsigned)adj_row & 63u;
  const u64 plane = c == 1 ? b.planes.x : c == 2 ? b.planes.y : ~(b.planes.x | b.planes.y);
  const u64 base = c == 1 ? b.cnt.x : c == 2 ? b.cnt.y : (b.blk << 6) - b.cnt.x - b.cnt.y;
  const u64 head = part ? plane >> (64 - part) : 0ull;
  return first_row_of(ix, c) + base + (u64)__popcll(head);
}

__device__ __forceinline__ u64 adjust_row(const DevIndex& ix, u64 row) { return row - (u64)(row > ix.shapline); }

// one backward-extension step on [sp, ep): find_occ_fm_index_combine, bwt.h:1473-1596.
// Returns the number of distinct occ blocks touched (1 or 2) for the work counters.
// Straight-line code: both ends always load their block (the second load of an interval inside one block is the same
// sector again, a hit), and plane / count / first row of the symbol are picked with masks, not branches -- the lanes of a
// warp extend by different symbols over intervals of different widths, and every branch here splits them three or six ways.
__device__ __forceinline__ u64 rank_masked(const DevIndex& ix, const OccBlock& b, u64 adj_row, u64 m0, u64 m1, u64 m2) {
  const unsigned part = (unsigned)adj_row & 63u;
  const u64 plane = (b.planes.x & m1) | (b.planes.y & m2) | (~(b.planes.x | b.planes.y) & m0);
  const u64 base = (b.cnt.x & m1) | (b.cnt.y & m2) | (((b.blk << 6) - b.cnt.x - b.cnt.y) & m0);
  const u64 first = (ix.C[1] & m1) | (ix.C[2] & m2) | (ix.C[0] & m0);
  const u64 head = (plane >> 1) >> (63u - part);           // plane >> (64 - part), and 0 for part == 0
  return first + base + (u64)__popcll(head);
}
__device__ __forceinline__ int lf_pair(const DevIndex& ix, u64& sp, u64& ep, int c) {
  const u64 a = adjust_row(ix, sp), b = adjust_row(ix, ep);
  const OccBlock ba = load_occ(ix, a), bb = load_occ(ix, b);
  const u64 m1 = 0ull - (u64)(c == 1), m2 = 0ull - (u64)(c == 2), m0 = ~(m1 | m2);
  sp = rank_masked(ix, ba, a, m0, m1, m2);
  ep = rank_masked(ix, bb, b, m0, m1, m2);
  return bb.blk != ba.blk ? 2 : 1;
}

__device__ __forceinline__ void hash_query(const DevIndex& ix, u64 key, u64& sp, u64& ep) {
  const u64 a = __ldg(ix.hash + key), b = __ldg(ix.hash + key + 1);
  sp = a & 0xFFFFFFFFFull;
  ep = (b & 0xFFFFFFFFFull) - (b >> 60);
}

// Deep seed table (built on the device at load): entry of the K-mer whose first 16 symbols have table key
// `key16` and whose symbols 16..K-1 have base-3 value `ext` (symbol 16 least significant) records where the greedy seed
// loop of count_backward_as_much_1_terminate (bwt.h:2081-2209) stands after at most K symbols:
//   bits 0..35  first row of the interval I_m          bits 36..38  0: the 16-mer does not occur; else m - 15 (m = 16..K)
//   bits 39..63 rows in I_m (KTAB_SAT: too many to store, fall back to the 16-mer table)
// m < K means the loop stopped there (one row left, or the next symbol empties the interval); m = K means it goes on.
constexpr u64 KTAB_SAT = (1ull << 25) - 1;
__device__ __forceinline__ u64 kmer_entry(u32 m, u64 top, u64 bot) {
  const u64 size = bot - top;
  return (top & 0xFFFFFFFFFull) | ((u64)(m ? m - 15 : 0) << 36) | ((size >= KTAB_SAT ? KTAB_SAT : size) << 39);
}

// Single-row locate: walk LF until a sampled row.  bwt.h:2449-2560.  `steps_out` counts LF steps.
__device__ __forceinline__ u64 locate_row_walk(const DevIndex& ix, u64 row, int& steps_out) {
  int steps = 0;
  u64 sa;
  for (;;) {
    if (row == ix.shapline) { sa = (u64)steps; break; }
    const ulonglong2 f = __ldg(ix.flag + (row >> 6));
    const unsigned in = (unsigned)row & 63u;
    if ((f.x >> (63 - in)) & 1ull) {
      const u64 rank = f.y + (in ? (u64)__popcll(f.x >> (64 - in)) : 0ull);
      sa = (u64)(__ldg(ix.ssa + rank) & 0x3FFFFFFFu) * 8ull + (u64)steps;
      break;
    }
    const u64 a = adjust_row(ix, row);
    const OccBlock b = load_occ(ix, a);
    const unsigned sh = 63u - ((unsigned)a & 63u);
    const int c = ((b.planes.x >> sh) & 1ull) ? 1 : ((b.planes.y >> sh) & 1ull) ? 2 : 0;
    row = rank_in(ix, b, a, c);
    ++steps;
  }
  steps_out = steps;
  return sa;
}

// Single-row locate.  With 180 GB of HBM the suffix array does not have to stay sampled: at load every row's
// value is computed once on the device (densify_sa) and locate becomes one 4-byte (5-byte) gather instead of
// up to 7 dependent LF steps of two sectors each.  Same value as the walk by construction.
__device__ __forceinline__ u64 locate_row(const DevIndex& ix, u64 row, int& steps_out) {
  if (ix.dsa_lo) {
    steps_out = 0;
    u64 sa = __ldg(ix.dsa_lo + row);
    if (ix.dsa_hi) sa |= (u64)__ldg(ix.dsa_hi + row) << 32;
    return sa;
  }
  return locate_row_walk(ix, row, steps_out);
}

// 2-bit code (A0 C1 G2 T3) of base `pos` of the double-strand sequence.
__device__ __forceinline__ int strand_base(const DevIndex& ix, u64 pos) {
  const uint2 w = __ldg(ix.planes + (pos >> 5));
  const unsigned s = (unsigned)pos & 31u;
  return ((w.x >> s) & 1u) | (((w.y >> s) & 1u) << 1);
}

// The reference zero-fills any window that would leave its strand (Schema.cpp:5013-5019,
// :5076-5084, including the modulo-2^64 cases): true when [site, site+len) is usable.
__device__ __forceinline__ bool window_inside(const DevIndex& ix, u64 site, u64 len) {
  if (site < ix.N) return site + len <= ix.N;
  const u64 s = site - ix.N;
  return s < ix.N && s + len <= ix.N;
}

}  // namespace bmbs
