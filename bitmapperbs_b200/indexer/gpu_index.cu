// bmbs-index-gpu: the index writer of build_index.hpp with the heavy parts on the device -- same files, byte for byte.
//
// Data-prep tooling, not the mapping path: BASELINE's headline genome is 3.1 Gbp (6.2 G suffixes) and a benchmark box
// starts with an empty disk, so the index has to be built in minutes.  The suffix array is built with library radix
// sorts (cub): suffixes are split into classes by their first symbols, each class is sorted on the next 32 symbols
// (one 64-bit key), and groups of equal keys are refined 32 symbols at a time until every group is a single suffix
// (genomes whose repeats are diverged copies need a few dozen rounds over a quickly shrinking set).  BWT bit-planes,
// occ counters, sampled-row flags, the sampled suffix array and the 16-mer counts are then plain data-parallel passes.
// Contract and file formats: build_index.hpp (which cites the reference writer block by block).
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "build_index.hpp"

using namespace bmbs::indexer;
typedef unsigned long long u64;
typedef unsigned int u32;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "bmbs-index-gpu: %s: %s\n", #call, cudaGetErrorString(e_)); exit(1); } } while (0)

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// ---- text: 2 bits per symbol, value = code + 1 (G 1, T 2, A 3), 0 beyond the end; symbol i at bits 62 - 2 (i & 31) of word i >> 5
__global__ void pack_text(const unsigned char* __restrict__ g, u64 N, u64* __restrict__ tw, u64 n_words) {
  const u64 n = 2 * N;
  for (u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (u64)gridDim.x * blockDim.x) {
    u64 x = 0;
    for (u32 j = 0; j < 32; ++j) {
      const u64 i = w * 32 + j;
      if (i >= n) break;
      u32 c;
      if (i < N) { const unsigned char b = g[i]; c = b == 'A' ? 1 : b == 'C' ? 0 : b == 'G' ? 1 : 2; }        // complement, then C->T
      else { const unsigned char b = g[n - 1 - i]; c = b == 'A' ? 2 : b == 'G' ? 0 : 1; }                      // reversed, C->T
      x |= (u64)(c + 1) << (62 - 2 * j);
    }
    tw[w] = x;
  }
}
__device__ __forceinline__ u64 window(const u64* __restrict__ tw, u64 n, u64 p) {      // 32 symbols from p (zeros past the end)
  if (p >= n) return 0;
  const u64 i = p >> 5; const u32 s = 2 * (u32)(p & 31);
  const u64 a = tw[i];
  return s ? (a << s) | (tw[i + 1] >> (64 - s)) : a;
}
__device__ __forceinline__ u32 symbol(const u64* __restrict__ tw, u64 p) { return (u32)(tw[p >> 5] >> (62 - 2 * (p & 31))) & 3u; }   // code + 1

// ---- classes: suffixes by their first P symbols (base-4 digits incl. the padding 0)
__global__ void class_histogram(const u64* __restrict__ tw, u64 n, int P, u64* __restrict__ count) {
  __shared__ u64 s[64];
  if (threadIdx.x < 64) s[threadIdx.x] = 0;
  __syncthreads();
  for (u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (u64)gridDim.x * blockDim.x)
    atomicAdd(&s[P ? (u32)(window(tw, n, p) >> (64 - 2 * P)) : 0u], 1ull);
  __syncthreads();
  if (threadIdx.x < 64 && s[threadIdx.x]) atomicAdd(&count[threadIdx.x], s[threadIdx.x]);
}
__global__ void class_collect(const u64* __restrict__ tw, u64 n, int P, u32 cls, u64* __restrict__ keys, u64* __restrict__ pos, u64* __restrict__ cursor) {
  for (u64 p0 = (u64)blockIdx.x * blockDim.x; p0 < n; p0 += (u64)gridDim.x * blockDim.x) {
    const u64 p = p0 + threadIdx.x;
    const bool mine = p < n && (P ? (u32)(window(tw, n, p) >> (64 - 2 * P)) : 0u) == cls;
    const u32 m = __ballot_sync(0xffffffffu, mine);
    if (!m) continue;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    u64 base = 0;
    if (lane == leader) base = atomicAdd(cursor, (u64)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (mine) { const u64 at = base + __popc(m & ((1u << lane) - 1u)); keys[at] = window(tw, n, p + P); pos[at] = p; }
  }
}
// after a sort by key: every slot gets its suffix; members of groups of equal keys go to the refinement list
struct Tie { u64 pos, seg; };
__global__ void place_and_flag(const u64* __restrict__ keys, const u64* __restrict__ pos, u64 m, u64 row0, u64* __restrict__ sa, unsigned char* __restrict__ tie) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x) {
    sa[row0 + i] = pos[i];
    const bool same_prev = i > 0 && keys[i - 1] == keys[i], same_next = i + 1 < m && keys[i + 1] == keys[i];
    tie[i] = same_prev || same_next;
  }
}
// group id of a tied slot = the first slot of its run of equal keys (global row index)
__global__ void run_first(const u64* __restrict__ keys, u64 m, u64* __restrict__ first) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x)
    first[i] = (i == 0 || keys[i - 1] != keys[i]) ? i : 0;
}
__global__ void make_ties(const u64* __restrict__ pos, const u64* __restrict__ first, const u64* __restrict__ tie_off, const unsigned char* __restrict__ tie, u64 m, u64 row0, Tie* __restrict__ out) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x)
    if (tie[i]) { Tie t; t.pos = pos[i]; t.seg = row0 + first[i]; out[tie_off[i]] = t; }
}
__global__ void tie_keys(const u64* __restrict__ tw, u64 n, const Tie* __restrict__ e, u64 ne, u64 depth, u64* __restrict__ key) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < ne; i += (u64)gridDim.x * blockDim.x) key[i] = window(tw, n, e[i].pos + depth);
}
struct TieK { u64 pos, key; };
__global__ void tie_to_segkeys(const Tie* __restrict__ e, const u64* __restrict__ key, u64 ne, u64* __restrict__ seg, TieK* __restrict__ v) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < ne; i += (u64)gridDim.x * blockDim.x) { seg[i] = e[i].seg; TieK t; t.pos = e[i].pos; t.key = key[i]; v[i] = t; }
}
// sorted by (seg, key): heads of groups and of runs as indices for the max-scans
__global__ void tie_heads(const u64* __restrict__ seg, const TieK* __restrict__ v, u64 ne, u64* __restrict__ seg_first, u64* __restrict__ run_first_i) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < ne; i += (u64)gridDim.x * blockDim.x) {
    const bool hs = i == 0 || seg[i - 1] != seg[i];
    seg_first[i] = hs ? i : 0;
    run_first_i[i] = (hs || v[i - 1].key != v[i].key) ? i : 0;
  }
}
__global__ void tie_place(const u64* __restrict__ seg, const TieK* __restrict__ v, const u64* __restrict__ seg_first, const u64* __restrict__ run_first_i, u64 ne,
                          u64* __restrict__ sa, unsigned char* __restrict__ tie, Tie* __restrict__ next) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < ne; i += (u64)gridDim.x * blockDim.x) {
    const u64 slot = seg[i] + (i - seg_first[i]);
    sa[slot] = v[i].pos;
    const bool same_prev = run_first_i[i] != i;
    const bool same_next = i + 1 < ne && seg[i + 1] == seg[i] && v[i + 1].key == v[i].key;
    tie[i] = same_prev || same_next;
    Tie t; t.pos = v[i].pos; t.seg = seg[i] + (run_first_i[i] - seg_first[i]);
    next[i] = t;       // compacted by the caller with the tie flags
  }
}
struct MaxOp { __device__ __forceinline__ u64 operator()(u64 a, u64 b) const { return a > b ? a : b; } };

// ---- BWT and friends from the suffix array (row 0 is the empty suffix, row r > 0 holds sa[r - 1])
__device__ __forceinline__ u64 sa_of(const u64* __restrict__ sa, u64 n, u64 row) { return row == 0 ? n : sa[row - 1]; }
__global__ void find_shapline(const u64* __restrict__ sa, u64 n, u64* __restrict__ out) {
  for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (u64)gridDim.x * blockDim.x) if (sa[r] == 0) *out = r + 1;
}
// one warp per 64 BWT symbols: plane words and the block's counts of symbols 1 and 2
__global__ void bwt_blocks(const u64* __restrict__ tw, const u64* __restrict__ sa, u64 n, u64 shapline, u64 n_blocks, u64* __restrict__ bwt, u32* __restrict__ c1, u32* __restrict__ c2) {
  const int lane = threadIdx.x & 31;
  for (u64 blk = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; blk < n_blocks; blk += ((u64)gridDim.x * blockDim.x) >> 5) {
    u32 lo[2], hi[2];
    for (int h = 0; h < 2; ++h) {
      const u64 j = blk * 64 + h * 32 + lane;          // BWT index; its row skips the one whose suffix is the whole text
      u32 ch = 0; bool live = j < n;
      if (live) { const u64 row = j + (j >= shapline ? 1 : 0); const u64 s = sa_of(sa, n, row); ch = symbol(tw, s - 1) - 1; }
      lo[h] = __brev(__ballot_sync(0xffffffffu, live && (ch & 1)));        // symbol j at bit 63 - (j & 63)
      hi[h] = __brev(__ballot_sync(0xffffffffu, live && (ch >> 1)));
    }
    if (lane == 0) {
      const u64 w = (blk >> 1) * 5 + 1 + (blk & 1) * 2;
      const u64 plo = ((u64)lo[0] << 32) | lo[1], phi = ((u64)hi[0] << 32) | hi[1];
      bwt[w] = plo; bwt[w + 1] = phi;
      c1[blk] = __popcll(plo & ~phi); c2[blk] = __popcll(phi);             // code 1 = T (lo), code 2 = A (hi)
    }
  }
}
// block headers (16-bit counts relative to the 65536-symbol table) and that table; s1 / s2: exclusive scans of the block counts
__global__ void bwt_high_occ(const u64* __restrict__ s1, const u64* __restrict__ s2, u64 n, u64 n_blocks, u64* __restrict__ high_occ) {
  for (u64 blk = (u64)blockIdx.x * blockDim.x + threadIdx.x; blk <= n_blocks; blk += (u64)gridDim.x * blockDim.x) {
    const u64 j = blk * 64;
    if (j == 0 || j > n) continue;                      // headers are written when the running index reaches a multiple of 64
    const u64 a1 = s1[blk], a2 = s2[blk];                // counts of symbols before position j
    if ((j & 65535) == 0) { high_occ[(j >> 16) * 2] = a1; high_occ[(j >> 16) * 2 + 1] = a2; }
  }
}
__global__ void bwt_headers(const u64* __restrict__ s1, const u64* __restrict__ s2, u64 n, u64 n_blocks, u64* __restrict__ bwt) {
  for (u64 sb = (u64)blockIdx.x * blockDim.x + threadIdx.x; sb * 2 <= n_blocks; sb += (u64)gridDim.x * blockDim.x) {
    u64 hdr = 0;
    for (int h = 0; h < 2; ++h) {
      const u64 blk = sb * 2 + h, j = blk * 64;
      if (j == 0 || j > n) continue;
      const u64 base_blk = (j >> 16) << 10;            // the block at the last multiple of 65536
      const u64 r1 = s1[blk] - s1[base_blk], r2 = s2[blk] - s2[base_blk];
      hdr |= h ? (r1 << 16) | r2 : (r1 << 48) | (r2 << 32);
    }
    if (hdr) bwt[sb * 5] = hdr;
  }
}
// sampled rows (suffix a multiple of 8): flag words per 64 rows and their counts
__global__ void flag_words(const u64* __restrict__ sa, u64 n, u64 R, u64 n_words, u64* __restrict__ bits, u32* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  for (u64 w = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_words; w += ((u64)gridDim.x * blockDim.x) >> 5) {
    u32 b[2];
    for (int h = 0; h < 2; ++h) {
      const u64 row = w * 64 + h * 32 + lane;
      const bool on = row < R && (sa_of(sa, n, row) & 7) == 0;
      b[h] = __brev(__ballot_sync(0xffffffffu, on));
    }
    if (lane == 0) { const u64 x = ((u64)b[0] << 32) | b[1]; bits[w] = x; cnt[w] = __popcll(x); }
  }
}
__global__ void flag_layout(const u64* __restrict__ bits, const u64* __restrict__ rank, u64 n_words, u64 R, u64 total, u64* __restrict__ flag) {
  for (u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (u64)gridDim.x * blockDim.x) {
    flag[(w >> 2) * 5 + 1 + (w & 3)] = bits[w];
    if ((w & 3) == 0) flag[(w >> 2) * 5] = rank[w];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && (R & 255) == 0) flag[(R >> 8) * 5] = total;
}
__global__ void sampled_sa(const u64* __restrict__ tw, const u64* __restrict__ sa, u64 n, u64 R, const u64* __restrict__ bits, const u64* __restrict__ rank, u32* __restrict__ ssa) {
  for (u64 row = (u64)blockIdx.x * blockDim.x + threadIdx.x; row < R; row += (u64)gridDim.x * blockDim.x) {
    const u64 s = sa_of(sa, n, row);
    if (s & 7) continue;
    const u64 x = bits[row >> 6]; const u32 in = (u32)(row & 63);
    const u64 at = rank[row >> 6] + (in ? __popcll(x >> (64 - in)) : 0);
    const u32 ch = s ? symbol(tw, s - 1) - 1 : 1;
    ssa[at] = (ch << 30) | (u32)(s >> 3);
  }
}
// occurrences of every 16-mer (base-3 value, first symbol most significant)
__global__ void count_16mers(const u64* __restrict__ tw, u64 n, u32* __restrict__ kc) {
  for (u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x; p + 16 <= n; p += (u64)gridDim.x * blockDim.x) {
    const u64 w = window(tw, n, p);
    u32 key = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) key = key * 3 + ((u32)(w >> (62 - 2 * i)) & 3u) - 1;
    atomicAdd(kc + key, 1u);
  }
}
__global__ void widen8(const unsigned char* __restrict__ in, u64 m, u64* __restrict__ out) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x) out[i] = in[i];
}
__global__ void widen(const u32* __restrict__ in, u64 m, u64* __restrict__ out) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x) out[i] = in[i];
}

template <class T> static T* dmalloc(size_t n) { T* p = nullptr; CK(cudaMalloc(&p, (n + 8) * sizeof(T))); return p; }

struct Temp {      // cub scratch, grown on demand
  void* p = nullptr; size_t cap = 0;
  void* need(size_t b) { if (b > cap) { if (p) cudaFree(p); cap = b + (b >> 3) + 256; CK(cudaMalloc(&p, cap)); } return p; }
};
static Temp g_tmp;

template <class V> static void sort_pairs(u64*& k_in, u64*& k_out, V*& v_in, V*& v_out, u64 m, int begin_bit, int end_bit) {
  size_t tb = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, k_in, k_out, v_in, v_out, (long long)m, begin_bit, end_bit));
  void* t = g_tmp.need(tb);
  CK(cub::DeviceRadixSort::SortPairs(t, tb, k_in, k_out, v_in, v_out, (long long)m, begin_bit, end_bit));
  std::swap(k_in, k_out); std::swap(v_in, v_out);
}
static void max_scan(u64* d, u64 m) {
  size_t tb = 0;
  CK(cub::DeviceScan::InclusiveScan(nullptr, tb, d, d, MaxOp(), (long long)m));
  CK(cub::DeviceScan::InclusiveScan(g_tmp.need(tb), tb, d, d, MaxOp(), (long long)m));
}
static u64 exclusive_sum(const u64* in, u64* out, u64 m) {     // out[m] = total
  size_t tb = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, (long long)m));
  CK(cub::DeviceScan::ExclusiveSum(g_tmp.need(tb), tb, in, out, (long long)m));
  u64 last_in = 0, last_out = 0;
  if (m) { CK(cudaMemcpy(&last_in, in + m - 1, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&last_out, out + m - 1, 8, cudaMemcpyDeviceToHost)); }
  const u64 total = last_in + last_out;
  CK(cudaMemcpy(out + m, &total, 8, cudaMemcpyHostToDevice));
  return total;
}
template <class T> static u64 compact(const T* in, const unsigned char* flags, T* out, u64 m) {
  u64* d_n = dmalloc<u64>(1);
  size_t tb = 0;
  CK(cub::DeviceSelect::Flagged(nullptr, tb, in, flags, out, d_n, (long long)m));
  CK(cub::DeviceSelect::Flagged(g_tmp.need(tb), tb, in, flags, out, d_n, (long long)m));
  u64 k = 0; CK(cudaMemcpy(&k, d_n, 8, cudaMemcpyDeviceToHost)); cudaFree(d_n);
  return k;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: bmbs-index-gpu genome.fa [device]\n"); return 2; }
  const std::string fa = argv[1];
  const bool verbose = getenv("BMBS_VERBOSE") != nullptr;
  double t0 = now_s();
  auto lap = [&](const char* what) { if (verbose) { CK(cudaDeviceSynchronize()); const double t = now_s(); fprintf(stderr, "[bmbs-index-gpu] %-34s %.2f s\n", what, t - t0); t0 = t; } };
  if (argc > 2) CK(cudaSetDevice(atoi(argv[2])));
  CK(cudaFree(0));
  std::vector<Chrom> chroms; std::vector<uint8_t> g;
  read_fasta(fa.c_str(), chroms, g);
  const u64 N = g.size(), n = 2 * N;
  fprintf(stderr, "bmbs-index-gpu: %zu chromosomes, %llu bases\n", chroms.size(), (unsigned long long)N);
  write_chrom_table(fa, chroms, N);
  write_pac(fa, g);
  lap("fasta, chromosome table, 2-bit genome");
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int G = prop.multiProcessorCount * 8, B = 256;

  // ---- text on the device
  const u64 n_words = (n + 31) / 32 + 4;
  u64* tw = dmalloc<u64>(n_words);
  {
    unsigned char* d_g = dmalloc<unsigned char>(N);
    CK(cudaMemcpy(d_g, g.data(), N, cudaMemcpyHostToDevice));
    CK(cudaMemset(tw, 0, (n_words + 8) * 8));
    pack_text<<<G, B>>>(d_g, N, tw, (n + 31) / 32);
    CK(cudaDeviceSynchronize()); cudaFree(d_g);
  }
  std::vector<uint8_t> tail;        // the last 15 symbols (codes) for the 16-mer table's short suffixes
  for (u64 i = n >= 15 ? n - 15 : 0; i < n; ++i) { const uint8_t c = i < N ? g[i] : g[n - 1 - i]; tail.push_back(i < N ? (c == 'A' ? 1 : c == 'C' ? 0 : c == 'G' ? 1 : 2) : (c == 'A' ? 2 : c == 'G' ? 0 : 1)); }
  std::vector<uint8_t>().swap(g);
  lap("text packed on the device");

  // ---- suffix array
  u64* sa = dmalloc<u64>(n);
  {
    int P = 0; { double per = (double)n; while (P < 3 && per > 4.0e8) { per /= 3; ++P; } }       // three symbols: classes of ~n / 3^P, at most ~0.4 G suffixes
    u64* d_count = dmalloc<u64>(64); CK(cudaMemset(d_count, 0, 64 * 8));
    class_histogram<<<G, B>>>(tw, n, P, d_count);
    u64 count[64]; CK(cudaMemcpy(count, d_count, 64 * 8, cudaMemcpyDeviceToHost));
    u64 biggest = 0; for (int c = 0; c < (1 << (2 * P)); ++c) biggest = std::max(biggest, count[c]);
    u64 *k0 = dmalloc<u64>(biggest), *k1 = dmalloc<u64>(biggest), *p0 = dmalloc<u64>(biggest), *p1 = dmalloc<u64>(biggest);
    u64 *first = dmalloc<u64>(biggest), *tie_off = dmalloc<u64>(biggest + 1), *tie64 = dmalloc<u64>(biggest);
    unsigned char* tie = dmalloc<unsigned char>(biggest);
    u64* d_cursor = dmalloc<u64>(1);
    u64 row0 = 0; int rounds_max = 0; u64 ties_total = 0;
    for (u32 cls = 0; cls < (1u << (2 * P)); ++cls) {
      const u64 m = count[cls];
      if (!m) continue;
      CK(cudaMemset(d_cursor, 0, 8));
      class_collect<<<G, B>>>(tw, n, P, cls, k0, p0, d_cursor);
      sort_pairs(k0, k1, p0, p1, m, 0, 64);
      place_and_flag<<<G, B>>>(k0, p0, m, row0, sa, tie);
      run_first<<<G, B>>>(k0, m, first); max_scan(first, m);
      widen8<<<G, B>>>(tie, m, tie64);                                       // positions of the tied slots in the refinement list
      u64 ne = exclusive_sum(tie64, tie_off, m);
      ties_total += ne;
      if (ne) {
        Tie *e0 = dmalloc<Tie>(ne), *e1 = dmalloc<Tie>(ne);
        TieK *v0 = dmalloc<TieK>(ne), *v1 = dmalloc<TieK>(ne);
        u64 *key = dmalloc<u64>(ne), *key2 = dmalloc<u64>(ne), *seg0 = dmalloc<u64>(ne), *seg1 = dmalloc<u64>(ne), *sf = dmalloc<u64>(ne), *rf = dmalloc<u64>(ne);
        unsigned char* tflag = dmalloc<unsigned char>(ne);
        make_ties<<<G, B>>>(p0, first, tie_off, tie, m, row0, e0);
        u64 depth = (u64)P + 32; int rounds = 0;
        const int seg_bits = 64 - __builtin_clzll(n | 1);
        while (ne) {
          tie_keys<<<G, B>>>(tw, n, e0, ne, depth, key);
          sort_pairs(key, key2, e0, e1, ne, 0, 64);                        // by the next 32 symbols ...
          tie_to_segkeys<<<G, B>>>(e0, key, ne, seg0, v0);
          sort_pairs(seg0, seg1, v0, v1, ne, 0, seg_bits);                 // ... then, stably, by group
          tie_heads<<<G, B>>>(seg0, v0, ne, sf, rf); max_scan(sf, ne); max_scan(rf, ne);
          tie_place<<<G, B>>>(seg0, v0, sf, rf, ne, sa, tflag, e1);
          ne = compact(e1, tflag, e0, ne);
          depth += 32; ++rounds;
          if (rounds > 200000) die("suffix sorting does not converge (megabase exact repeats?)");
        }
        rounds_max = std::max(rounds_max, rounds);
        cudaFree(e0); cudaFree(e1); cudaFree(v0); cudaFree(v1); cudaFree(key); cudaFree(key2); cudaFree(seg0); cudaFree(seg1); cudaFree(sf); cudaFree(rf); cudaFree(tflag);
      }
      row0 += m;
    }
    if (row0 != n) die("class sizes do not add up");
    if (verbose) fprintf(stderr, "[bmbs-index-gpu] %d classes, %llu suffixes tied after the first 32-symbol key, at most %d refinement rounds\n", 1 << (2 * P), (unsigned long long)ties_total, rounds_max);
    cudaFree(k0); cudaFree(k1); cudaFree(p0); cudaFree(p1); cudaFree(first); cudaFree(tie_off); cudaFree(tie64); cudaFree(tie); cudaFree(d_cursor); cudaFree(d_count);
  }
  lap("suffix array");
  fprintf(stderr, "bmbs-index-gpu: suffix array done\n");

  // ---- BWT bit-planes + counters + 65536-symbol table
  const u64 R = n + 1, S = n;
  u64* d_shap = dmalloc<u64>(1); CK(cudaMemset(d_shap, 0, 8));
  find_shapline<<<G, B>>>(sa, n, d_shap);
  u64 shapline = 0; CK(cudaMemcpy(&shapline, d_shap, 8, cudaMemcpyDeviceToHost)); cudaFree(d_shap);
  const u64 bwt_words = 1 + 2 * (S / 64) + (S / 128) + 2;
  const u64 n_blocks = (S + 63) / 64;
  std::vector<uint64_t> bwt(bwt_words + 8, 0), high_occ(2 * (S / 65536 + 1), 0);
  u64 cnt1 = 0, cnt2 = 0;
  {
    const u64 alloc_words = ((n_blocks + 1) / 2 + 1) * 5 + 16;
    u64* d_bwt = dmalloc<u64>(alloc_words); CK(cudaMemset(d_bwt, 0, alloc_words * 8));
    u32 *c1 = dmalloc<u32>(n_blocks + 1), *c2 = dmalloc<u32>(n_blocks + 1);
    CK(cudaMemset(c1, 0, (n_blocks + 1) * 4)); CK(cudaMemset(c2, 0, (n_blocks + 1) * 4));
    bwt_blocks<<<G, B>>>(tw, sa, n, shapline, n_blocks, d_bwt, c1, c2);
    u64 *w1 = dmalloc<u64>(n_blocks + 2), *w2 = dmalloc<u64>(n_blocks + 2), *s1 = dmalloc<u64>(n_blocks + 2), *s2 = dmalloc<u64>(n_blocks + 2);
    widen<<<G, B>>>(c1, n_blocks, w1); widen<<<G, B>>>(c2, n_blocks, w2);
    cnt1 = exclusive_sum(w1, s1, n_blocks); cnt2 = exclusive_sum(w2, s2, n_blocks);
    u64* d_high = dmalloc<u64>(high_occ.size()); CK(cudaMemset(d_high, 0, high_occ.size() * 8));
    bwt_high_occ<<<G, B>>>(s1, s2, n, n_blocks, d_high);
    bwt_headers<<<G, B>>>(s1, s2, n, n_blocks, d_bwt);
    CK(cudaMemcpy(bwt.data(), d_bwt, std::min<u64>(bwt_words + 8, alloc_words) * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(high_occ.data(), d_high, high_occ.size() * 8, cudaMemcpyDeviceToHost));
    cudaFree(d_bwt); cudaFree(c1); cudaFree(c2); cudaFree(w1); cudaFree(w2); cudaFree(s1); cudaFree(s2); cudaFree(d_high);
  }
  lap("bwt planes, counters");
  // ---- flag bit-vector + sampled suffix array
  const u64 flag_words_n = 1 + R / 64 + (R % 64 ? 1 : 0) + R / 256 + 1;
  std::vector<uint64_t> flag(flag_words_n + 8, 0);
  std::vector<uint32_t> ssa;
  {
    const u64 n_fw = (R + 63) / 64;
    u64* bits = dmalloc<u64>(n_fw + 1); u32* fc = dmalloc<u32>(n_fw + 1);
    flag_words<<<G, B>>>(sa, n, R, n_fw, bits, fc);
    u64 *wc = dmalloc<u64>(n_fw + 2), *rank = dmalloc<u64>(n_fw + 2);
    widen<<<G, B>>>(fc, n_fw, wc);
    const u64 total = exclusive_sum(wc, rank, n_fw);
    const u64 alloc = ((n_fw + 3) / 4 + 2) * 5 + 16;
    u64* d_flag = dmalloc<u64>(alloc); CK(cudaMemset(d_flag, 0, alloc * 8));
    flag_layout<<<G, B>>>(bits, rank, n_fw, R, total, d_flag);
    u32* d_ssa = dmalloc<u32>(total + 1);
    sampled_sa<<<G, B>>>(tw, sa, n, R, bits, rank, d_ssa);
    CK(cudaMemcpy(flag.data(), d_flag, std::min<u64>(flag_words_n + 8, alloc) * 8, cudaMemcpyDeviceToHost));
    ssa.resize(total); CK(cudaMemcpy(ssa.data(), d_ssa, total * 4, cudaMemcpyDeviceToHost));
    cudaFree(bits); cudaFree(fc); cudaFree(wc); cudaFree(rank); cudaFree(d_flag); cudaFree(d_ssa);
  }
  cudaFree(sa);
  lap("flags, sampled suffix array");
  // ---- 16-mer counts
  const u64 H = 43046721ull;
  std::vector<uint32_t> kc(H, 0);
  {
    u32* d_kc = dmalloc<u32>(H); CK(cudaMemset(d_kc, 0, H * 4));
    if (n >= 16) count_16mers<<<G, B>>>(tw, n, d_kc);
    CK(cudaMemcpy(kc.data(), d_kc, H * 4, cudaMemcpyDeviceToHost)); cudaFree(d_kc);
  }
  cudaFree(tw);
  lap("16-mer counts");
  const u64 cnt0 = S - cnt1 - cnt2;
  write_bwt_files(fa, R, shapline, cnt0, cnt1, cnt2, bwt, bwt_words, high_occ, ssa, flag, flag_words_n, kc, tail, n);
  lap("files written");
  return 0;
}
