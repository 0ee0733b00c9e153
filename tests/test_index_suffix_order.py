"""CPU: the suffix order inside the index files against an INDEPENDENT suffix sort.

The sha256 test (test_oracle_golden.py) compares bmbs-index with a reference build whose psascan was replaced by a shim
over the same sorter, so it cannot see a sorter bug.  Here the text (complement + reverse of the genome, C->T, alphabet
G < T < A; Index.cpp:591-692, bwt.cpp:1135-1140) is suffix-sorted by numpy prefix doubling -- no code shared with
indexer/suffix_array.hpp -- and the files' contents are rebuilt from that order:
  .bs.index.bwt  every BWT symbol (bit-planes with interleaved counters, bwt.cpp:1345-1531) and the row of the whole text
  .bs.index.sa   every sampled suffix-array entry, in row order (bwt.cpp:1751-1816)
"""
import numpy as np


def suffix_array(t: np.ndarray) -> np.ndarray:
    """plain prefix doubling; a suffix that ends earlier sorts first"""
    n = len(t)
    rank = t.astype(np.int64)
    k = 1
    while True:
        nxt = np.full(n, -1, dtype=np.int64)
        nxt[: n - k] = rank[k:] if k < n else rank[:0]
        key = rank * (n + 2) + (nxt + 1)
        order = np.argsort(key, kind="stable")
        sk = key[order]
        new = np.zeros(n, dtype=np.int64)
        new[order] = np.cumsum(np.concatenate(([0], (sk[1:] != sk[:-1]).astype(np.int64))))
        rank = new
        if rank.max() == n - 1:
            return order
        k *= 2


def read_genome(path):
    seq = []
    for line in open(path, "rb"):
        if not line.startswith(b">"):
            seq.append(line.strip().upper())
    return np.frombuffer(b"".join(seq), dtype=np.uint8)


def test_index_suffix_order_matches_independent_sort(golden):
    g = read_genome(golden / "genome.fa")
    N = len(g); n = 2 * N
    code_fwd = np.zeros(256, dtype=np.uint8); code_rev = np.zeros(256, dtype=np.uint8)
    for c, v in zip(b"ACGT", (1, 0, 1, 2)): code_fwd[c] = v       # complement, then C->T:  A->T(1) C->G(0) G->C->T(1) T->A(2)
    for c, v in zip(b"ACGT", (2, 1, 0, 1)): code_rev[c] = v       # reversed strand, C->T: A(2) C->T(1) G(0) T(1)
    t = np.concatenate((code_fwd[g], code_rev[g][::-1]))
    sa = suffix_array(t)                                          # rows 1..n; row 0 is the empty suffix (SA = n)
    rows = np.concatenate(([n], sa))
    # ---- BWT symbols in row order, the row with SA = 0 dropped
    keep = rows != 0
    want_bwt = t[rows[keep] - 1]
    raw = np.fromfile(golden / "genome.fa.index.bs.index.bwt", dtype=np.uint64)
    words = int(raw[0]); bw = raw[1:1 + words]
    j = np.arange(n, dtype=np.int64)
    w = (j >> 7) * 5 + 1 + ((j >> 6) & 1) * 2
    bit = (63 - (j & 63)).astype(np.uint64)
    got_bwt = ((bw[w] >> bit) & np.uint64(1)) | (((bw[w + 1] >> bit) & np.uint64(1)) << np.uint64(1))
    assert np.array_equal(got_bwt.astype(np.uint8), want_bwt)
    hdr = np.fromfile(golden / "genome.fa.index.bs.index", dtype=np.uint64, count=2)
    assert int(hdr[0]) == n + 1 and int(hdr[1]) == int(np.nonzero(rows == 0)[0][0])      # rows, row of the whole text
    # ---- sampled suffix array: rows whose SA is a multiple of 8, in row order; bits 30-31 hold the BWT symbol
    sraw = np.fromfile(golden / "genome.fa.index.bs.index.sa", dtype=np.uint8)
    cnt = int(np.frombuffer(sraw[:8].tobytes(), dtype=np.uint64)[0])
    ssa = np.frombuffer(sraw[8:8 + 4 * cnt].tobytes(), dtype=np.uint32)
    sampled = rows[(rows & 7) == 0]
    assert cnt == len(sampled)
    assert np.array_equal(ssa & np.uint32(0x3FFFFFFF), (sampled >> 3).astype(np.uint32))
    ch = np.where(sampled == 0, 1, t[np.maximum(sampled, 1) - 1])
    assert np.array_equal(ssa >> np.uint32(30), ch.astype(np.uint32))
