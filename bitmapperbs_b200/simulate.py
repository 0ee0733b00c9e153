"""Synthetic genomes and simulated directional bisulfite reads (seeded numpy).

Data preparation for tests and bench.py -- the inputs SURVEY.md §8(d) specifies:
uniform-random chromosomes (optionally with diverged repeat families so that
candidate voting and verification get real work), directional WGBS reads
(C->T on the sequenced strand with probability `conv`), substitutions,
optional single-base indels, FASTQ with constant quality 'I' (or random
qualities), single-end or paired-end (mate 2 = reverse complement of the
fragment suffix, as Illumina reports it).
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


def random_genome(chrom_lengths, seed, repeat_fraction=0.0, repeat_div=(0.01, 0.15),
                  repeat_len=(1000, 10000), repeat_copies=(10, 200)):
    """Return list of (name, uint8 ASCII array). `repeat_fraction` of every
    chromosome is overwritten by copies of repeat-family consensus elements,
    each copy independently diverged (substitutions)."""
    rng = np.random.default_rng(seed)
    chroms = []
    for ci, n in enumerate(chrom_lengths):
        seq = _ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]
        chroms.append([f"chr{ci + 1}", seq])
    if repeat_fraction > 0:
        total = sum(chrom_lengths)
        budget = int(total * repeat_fraction)
        while budget > 0:
            flen = int(rng.integers(repeat_len[0], repeat_len[1] + 1))
            copies = int(rng.integers(repeat_copies[0], repeat_copies[1] + 1))
            cons = _ACGT[rng.integers(0, 4, size=flen, dtype=np.uint8)]
            div = rng.uniform(*repeat_div)
            for _ in range(copies):
                ci = int(rng.integers(0, len(chroms)))
                seq = chroms[ci][1]
                if len(seq) <= flen:
                    continue
                p = int(rng.integers(0, len(seq) - flen))
                cp = cons.copy()
                m = rng.random(flen) < div
                cp[m] = _ACGT[rng.integers(0, 4, size=int(m.sum()), dtype=np.uint8)]
                if rng.random() < 0.5:
                    cp = _COMP[cp[::-1]]
                seq[p:p + flen] = cp
                budget -= flen
                if budget <= 0:
                    break
    return [(n, s) for n, s in chroms]


def write_fasta(path, chroms, width=80):
    with open(path, "wb") as f:
        for name, seq in chroms:
            f.write(b">" + name.encode() + b"\n")
            n = len(seq)
            full = (n // width) * width
            if full:
                body = np.empty((n // width, width + 1), dtype=np.uint8)
                body[:, :width] = seq[:full].reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if full < n:
                f.write(seq[full:].tobytes() + b"\n")


def _mutate(rng, frag, conv, sub, indel):
    """Bisulfite-convert and add sequencing errors to one fragment (uint8)."""
    r = frag.copy()
    c = (r == ord("C")) & (rng.random(len(r)) < conv)
    r[c] = ord("T")
    if sub > 0:
        m = rng.random(len(r)) < sub
        k = int(m.sum())
        if k:
            r[m] = _ACGT[(np.searchsorted(_ACGT, r[m]) + rng.integers(1, 4, size=k)) % 4]
    if indel > 0:
        k = rng.binomial(len(r), indel)
        for _ in range(k):
            p = int(rng.integers(1, len(r) - 1))
            if rng.random() < 0.5:
                r = np.delete(r, p)
            else:
                r = np.insert(r, p, _ACGT[rng.integers(0, 4)])
    return r


def simulate_reads(chroms, n_reads, read_len, seed, paired=False, conv=0.98, sub=0.01,
                   indel=0.0, frag_range=(200, 480), n_rate=0.0, random_qual=False,
                   junk_fraction=0.0):
    """Return (reads1, reads2|None); each read is (name, seq bytes, qual bytes).
    A fragment is taken from the forward strand or (50%) the reverse
    complement, then converted -- directional protocol."""
    rng = np.random.default_rng(seed)
    lens = np.array([len(s) for _, s in chroms])
    prob = lens / lens.sum()
    r1, r2 = [], []
    slack = 8  # room for deletions so that reads keep their length
    for i in range(n_reads):
        ci = int(rng.choice(len(chroms), p=prob))
        seq = chroms[ci][1]
        flen = int(rng.integers(frag_range[0], frag_range[1] + 1)) if paired else read_len
        flen = max(flen, read_len)
        flen_x = flen + slack
        if len(seq) <= flen_x:
            raise ValueError("chromosome shorter than fragment")
        p = int(rng.integers(0, len(seq) - flen_x))
        frag = seq[p:p + flen_x]
        strand = int(rng.random() < 0.5)
        if strand:
            frag = _COMP[frag[::-1]]
        if junk_fraction > 0 and rng.random() < junk_fraction:
            frag = _ACGT[rng.integers(0, 4, size=flen_x, dtype=np.uint8)]
        frag = _mutate(rng, frag, conv, sub, indel)
        if len(frag) < flen:
            frag = np.concatenate([frag, _ACGT[rng.integers(0, 4, size=flen - len(frag), dtype=np.uint8)]])
        frag = frag[:flen]
        a = frag[:read_len].copy()
        if n_rate > 0:
            a[rng.random(read_len) < n_rate] = ord("N")
        q = (rng.integers(35, 74, size=read_len, dtype=np.uint8) if random_qual
             else np.full(read_len, ord("I"), dtype=np.uint8))
        name = f"r{i}_{chroms[ci][0]}_{p}_{'-' if strand else '+'}"
        if paired:
            b = _COMP[frag[flen - read_len:][::-1]].copy()
            if n_rate > 0:
                b[rng.random(read_len) < n_rate] = ord("N")
            q2 = (rng.integers(35, 74, size=read_len, dtype=np.uint8) if random_qual
                  else np.full(read_len, ord("I"), dtype=np.uint8))
            r1.append((name + "/1", a.tobytes(), q.tobytes()))
            r2.append((name + "/2", b.tobytes(), q2.tobytes()))
        else:
            r1.append((name, a.tobytes(), q.tobytes()))
    return r1, (r2 if paired else None)


def write_fastq(path, reads):
    with open(path, "wb") as f:
        for name, seq, qual in reads:
            f.write(b"@" + name.encode() + b"\n" + seq + b"\n+\n" + qual + b"\n")


# ---------------------------------------------------------------------------------------------
# vectorised generators for bench.py (millions of reads in seconds)
def concat_genome(chroms):
    """-> (uint8 genome, int64 chromosome start offsets incl. the end)"""
    seq = np.concatenate([s for _, s in chroms])
    starts = np.zeros(len(chroms) + 1, dtype=np.int64)
    starts[1:] = np.cumsum([len(s) for _, s in chroms])
    return seq, starts


def _revcomp_rows(m):
    return _COMP[m[:, ::-1]]


def simulate_fast(genome, starts, n, L, seed, paired=True, frag_range=(200, 480), conv=0.98, sub=0.01, indel_reads=0.0):
    """Directional bisulfite reads as uint8 matrices.
    Returns (mate1 [n,L], mate2 [n,L] or None).  mate2 is in FASTQ orientation (reverse complement of the fragment
    suffix); pass revcomp rows to the C ABI.  `indel_reads` (single end only): fraction of reads that carry one
    single-base insertion or deletion (0.001 indels per base over 150 bases ~ 0.14)."""
    rng = np.random.default_rng(seed)
    lens = np.diff(starts)
    c = rng.choice(len(lens), size=n, p=lens / lens.sum())
    flen = rng.integers(frag_range[0], frag_range[1] + 1, size=n) if paired else np.full(n, L)
    flen = np.maximum(flen, L)
    p = starts[c] + (rng.random(n) * (lens[c] - flen - 1)).astype(np.int64)
    strand = rng.random(n) < 0.5
    ar = np.arange(L, dtype=np.int64)
    left = genome[p[:, None] + ar]                      # G[p : p+L]
    right = genome[(p + flen - L)[:, None] + ar]        # G[p+flen-L : p+flen]
    m1 = np.where(strand[:, None], _revcomp_rows(right), left)
    tail = np.where(strand[:, None], _revcomp_rows(left), right)   # fragment suffix on the sequenced strand

    def damage(m):
        cc = (m == ord("C")) & (rng.random(m.shape) < conv)
        m = np.where(cc, ord("T"), m).astype(np.uint8)
        if sub > 0:
            e = rng.random(m.shape) < sub
            k = int(e.sum())
            m[e] = _ACGT[(np.searchsorted(_ACGT, m[e]) + rng.integers(1, 4, size=k)) % 4]
        return m

    m1 = damage(m1)
    if not paired and indel_reads > 0:
        # one indel per chosen read: deletion = skip one fragment base (the read takes the next genome base at its end),
        # insertion = one random base pushed in (the last base falls off); both by index arithmetic on the forward read
        pick = np.flatnonzero(rng.random(n) < indel_reads)
        pos = rng.integers(5, L - 5, size=len(pick))
        is_del = rng.random(len(pick)) < 0.5
        sub_m = m1[pick]
        ar2 = np.arange(L)[None, :]
        src_del = np.minimum(ar2 + (ar2 >= pos[:, None]), L - 1)
        src_ins = ar2 - (ar2 > pos[:, None])
        src = np.where(is_del[:, None], src_del, src_ins)
        out = np.take_along_axis(sub_m, src, axis=1)
        ins_rows = np.flatnonzero(~is_del)
        out[ins_rows, pos[ins_rows]] = _ACGT[rng.integers(0, 4, size=len(ins_rows))]
        m1[pick] = out
    if not paired:
        return m1, None
    tail = damage(tail)
    return m1, _revcomp_rows(tail)


def write_fastq_matrix(path, m, tag, qual=ord("I")):
    """fixed-width FASTQ records straight from a read matrix"""
    n, L = m.shape
    names = np.char.add(np.char.add("@r", np.char.zfill(np.arange(n).astype(str), 9)), tag)
    nb = np.frombuffer("".join(names.tolist()).encode(), dtype=np.uint8).reshape(n, -1)
    w = nb.shape[1]
    rec = np.empty((n, w + 1 + L + 3 + L + 1), dtype=np.uint8)
    rec[:, :w] = nb; rec[:, w] = 10
    rec[:, w + 1:w + 1 + L] = m; rec[:, w + 1 + L] = 10; rec[:, w + 2 + L] = ord("+"); rec[:, w + 3 + L] = 10
    rec[:, w + 4 + L:w + 4 + 2 * L] = qual; rec[:, -1] = 10
    rec.tofile(path)
