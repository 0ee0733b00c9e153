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
