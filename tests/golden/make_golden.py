#!/usr/bin/env python3
"""Regenerates tests/golden/* from the REAL reference (oracle/_ref/bitmapperBS built by oracle/build_ref.sh from
/root/reference).  Run in the build container only; the fixtures are committed and travel to the GPU box.

  genome.fa.gz                      220 kbp, 3 chromosomes, 35 % diverged repeats
  se100.fq.gz  -> se100.sam.gz      3000 x 100 bp single end (subs, indels, N, random qualities, 2 % junk)
  se250.fq.gz  -> se250.sam.gz      600 x 250 bp single end (k = 20: 64-bit bands)
  pe150_[12].fq.gz -> pe150.sam.gz  1500 pairs x 150 bp, --pe (fast mode)
  pe150s.sam.gz                     same pairs, --pe --sensitive
  pe100h_[12].fq.gz -> pe100h.sam.gz (--pe) and pe100hs.sam.gz (--pe --sensitive): 2500 pairs x 100 bp with 5 % substitutions,
                                    0.8 % indels, 3 % junk -- hard enough that the sensitive mode's mate filter and
                                    re-seeding change the outcome of several hundred pairs
  *.stats                           the five --mapstats lines of each run
  index_sha256.json                 sha256 of every file `--index` wrote (the .sa hash skips its last 8 bytes,
                                    which the reference leaves uninitialised)
SAM bodies are stored without @ header lines, in the reference's -t 1 (input) order.
"""
import gzip
import hashlib
import json
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
from bitmapperbs_b200 import simulate as S  # noqa: E402

REF = ROOT / "oracle/_ref/bitmapperBS"
SHIM = ROOT / "oracle/_ref/psascan"


def gz_write(path, data: bytes):
    with open(path, "wb") as raw, gzip.GzipFile(fileobj=raw, mode="wb", mtime=0, compresslevel=9) as f:
        f.write(data)


def body(path):
    return b"".join(l for l in open(path, "rb") if not l.startswith(b"@"))


def main():
    assert REF.exists(), "build the reference first: bash oracle/build_ref.sh"
    chroms = S.random_genome([120000, 70000, 30000], seed=4242, repeat_fraction=0.35, repeat_copies=(3, 40), repeat_len=(200, 2500))
    with tempfile.TemporaryDirectory() as td:
        w = Path(td)
        S.write_fasta(w / "genome.fa", chroms)
        se100, _ = S.simulate_reads(chroms, 3000, 100, seed=1, sub=0.02, indel=0.003, n_rate=0.002, random_qual=True, junk_fraction=0.02)
        se250, _ = S.simulate_reads(chroms, 600, 250, seed=2, sub=0.03, indel=0.004, n_rate=0.001, random_qual=True, junk_fraction=0.02)
        p1, p2 = S.simulate_reads(chroms, 1500, 150, seed=3, paired=True, sub=0.015, indel=0.002, n_rate=0.001, random_qual=True, junk_fraction=0.01)
        S.write_fastq(w / "se100.fq", se100); S.write_fastq(w / "se250.fq", se250)
        S.write_fastq(w / "pe150_1.fq", p1); S.write_fastq(w / "pe150_2.fq", p2)
        h1, h2 = S.simulate_reads(chroms, 2500, 100, seed=16, paired=True, sub=0.05, indel=0.008, n_rate=0.003, random_qual=True, frag_range=(150, 420), junk_fraction=0.03)
        S.write_fastq(w / "pe100h_1.fq", h1); S.write_fastq(w / "pe100h_2.fq", h2)
        shutil.copy(SHIM, w / "psascan")
        run = lambda *a: subprocess.run([str(REF), *a], cwd=w, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        run("--index", "genome.fa")
        run("--search", "genome.fa", "--seq", "se100.fq", "-t", "1", "-o", "se100.sam", "--mapstats", "se100.stats")
        run("--search", "genome.fa", "--seq", "se250.fq", "-t", "1", "-o", "se250.sam", "--mapstats", "se250.stats")
        run("--search", "genome.fa", "--seq1", "pe150_1.fq", "--seq2", "pe150_2.fq", "--pe", "-t", "1", "-o", "pe150.sam", "--mapstats", "pe150.stats")
        run("--search", "genome.fa", "--seq1", "pe150_1.fq", "--seq2", "pe150_2.fq", "--pe", "--sensitive", "-t", "1", "-o", "pe150s.sam", "--mapstats", "pe150s.stats")
        run("--search", "genome.fa", "--seq1", "pe100h_1.fq", "--seq2", "pe100h_2.fq", "--pe", "-t", "1", "-o", "pe100h.sam", "--mapstats", "pe100h.stats")
        run("--search", "genome.fa", "--seq1", "pe100h_1.fq", "--seq2", "pe100h_2.fq", "--pe", "--sensitive", "-t", "1", "-o", "pe100hs.sam", "--mapstats", "pe100hs.stats")
        for f in ["genome.fa", "se100.fq", "se250.fq", "pe150_1.fq", "pe150_2.fq", "pe100h_1.fq", "pe100h_2.fq"]:
            gz_write(HERE / (f + ".gz"), (w / f).read_bytes())
        for f in ["se100", "se250", "pe150", "pe150s", "pe100h", "pe100hs"]:
            gz_write(HERE / (f + ".sam.gz"), body(w / (f + ".sam")))
            shutil.copy(w / (f + ".stats"), HERE / (f + ".stats"))
        hashes = {}
        for suf in [".index", ".index.bs.pac", ".index.bs.index", ".index.bs.index.bwt", ".index.bs.index.occ", ".index.bs.index.sa"]:
            data = (w / ("genome.fa" + suf)).read_bytes()
            if suf.endswith(".sa"):
                data = data[:-8]
            hashes[suf] = hashlib.sha256(data).hexdigest()
        (HERE / "index_sha256.json").write_text(json.dumps(hashes, indent=1) + "\n")
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
