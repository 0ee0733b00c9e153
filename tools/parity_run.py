#!/usr/bin/env python3
"""End-to-end parity run: synthetic genome + reads -> index -> reference CPU mapper vs GPU mapper, SAM diff.

usage: parity_run.py [--work DIR] [--genome-mbp 1] [--reads 20000] [--len 100] [--pairs 10000] [--plen 150] [--seed 11]
The CPU side is oracle/_ref/bitmapperBS (the real reference) when present, else oracle/_build/oracle_cli.
"""
import argparse
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bitmapperbs_b200 import simulate as S  # noqa: E402

REF = ROOT / "oracle/_ref/bitmapperBS"
ORACLE = ROOT / "oracle/_build/oracle_cli"
BMBS = ROOT / "bitmapperbs_b200/_build/bmbs"
INDEXER = ROOT / "bitmapperbs_b200/_build/bmbs-index"


def sam_body(path):
    with open(path, "rb") as f:
        return sorted(l for l in f if not l.startswith(b"@"))


def run(cmd, cwd=None, quiet=True):
    t = time.time()
    r = subprocess.run([str(c) for c in cmd], cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if r.returncode != 0:
        sys.stderr.write(r.stderr.decode(errors="replace")[-2000:])
        raise SystemExit(f"command failed: {cmd}")
    return time.time() - t, r.stderr.decode(errors="replace")


def diff(a, b, label, show=6):
    la, lb = sam_body(a), sam_body(b)
    sa, sb = set(la), set(lb)
    only_a, only_b = [x for x in la if x not in sb], [x for x in lb if x not in sa]
    print(f"[{label}] cpu records {len(la)}  gpu records {len(lb)}  only-cpu {len(only_a)}  only-gpu {len(only_b)}")
    for x in only_a[:show]:
        print("   cpu:", x.decode()[:200].replace("\t", " "))
    for x in only_b[:show]:
        print("   gpu:", x.decode()[:200].replace("\t", " "))
    return len(only_a) + len(only_b)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--work", default="/tmp/bmbs_parity")
    ap.add_argument("--genome-mbp", type=float, default=1.0)
    ap.add_argument("--reads", type=int, default=20000)
    ap.add_argument("--len", type=int, default=100)
    ap.add_argument("--pairs", type=int, default=10000)
    ap.add_argument("--plen", type=int, default=150)
    ap.add_argument("--seed", type=int, default=11)
    ap.add_argument("--repeat", type=float, default=0.3)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    w = Path(a.work); w.mkdir(parents=True, exist_ok=True)
    n = int(a.genome_mbp * 1e6)
    chroms = S.random_genome([int(n * 0.6), n - int(n * 0.6)], seed=a.seed, repeat_fraction=a.repeat, repeat_copies=(3, 30), repeat_len=(300, 3000))
    S.write_fasta(w / "g.fa", chroms)
    r1, _ = S.simulate_reads(chroms, a.reads, a.len, seed=a.seed + 10, sub=0.02, indel=0.002, n_rate=0.001, random_qual=True, junk_fraction=0.02)
    S.write_fastq(w / "r.fq", r1)
    p1, p2 = S.simulate_reads(chroms, a.pairs, a.plen, seed=a.seed + 11, paired=True, sub=0.01, indel=0.001, random_qual=True)
    S.write_fastq(w / "p1.fq", p1); S.write_fastq(w / "p2.fq", p2)
    t, _ = run([INDEXER, w / "g.fa"]); print(f"index built in {t:.1f}s")
    bad = 0
    cpu = REF if REF.exists() else None
    # single end
    if cpu:
        t, _ = run([cpu, "--search", "g.fa", "--seq", "r.fq", "-t", a.threads, "-o", "cpu_se.sam", "--mapstats", "cpu_se.stats"], cwd=w)
    else:
        t, _ = run([ORACLE, "se", "g.fa", "r.fq", "cpu_se.sam"], cwd=w)
    print(f"cpu SE {t:.2f}s ({'reference' if cpu else 'oracle port'})")
    t, err = run([BMBS, "--search", "g.fa", "--seq", "r.fq", "-t", a.threads, "-o", "gpu_se.sam", "--mapstats", "gpu_se.stats"], cwd=w)
    print(f"gpu SE {t:.2f}s"); print(err[-700:])
    bad += diff(w / "cpu_se.sam", w / "gpu_se.sam", "SE")
    if cpu:
        same = open(w / "cpu_se.stats").read() == open(w / "gpu_se.stats").read()
        print("[SE] mapstats identical:", same); bad += 0 if same else 1
    # paired end
    if cpu:
        t, _ = run([cpu, "--search", "g.fa", "--seq1", "p1.fq", "--seq2", "p2.fq", "--pe", "-t", a.threads, "-o", "cpu_pe.sam", "--mapstats", "cpu_pe.stats"], cwd=w)
    else:
        t, _ = run([ORACLE, "pe", "g.fa", "p1.fq", "p2.fq", "cpu_pe.sam"], cwd=w)
    print(f"cpu PE {t:.2f}s")
    t, err = run([BMBS, "--search", "g.fa", "--seq1", "p1.fq", "--seq2", "p2.fq", "--pe", "-t", a.threads, "-o", "gpu_pe.sam", "--mapstats", "gpu_pe.stats"], cwd=w)
    print(f"gpu PE {t:.2f}s"); print(err[-700:])
    bad += diff(w / "cpu_pe.sam", w / "gpu_pe.sam", "PE")
    if cpu:
        same = open(w / "cpu_pe.stats").read() == open(w / "gpu_pe.stats").read()
        print("[PE] mapstats identical:", same); bad += 0 if same else 1
    print("PARITY", "OK" if bad == 0 else f"FAILED ({bad})")
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
