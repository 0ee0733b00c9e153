#!/usr/bin/env python3
"""Whole-program parity and wall-clock at BASELINE.json's config sizes: the real reference (oracle/_ref/bitmapperBS) and
the GPU mapper (bitmapperbs_b200/_build/bmbs) on the same FASTQ files and index, SAM records diffed after sorting
(the reference's -t N output order is not deterministic) and the five --mapstats lines compared.

  cfg1   10 Mbp genome, 100 k x 100 bp single-end reads                       (BASELINE.json configs[0])
  cfg2   100 Mbp genome, 1 M x 2 x 150 bp pairs, --pe                         (configs[1])
  cfg2s  same pairs, --pe --sensitive
  cfg3r  reduced stand-in for configs[2]: 100 Mbp genome with 50 % repeat families, 200 k x 150 bp single-end reads with
         1 % substitutions and 0.1 % indels
  cfg3   configs[2] itself: 3.1 Gbp genome (24 chromosomes, half repeat families), 10 M x 150 bp single-end reads (--pairs)
  cfg4   configs[3]: the same genome, 10 M x 2 x 150 bp pairs, --pe --sensitive (cfg4f: --pe fast)

  python tools/cli_compare.py [cfg1 cfg2 cfg2s cfg3r] [--pairs N] [--out profiles/rNN_cli_compare.json]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import re
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bitmapperbs_b200 import simulate as S  # noqa: E402

REF = ROOT / "oracle/_ref/bitmapperBS"
BMBS = ROOT / "bitmapperbs_b200/_build/bmbs"
INDEXER = ROOT / "bitmapperbs_b200/_build/bmbs-index"


def sorted_body_digest(path):
    lines = sorted(l for l in open(path, "rb") if not l.startswith(b"@"))
    h = hashlib.sha256()
    for l in lines:
        h.update(l)
    return len(lines), h.hexdigest()


def run(cmd, cwd):
    t = time.time()
    r = subprocess.run([str(c) for c in cmd], cwd=cwd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    wall = time.time() - t
    m = re.search(r"Total:\s+([0-9.]+)\s+([0-9.]+)", r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"{cmd[0]} failed: {r.stderr[-2000:]}")
    return wall, (float(m.group(1)), float(m.group(2))) if m else (None, None)


def dataset(cache: Path, name: str, chrom_lens, seed, **repeat):
    d = cache / name
    if not (d / ".done").exists():
        d.mkdir(parents=True, exist_ok=True)
        chroms = S.random_genome(chrom_lens, seed=seed, **repeat)
        S.write_fasta(d / "g.fa", chroms)
        g, st = S.concat_genome(chroms)
        np.save(d / "genome.npy", g); np.save(d / "starts.npy", st)
        subprocess.run([str(INDEXER), str(d / "g.fa")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        (d / ".done").write_text("ok")
    return d, np.load(d / "genome.npy"), np.load(d / "starts.npy")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cfgs", nargs="*", default=["cfg1", "cfg2", "cfg2s"])
    ap.add_argument("--pairs", type=int, default=1_000_000)
    ap.add_argument("--cache", default=os.environ.get("BMBS_BENCH_CACHE", "/tmp/bmbs_bench"))
    ap.add_argument("--out", default="")
    ap.add_argument("--genome-scale", type=float, default=1.0, help="cfg3: fraction of the 3.1 Gbp genome")
    a = ap.parse_args()
    cache = Path(a.cache); cores = os.cpu_count() or 1
    results = []
    for cfg in a.cfgs:
        if cfg == "cfg1":
            d, g, st = dataset(cache, "cfg1_s1001", [10_000_000], 1001)
            m1, _ = S.simulate_fast(g, st, 100_000, 100, 2001, paired=False)
            S.write_fastq_matrix(d / "r.fq", m1, "")
            args = ["--seq", "r.fq"]; n_reads = len(m1)
        elif cfg in ("cfg2", "cfg2s"):
            d, g, st = dataset(cache, "cfg2_s1002_x1", [40_000_000, 30_000_000, 20_000_000, 10_000_000], 1002)
            fa, fb = d / f"cmp_{a.pairs}_1.fq", d / f"cmp_{a.pairs}_2.fq"
            if not fa.exists():
                m1, m2 = S.simulate_fast(g, st, a.pairs, 150, 2002)
                S.write_fastq_matrix(fa, m1, "/1"); S.write_fastq_matrix(fb, m2, "/2")
            args = ["--seq1", fa.name, "--seq2", fb.name, "--pe"] + (["--sensitive"] if cfg == "cfg2s" else [])
            n_reads = 2 * a.pairs
        elif cfg == "cfg3r":
            d, g, st = dataset(cache, "cfg3r_s1003", [40_000_000, 30_000_000, 20_000_000, 10_000_000], 1003,
                               repeat_fraction=0.5, repeat_len=(1000, 10000), repeat_copies=(10, 2000), repeat_div=(0.01, 0.15))
            f = d / f"se_{a.pairs}.fq"
            if not f.exists():
                chroms = [(f"chr{i + 1}", g[st[i]:st[i + 1]]) for i in range(len(st) - 1)]
                n_slow = min(a.pairs, 200_000)      # reads with indels come from the per-read simulator
                r1, _ = S.simulate_reads(chroms, n_slow, 150, seed=2003, sub=0.01, indel=0.001)
                S.write_fastq(f, r1)
            args = ["--seq", f.name]; n_reads = min(a.pairs, 200_000)
        elif cfg in ("cfg4", "cfg4f"):
            # BASELINE.json configs[3]: the 3.1 Gbp genome of cfg3, 10 M x 2 x 150 bp pairs, --pe --sensitive (cfg4f: --pe fast)
            mbp = [250, 243, 198, 190, 181, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57]
            scale = a.genome_scale
            t0 = time.time()
            d, g, st = dataset(cache, f"cfg3_s1003_x{scale:g}", [int(x * 1_000_000 * scale) for x in mbp], 1003,
                               repeat_fraction=0.5, repeat_len=(1000, 10000), repeat_copies=(10, 10000), repeat_div=(0.01, 0.15))
            out_extra = {"genome_bases": int(len(g)), "dataset_s": time.time() - t0}
            fa, fb = d / f"pe_{a.pairs}_1.fq", d / f"pe_{a.pairs}_2.fq"
            if not fa.exists():
                with open(fa, "wb") as f1, open(fb, "wb") as f2:
                    done = 0
                    while done < a.pairs:
                        m = min(1_000_000, a.pairs - done)
                        m1, m2 = S.simulate_fast(g, st, m, 150, 2004 + done)
                        t1, t2 = d / ".chunk1.fq", d / ".chunk2.fq"
                        S.write_fastq_matrix(t1, m1, f"_{done // 1_000_000}/1"); S.write_fastq_matrix(t2, m2, f"_{done // 1_000_000}/2")
                        f1.write(t1.read_bytes()); f2.write(t2.read_bytes()); done += m
            args = ["--seq1", fa.name, "--seq2", fb.name, "--pe"] + (["--sensitive"] if cfg == "cfg4" else [])
            n_reads = 2 * a.pairs
        elif cfg == "cfg3":
            # BASELINE.json configs[2]: 3.1 Gbp, 24 chromosomes, half of it repeat families; 10 M x 150 bp single-end reads
            mbp = [250, 243, 198, 190, 181, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57]
            scale = a.genome_scale
            t0 = time.time()
            d, g, st = dataset(cache, f"cfg3_s1003_x{scale:g}", [int(x * 1_000_000 * scale) for x in mbp], 1003,
                               repeat_fraction=0.5, repeat_len=(1000, 10000), repeat_copies=(10, 10000), repeat_div=(0.01, 0.15))
            out_extra = {"genome_bases": int(len(g)), "dataset_s": time.time() - t0}
            f = d / f"se_{a.pairs}.fq"
            if not f.exists():
                with open(f, "wb") as fo:
                    done = 0
                    while done < a.pairs:
                        m = min(1_000_000, a.pairs - done)
                        m1, _ = S.simulate_fast(g, st, m, 150, 2003 + done, paired=False, indel_reads=0.14)
                        tmp = d / ".chunk.fq"
                        S.write_fastq_matrix(tmp, m1, f"_{done // 1_000_000}")
                        fo.write(tmp.read_bytes()); done += m
            args = ["--seq", f.name]; n_reads = a.pairs
        else:
            raise SystemExit(f"unknown config {cfg}")
        out = {"config": cfg, "reads": n_reads, "host_cores": cores}
        if cfg in ("cfg3", "cfg4", "cfg4f"):
            out.update(out_extra)
        if n_reads <= 4_000_000:
            run([BMBS, "--search", "g.fa", *args, "-t", cores, "-o", "warm.sam"], d)    # page cache + CUDA context warm-up
        w, (load, mp) = run([BMBS, "--search", "g.fa", *args, "-t", cores, "-o", "gpu.sam", "--mapstats", "gpu.st"], d)
        out["gpu"] = {"wall_s": w, "load_s": load, "map_s": mp, "reads_per_s_map": n_reads / mp if mp else None}
        if REF.exists():
            w, (load, mp) = run([REF, "--search", "g.fa", *args, "-t", cores, "-o", "ref.sam", "--mapstats", "ref.st"], d)
            out["reference"] = {"wall_s": w, "load_s": load, "map_s": mp, "reads_per_s_map": n_reads / mp if mp else None}
            ng, hg = sorted_body_digest(d / "gpu.sam"); nr, hr = sorted_body_digest(d / "ref.sam")
            out["sam_records"] = ng; out["sam_identical"] = bool(ng == nr and hg == hr)
            out["mapstats_identical"] = (d / "gpu.st").read_text() == (d / "ref.st").read_text()
            out["mapstats"] = (d / "ref.st").read_text().splitlines()
        print(json.dumps(out), flush=True)
        results.append(out)
    if a.out:
        Path(a.out).write_text(json.dumps(results, indent=1) + "\n")
    bad = [r["config"] for r in results if r.get("sam_identical") is False or r.get("mapstats_identical") is False]
    if bad:
        print("MISMATCH:", bad, file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
