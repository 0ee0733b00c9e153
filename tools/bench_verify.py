#!/usr/bin/env python3
"""Verification-kernel microbench (BASELINE.json configs[4], SURVEY.md §8d cfg 5).

Per point (read length L in {100,150,250} x edit rate e in {0,1,2,4,6,8} %): 2^21 reads simulated from the 100 Mbp
synthetic genome (98 % C->T, then e edits: 2/3 substitutions, 1/6 insertions, 1/6 deletions; half on the reverse-complement
strand), each verified against 8 candidate windows -- the true locus shifted by 0, +-1, +-2, +-3 bases plus one decoy
window elsewhere -- = 2^24 windows, through bmbs_batch_verify (kernel 3 alone).  k = floor(0.08 L) = 8 / 12 / 20.

Reported per point: GCUPS = windows * L * (2k+1) / verify_windows time (CUDA events on the library stream), the
integer-op rate (14 word-ops per column per band word) against the measured LOP3+IADD peak of the
device, bit-exact parity of ALL 2^24 results against the oracle restatement, and the reference's own AVX2 kernels
(BS_Reserve_Banded_BPM_8_SSE / _4_SSE through oracle/_ref/libref_bpm.so) on the host cores over a bounded sample.

  python tools/bench_verify.py [--reads-log2 21] [--out profiles/rNN_verify_microbench.json]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))


def make_point(torch, D, N, L, k, e, n_reads, seed, dev):
    """-> reads [n,L] uint8 (GPU), sites [n,8] int64 (GPU)"""
    g = torch.Generator(device=dev); g.manual_seed(seed)
    pad = k + 8
    pos = (torch.rand(n_reads, generator=g, device=dev, dtype=torch.float64) * (N - L - 2 * pad - 16)).long() + pad
    strand = torch.rand(n_reads, generator=g, device=dev) < 0.5
    s0 = pos + strand.long() * N                                   # double-strand coordinate of the read start
    ev = torch.rand(n_reads, L, generator=g, device=dev)
    is_sub = ev < e * 2 / 3
    is_ins = (ev >= e * 2 / 3) & (ev < e * 5 / 6)
    is_del = (ev >= e * 5 / 6) & (ev < e)
    shift = torch.cumsum(is_del.long(), 1) - torch.cumsum(is_ins.long(), 1)
    src = (s0[:, None] + torch.arange(L, device=dev)[None, :] + shift).clamp_(0, 2 * N - 1)
    base = D[src]
    conv = (base == ord("C")) & (torch.rand(n_reads, L, generator=g, device=dev) < 0.98)
    base = torch.where(conv, torch.full_like(base, ord("T")), base)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    rnd = acgt[torch.randint(0, 4, (n_reads, L), generator=g, device=dev)]
    code = ((base >> 1) & 3) ^ ((base >> 2) & 1)                    # A0 C1 G2 T3
    other = acgt[(code.long() + torch.randint(1, 4, (n_reads, L), generator=g, device=dev)) % 4]
    base = torch.where(is_sub, other, base)
    base = torch.where(is_ins, rnd, base)
    delta = torch.tensor([0, 1, -1, 2, -2, 3, -3, 0], device=dev)
    sites = s0[:, None] - k + delta[None, :]
    decoy = (torch.rand(n_reads, generator=g, device=dev, dtype=torch.float64) * (N - L - 2 * pad - 16)).long() + pad
    sites[:, 7] = decoy + (~strand).long() * N
    return base.contiguous(), sites.contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads-log2", type=int, default=21)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu-sample-log2", type=int, default=16, help="reads (x8 windows) per CPU reference run")
    ap.add_argument("--lengths", default="100,150,250")
    ap.add_argument("--rates", default="0,0.01,0.02,0.04,0.06,0.08")
    ap.add_argument("--cache", default=os.environ.get("BMBS_BENCH_CACHE", "/tmp/bmbs_bench"))
    ap.add_argument("--out", default="")
    ap.add_argument("--no-oracle", action="store_true")
    a = ap.parse_args()

    import torch
    import bitmapperbs_b200 as B
    from bitmapperbs_b200 import capi
    import bench
    from oracle_binding import OracleIndex
    dev = "cuda:0"
    d = bench.ensure_dataset(Path(a.cache), 1.0, True)
    genome = np.load(d / "genome.npy")
    N = len(genome)
    G = torch.from_numpy(genome).to(dev)
    comp = torch.zeros(256, dtype=torch.uint8, device=dev)
    for x, y in zip(b"ACGT", b"TGCA"):
        comp[x] = y
    D = torch.cat([G, comp[G.flip(0).long()]])
    index = B.Index(d / "g.fa.index", devices=(0,))
    oidx = None if a.no_oracle else OracleIndex(d / "g.fa.index")
    cores = os.cpu_count() or 1
    peak = capi.int_pipe_peak(0)
    print(f"[verify-bench] measured integer pipe peak (LOP3+IADD mix): {peak / 1e12:.2f} T ops/s", file=sys.stderr)
    refl = None
    p = ROOT / "oracle/_ref/libref_bpm.so"
    if p.exists():
        refl = C.CDLL(str(p))
        refl.ref_bpm_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_ushort, C.c_int, C.c_void_p, C.c_void_p]
    n_reads = 1 << a.reads_log2
    n_items = n_reads * 8
    points = []
    batch = None
    for L in [int(x) for x in a.lengths.split(",")]:
        k = min(31, int(0.08 * L))
        if batch is not None:
            batch.close()
        batch = B.Batch(index, 0, n_reads, n_reads * L + 64, n_items + 1024)
        h_reads = torch.empty(n_reads * L + 64, dtype=torch.uint8, pin_memory=True)
        offs = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(L))
        ridx = np.repeat(np.arange(n_reads, dtype=np.uint32), 8)
        for e in [float(x) for x in a.rates.split(",")]:
            reads, sites = make_point(torch, D, N, L, k, e, n_reads, 3000 + L + int(e * 1000), dev)
            h_reads[: n_reads * L].copy_(reads.view(-1)); torch.cuda.synchronize()
            flat = h_reads.numpy()[: n_reads * L]
            sites_h = sites.view(-1).cpu().numpy().astype(np.uint64)
            batch.upload(flat, offs)
            ms = []
            for rep in range(a.reps + 2):
                batch.verify(ridx, sites_h, 0.08)
                t = batch.timings()
                if rep >= 2:
                    ms.append(t["verify"])
            cnt = batch.counters()
            end, err = batch.download_verify()
            t_ms = float(np.median(ms))
            cells = n_items * L * (2 * k + 1)
            assert cnt["cells"] == cells, (cnt["cells"], cells)
            words = 1 if k <= 15 else 2
            int_ops = n_items * L * 14 * words
            pt = {"L": L, "k": k, "edit_rate": e, "windows": n_items, "verify_ms": t_ms, "gcups": cells / (t_ms / 1e3) / 1e9,
                  "int_ops_per_s": int_ops / (t_ms / 1e3), "int_roofline_frac": int_ops / (t_ms / 1e3) / peak,
                  "hits": int((end >= 0).sum()), "window_bytes": cnt["window_bytes"]}
            if oidx is not None:
                t0 = time.time()
                oend, oerr = oidx.verify((flat, offs), ridx, sites_h, threads=cores)
                pt["oracle_identical"] = bool(np.array_equal(end, oend) and np.array_equal(err, oerr))
                pt["oracle_s"] = time.time() - t0
                assert pt["oracle_identical"], f"GPU verification differs from the oracle at L={L} e={e}"
            if refl is not None:
                ns = min(n_reads, 1 << a.cpu_sample_log2)
                plen = L + 2 * k; stride = plen + 2
                widx = sites[:ns].reshape(-1)[:, None] + torch.arange(plen, device=dev)[None, :]
                wins = torch.zeros(ns * 8, stride, dtype=torch.uint8, device=dev)
                wins[:, :plen] = D[widx.clamp_(0, 2 * N - 1)]
                wins_h = np.ascontiguousarray(wins.cpu().numpy()); rd_h = np.ascontiguousarray(reads[:ns].cpu().numpy())
                s_out = np.zeros(ns * 8, dtype=np.int32); e_out = np.zeros(ns * 8, dtype=np.uint32)
                best = 1e9
                for _ in range(3):
                    t0 = time.perf_counter()
                    refl.ref_bpm_batch(wins_h.ctypes.data, ns, stride, plen, rd_h.ctypes.data, L, k, cores, s_out.ctypes.data, e_out.ctypes.data)
                    best = min(best, time.perf_counter() - t0)
                pt["cpu_reference_gcups"] = ns * 8 * L * (2 * k + 1) / best / 1e9
                pt["cpu_reference_identical"] = bool(np.array_equal(s_out, end[: ns * 8]) and np.array_equal(e_out, err[: ns * 8]))
                pt["cpu_cores"] = cores
                assert pt["cpu_reference_identical"], f"GPU verification differs from the reference's AVX2 kernel at L={L} e={e}"
            print(json.dumps(pt), flush=True)
            points.append(pt)
    out = {"workload": "cfg5 verification microbench: 2^%d reads x 8 windows per point, 100 Mbp genome, k = floor(0.08 L)" % a.reads_log2,
           "int_pipe_peak_ops_per_s": peak, "int_ops_per_cell_column": "14 word-ops per column per band word (SURVEY.md 8d)", "points": points}
    if a.out:
        Path(a.out).write_text(json.dumps(out, indent=1) + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
