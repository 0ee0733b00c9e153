#!/usr/bin/env python3
"""bench.py -- throughput of the seed-and-verify hot path on B200, next to the reference CPU mapper.

Workload (BASELINE.json configs[1], SURVEY.md §8d cfg 2): synthetic 100 Mbp genome (4 chromosomes, uniform,
seed 1002), 1 M simulated 150 bp paired-end directional bisulfite reads (fragments U[200,480], 98 % C->T,
1 % substitutions, seed 2002 + rank), `--pe` fast mode.  A "step" is one pass of the whole device pipeline
(pack -> seed -> locate -> votes -> pair filter -> verify) over the batch of 1 M pairs = 2 M reads.

  value  reads/s (mates, 2 per pair), inputs already resident in HBM, CUDA-event time on the library's stream
  e2e    the same through the C ABI with pinned HOST buffers: H2D of the reads and D2H of the per-read records
         and verified candidate lists inside the timed region (what the host mapper calls per batch)
  roofline      dominant kernel of the step against the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the REAL reference (oracle/_ref/bitmapperBS, compiled from /root/reference) with -t <host cores>
                on a bounded sample of the same reads, its own mapping timer (Bitmapper_main.cpp:262)

`--impl reference` times only that CPU run.  Multi-GPU (torchrun, one rank per GPU): read batches are sharded,
the index is replicated, no collective on the data path; per-GPU work is fixed ("weak").
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

GENOME_CHROMS = [40_000_000, 30_000_000, 20_000_000, 10_000_000]
READ_LEN = 150


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------ dataset
def ensure_dataset(cache: Path, scale: float, build: bool):
    """genome + index under `cache`; built once per box (rank 0), reused by every later run"""
    from bitmapperbs_b200 import simulate as S
    d = cache / f"cfg2_s1002_x{scale:g}"
    done = d / ".done"
    if build and not done.exists():
        d.mkdir(parents=True, exist_ok=True)
        t = time.time()
        chroms = S.random_genome([int(c * scale) for c in GENOME_CHROMS], seed=1002)
        S.write_fasta(d / "g.fa", chroms)
        g, st = S.concat_genome(chroms)
        np.save(d / "genome.npy", g); np.save(d / "starts.npy", st)
        from bitmapperbs_b200 import build as B
        idx, _ = B.build_tools()
        subprocess.run([str(idx), str(d / "g.fa")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        done.write_text("ok")
        log(f"[bench] dataset + index built in {time.time() - t:.1f}s at {d}")
    while not done.exists():
        time.sleep(0.5)
    return d


def make_reads(d: Path, pairs: int, seed: int):
    from bitmapperbs_b200 import simulate as S
    f = d / f"reads_p{pairs}_s{seed}.npz"
    if f.exists():
        z = np.load(f)
        return z["m1"], z["m2"]
    g = np.load(d / "genome.npy"); st = np.load(d / "starts.npy")
    m1, m2 = S.simulate_fast(g, st, pairs, READ_LEN, seed)
    tmp = d / f".tmp_{os.getpid()}_{seed}.npz"
    np.savez(tmp, m1=m1, m2=m2); os.replace(tmp, f)
    return m1, m2


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons every 5 ms from a thread (NVML); `nvidia-smi -lms` as the fallback."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.rows = []; self.p = None; self.gpu = gpu; self.run = False; self.max_mhz = None; self.src = None

    def start(self):
        try:
            import pynvml as N
            N.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.gpu
            h = N.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            reasons_fn = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM); reasons_fn(h)
            self.run = True; self.src = "nvml"

            def loop():
                while self.run:
                    try:
                        mhz = float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)); r = int(reasons_fn(h))
                        self.rows.append((time.time(), mhz, [k for k, b in bits.items() if r & b]))
                    except Exception:
                        pass
                    time.sleep(0.005)
            threading.Thread(target=loop, daemon=True).start()
            return
        except Exception:
            self.run = False
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.src = "nvidia-smi"
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.p.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                self.max_mhz = max(self.max_mhz or 0.0, float(f[1]))
                self.rows.append((time.time(), float(f[0]), [n for n, v in zip(names, f[2:6]) if v.lower().startswith("active")]))
            except Exception:
                continue

    def stop(self, t0=0.0, t1=1e30):
        """samples taken inside [t0, t1] (the timed region); if the region was shorter than a few sampling periods,
        every sample since start() -- the sampler is started before the warm-up, so those are under load too"""
        self.run = False
        if self.p:
            self.p.terminate()
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        rows = inside if len(inside) >= 3 else self.rows
        sm = [r[1] for r in rows]; reasons = sorted({x for r in rows for x in r[2]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm), "source": self.src,
                "window": "timed region" if len(inside) >= 3 else "warm-up + timed region"}


# ------------------------------------------------------------------------------------------------ reference CPU arm
def run_reference(d: Path, m1, m2, sample_pairs: int, threads: int, repeats: int):
    """oracle/_ref/bitmapperBS --search --pe -t threads on the first `sample_pairs` pairs; returns list of (map_s, wall_s)"""
    from bitmapperbs_b200 import simulate as S
    ref = ROOT / "oracle/_ref/bitmapperBS"
    if not ref.exists():
        return None
    n = min(sample_pairs, len(m1))
    fa, fb = d / f"ref_{n}_1.fq", d / f"ref_{n}_2.fq"
    if not fa.exists() or not fb.exists():
        S.write_fastq_matrix(fa, m1[:n], "/1"); S.write_fastq_matrix(fb, m2[:n], "/2")
    out = []
    for _ in range(repeats):
        t = time.time()
        r = subprocess.run([str(ref), "--search", "g.fa", "--seq1", fa.name, "--seq2", fb.name, "--pe", "-t", str(threads), "-o", "/dev/null"],
                           cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        wall = time.time() - t
        m = re.search(r"Total:\s+([0-9.]+)\s+([0-9.]+)", r.stderr)
        if r.returncode != 0 or not m:
            return None
        out.append((float(m.group(2)), wall))
    return n, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1_000_000, help="pairs per GPU per step")
    ap.add_argument("--scale", type=float, default=1.0, help="genome scale (1.0 = 100 Mbp)")
    ap.add_argument("--cache", default=os.environ.get("BMBS_BENCH_CACHE", "/tmp/bmbs_bench"))
    ap.add_argument("--ref-sample", type=int, default=250_000, help="pairs per reference-CPU run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inflight", type=int, default=3, help="batches in flight in the end-to-end loop")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    cache = Path(a.cache)
    cores = os.cpu_count() or 1
    workload = f"cfg2: synthetic {100 * a.scale:g} Mbp genome (4 chr, uniform, seed 1002), {a.pairs} x 2 x {READ_LEN} bp paired-end directional bisulfite reads per GPU, --pe fast mode"

    # ---------------------------------------------------------------- reference arm
    if a.impl == "reference":
        if rank != 0:
            return 0
        d = ensure_dataset(cache, a.scale, True)
        m1, m2 = make_reads(d, a.pairs, 2002)
        r = run_reference(d, m1, m2, a.ref_sample, cores, a.warmup + a.steps)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/bitmapperBS missing or failed (built by oracle/build_ref.sh from /root/reference)"}))
            return 0
        n, runs = r
        runs = runs[a.warmup:]
        map_s = sum(x[0] for x in runs)
        value = 2 * n * len(runs) / map_s
        sample = f"first {n} pairs of the workload, reference's own mapping timer (index load excluded), -t {cores}"
        print(json.dumps({
            "impl": "reference", "metric": "mapped_reads_per_sec", "value": value, "unit": "reads/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1000 * map_s / len(runs), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload, "sample": sample, "reads_per_step": 2 * n},
            "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s_per_step": sum(x[1] for x in runs) / len(runs)}))
        return 0

    # ---------------------------------------------------------------- our arm
    # libraries (NCCL's version banner) may write to stdout: keep fd 1 for the one JSON line
    sys.stdout.flush()
    json_fd = os.dup(1); os.dup2(2, 1)
    import torch
    import bitmapperbs_b200 as B
    from bitmapperbs_b200 import capi
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = local
    torch.cuda.set_device(dev)
    d = ensure_dataset(cache, a.scale, rank == 0)
    if dist:
        dist.barrier()
    from bitmapperbs_b200 import shard
    m1, m2 = make_reads(d, a.pairs, shard.rank_seed(2002, rank))
    n_pairs = len(m1); n_reads = 2 * n_pairs
    from bitmapperbs_b200.simulate import _revcomp_rows
    mates = np.empty((n_reads, READ_LEN), dtype=np.uint8)
    mates[0::2] = m1; mates[1::2] = _revcomp_rows(m2)
    bases = mates.size
    # pinned host staging (torch is plumbing here: pinned memory, device selection, rendezvous)
    h_flat = torch.empty(bases + 64, dtype=torch.uint8, pin_memory=True); h_flat.numpy()[:bases] = mates.ravel()
    h_offs = torch.empty(n_reads + 1, dtype=torch.int64, pin_memory=True); h_offs.numpy()[:] = np.arange(n_reads + 1, dtype=np.int64) * READ_LEN
    flat = h_flat.numpy()[:bases]; offs = h_offs.numpy().view(np.uint64)

    t0 = time.time()
    index = B.Index(d / "g.fa.index", devices=(dev,))
    log(f"[bench r{rank}] index resident: {index.device_bytes / 1e9:.2f} GB in {time.time() - t0:.1f}s")
    prm = capi.default_params()
    cand_cap = 10 * n_reads
    while True:
        batch = B.Batch(index, dev, n_reads, bases + 64, cand_cap)
        batch.upload(flat, offs, pe=True); batch.run(prm)
        try:
            batch.sync()
            h_res = torch.empty(n_reads * capi.ReadResult.itemsize, dtype=torch.uint8, pin_memory=True)
            h_cand = torch.empty(cand_cap * capi.Cand.itemsize, dtype=torch.uint8, pin_memory=True)
            res = h_res.numpy().view(capi.ReadResult); cand = h_cand.numpy().view(capi.Cand)
            _, _, used = batch.download(res, cand)
            break
        except B.BmbsError as e:
            if e.code != -4:
                raise
            batch.close(); cand_cap *= 2
    states = np.bincount(res["state"], minlength=5)
    # more in-flight batches for the end-to-end loop (each with its own stream and pinned result buffers): the H2D copy of
    # one batch overlaps the kernels of the previous one and the D2H copy of the one before
    outs = [(batch, res, cand)]
    keep = [h_res, h_cand]
    for _ in range(a.inflight - 1):
        hr = torch.empty(n_reads * capi.ReadResult.itemsize, dtype=torch.uint8, pin_memory=True)
        hc = torch.empty(cand_cap * capi.Cand.itemsize, dtype=torch.uint8, pin_memory=True)
        keep += [hr, hc]
        outs.append((B.Batch(index, dev, n_reads, bases + 64, cand_cap), hr.numpy().view(capi.ReadResult), hc.numpy().view(capi.Cand)))
    NB = len(outs)
    clocks = ClockSampler(dev); clocks.start()
    for _ in range(max(0, a.warmup - 1)):
        batch.run(prm)
    batch.sync()
    for i in range(max(a.warmup, NB)):
        bb, rr, cc = outs[i % NB]
        bb.upload(flat, offs, pe=True); bb.run(prm); bb.download(rr, cc)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed: device-resident steps
    barrier()
    t_begin = time.time()
    stage = {}
    dev_ms = 0.0
    w0 = time.perf_counter()
    for _ in range(a.steps):
        batch.run(prm)
        t = batch.timings()           # waits on the step's last event; CUDA-event times on the launching stream
        dev_ms += t["total"]
        for k, v in t.items():
            stage[k] = stage.get(k, 0.0) + v
    batch.sync()
    wall_ms = (time.perf_counter() - w0) * 1000
    barrier()
    counters = batch.counters()
    launches = batch.launches() * a.steps
    # ---- timed: end to end through the C ABI, host buffers in / out
    # (every step: H2D of the step's reads from pinned host memory, kernels, D2H of its records; two batches in
    # flight on two streams, as the host mapper drives them)
    barrier()
    e0 = time.perf_counter()
    for i in range(a.steps):
        bb, rr, cc = outs[i % NB]
        bb.upload(flat, offs, pe=True); bb.run(prm)
        if i >= NB - 1:
            pb, pr, pc = outs[(i - (NB - 1)) % NB]
            _, _, used = pb.download(pr, pc)
    for i in range(max(0, a.steps - (NB - 1)), a.steps):
        pb, pr, pc = outs[i % NB]
        _, _, used = pb.download(pr, pc)
    e2e_ms = (time.perf_counter() - e0) * 1000
    barrier()
    clk = clocks.stop(t_begin, time.time())
    assert all(np.array_equal(outs[0][1]["state"], o[1]["state"]) for o in outs[1:])
    h2d = int(bases + 8 * (n_reads + 1)); d2h = int(n_reads * capi.ReadResult.itemsize + used * capi.Cand.itemsize)

    if dist:
        t = torch.tensor([dev_ms, e2e_ms, wall_ms], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms, wall_ms = (float(x) for x in t.tolist())
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return 0

    value = n_reads * world * a.steps / (dev_ms / 1000)
    e2e = n_reads * world * a.steps / (e2e_ms / 1000)
    # ---- roofline of the dominant kernel (SURVEY.md §8d algorithmic bytes per unit)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0)); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    per_step = {k: v / a.steps for k, v in stage.items()}
    # algorithmic bytes per unit: 8 B per table query (one deep-table / 16-mer entry), 32 B per occ-block lookup
    seed_bytes = 8 * counters["hash_queries"] + 32 * counters["occ_lookups"]
    # sampled suffix array: 80 B per LF step + 44 B per row; dense suffix array (default): one 4-byte entry per row
    loc_bytes = counters["locate_lf_steps"] * 80 + counters["located_rows"] * (44 if counters["locate_lf_steps"] else 4)
    ver_bytes = counters["window_bytes"]
    kernels = {"seed_first+second+rest": (per_step["seed"], seed_bytes), "expand_locate": (per_step["locate"], loc_bytes), "verify_windows": (per_step["verify"], ver_bytes)}
    dom = max(kernels, key=lambda k: kernels[k][0])
    dms, dbytes = kernels[dom]
    achieved = dbytes / (dms / 1000) / 1e9 if dms > 0 else 0.0
    # DRAM traffic of the same kernels from the committed `ncu --set full` capture (profiles/ncu_traffic.json), per launch
    traffic = None
    try:
        traffic = json.loads((ROOT / "profiles/ncu_traffic.json").read_text()).get(dom, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    gcups = counters["cells"] / (per_step["verify"] / 1000) / 1e9 if per_step["verify"] > 0 else 0.0
    # the bound that applies to the seeding kernels: independent random 32-byte sectors per second (measured live)
    try:
        rs_peak = capi.random_sector_peak(dev)
    except Exception:
        rs_peak = None
    seed_sectors = counters["hash_queries"] + counters["occ_lookups"] + counters["located_rows"]
    seed_sector_rate = seed_sectors / (per_step["seed"] / 1000) if per_step["seed"] > 0 else 0.0
    out = {
        "metric": "mapped_reads_per_sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload, "reads_per_step_per_gpu": n_reads, "l2": "inputs (300 MB of reads, 0.6 GB index) exceed the 126 MB L2; no flush needed",
                   "read_states": {"none": int(states[0]), "exact_unique": int(states[1]), "multi_exact": int(states[2]), "one_mismatch": int(states[3]), "verify": int(states[4])},
                   "index_hbm_bytes": index.device_bytes},
        "e2e": {"value": e2e, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / a.steps, "batches_in_flight": NB},
        "gpu_launches": launches,
        "clocks": clk,
        "stage_ms_per_step": per_step,
        "wall_ms_per_step": wall_ms / a.steps,
        "verify_gcups": gcups,
        "work_per_step": counters,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": dbytes, "kernel_ms": dms,
                     # DRAM bytes the kernels really move (ncu, profiles/ncu_traffic.json) over the live kernel time: random 8-byte table entries and
                     # 32-byte occ blocks cost a 64-byte DRAM access each, so this -- not the algorithmic figure -- is what loads HBM
                     "traffic_gbs": traffic / (dms / 1000) / 1e9 if traffic and dms > 0 else None,
                     "traffic_frac_of_peak": traffic / (dms / 1000) / 1e9 / peak if traffic and dms > 0 and peak else None,
                     "random_sector_peak_gbs": rs_peak * 32 / 1e9 if rs_peak else None, "seed_random_sector_gbs": seed_sector_rate * 32 / 1e9,
                     "frac_of_random_sector_peak": seed_sector_rate / rs_peak if rs_peak else None,
                     "note": "after the deep seed table a seed is one 8-byte entry: the seed kernels are bound by dependent random sectors and instruction issue, not by bytes (DESIGN.md 7)",
                     "all_kernels": {k: {"ms": v[0], "algorithmic_bytes": v[1], "GBps": (v[1] / (v[0] / 1000) / 1e9 if v[0] > 0 else 0.0)} for k, v in kernels.items()}},
    }
    # ---- reference CPU baseline on this box (rank 0, N == 1 only)
    if world == 1 and not a.no_cpu_baseline:
        r = run_reference(d, m1, m2, a.ref_sample, cores, 2)
        if r is not None:
            n, runs = r
            out["cpu_baseline"] = {"value": 2 * n / runs[-1][0], "unit": "reads/s", "cores": cores, "kind": "reference",
                                   "sample": f"first {n} pairs of the workload, oracle/_ref/bitmapperBS --pe -t {cores}, its own mapping timer, 2nd of 2 runs",
                                   "wall_s": runs[-1][1]}
        else:
            out["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": cores, "kind": "reference", "sample": "oracle/_ref/bitmapperBS unavailable"}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(out) + "\n").encode())
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
