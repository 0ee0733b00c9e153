#!/usr/bin/env python3
"""Whole program at the headline size: bench.py's cfg3 genome (3.087 Gbp unless --scale), N x 150 bp single-end reads from one
FASTQ file to one SAM file through bitmapperbs_b200/_build/bmbs and through the reference (oracle/_ref/bitmapperBS), each
program's own `Total:` timers, SAM records compared after sorting, the mapper's per-stage busy times (BMBS_TIMING).

  python tools/cli_scale.py [--reads 10000000] [--scale 1.0] [--out gpurun_out/cli_scale.json]
"""
import argparse, hashlib, json, os, re, subprocess, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench as BN

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=10_000_000); ap.add_argument("--scale", type=float, default=None)
ap.add_argument("--workload", default="cfg3"); ap.add_argument("--out", default="gpurun_out/cli_scale.json"); ap.add_argument("--no-reference", action="store_true")
ap.add_argument("--threads", type=int, default=os.cpu_count() or 16)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--variants", default="", help="extra runs of our mapper to /dev/null and to a file under environment settings: 'name:K=V,K=V;name2:K=V'")
a = ap.parse_args()
scale = a.scale if a.scale is not None else float(os.environ.get("BMBS_BENCH_SCALE", 1.0))
wl = BN.Workload(a.workload, scale)
cache = Path(os.environ.get("BMBS_BENCH_CACHE", "/tmp/bmbs_bench"))
d = BN.ensure_dataset(cache, wl, True)
t = time.time(); m1, m2 = BN.make_reads(d, wl, a.reads, wl.read_seed + 77); n, seq_args = BN.write_fastq_files(d, wl, m1, m2, a.reads, "scale")
BN.log(f"[cli_scale] {n} reads simulated and written in {time.time() - t:.1f}s")
EXE, REF = ROOT / "bitmapperbs_b200/_build/bmbs", ROOT / "oracle/_ref/bitmapperBS"


def run(exe, out, env=None):
    """best of --reps runs by the program's own mapping timer (the boxes are shared: single runs scatter by a factor of two)"""
    rs = [run_once(exe, out, env) for _ in range(a.reps)]
    best = min(rs, key=lambda r: r["map_s"])
    best["map_s_all"] = [r["map_s"] for r in rs]; best["load_s_all"] = [r["load_s"] for r in rs]
    return best


def run_once(exe, out, env=None):
    t = time.time()
    r = subprocess.run([str(exe), "--search", "g.fa", *seq_args, *wl.cli_flags(), "-t", str(a.threads), "-o", out], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True,
                       env={**os.environ, **(env or {})})
    wall = time.time() - t
    m = re.search(r"Total:\s+([0-9.]+)\s+([0-9.]+)", r.stderr)
    if r.returncode or not m:
        raise SystemExit(f"{exe} failed: {r.stderr[-1500:]}")
    return {"load_s": float(m.group(1)), "map_s": float(m.group(2)), "wall_s": wall, "reads_per_s_map": (2 if wl.pe else 1) * n / float(m.group(2)),
            "timing": [l for l in r.stderr.splitlines() if l.startswith("[bmbs timing]")]}


def digest(p):
    h = hashlib.sha256(); k = 0
    for line in sorted(x for x in open(p, "rb") if not x.startswith(b"@")):
        h.update(line); k += 1
    return k, h.hexdigest()


out = {"workload": wl.describe(n), "reads": n, "threads": a.threads, "runs": []}
run_once(EXE, "/dev/null")                                                 # page cache, driver
out["ours_devnull"] = run(EXE, "/dev/null", {"BMBS_TIMING": "1"})
out["ours"] = run(EXE, "scale_gpu.sam", {"BMBS_TIMING": "1"})
out["ours_host_finish"] = run(EXE, "/dev/null", {"BMBS_TIMING": "1", "BMBS_HOST_FINISH": "1"})
for spec in filter(None, a.variants.split(";")):
    name, _, kv = spec.partition(":")
    env = {"BMBS_TIMING": "1", **dict(x.split("=", 1) for x in kv.split(",") if x)}
    out["runs"].append({"name": name, "env": env, "devnull": run(EXE, "/dev/null", env), "file": run(EXE, "scale_var.sam", env)})
    (d / "scale_var.sam").unlink()
if REF.exists() and not a.no_reference:
    out["reference"] = run_once(REF, "scale_ref.sam")
    g, r = digest(d / "scale_gpu.sam"), digest(d / "scale_ref.sam")
    out["sam_records"] = g[0]; out["sam_identical"] = g == r
    out["map_speedup"] = out["reference"]["map_s"] / out["ours"]["map_s"]; out["wall_speedup"] = out["reference"]["wall_s"] / out["ours"]["wall_s"]
for f in ("scale_gpu.sam", "scale_ref.sam"):
    try:
        (d / f).unlink()
    except OSError:
        pass
Path(a.out).parent.mkdir(exist_ok=True)
Path(a.out).write_text(json.dumps(out, indent=1))
print(json.dumps(out, indent=1))
