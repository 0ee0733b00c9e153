#!/usr/bin/env python3
"""Turns one gpurun evidence set (tools/gpu_round.sh TAG ...) into the files committed under profiles/:

  profiles/<tag>_launches.csv      ncu launch list (gpu__time_duration.sum, --clock-control none) of `bench.py --steps 2 --warmup 3`
  profiles/<tag>_bench.json        the bench line of the same build (not run under a profiler)
  profiles/<tag>_ncu_summary.md    per-kernel share of the step + the `ncu --set full` table of one launch of each kernel
  profiles/ncu_traffic.json        DRAM bytes per launch of each kernel (bench.py reads it for roofline.traffic)

usage: tools/profile_summary.py TAG [title...]      (reads gpurun_out/TAG_*; needs `ncu` for the .ncu-rep import)
"""
import collections, csv, json, shutil, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
G, P = ROOT / "gpurun_out", ROOT / "profiles"
tag = sys.argv[1]; title = " ".join(sys.argv[2:]) or tag
SETUP = ("build_ktab", "densify_sa", "random_sector_ubench", "int_pipe_ubench")

FULL = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def short(name):
    return name.split("(")[0].replace("void ", "").replace("bmbs::", "").replace("<unnamed>::", "")


out = [f"# {title}\n"]
# ---- launch list
lf = G / f"{tag}_launches.csv"
if lf.exists():
    rows = [r for r in csv.reader(l for l in open(lf) if l.startswith('"'))]
    h = rows[0]; ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    d = collections.defaultdict(list)
    for r in rows[1:]:
        v = float(r[vi].replace(",", "")); v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        d[short(r[ki])].append(v)
    step = {k: v for k, v in d.items() if not k.startswith(SETUP)}
    tot = sum(sum(v) for v in step.values())
    out += ["Launch list: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline`",
            f"(`profiles/{tag}_launches.csv`; per-launch times are serialised and cold-cache, so shares are what compares with the live stage times)\n",
            "| kernel | launches | avg us | share of the step |", "|---|---|---|---|"]
    for k, v in sorted(step.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| {k} | {len(v)} | {sum(v) / len(v):.1f} | {100 * sum(v) / tot:.1f}% |")
    setup = {k: v for k, v in d.items() if k.startswith(SETUP)}
    if setup:
        out.append("\nOnce per index load / ubench (outside the step): " + ", ".join(f"{k} {sum(v) / 1e3:.1f} ms" for k, v in setup.items()))
    shutil.copy(lf, P / f"{tag}_launches.csv")
# ---- bench line
bf = G / f"{tag}_bench.json"
if bf.exists() and bf.stat().st_size:
    j = json.loads([l for l in open(bf) if l.startswith("{")][-1])
    shutil.copy(bf, P / f"{tag}_bench.json")
    st = j["stage_ms_per_step"]
    out += ["", f"bench.py of the same build (CUDA events, not under a profiler): {j['ms_per_step']:.3f} ms/step, {j['value'] / 1e6:.0f} M reads/s device, "
            f"{j['e2e']['value'] / 1e6:.0f} M reads/s e2e; stage ms: " + ", ".join(f"{k} {v:.3f}" for k, v in st.items() if k != "total")]
    if j.get("cpu_baseline"):
        out.append(f"reference CPU arm on the same box: {j['cpu_baseline']['value'] / 1e6:.2f} M reads/s with {j['cpu_baseline']['cores']} threads.")
# ---- full set
rep = G / f"{tag}_full.ncu-rep"
if rep.exists():
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units, data = rows[0], rows[1], rows[2:]
    names = [short(r[h.index("Kernel Name")]) for r in data]
    out += ["", f"## `ncu --set full --clock-control none --import-source on`, one launch of each kernel (step 4 of the same command)\n",
            "| metric | " + " | ".join(names) + " |", "|---|" + "---|" * len(names)]
    for m in FULL:
        if m in h:
            i = h.index(m)
            out.append(f"| {m} [{units[i]}] | " + " | ".join(r[i] for r in data) + " |")
    traffic = {}
    ir, iw, it = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for n, r in zip(names, data):
        b = float(r[ir]) * mult[units[ir]] + float(r[iw]) * mult[units[iw]]
        traffic[n] = {"dram_bytes_per_launch": b, "duration_us": float(r[it]) * {"us": 1, "ms": 1e3, "ns": 1e-3}.get(units[it], 1), "capture": f"profiles/{tag}_ncu_summary.md"}
    # keyed by workload (bench.py looks up ncu_traffic[workload][kernel]); other workloads' entries are kept
    wl_name = "cfg3"
    if bf.exists() and bf.stat().st_size:
        wl_name = json.loads([l for l in open(bf) if l.startswith("{")][-1])["config"]["workload"].split(":")[0]
    allt = {}
    try:
        allt = json.loads((P / "ncu_traffic.json").read_text())
        if not all(isinstance(v, dict) and all(isinstance(x, dict) for x in v.values()) for v in allt.values()):
            allt = {}
    except Exception:
        allt = {}
    allt[wl_name] = traffic
    (P / "ncu_traffic.json").write_text(json.dumps(allt, indent=1) + "\n")
# ---- DRAM traffic at the bench's own scale (tools/gpu_round.sh traffic: one pass, metrics only)
tf = G / f"{tag}_traffic.csv"
if tf.exists():
    rows = [r for r in csv.reader(l for l in open(tf) if l.startswith('"'))]
    h = rows[0]; ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}
    per = collections.OrderedDict()
    for r in rows[1:]:
        k = short(r[ki]); v = float(r[vi].replace(",", "")) * mult.get(r[ui], 1)
        e = per.setdefault(k, {"dram_bytes_per_launch": 0.0, "duration_us": 0.0, "capture": f"profiles/{tag}_traffic.csv"})
        if r[mi].startswith("dram__bytes"):
            e["dram_bytes_per_launch"] += v
        elif r[mi].startswith("gpu__time_duration"):
            e["duration_us"] += v
    wl_name = "cfg3"
    if bf.exists() and bf.stat().st_size:
        wl_name = json.loads([l for l in open(bf) if l.startswith("{")][-1])["config"]["workload"].split(":")[0]
    try:
        allt = json.loads((P / "ncu_traffic.json").read_text())
        if not all(isinstance(v, dict) and all(isinstance(x, dict) for x in v.values()) for v in allt.values()):
            allt = {}
    except Exception:
        allt = {}
    allt[wl_name] = per
    (P / "ncu_traffic.json").write_text(json.dumps(allt, indent=1) + "\n")
    shutil.copy(tf, P / f"{tag}_traffic.csv")
    out += ["", f"## DRAM traffic per launch at the bench's scale (`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`, one pass; `profiles/{tag}_traffic.csv`)\n",
            "| kernel | DRAM bytes | us | GB/s |", "|---|---|---|---|"]
    for k, e in per.items():
        out.append(f"| {k} | {e['dram_bytes_per_launch'] / 1e6:.1f} MB | {e['duration_us']:.1f} | {e['dram_bytes_per_launch'] / max(e['duration_us'], 1e-9) / 1e3:.0f} |")
notes = G / f"{tag}_notes.md"
if notes.exists():
    out += ["", notes.read_text()]
(P / f"{tag}_ncu_summary.md").write_text("\n".join(out) + "\n")
print("\n".join(out))
