#!/usr/bin/env python3
"""Per-source-line share of a kernel's executed instructions, active lanes and stall samples from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X.csv`.   usage: tools/ncu_lines.py X.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]))); top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
cur, hdr, out = None, None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] in ("Function Name", "") or not hdr: continue
    try:
        ie, te, sm = int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("Thread Instructions Executed")]), int(r[hdr.index("# Samples")])
    except ValueError:
        continue
    out.append((ie, te, sm, cur, r[0], r[1].strip()[:120]))
tot, tt, ts = sum(o[0] for o in out), sum(o[1] for o in out), sum(o[2] for o in out)
print(f"warp instructions {tot}, thread instructions {tt}, lanes per instruction {tt / tot:.1f}, samples {ts}")
for ie, te, sm, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{100 * ie / tot:5.1f}% inst {te / max(ie, 1):5.1f} lanes {100 * sm / ts:5.1f}% smp  {f}:{ln}  {src}")
