#!/usr/bin/env python3
"""Random-sector load experiment: rate per load-instruction variant (BMBS_UBENCH_VARIANT); run under ncu for the DRAM bytes each moves."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bitmapperbs_b200 import capi
names = ["ld.global.nc.v4.u64", "+L2::64B", "+L2::128B", "ld.global.cv", "ld.global.cg", "nc.L1::no_allocate", "ld.global.lu", "ld.global.nc.u64 (8 B)", "nc.L1::no_allocate.L2::evict_first"]
for v in [int(x) for x in (sys.argv[1:] or range(9))]:
    os.environ["BMBS_UBENCH_VARIANT"] = str(v)
    r = capi.random_sector_peak(0, 8 << 30)
    print(f"variant {v} {names[v]:38s} {r / 1e9:7.2f} G loads/s", flush=True)
