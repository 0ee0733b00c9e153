#!/usr/bin/env python3
"""Test-infrastructure only (oracle/): make a scratch copy of the reference's
top-level sources compile-safe under g++ -O3.

The reference has non-void functions that fall off their end (UB; g++ >= 8
turns that into fall-through at -O1+ and the binary crashes after index load,
SURVEY.md §5 / §8c-3).  This script works on a SCRATCH COPY (never inside the
repo, never in /root/reference): it asks g++ for every -Wreturn-type site and
inserts `return 0;` in front of the closing brace the warning points at.  No
semantic change: none of the callers consume these return values.
"""
import re
import subprocess
import sys
from pathlib import Path

FLAGS = ["-mavx2", "-mpopcnt", "-D__AVX2__", "-Wreturn-type"]


def sites(src: Path, inc: str, mode):
    cmd = ["g++", *mode, *FLAGS, "-I", inc, str(src)]
    err = subprocess.run(cmd, capture_output=True, text=True, cwd=src.parent).stderr
    out = set()
    for m in re.finditer(r"^([^:\s]+):(\d+):(\d+): warning: (no return statement|control reaches end)", err, re.M):
        if Path(m.group(1)).name == src.name:
            out.add((int(m.group(2)), int(m.group(3))))
    return sorted(out)


def patch(src: Path, inc: str) -> int:
    total = 0
    for mode in (["-fsyntax-only"], ["-O1", "-c", "-o", "/dev/null"]):
        for _ in range(4):  # iterate: fixing one site never creates another, but be safe
            found = sites(src, inc, mode)
            if not found:
                break
            lines = src.read_text(errors="surrogateescape").split("\n")
            for line, col in sorted(found, reverse=True):
                s = lines[line - 1]
                if col <= len(s) and s[col - 1] == "}":
                    lines[line - 1] = s[: col - 1] + "return 0; " + s[col - 1:]
                else:
                    # "control reaches end": the warning points into the body;
                    # the function ends at the next column-0 closing brace.
                    j = line
                    while not lines[j].startswith("}"):
                        j += 1
                    lines[j] = "return 0; " + lines[j]
                total += 1
            src.write_text("\n".join(lines), errors="surrogateescape")
    return total


if __name__ == "__main__":
    d, inc = Path(sys.argv[1]), sys.argv[2]
    n = 0
    for f in sorted(d.glob("*.cpp")):
        k = patch(f, inc)
        if k:
            print(f"{f.name}: {k} return(s) inserted")
        n += k
    print("total", n)
