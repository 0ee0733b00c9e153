"""bmbs_band_walk.h (the run-to-run walk of the last band column that verify_windows runs on the device) gives the end
position and edit distance of the oracle's cell-by-cell restatement of BS_Reserve_Banded_BPM (Levenshtein_Cal.h:351-567):
300 000 windows, both band widths, checked on the CPU."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_band_walk_equals_oracle(tmp_path):
    exe = tmp_path / "band_walk_harness"
    subprocess.run(["g++", "-O2", "-std=c++17", str(ROOT / "tests/band_walk_harness.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], check=True, capture_output=True, text=True)
    assert "mismatching 0" in r.stdout and "tests 300000" in r.stdout, r.stdout
