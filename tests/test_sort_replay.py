"""bmbs_sort_replay.h (what finish_sorted runs on the device) leaves vote lists in exactly the order std::sort leaves the
reference's vote records in: 200 000 lists, checked element by element on the CPU."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_sort_replay_equals_std_sort(tmp_path):
    exe = tmp_path / "sort_replay_harness"
    subprocess.run(["g++", "-O2", "-std=c++17", str(ROOT / "tests/sort_replay_harness.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], check=True, capture_output=True, text=True)
    assert "mismatching 0 fallback 0" in r.stdout, r.stdout
