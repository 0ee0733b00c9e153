"""Host finishing code of the product (bitmapperbs_b200/csrc/host/mapper.hpp: vote-ordered reduction, pair pick, CIGAR
refinement, MAPQ, SAM text), on the CPU: fed with the ORACLE's per-read records -- the same arrays the GPU library returns
(tests/test_gpu_parity.py) -- its SAM must equal the golden SAM of the real reference, record for record, and the
--mapstats counters with it.  Covers single end, paired end fast and sensitive, and --unmapped_out against the live reference."""
import subprocess
from pathlib import Path

import pytest

from conftest import sam_body

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def harness(built, tmp_path_factory):
    exe = tmp_path_factory.mktemp("hf") / "host_finish_harness"
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-Wno-sign-compare", str(ROOT / "tests/host_finish_harness.cpp"), "-o", str(exe), "-lz"], check=True)
    return exe


CASES = [("se100", "se", ["se100.fq"]), ("se250", "se", ["se250.fq"]), ("pe150", "pe", ["pe150_1.fq", "pe150_2.fq"]),
         ("pe150s", "pes", ["pe150_1.fq", "pe150_2.fq"]), ("pe100h", "pe", ["pe100h_1.fq", "pe100h_2.fq"]), ("pe100hs", "pes", ["pe100h_1.fq", "pe100h_2.fq"])]


@pytest.mark.parametrize("name,mode,files", CASES)
def test_host_finish_matches_reference_golden(golden, harness, name, mode, files):
    out = golden / f"hf_{name}.sam"
    r = subprocess.run([str(harness), mode, "genome.fa", out.name, *files], cwd=golden, check=True, capture_output=True, text=True)
    assert sam_body(out) == sam_body(golden / f"{name}.sam")
    n, uq, am, rate = r.stdout.split()
    st = (golden / f"ref_{name}.stats").read_text().splitlines()
    assert int(st[0].split()[-1]) == int(n) and int(st[1].split(":")[1].split()[0]) == int(uq) and int(st[2].split(":")[1].split()[0]) == int(am)
    assert st[4].split()[-1] == f"{float(rate):.2f}%"


@pytest.mark.parametrize("mode,files", [("se", ["se100.fq"]), ("pe", ["pe100h_1.fq", "pe100h_2.fq"]), ("pes", ["pe100h_1.fq", "pe100h_2.fq"])])
def test_unmapped_out_matches_live_reference(golden, built, harness, mode, files):
    if not built["ref"].exists():
        pytest.skip("compiled reference absent")
    args = ["--seq", files[0]] if mode == "se" else ["--seq1", files[0], "--seq2", files[1], "--pe"] + (["--sensitive"] if mode == "pes" else [])
    subprocess.run([str(built["ref"]), "--search", "genome.fa", *args, "--unmapped_out", "-t", "1", "-o", f"ref_un_{mode}.sam"],
                   cwd=golden, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([str(harness), mode, "genome.fa", f"hf_un_{mode}.sam", *files, "--unmapped_out"], cwd=golden, check=True, capture_output=True)
    assert sam_body(golden / f"hf_un_{mode}.sam") == sam_body(golden / f"ref_un_{mode}.sam")
