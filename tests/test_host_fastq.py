"""Host FASTQ block reader (bitmapperbs_b200/csrc/host/fastq.hpp): the memory-mapped path for plain files and the gzread
path must hand out the same whole records for every block size, with or without a final newline, with CRLF line ends and
with a truncated last record (Process_Reads.cpp:62-90 reads line by line; the reference needs the final newline, we do not)."""
import gzip
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = tmp_path_factory.mktemp("fq") / "fq_harness"
    subprocess.run(["g++", "-O2", "-std=c++17", str(ROOT / "tests/host_fastq_harness.cpp"), "-o", str(exe), "-lz"], check=True)
    return exe


def records(n, crlf=False):
    nl = "\r\n" if crlf else "\n"
    recs = [(f"@r{i} extra/1", "ACGTN"[i % 5] * (1 + (i * 7) % 150), "+", "I" * (1 + (i * 7) % 150)) for i in range(n)]
    return recs, "".join(nl.join(r) + nl for r in recs)


def run(harness, path, per):
    out = subprocess.run([str(harness), str(path), str(per)], capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[-1].startswith("TOTAL")
    return int(out[-1].split()[1]), [tuple(l.split("\t")) for l in out[:-1]]


@pytest.mark.parametrize("n", [0, 1, 5, 1000])
@pytest.mark.parametrize("per", [1, 3, 64, 100000])
@pytest.mark.parametrize("variant", ["plain", "no_final_newline", "crlf", "gz", "truncated"])
def test_block_reader(harness, tmp_path, n, per, variant):
    recs, text = records(n, crlf=variant == "crlf")
    expect = [(r[0], r[1], r[3]) for r in recs]
    if variant == "no_final_newline":
        text = text[:-1]
    if variant == "truncated":
        text += "@partial\nACGT\n"
    p = tmp_path / ("r.fq.gz" if variant == "gz" else "r.fq")
    if variant == "gz":
        with gzip.open(p, "wt") as f:
            f.write(text)
    else:
        p.write_text(text)
    if n == 0 and variant in ("plain", "crlf", "no_final_newline"):
        p.write_text("" if variant != "truncated" else text)
    total, got = run(harness, p, per)
    assert total == n
    assert got == [tuple(x) for x in expect]


@pytest.mark.parametrize("n", [1, 5, 1000])
@pytest.mark.parametrize("target", [1, 7, 50, 333, 4096, 10 ** 9])
@pytest.mark.parametrize("variant", ["plain", "no_final_newline", "crlf", "truncated", "at_quality"])
def test_blocks_cut_by_size_hold_whole_records(harness, tmp_path, n, target, variant):
    """the single-end splitter cuts a mapped file every `target` bytes at the next record boundary without scanning the text:
    every record comes out once, in order, whatever the block size -- also when quality lines begin with '@' or '+'"""
    recs, text = records(n, crlf=variant == "crlf")
    if variant == "at_quality":      # qualities that look like header / separator lines
        recs = [(r[0], r[1], r[2], ("@+"[i % 2] + r[3][1:])) for i, r in enumerate(recs)]
        text = "".join("\n".join(r) + "\n" for r in recs)
    expect = [(r[0], r[1], r[3]) for r in recs]
    if variant == "no_final_newline":
        text = text[:-1]
    if variant == "truncated":
        text += "@partial\nACGT\n"
    p = tmp_path / "r.fq"
    p.write_text(text)
    out = subprocess.run([str(harness), str(p), "1", str(target)], capture_output=True, text=True, check=True).stdout.splitlines()
    assert int(out[-1].split()[1]) == n
    assert [tuple(l.split("\t")) for l in out[:-1]] == [tuple(x) for x in expect]
