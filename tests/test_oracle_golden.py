"""CPU: the index writer and the oracle restatement against golden output of the REAL reference
(tests/golden/, produced by tests/golden/make_golden.py from oracle/_ref/bitmapperBS)."""
import hashlib
import json
import subprocess

import pytest

from conftest import GOLDEN, sam_body


def test_index_files_identical_to_reference(golden):
    want = json.loads((GOLDEN / "index_sha256.json").read_text())
    for suf, h in want.items():
        data = (golden / ("genome.fa" + suf)).read_bytes()
        if suf.endswith(".sa"):
            data = data[:-8]   # last flag word is uninitialised memory in reference-built indexes
        assert hashlib.sha256(data).hexdigest() == h, f"index file {suf} differs from the reference's"


@pytest.mark.parametrize("name,args", [
    ("se100", ["se", "genome.fa", "se100.fq", "out.sam"]),
    ("se250", ["se", "genome.fa", "se250.fq", "out.sam"]),
    ("pe150", ["pe", "genome.fa", "pe150_1.fq", "pe150_2.fq", "out.sam"]),
    ("pe150s", ["pe", "genome.fa", "pe150_1.fq", "pe150_2.fq", "out.sam", "1"]),     # --pe --sensitive
    ("pe100h", ["pe", "genome.fa", "pe100h_1.fq", "pe100h_2.fq", "out.sam"]),
    ("pe100hs", ["pe", "genome.fa", "pe100h_1.fq", "pe100h_2.fq", "out.sam", "1"]),  # --pe --sensitive, mate filter + re-seeding matter
])
def test_oracle_sam_identical_to_reference(golden, built, name, args):
    r = subprocess.run([str(built["oracle_cli"]), *args], cwd=golden, stderr=subprocess.PIPE, check=True)
    got = sam_body(golden / "out.sam")
    want = sam_body(golden / f"{name}.sam")
    assert len(got) == len(want)
    assert got == want            # byte-identical records, reference -t 1 order
    # the five mapstats numbers
    ref = dict(l.split(":", 1) for l in (golden / f"ref_{name}.stats").read_text().splitlines())
    err = r.stderr.decode()
    fields = dict(l.split(": ", 1) for l in err.strip().splitlines())
    assert int(fields["No. of Reads"]) == int(ref["No. of Reads"].split()[0])
    assert int(fields["Unique"]) == int(ref["No. of Unique Mapped Reads"].split()[0])
    assert int(fields["Ambiguous"]) == int(ref["No. of Ambiguous Mapped Reads"].split()[0])
    assert int(fields["Unmapped"]) == int(ref["No. of Unmapped Reads"].split()[0])
    e, b = (int(x) for x in fields["ErrBases/Bases"].split("/"))
    assert f"{e / b * 100:0.2f}%" == ref["Mismatch and Indel Rate"].split()[0]
