import gzip
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Native pieces: product (library, bmbs, bmbs-index) and oracle (cli, liboracle.so)."""
    from bitmapperbs_b200 import build as B
    sys.path.insert(0, str(ROOT / "oracle"))
    import build_oracle
    lib = ROOT / "bitmapperbs_b200/libbmbs_gpu.so"
    if not lib.exists() or not (ROOT / "bitmapperbs_b200/_build/bmbs").exists():
        B.build_all()
    else:
        B.build_tools()
    if not (ROOT / "oracle/_build/liboracle.so").exists() or not (ROOT / "oracle/_build/oracle_cli").exists():
        build_oracle.build()
    return {"lib": lib, "bmbs": ROOT / "bitmapperbs_b200/_build/bmbs", "indexer": ROOT / "bitmapperbs_b200/_build/bmbs-index",
            "oracle_cli": ROOT / "oracle/_build/oracle_cli", "oracle_lib": ROOT / "oracle/_build/liboracle.so",
            "ref": ROOT / "oracle/_ref/bitmapperBS"}


@pytest.fixture(scope="session")
def golden(built, tmp_path_factory):
    """Unpacked golden fixtures + an index built by bmbs-index next to genome.fa."""
    d = tmp_path_factory.mktemp("golden")
    for gz in GOLDEN.glob("*.gz"):
        with gzip.open(gz, "rb") as f, open(d / gz.name[:-3], "wb") as o:
            shutil.copyfileobj(f, o)
    for st in GOLDEN.glob("*.stats"):
        shutil.copy(st, d / ("ref_" + st.name))
    subprocess.run([str(built["indexer"]), str(d / "genome.fa")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return d


def sam_body(path):
    return [l for l in open(path, "rb") if not l.startswith(b"@")]


def read_fastq(path):
    out = []
    with open(path, "rb") as f:
        while True:
            name = f.readline()
            if not name:
                break
            seq = f.readline().rstrip(b"\n"); f.readline(); qual = f.readline().rstrip(b"\n")
            out.append((name.rstrip(b"\n")[1:], seq.upper(), qual))
    return out


_COMP = bytes.maketrans(b"ACGT", b"TGCA")


def revcomp(s: bytes) -> bytes:
    return s.translate(_COMP)[::-1]
