"""CPU: the C-ABI library loads without a GPU and exports every symbol include/bmbs.h declares; the Python
host mirror refuses to run without it (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import pytest

import bitmapperbs_b200 as B
from bitmapperbs_b200 import capi

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include/bmbs.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bmbs_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    syms = declared_symbols()
    assert len(syms) >= 18
    lib = ctypes.CDLL(str(built["lib"]))
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/bmbs.h but not exported"
    assert sorted(capi.EXPORTS) == syms      # the Python mirror binds exactly the declared surface


def test_struct_layouts_match_header():
    assert capi.ReadResult.itemsize == 24 and capi.Cand.itemsize == 16
    assert capi.ReadResult.fields["first_cand"][1] == 8 and capi.ReadResult.fields["state"][1] == 18
    assert capi.Cand.fields["end_site"][1] == 12 and capi.Cand.fields["err"][1] == 14


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "lib_path", lambda: tmp_path / "libbmbs_gpu.so")
    with pytest.raises(B.BmbsError, match="no CPU fallback"):
        capi.load_library()


def test_errors_are_reported_not_fatal(built):
    lib = capi.load_library()
    h = ctypes.c_void_p()
    rc = lib.bmbs_index_load(b"/nonexistent/genome.fa.index", None, 0, ctypes.byref(h))
    assert rc == -1 and b"cannot open" in lib.bmbs_last_error()
