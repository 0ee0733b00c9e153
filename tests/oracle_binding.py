"""ctypes access to oracle/_build/liboracle.so (the CPU restatement) -- tests only."""
import ctypes as C
from pathlib import Path

import numpy as np

from bitmapperbs_b200.capi import Cand, Final, ReadResult, flatten

ROOT = Path(__file__).resolve().parent.parent
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(str(ROOT / "oracle/_build/liboracle.so"))
        vp = C.c_void_p
        L.orc_load.restype = vp; L.orc_load.argtypes = [C.c_char_p]
        L.orc_free.argtypes = [vp]
        L.orc_genome_length.restype = C.c_uint64; L.orc_genome_length.argtypes = [vp]
        L.orc_bpm.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_uint, C.POINTER(C.c_uint32)]
        L.orc_window.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_char_p]
        L.orc_lf.restype = C.c_uint64; L.orc_lf.argtypes = [vp, C.c_uint64, C.c_int]
        L.orc_locate.argtypes = [vp, C.c_uint64, C.POINTER(C.c_uint64)]
        L.orc_seed.restype = C.c_uint64; L.orc_seed.argtypes = [vp, C.c_char_p, C.c_uint64] + [C.POINTER(C.c_uint64)] * 3
        L.orc_count.restype = C.c_uint64; L.orc_count.argtypes = [vp, C.c_char_p, C.c_uint64] + [C.POINTER(C.c_uint64)] * 2
        L.orc_map_se.argtypes = [vp, vp, vp, C.c_int, C.c_double, C.c_int, vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.orc_map_pe.argtypes = [vp, vp, vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.orc_map_pe_sensitive.argtypes = [vp, vp, vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, vp, vp, C.c_size_t, C.POINTER(C.c_size_t), vp]
        L.orc_verify.argtypes = [vp, vp, vp, C.c_int, vp, vp, C.c_size_t, C.c_double, vp, vp, C.c_int]
        L.orc_refine_final.argtypes = [vp, C.c_uint64, C.c_char_p, C.c_char_p] + [C.c_int] * 8 + [C.POINTER(C.c_int)] * 3 + [C.POINTER(C.c_uint), vp, C.c_int]
        L.orc_std_sort_order.argtypes = [vp, C.c_uint32, vp]
        L.orc_finish_se.argtypes = [vp, vp, vp, C.c_int, C.c_double, C.c_int, vp, vp, vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.orc_finish_pe.argtypes = [vp, vp, vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.orc_banded_align.argtypes = [vp, C.c_uint64, C.c_char_p, C.c_char_p] + [C.c_int] * 8 + [C.POINTER(C.c_int)] * 3 + [vp, C.c_int]
        _lib = L
    return _lib


def std_sort_order(votes):
    """positions in the order std::sort (by vote, descending) leaves the reference's vote records in"""
    votes = np.ascontiguousarray(votes, dtype=np.uint32); order = np.zeros(len(votes), dtype=np.uint32)
    lib().orc_std_sort_order(votes.ctypes.data, len(votes), order.ctypes.data)
    return order


class OracleIndex:
    def __init__(self, prefix):
        self.h = lib().orc_load(str(prefix).encode())
        assert self.h, f"oracle cannot load {prefix}"
        self.N = lib().orc_genome_length(self.h)

    def window(self, site, n):
        buf = C.create_string_buffer(n + 8)
        lib().orc_window(self.h, site, n, buf)
        return buf.raw[:n]

    def map_se(self, reads, e_rate=0.08, seed_len=30, cap=None):
        flat, offs = reads if isinstance(reads, tuple) else flatten(reads)
        n = len(offs) - 1
        cap = cap or max(1 << 16, 64 * n)
        res = np.zeros(n, dtype=ReadResult); cand = np.zeros(cap, dtype=Cand); used = C.c_size_t(0)
        rc = lib().orc_map_se(self.h, flat.ctypes.data, offs.ctypes.data, n, e_rate, seed_len, res.ctypes.data, cand.ctypes.data, cap, C.byref(used))
        assert rc == 0
        return res, cand[: used.value]

    def finish_se(self, reads, res, cand, e_rate=0.08, ambiguous_out=False):
        """the reference's reduction + ungapped CIGAR check + coordinates over (res, cand) -> (Final[n], mismatch positions)"""
        flat, offs = reads if isinstance(reads, tuple) else flatten(reads)
        n = len(offs) - 1
        fin = np.zeros(n, dtype=Final); mism = np.zeros(32 * n + 64, dtype=np.uint16); used = C.c_size_t(0)
        res = np.ascontiguousarray(res); cand = np.ascontiguousarray(cand)
        rc = lib().orc_finish_se(self.h, flat.ctypes.data, offs.ctypes.data, n, e_rate, 1 if ambiguous_out else 0, res.ctypes.data, cand.ctypes.data,
                                 fin.ctypes.data, mism.ctypes.data, len(mism), C.byref(used))
        assert rc == 0
        return fin, mism[: used.value]

    def finish_pe(self, mates, res, cand, e_rate=0.08, min_ins=0, max_ins=500, sensitive=False, ambiguous_out=False):
        """the reference's hit compaction + single-side filter + pair pick + ungapped CIGAR check + coordinates over the records
        and verified hit lists of a paired batch -> (Final[2 * pairs], mismatch positions)"""
        flat, offs = mates if isinstance(mates, tuple) else flatten(mates)
        n = len(offs) - 1
        fin = np.zeros(n, dtype=Final); mism = np.zeros(32 * n + 64, dtype=np.uint16); used = C.c_size_t(0)
        res = np.ascontiguousarray(res); cand = np.ascontiguousarray(cand)
        rc = lib().orc_finish_pe(self.h, flat.ctypes.data, offs.ctypes.data, n // 2, e_rate, min_ins, max_ins, 1 if sensitive else 0, 1 if ambiguous_out else 0,
                                 res.ctypes.data, cand.ctypes.data, fin.ctypes.data, mism.ctypes.data, len(mism), C.byref(used))
        assert rc == 0
        return fin, mism[: used.value]

    def map_pe(self, mates, e_rate=0.08, seed_len=30, min_ins=0, max_ins=500, cap=None):
        flat, offs = mates if isinstance(mates, tuple) else flatten(mates)
        n = len(offs) - 1
        cap = cap or max(1 << 16, 64 * n)
        res = np.zeros(n, dtype=ReadResult); cand = np.zeros(cap, dtype=Cand); used = C.c_size_t(0)
        rc = lib().orc_map_pe(self.h, flat.ctypes.data, offs.ctypes.data, n // 2, e_rate, seed_len, min_ins, max_ins, res.ctypes.data, cand.ctypes.data, cap, C.byref(used))
        assert rc == 0
        return res, cand[: used.value]

    def map_pe_sensitive(self, mates, e_rate=0.08, seed_len=30, min_ins=0, max_ins=500, cap=None):
        """-> (records, final hit lists, per-read flag: went through the re-seeding round)"""
        flat, offs = mates if isinstance(mates, tuple) else flatten(mates)
        n = len(offs) - 1
        cap = cap or max(1 << 16, 64 * n)
        res = np.zeros(n, dtype=ReadResult); cand = np.zeros(cap, dtype=Cand); used = C.c_size_t(0); reseeded = np.zeros(n, dtype=np.uint8)
        rc = lib().orc_map_pe_sensitive(self.h, flat.ctypes.data, offs.ctypes.data, n // 2, e_rate, seed_len, min_ins, max_ins, res.ctypes.data,
                                        cand.ctypes.data, cap, C.byref(used), reseeded.ctypes.data)
        assert rc == 0
        return res, cand[: used.value], reseeded

    def verify(self, reads, read_idx, sites, e_rate=0.08, threads=8):
        flat, offs = reads if isinstance(reads, tuple) else flatten(reads)
        read_idx = np.ascontiguousarray(read_idx, dtype=np.uint32); sites = np.ascontiguousarray(sites, dtype=np.uint64)
        n = len(sites)
        end = np.zeros(n, dtype=np.int32); err = np.zeros(n, dtype=np.uint32)
        lib().orc_verify(self.h, flat.ctypes.data, offs.ctypes.data, len(offs) - 1, read_idx.ctypes.data, sites.ctypes.data, n, e_rate, end.ctypes.data, err.ctypes.data, threads)
        return end, err


def refine_final(oidx, site, read: bytes, qual: bytes, k, scoring=(6, 2, 1, 5, 3, 33)):
    """the reference's whole refinement of an alignment with indels on the CPU (DP, end fix-ups, NM recount):
    (score, qb, qe, nm, final ops in read order) -- what bmbs_refine must return value for value"""
    score, qb, qe, nm = C.c_int(), C.c_int(), C.c_int(), C.c_uint()
    ops = np.zeros(2 * len(read) + 2 * k + 2, dtype=np.uint32)
    n = lib().orc_refine_final(oidx.h, int(site), read, qual, len(read), int(k), *scoring, C.byref(score), C.byref(qb), C.byref(qe), C.byref(nm), ops.ctypes.data, len(ops))
    assert n >= 0
    return score.value, qb.value, qe.value, nm.value, ops[:n].copy()


def banded_align(oidx, site, read: bytes, qual: bytes, k, scoring=(6, 2, 1, 5, 3, 33)):
    """CPU banded affine-gap DP with traceback (host/postprocess.hpp banded_affine_align) on the window at `site`:
    (score, qb, qe, ops) -- what bmbs_refine must return value for value"""
    score, qb, qe = C.c_int(), C.c_int(), C.c_int()
    ops = np.zeros(2 * len(read) + 2 * k + 2, dtype=np.uint32)
    n = lib().orc_banded_align(oidx.h, int(site), read, qual, len(read), int(k), *scoring, C.byref(score), C.byref(qb), C.byref(qe), ops.ctypes.data, len(ops))
    assert n >= 0
    return score.value, qb.value, qe.value, ops[:n].copy()
