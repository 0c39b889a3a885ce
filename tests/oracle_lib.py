"""ctypes doorway to the CPU checker (oracle/liboracle.so) and, when present, to the unmodified
reference build (oracle/_ref/libref*.so).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"


class OracleParams(C.Structure):
    _fields_ = [("match", C.c_int), ("mismatch", C.c_int), ("gap_init", C.c_int), ("gap_ext", C.c_int)]


DEFAULT = (1, -1, 1, 1)  # match, mismatch, gap_init, gap_ext  (main.cpp:20-23)


def _params(p):
    m, x, gi, ge = p
    return OracleParams(m, x, gi, ge)


def _u8(a):
    if isinstance(a, (bytes, bytearray)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    elif isinstance(a, str):
        a = np.frombuffer(a.encode("latin1"), dtype=np.uint8)
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def _ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_ubyte))


_oracle = None


def build_oracle():
    subprocess.run(["make", "-s", "-C", str(ORACLE_DIR)], check=True)
    if Path("/root/reference").is_dir():
        subprocess.run(["make", "-s", "-C", str(ORACLE_DIR), "ref"], check=True)


def oracle():
    global _oracle
    if _oracle is None:
        so = ORACLE_DIR / "liboracle.so"
        if not so.exists():
            build_oracle()
        lib = C.CDLL(str(so))
        sig = [C.POINTER(C.c_ubyte), C.POINTER(C.c_ubyte), C.c_int, C.c_int, C.POINTER(OracleParams)]
        for name in ("oracle_gotoh_full", "oracle_gotoh_rolling", "oracle_lazy_smith", "oracle_linear_gap"):
            getattr(lib, name).argtypes = sig
            getattr(lib, name).restype = C.c_int
        lib.oracle_gotoh_last_row.argtypes = sig + [C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.oracle_gotoh_last_row.restype = C.c_int
        lib.oracle_gotoh_end.argtypes = sig + [C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.oracle_gotoh_end.restype = C.c_int
        lib.oracle_gotoh_anchored_end.argtypes = sig + [C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.oracle_gotoh_anchored_end.restype = C.c_int
        lib.oracle_gotoh_span.argtypes = sig + [C.POINTER(C.c_int)] * 4
        lib.oracle_gotoh_span.restype = C.c_int
        lib.oracle_gotoh_mt.argtypes = sig + [C.c_int]
        lib.oracle_gotoh_mt.restype = C.c_int
        lib.oracle_gotoh_fast.argtypes = sig + [C.c_int]
        lib.oracle_gotoh_fast.restype = C.c_int
        lib.oracle_gotoh_banded.argtypes = sig[:4] + [C.c_int, C.c_int, C.POINTER(OracleParams), C.POINTER(C.c_int64)]
        lib.oracle_gotoh_banded.restype = C.c_int
        lib.oracle_gotoh_batch.argtypes = [C.POINTER(C.c_ubyte), C.POINTER(C.c_int64), C.POINTER(C.c_int),
                                           C.POINTER(C.c_ubyte), C.POINTER(C.c_int64), C.POINTER(C.c_int),
                                           C.c_int64, C.POINTER(OracleParams), C.c_int, C.POINTER(C.c_int)]
        lib.oracle_gotoh_batch.restype = None
        lib.oracle_gotoh_banded_batch.argtypes = [C.POINTER(C.c_ubyte), C.POINTER(C.c_int64), C.POINTER(C.c_int),
                                                  C.POINTER(C.c_ubyte), C.POINTER(C.c_int64), C.POINTER(C.c_int),
                                                  C.c_int64, C.c_int, C.c_int, C.POINTER(OracleParams), C.c_int,
                                                  C.POINTER(C.c_int)]
        lib.oracle_gotoh_banded_batch.restype = None
        lib.oracle_mix64.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        lib.oracle_mix64.restype = C.c_uint64
        lib.oracle_random_acgt.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.POINTER(C.c_ubyte)]
        lib.oracle_random_acgt.restype = None
        lib.oracle_max_threads.restype = C.c_int
        _oracle = lib
    return _oracle


def _call(name, s1, s2, p=DEFAULT, *extra):
    a, b = _u8(s1), _u8(s2)
    pp = _params(p)
    return getattr(oracle(), name)(_ptr(a), _ptr(b), len(a), len(b), C.byref(pp), *extra)


def gotoh_full(s1, s2, p=DEFAULT):
    return _call("oracle_gotoh_full", s1, s2, p)


def gotoh_rolling(s1, s2, p=DEFAULT):
    return _call("oracle_gotoh_rolling", s1, s2, p)


def lazy_smith(s1, s2, p=DEFAULT):
    return _call("oracle_lazy_smith", s1, s2, p)


def linear_gap(s1, s2, p=DEFAULT):
    return _call("oracle_linear_gap", s1, s2, p)


def gotoh_last_row(s1, s2, p=DEFAULT):
    """(best, H[m][0..n], F[m][0..n]): columns index seq1, the row is the last row of seq2."""
    a, b = _u8(s1), _u8(s2)
    pp = _params(p)
    H = np.zeros(len(a) + 1, dtype=np.int32)
    F = np.zeros(len(a) + 1, dtype=np.int32)
    best = oracle().oracle_gotoh_last_row(_ptr(a), _ptr(b), len(a), len(b), C.byref(pp), H.ctypes.data_as(C.POINTER(C.c_int)),
                                          F.ctypes.data_as(C.POINTER(C.c_int)))
    return best, H, F


def gotoh_end(s1, s2, p=DEFAULT):
    """(score, i_end, j_end): 1-based end cell of the best local alignment, i in seq2, j in seq1; smallest j, then i."""
    a, b = _u8(s1), _u8(s2)
    pp = _params(p)
    ie, je = C.c_int(0), C.c_int(0)
    best = oracle().oracle_gotoh_end(_ptr(a), _ptr(b), len(a), len(b), C.byref(pp), C.byref(ie), C.byref(je))
    return best, ie.value, je.value


def gotoh_anchored_end(s1, s2, p=DEFAULT):
    """Anchored recurrence (alignments start at the origin): (max score, i, j) by the same tie rule."""
    a, b = _u8(s1), _u8(s2)
    pp = _params(p)
    ie, je = C.c_int(0), C.c_int(0)
    best = oracle().oracle_gotoh_anchored_end(_ptr(a), _ptr(b), len(a), len(b), C.byref(pp), C.byref(ie), C.byref(je))
    return best, ie.value, je.value


def gotoh_span(s1, s2, p=DEFAULT):
    """(score, i_start, j_start, i_end, j_end), 1-based inclusive; i in seq2, j in seq1."""
    a, b = _u8(s1), _u8(s2)
    pp = _params(p)
    v = [C.c_int(0) for _ in range(4)]
    best = oracle().oracle_gotoh_span(_ptr(a), _ptr(b), len(a), len(b), C.byref(pp), *[C.byref(x) for x in v])
    return (best,) + tuple(x.value for x in v)


def gotoh_mt(s1, s2, p=DEFAULT, threads=0):
    return _call("oracle_gotoh_mt", s1, s2, p, threads)


def gotoh_fast(s1, s2, p=DEFAULT, threads=0):
    """oracle/gotoh_fast.c: one tile per SIMD lane; identical to gotoh_rolling, fast enough for BASELINE config 3."""
    return _call("oracle_gotoh_fast", s1, s2, p, threads)


def gotoh_banded(s1, s2, band_lo, band_hi, p=DEFAULT, want_cells=False):
    a, b = _u8(s1), _u8(s2)
    pp = _params(p)
    cells = C.c_int64(0)
    r = oracle().oracle_gotoh_banded(_ptr(a), _ptr(b), len(a), len(b), band_lo, band_hi, C.byref(pp), C.byref(cells))
    return (r, cells.value) if want_cells else r


def _batch_args(seqs):
    lens = np.array([len(s) for s in seqs], dtype=np.int32)
    offs = np.zeros(len(seqs), dtype=np.int64)
    if len(seqs):
        offs[1:] = np.cumsum(lens[:-1], dtype=np.int64)
    flat = np.concatenate([_u8(s) for s in seqs]) if len(seqs) and lens.sum() else np.zeros(1, dtype=np.uint8)
    return flat, offs, lens


def gotoh_batch(seqs1, seqs2, p=DEFAULT, threads=0):
    f1, o1, l1 = _batch_args(seqs1)
    f2, o2, l2 = _batch_args(seqs2)
    out = np.zeros(len(seqs1), dtype=np.int32)
    pp = _params(p)
    oracle().oracle_gotoh_batch(_ptr(f1), o1.ctypes.data_as(C.POINTER(C.c_int64)), l1.ctypes.data_as(C.POINTER(C.c_int)),
                                _ptr(f2), o2.ctypes.data_as(C.POINTER(C.c_int64)), l2.ctypes.data_as(C.POINTER(C.c_int)),
                                len(seqs1), C.byref(pp), threads, out.ctypes.data_as(C.POINTER(C.c_int)))
    return out


def gotoh_banded_batch(seqs1, seqs2, band_lo, band_hi, p=DEFAULT, threads=0):
    f1, o1, l1 = _batch_args(seqs1)
    f2, o2, l2 = _batch_args(seqs2)
    out = np.zeros(len(seqs1), dtype=np.int32)
    pp = _params(p)
    oracle().oracle_gotoh_banded_batch(_ptr(f1), o1.ctypes.data_as(C.POINTER(C.c_int64)), l1.ctypes.data_as(C.POINTER(C.c_int)),
                                       _ptr(f2), o2.ctypes.data_as(C.POINTER(C.c_int64)), l2.ctypes.data_as(C.POINTER(C.c_int)),
                                       len(seqs1), band_lo, band_hi, C.byref(pp), threads,
                                       out.ctypes.data_as(C.POINTER(C.c_int)))
    return out


def random_acgt(seed, stream, length):
    out = np.zeros(max(length, 1), dtype=np.uint8)
    oracle().oracle_random_acgt(seed, stream, length, _ptr(out))
    return out[:length]


# ---- the unmodified reference, when oracle/_ref/libref*.so exists ------------------------------
_ref = {}


def ref_available(p=DEFAULT):
    return _ref_path(p).exists()


def _ref_path(p):
    if tuple(p) == DEFAULT:
        return ORACLE_DIR / "_ref" / "libref.so"
    m, x, gi, ge = p
    return ORACLE_DIR / "_ref" / f"libref_{gi}_{ge}_{m}_{x}.so"


def ref(p=DEFAULT):
    key = tuple(p)
    if key not in _ref:
        lib = C.CDLL(str(_ref_path(p)))
        sig = [C.POINTER(C.c_ubyte), C.POINTER(C.c_ubyte), C.c_int, C.c_int]
        for name in ("ref_SmithWatermanScore", "ref_LazySmith", "ref_ParallelLazySmith_threads"):
            getattr(lib, name).argtypes = sig
            getattr(lib, name).restype = C.c_int
        lib.ref_batch.argtypes = [C.c_int, C.POINTER(C.c_ubyte), C.POINTER(C.c_longlong), C.POINTER(C.c_int),
                                  C.POINTER(C.c_ubyte), C.POINTER(C.c_longlong), C.POINTER(C.c_int),
                                  C.c_longlong, C.c_int, C.POINTER(C.c_int)]
        lib.ref_batch.restype = None
        lib.ref_hardware_concurrency.restype = C.c_int
        _ref[key] = lib
    return _ref[key]


def ref_call(name, s1, s2, p=DEFAULT):
    a, b = _u8(s1), _u8(s2)
    return getattr(ref(p), name)(_ptr(a), _ptr(b), len(a), len(b))
