"""ctypes binding of oracle/liborc.so (the CPU restatement).  Test infrastructure: imported only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict

import numpy as np

from stringtie_b200.abi import StgPackedBatch, make_packed, c_f32p, c_f64p, c_i64p

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "liborc.so")


class OrcParams(C.Structure):
    _fields_ = [("junctionsupport", C.c_uint32), ("junctionthr", C.c_int32), ("longreads", C.c_uint8)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "orc"], stdout=subprocess.DEVNULL)
        L = C.CDLL(LIB)
        L.orc_batch_run.argtypes = [C.POINTER(StgPackedBatch), C.POINTER(OrcParams), c_i64p, c_f32p, c_f64p, c_f64p,
                                    c_f64p, C.c_int]
        L.orc_batch_run.restype = None
        L.orc_get_cov.argtypes = [c_f32p, C.c_int64, C.c_int, C.c_uint32, C.c_uint32]
        L.orc_get_cov.restype = C.c_double
        L.orc_get_cov_sign.argtypes = [c_f32p, C.c_int64, C.c_int, C.c_uint32, C.c_uint32]
        L.orc_get_cov_sign.restype = C.c_double
        L.orc_batch_stats.argtypes = [C.POINTER(StgPackedBatch), c_i64p, c_i64p, c_i64p]
        L.orc_batch_stats.restype = None
        _lib = L
    return _lib


def cov_layout(arrays: Dict[str, np.ndarray]):
    """Dense oracle layout: bundle b occupies 3*len_b floats, lane-major, at cov_off[b]."""
    lens = arrays["b_end"].astype(np.int64) - arrays["b_start"].astype(np.int64) + 3
    off = np.zeros(lens.size + 1, dtype=np.int64)
    np.cumsum(3 * lens, out=off[1:])
    return lens, off


def run(arrays: Dict[str, np.ndarray], want_cov: bool = True, threads: int = 0, junctionsupport: int = 10,
        junctionthr: int = 1):
    """Returns (cov [concatenated, lane-major per bundle] or None, cov_off, good, left, right)."""
    L = lib()
    pb, keep = make_packed(arrays)
    lens, off = cov_layout(arrays)
    cov = np.zeros(int(off[-1]), dtype=np.float32) if want_cov else None
    nj = int(pb.n_juncs)
    good = np.zeros(nj, dtype=np.float64)
    left = np.zeros(nj, dtype=np.float64)
    right = np.zeros(nj, dtype=np.float64)
    P = OrcParams(junctionsupport, junctionthr, 0)
    L.orc_batch_run(C.byref(pb), C.byref(P), off.ctypes.data_as(c_i64p),
                    cov.ctypes.data_as(c_f32p) if want_cov else None, good.ctypes.data_as(c_f64p),
                    left.ctypes.data_as(c_f64p), right.ctypes.data_as(c_f64p), threads)
    return cov, off, good, left, right


def get_cov(cov_bundle: np.ndarray, length: int, s: int, start: int, end: int, sign: bool = False) -> float:
    L = lib()
    f = L.orc_get_cov_sign if sign else L.orc_get_cov
    return f(cov_bundle.ctypes.data_as(c_f32p), length, s, start, end)


def stats(arrays):
    L = lib()
    pb, keep = make_packed(arrays)
    a = (C.c_int64 * 3)()
    L.orc_batch_stats(C.byref(pb), C.cast(C.byref(a, 0), c_i64p), C.cast(C.byref(a, 8), c_i64p),
                      C.cast(C.byref(a, 16), c_i64p))
    return int(a[0]), int(a[1]), int(a[2])


def junction_pass(arrays: Dict[str, np.ndarray], cov: np.ndarray, cov_off: np.ndarray, good: np.ndarray,
                  left: np.ndarray, right: np.ndarray, junctionsupport: int = 10, junctionthr: int = 1):
    """SURVEY §8 row a7 on every bundle (oracle/stg_oracle_junc.c).  Returns a dict of per-junction-row arrays:
    strand, nreads, nreads_good, leftsupport, rightsupport, mm, ej (bundle-local row of ejunction[k]), good,
    good_strand."""
    L = lib()
    if not hasattr(L, "_jp_ready"):
        i8p, i32p = C.POINTER(C.c_int8), C.POINTER(C.c_int32)
        L.orc_junction_pass.argtypes = [C.POINTER(StgPackedBatch), C.POINTER(OrcParams), C.c_int64, c_f32p, c_f64p,
                                        c_f64p, c_f64p, i8p, c_f64p, c_f64p, c_f64p, c_f64p, c_f64p, i32p, i8p, i8p]
        L.orc_junction_pass.restype = None
        L._jp_ready = True
    pb, keep = make_packed(arrays)
    nj = int(pb.n_juncs)
    out = {"strand": np.zeros(nj, np.int8), "nreads": np.zeros(nj), "nreads_good": np.zeros(nj),
           "leftsupport": np.zeros(nj), "rightsupport": np.zeros(nj), "mm": np.zeros(nj),
           "ej": np.zeros(nj, np.int32), "good": np.zeros(nj, np.int8), "good_strand": np.zeros(nj, np.int8)}
    P = OrcParams(junctionsupport, junctionthr, 0)
    joff = arrays["b_junc_off"].astype(np.int64)
    good = np.ascontiguousarray(good, dtype=np.float64)
    left = np.ascontiguousarray(left, dtype=np.float64)
    right = np.ascontiguousarray(right, dtype=np.float64)
    i8p, i32p = C.POINTER(C.c_int8), C.POINTER(C.c_int32)

    def at(a, o, t):
        return C.cast(a.ctypes.data + o * a.itemsize, t)

    for b in range(int(pb.n_bundles)):
        j0 = int(joff[b])
        if joff[b + 1] == j0:
            continue
        L.orc_junction_pass(C.byref(pb), C.byref(P), b, at(cov, int(cov_off[b]), c_f32p), at(good, j0, c_f64p),
                            at(left, j0, c_f64p), at(right, j0, c_f64p), at(out["strand"], j0, i8p),
                            at(out["nreads"], j0, c_f64p), at(out["nreads_good"], j0, c_f64p),
                            at(out["leftsupport"], j0, c_f64p), at(out["rightsupport"], j0, c_f64p),
                            at(out["mm"], j0, c_f64p), at(out["ej"], j0, i32p), at(out["good"], j0, i8p),
                            at(out["good_strand"], j0, i8p))
    return out
