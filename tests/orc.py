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


class OrcGroupsOut(C.Structure):
    _fields_ = ([(k, C.c_int32) for k in ("n_reads", "n_segs", "n_pairs", "n_juncs", "n_groups", "n_colors", "color")] +
                [("startgroup", C.c_int32 * 3), ("num_fragments", C.c_double), ("frag_len", C.c_double)] +
                [("r_strand", C.POINTER(C.c_int8)), ("pair_idx", C.POINTER(C.c_int32)),
                 ("seg_start", C.POINTER(C.c_uint32)), ("seg_end", C.POINTER(C.c_uint32)), ("rjunc", C.POINTER(C.c_int32)),
                 ("j_start", C.POINTER(C.c_uint32)), ("j_end", C.POINTER(C.c_uint32)), ("j_strand", C.POINTER(C.c_int8)),
                 ("j_nreads", c_f64p), ("j_nreads_good", c_f64p), ("j_left", c_f64p), ("j_right", c_f64p),
                 ("j_nm", c_f64p), ("j_mm", c_f64p),
                 ("g_start", C.POINTER(C.c_uint32)), ("g_end", C.POINTER(C.c_uint32)), ("g_color", C.POINTER(C.c_int32)),
                 ("g_next", C.POINTER(C.c_int32)), ("g_cov_sum", c_f32p), ("g_multi", c_f32p),
                 ("g_alive", C.POINTER(C.c_int8)), ("merge", C.POINTER(C.c_int32)), ("eqcol", C.POINTER(C.c_int32)),
                 ("rg_off", C.POINTER(C.c_uint32)), ("rg", C.POINTER(C.c_int32)),
                 ("g_color2", C.POINTER(C.c_int32)), ("g_neg_prop", c_f32p), ("eqcol2", C.POINTER(C.c_int32)),
                 ("eqnegcol", C.POINTER(C.c_int32)), ("eqposcol", C.POINTER(C.c_int32)),
                 ("bun_off", C.c_int32 * 4), ("bn_off", C.c_int32 * 4),
                 ("bun_len", C.POINTER(C.c_int32)), ("bun_start", C.POINTER(C.c_int32)), ("bun_last", C.POINTER(C.c_int32)),
                 ("bun_cov", c_f32p), ("bun_multi", c_f32p),
                 ("bn_start", C.POINTER(C.c_uint32)), ("bn_end", C.POINTER(C.c_uint32)), ("bn_cov", c_f32p),
                 ("bn_next", C.POINTER(C.c_int32)), ("group2bundle", C.POINTER(C.c_int32))])


def _np(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def groups(arrays: Dict[str, np.ndarray], cov: np.ndarray, cov_off: np.ndarray, good: np.ndarray, left: np.ndarray,
           right: np.ndarray, bundles=None, junctionsupport: int = 10, junctionthr: int = 1, bundledist: int = 50):
    """SURVEY §8 rows a8 + a9 (first half) per bundle (oracle/stg_oracle_groups.c): list of dicts, one per bundle."""
    L = lib()
    if not hasattr(L, "_gr_ready"):
        L.orc_groups_bundle.argtypes = [C.POINTER(StgPackedBatch), C.POINTER(OrcParams), C.c_int64, c_f32p, c_f64p, c_f64p,
                                        c_f64p, C.c_uint32, C.POINTER(OrcGroupsOut)]
        L.orc_groups_bundle.restype = None
        L.orc_groups_free.argtypes = [C.POINTER(OrcGroupsOut)]
        L.orc_groups_free.restype = None
        L._gr_ready = True
    pb, keep = make_packed(arrays)
    P = OrcParams(junctionsupport, junctionthr, 0)
    joff = arrays["b_junc_off"].astype(np.int64)
    good = np.ascontiguousarray(good, dtype=np.float64)
    left = np.ascontiguousarray(left, dtype=np.float64)
    right = np.ascontiguousarray(right, dtype=np.float64)

    def at(a, o, t):
        return C.cast(a.ctypes.data + o * a.itemsize, t)

    res = []
    for b in (range(int(pb.n_bundles)) if bundles is None else bundles):
        o = OrcGroupsOut()
        j0 = int(joff[b])
        L.orc_groups_bundle(C.byref(pb), C.byref(P), b, at(cov, int(cov_off[b]), c_f32p), at(good, j0, c_f64p),
                            at(left, j0, c_f64p), at(right, j0, c_f64p), bundledist, C.byref(o))
        nintr = o.n_segs - o.n_reads
        d = {"startgroup": np.array(list(o.startgroup), dtype=np.int32), "num_fragments": o.num_fragments,
             "frag_len": o.frag_len, "color": o.color,
             "r_strand": _np(o.r_strand, o.n_reads, np.int8), "pair_idx": _np(o.pair_idx, o.n_pairs, np.int32),
             "seg_start": _np(o.seg_start, o.n_segs, np.uint32), "seg_end": _np(o.seg_end, o.n_segs, np.uint32),
             "rjunc": _np(o.rjunc, nintr, np.int32),
             "j_start": _np(o.j_start, o.n_juncs, np.uint32), "j_end": _np(o.j_end, o.n_juncs, np.uint32),
             "j_strand": _np(o.j_strand, o.n_juncs, np.int8), "j_nreads": _np(o.j_nreads, o.n_juncs, np.float64),
             "j_nreads_good": _np(o.j_nreads_good, o.n_juncs, np.float64), "j_left": _np(o.j_left, o.n_juncs, np.float64),
             "j_right": _np(o.j_right, o.n_juncs, np.float64), "j_nm": _np(o.j_nm, o.n_juncs, np.float64),
             "j_mm": _np(o.j_mm, o.n_juncs, np.float64),
             "g_alive": _np(o.g_alive, o.n_groups, np.int8), "g_start": _np(o.g_start, o.n_groups, np.uint32),
             "g_end": _np(o.g_end, o.n_groups, np.uint32), "g_color": _np(o.g_color, o.n_groups, np.int32),
             "g_next": _np(o.g_next, o.n_groups, np.int32), "g_cov": _np(o.g_cov_sum, o.n_groups, np.float32),
             "g_multi": _np(o.g_multi, o.n_groups, np.float32), "merge": _np(o.merge, o.n_groups, np.int32),
             "eqcol": _np(o.eqcol, o.n_colors, np.int32), "rg_off": _np(o.rg_off, o.n_reads + 1, np.uint32),
             "rg": _np(o.rg, int(o.rg_off[o.n_reads]) if o.n_reads else 0, np.int32),
             # second half of a9: colour pass + bundles
             "g_color2": _np(o.g_color2, o.n_groups, np.int32), "g_neg_prop": _np(o.g_neg_prop, o.n_groups, np.float32),
             "eqcol2": _np(o.eqcol2, o.n_colors, np.int32), "eqneg": _np(o.eqnegcol, o.n_colors, np.int32),
             "eqpos": _np(o.eqposcol, o.n_colors, np.int32),
             "bun_off": np.array(list(o.bun_off), dtype=np.int32), "bn_off": np.array(list(o.bn_off), dtype=np.int32),
             "bun_len": _np(o.bun_len, o.bun_off[3], np.int32), "bun_start": _np(o.bun_start, o.bun_off[3], np.int32),
             "bun_last": _np(o.bun_last, o.bun_off[3], np.int32), "bun_cov": _np(o.bun_cov, o.bun_off[3], np.float32),
             "bun_multi": _np(o.bun_multi, o.bun_off[3], np.float32),
             "bn_start": _np(o.bn_start, o.bn_off[3], np.uint32), "bn_end": _np(o.bn_end, o.bn_off[3], np.uint32),
             "bn_cov": _np(o.bn_cov, o.bn_off[3], np.float32), "bn_next": _np(o.bn_next, o.bn_off[3], np.int32),
             "g2b": _np(o.group2bundle, 3 * o.n_groups, np.int32)}
        L.orc_groups_free(C.byref(o))
        res.append(d)
    return res


def find_trims(arrays: Dict[str, np.ndarray], cov: np.ndarray, cov_off: np.ndarray, bundle, sno, start, end):
    """find_all_trims (oracle/stg_oracle_trims.c) on regions (bundle[i], sno[i], start[i], end[i]).
    Returns (off [n+1], pos, abund, isstart): the reference's trimpoint vectors, concatenated."""
    L = lib()
    if not hasattr(L, "_ft_ready"):
        L.orc_trims_capacity.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_trims_capacity.restype = C.c_int
        L.orc_find_all_trims.argtypes = [c_f32p, C.c_int64, C.c_int64, C.c_int, C.c_uint32, C.c_uint32,
                                         C.POINTER(C.c_uint32), c_f32p, C.POINTER(C.c_uint8)]
        L.orc_find_all_trims.restype = C.c_int
        L._ft_ready = True
    lens = arrays["b_end"].astype(np.int64) - arrays["b_start"].astype(np.int64) + 3
    off = [0]
    pos, ab, st = [], [], []
    for b, s, a, z in zip(bundle, sno, start, end):
        cap = L.orc_trims_capacity(int(a), int(z))
        p = np.zeros(cap, np.uint32); f = np.zeros(cap, np.float32); t = np.zeros(cap, np.uint8)
        n = L.orc_find_all_trims(C.cast(cov.ctypes.data + int(cov_off[b]) * 4, c_f32p), int(lens[b]),
                                 int(arrays["b_start"][b]), int(s), int(a), int(z),
                                 p.ctypes.data_as(C.POINTER(C.c_uint32)), f.ctypes.data_as(c_f32p),
                                 t.ctypes.data_as(C.POINTER(C.c_uint8)))
        pos.append(p[:n]); ab.append(f[:n]); st.append(t[:n])
        off.append(off[-1] + n)
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    return np.array(off, dtype=np.uint32), cat(pos, np.uint32), cat(ab, np.float32), cat(st, np.uint8)


def neutral_preds(arrays: Dict[str, np.ndarray], cov: np.ndarray, cov_off: np.ndarray, good: np.ndarray, left: np.ndarray,
                  right: np.ndarray, bundles=None, junctionsupport: int = 10, junctionthr: int = 1, bundledist: int = 50,
                  mintranscriptlen: int = 200, mcov: float = 1.0):
    """SURVEY §8 row a19 (de novo part) per bundle (oracle/stg_oracle_neutral.c): list of dicts start, end, cov, tlen,
    geneno — the single-exon predictions of the neutral bundles, in the reference's order."""
    L = lib()
    groups(arrays, cov, cov_off, good, left, right, bundles=[])   # sets the argtypes of orc_groups_bundle
    if not hasattr(L, "_np_ready"):
        u32p, i32p = C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
        L.orc_neutral_preds.argtypes = [c_f32p, C.c_int64, C.c_int64, C.POINTER(OrcGroupsOut), C.c_int, C.c_float, C.c_int, u32p,
                                        u32p, c_f64p, i32p, i32p]
        L.orc_neutral_preds.restype = C.c_int
        L._np_ready = True
    pb, keep = make_packed(arrays)
    P = OrcParams(junctionsupport, junctionthr, 0)
    joff = arrays["b_junc_off"].astype(np.int64)
    good = np.ascontiguousarray(good, dtype=np.float64)
    left = np.ascontiguousarray(left, dtype=np.float64)
    right = np.ascontiguousarray(right, dtype=np.float64)
    lens, _ = cov_layout(arrays)

    def at(a, o, t):
        return C.cast(a.ctypes.data + o * a.itemsize, t)

    res = []
    for b in (range(int(pb.n_bundles)) if bundles is None else bundles):
        o = OrcGroupsOut()
        j0 = int(joff[b])
        covp = at(cov, int(cov_off[b]), c_f32p)
        L.orc_groups_bundle(C.byref(pb), C.byref(P), b, covp, at(good, j0, c_f64p), at(left, j0, c_f64p),
                            at(right, j0, c_f64p), bundledist, C.byref(o))
        cap = 16
        while True:
            st, en = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
            cv, tl, gn = np.zeros(cap, np.float64), np.zeros(cap, np.int32), np.zeros(cap, np.int32)
            n = L.orc_neutral_preds(covp, int(lens[b]), int(arrays["b_start"][b]), C.byref(o), mintranscriptlen, mcov, cap,
                                    st.ctypes.data_as(C.POINTER(C.c_uint32)), en.ctypes.data_as(C.POINTER(C.c_uint32)),
                                    cv.ctypes.data_as(c_f64p), tl.ctypes.data_as(C.POINTER(C.c_int32)),
                                    gn.ctypes.data_as(C.POINTER(C.c_int32)))
            if n <= cap:
                break
            cap = n
        L.orc_groups_free(C.byref(o))
        res.append({"start": st[:n].copy(), "end": en[:n].copy(), "cov": cv[:n].copy(), "tlen": tl[:n].copy(), "geneno": gn[:n].copy()})
    return res
