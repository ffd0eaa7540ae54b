"""ctypes mirror of include/stgpu.h (the C ABI of libstgpu).  Field order must match the header exactly."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List

import numpy as np

from .stgb import INPUT_FIELDS

c_u32p = C.POINTER(C.c_uint32)
c_i32p = C.POINTER(C.c_int32)
c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)
c_i16p = C.POINTER(C.c_int16)
c_i8p = C.POINTER(C.c_int8)
c_u8p = C.POINTER(C.c_uint8)
c_i64p = C.POINTER(C.c_int64)

_PTR = {np.dtype(np.uint32): c_u32p, np.dtype(np.int32): c_i32p, np.dtype(np.float32): c_f32p,
        np.dtype(np.float64): c_f64p, np.dtype(np.int16): c_i16p, np.dtype(np.int8): c_i8p,
        np.dtype(np.uint8): c_u8p, np.dtype(np.int64): c_i64p}


class StgParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("junctionsupport", C.c_uint32), ("sserror", C.c_uint32),
        ("junctionthr", C.c_int32), ("bundledist", C.c_uint32), ("mintranscriptlen", C.c_int32),
        ("allowed_nodes", C.c_int32), ("readthr", C.c_float), ("singlethr", C.c_float), ("isofrac", C.c_float),
        ("mcov", C.c_float), ("trim", C.c_uint8), ("eonly", C.c_uint8), ("nomulti", C.c_uint8),
        ("viral", C.c_uint8), ("mixedMode", C.c_uint8), ("isnascent", C.c_uint8), ("printNascent", C.c_uint8),
        ("guided", C.c_uint8), ("isunitig", C.c_uint8), ("longreads", C.c_uint8), ("rawreads", C.c_uint8),
        ("includesource", C.c_uint8), ("geneabundance", C.c_uint8), ("havePtFeatures", C.c_uint8),
        ("have_gseq", C.c_uint8), ("reserved", C.c_uint8 * 17),
    ]


_PACKED_ORDER = ["b_start", "b_end", "b_read_off", "b_junc_off", "r_start", "r_end", "r_len", "r_count", "r_nh",
                 "r_strand", "r_flags", "r_seg_off", "r_pair_off", "seg_start", "seg_end", "rjunc_idx", "pair_idx",
                 "pair_cnt", "j_start", "j_end", "j_strand", "j_guide_match", "j_nreads", "j_nm", "j_mm"]


class StgPackedBatch(C.Structure):
    _fields_ = [("n_bundles", C.c_int64), ("n_reads", C.c_int64), ("n_segs", C.c_int64), ("n_pairs", C.c_int64),
                ("n_juncs", C.c_int64)] + [(k, _PTR[np.dtype(INPUT_FIELDS[k])]) for k in _PACKED_ORDER]


_VIEW_ORDER = ["r_start", "r_end", "r_len", "r_count", "r_nh", "r_strand", "r_flags", "r_seg_off", "seg_start",
               "seg_end", "rjunc_idx", "r_pair_off", "pair_idx", "pair_cnt", "j_start", "j_end", "j_strand",
               "j_guide_match", "j_nreads", "j_nm", "j_mm"]


class StgBundleView(C.Structure):
    _fields_ = [("start", C.c_int32), ("end", C.c_int32), ("n_reads", C.c_int32), ("n_juncs", C.c_int32)] + \
               [(k, _PTR[np.dtype(INPUT_FIELDS[k])]) for k in _VIEW_ORDER]


class StgBundleResult(C.Structure):
    _fields_ = [("bpcov", c_f32p), ("j_nreads_good", c_f64p), ("j_leftsupport", c_f64p), ("j_rightsupport", c_f64p),
                ("assemble", C.c_int32), ("n_pred", C.c_int32), ("n_exon", C.c_int32), ("num_fragments", C.c_double),
                ("frag_len", C.c_double), ("p_start", C.POINTER(C.c_int32)), ("p_end", C.POINTER(C.c_int32)), ("p_cov", c_f64p),
                ("p_tlen", C.POINTER(C.c_int32)), ("p_geneno", C.POINTER(C.c_int32)), ("p_strand", C.POINTER(C.c_int8)),
                ("p_exon_off", C.POINTER(C.c_uint32)), ("e_start", C.POINTER(C.c_uint32)), ("e_end", C.POINTER(C.c_uint32)),
                ("e_cov", c_f32p)]


def ptr(a: np.ndarray):
    return a.ctypes.data_as(_PTR[a.dtype])


def make_packed(arrays: Dict[str, np.ndarray]):
    """Build a StgPackedBatch over numpy arrays.  Returns (struct, keepalive list)."""
    keep: List[np.ndarray] = []
    pb = StgPackedBatch()
    a = {k: np.ascontiguousarray(arrays[k], dtype=INPUT_FIELDS[k]) for k in _PACKED_ORDER}
    pb.n_bundles = a["b_start"].size
    pb.n_reads = a["r_start"].size
    pb.n_segs = a["seg_start"].size
    pb.n_pairs = a["pair_idx"].size
    pb.n_juncs = a["j_start"].size
    for k in _PACKED_ORDER:
        setattr(pb, k, ptr(a[k]))
        keep.append(a[k])
    return pb, keep


def bundle_view(arrays: Dict[str, np.ndarray], b: int):
    """StgBundleView of bundle b of a packed batch (per-bundle arrays are copies with rebased offsets)."""
    from .stgb import slice_bundles
    s = slice_bundles(arrays, b, b + 1)
    v = StgBundleView()
    v.start = int(s["b_start"][0])
    v.end = int(s["b_end"][0])
    v.n_reads = s["r_start"].size
    v.n_juncs = s["j_start"].size
    keep = []
    for k in _VIEW_ORDER:
        setattr(v, k, ptr(s[k]))
        keep.append(s[k])
    return v, keep
