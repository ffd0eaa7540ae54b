"""Packed-batch files (.stgb): named little-endian arrays (format in include/stgpu.h).

A packed batch is the concatenated SoA form of many BundleData inputs (bundle.h:314-333) — the thing the bundle
batcher hands to the GPU.  The same container also carries stage outputs dumped by the oracle tools.
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

MAGIC = 0x42475453  # "STGB"
_DT = {1: np.int8, 2: np.uint8, 3: np.int16, 4: np.int32, 5: np.uint32, 6: np.float32, 7: np.float64, 8: np.int64}
_DT_REV = {np.dtype(v): k for k, v in _DT.items()}

# field name -> dtype of the packed-batch INPUT schema (stg_packed_batch in include/stgpu.h)
INPUT_FIELDS = {
    "b_start": np.int32, "b_end": np.int32, "b_read_off": np.uint32, "b_junc_off": np.uint32,
    "r_start": np.uint32, "r_end": np.uint32, "r_len": np.uint32, "r_count": np.float32, "r_nh": np.int16,
    "r_strand": np.int8, "r_flags": np.uint8, "r_seg_off": np.uint32, "r_pair_off": np.uint32,
    "seg_start": np.uint32, "seg_end": np.uint32, "rjunc_idx": np.int32, "pair_idx": np.int32,
    "pair_cnt": np.float32, "j_start": np.uint32, "j_end": np.uint32, "j_strand": np.int8,
    "j_guide_match": np.int8, "j_nreads": np.float64, "j_nm": np.float64, "j_mm": np.float64,
}


def read_stgb(path: str) -> Dict[str, np.ndarray]:
    with open(path, "rb") as f:
        buf = f.read()
    magic, version, n = struct.unpack_from("<III", buf, 0)
    if magic != MAGIC:
        raise ValueError(f"{path}: not an STGB file")
    off = 12
    out: Dict[str, np.ndarray] = {}
    for _ in range(n):
        name = buf[off:off + 24].split(b"\0", 1)[0].decode()
        dt, _pad, cnt = struct.unpack_from("<IIQ", buf, off + 24)
        off += 40
        dtype = np.dtype(_DT[dt])
        nb = cnt * dtype.itemsize
        out[name] = np.frombuffer(buf, dtype=dtype, count=cnt, offset=off).copy()
        off += nb + (8 - nb % 8) % 8
    return out


def write_stgb(path: str, arrays: Dict[str, np.ndarray]) -> None:
    with open(path, "wb") as f:
        f.write(struct.pack("<III", MAGIC, 1, len(arrays)))
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            f.write(name.encode()[:23].ljust(24, b"\0"))
            f.write(struct.pack("<IIQ", _DT_REV[a.dtype], 0, a.size))
            b = a.tobytes()
            f.write(b)
            f.write(b"\0" * ((8 - len(b) % 8) % 8))


def input_view(arrays: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Only the packed-batch input fields, with the schema dtypes enforced."""
    return {k: np.ascontiguousarray(arrays[k], dtype=dt) for k, dt in INPUT_FIELDS.items()}


def slice_bundles(arrays: Dict[str, np.ndarray], lo: int, hi: int) -> Dict[str, np.ndarray]:
    """Sub-batch [lo, hi) of a packed batch (inputs only), offsets rebased."""
    a = arrays
    r0, r1 = int(a["b_read_off"][lo]), int(a["b_read_off"][hi])
    j0, j1 = int(a["b_junc_off"][lo]), int(a["b_junc_off"][hi])
    s0, s1 = int(a["r_seg_off"][r0]), int(a["r_seg_off"][r1])
    p0, p1 = int(a["r_pair_off"][r0]), int(a["r_pair_off"][r1])
    out = {
        "b_start": a["b_start"][lo:hi], "b_end": a["b_end"][lo:hi],
        "b_read_off": a["b_read_off"][lo:hi + 1] - np.uint32(r0),
        "b_junc_off": a["b_junc_off"][lo:hi + 1] - np.uint32(j0),
        "r_seg_off": a["r_seg_off"][r0:r1 + 1] - np.uint32(s0),
        "r_pair_off": a["r_pair_off"][r0:r1 + 1] - np.uint32(p0),
        "seg_start": a["seg_start"][s0:s1], "seg_end": a["seg_end"][s0:s1],
        "rjunc_idx": a["rjunc_idx"][s0 - r0:s1 - r1],
        "pair_idx": a["pair_idx"][p0:p1], "pair_cnt": a["pair_cnt"][p0:p1],
    }
    for k in ("r_start", "r_end", "r_len", "r_count", "r_nh", "r_strand", "r_flags"):
        out[k] = a[k][r0:r1]
    for k in ("j_start", "j_end", "j_strand", "j_guide_match", "j_nreads", "j_nm", "j_mm"):
        out[k] = a[k][j0:j1]
    return input_view(out)
