"""Multi-GPU sharding of a packed batch: bundles are independent, so they shard by genomic locus with
host-side load balancing; every rank runs the whole per-bundle pipeline on its shard and the only exchange is the
sum of the global totals that the FPKM/TPM normalisation needs (the reference adds them up under countMutex,
stringtie.cpp:1490-1497, and uses them at stringtie.cpp:915-917)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import numpy as np

from .engine import load_library
from .stgb import INPUT_FIELDS, input_view


def assign(arrays: Dict[str, np.ndarray], world: int) -> Tuple[np.ndarray, np.ndarray]:
    """rank_of[b] for every bundle (stg_shard_assign: greedy least-loaded on reads + span/64) and the load per rank."""
    L = load_library()
    nb = int(arrays["b_start"].size)
    off = np.ascontiguousarray(arrays["b_read_off"], dtype=np.uint32)
    st = np.ascontiguousarray(arrays["b_start"], dtype=np.int32)
    en = np.ascontiguousarray(arrays["b_end"], dtype=np.int32)
    rank_of = np.zeros(nb, dtype=np.int32)
    cost = np.zeros(world, dtype=np.float64)
    rc = L.stg_shard_assign(nb, off.ctypes.data_as(C.POINTER(C.c_uint32)), st.ctypes.data_as(C.POINTER(C.c_int32)),
                            en.ctypes.data_as(C.POINTER(C.c_int32)), world, rank_of.ctypes.data_as(C.POINTER(C.c_int32)),
                            cost.ctypes.data_as(C.POINTER(C.c_double)))
    if rc != 0:
        raise RuntimeError("stg_shard_assign failed: %d" % rc)
    return rank_of, cost


def take_bundles(arrays: Dict[str, np.ndarray], idx: np.ndarray) -> Dict[str, np.ndarray]:
    """The packed batch made of bundles `idx` (kept in the given order), offsets rebuilt."""
    a = arrays
    idx = np.asarray(idx, dtype=np.int64)
    bro, bjo = a["b_read_off"].astype(np.int64), a["b_junc_off"].astype(np.int64)
    rso, rpo = a["r_seg_off"].astype(np.int64), a["r_pair_off"].astype(np.int64)

    def ranges(lo, hi):
        n = hi - lo
        tot = int(n.sum())
        if tot == 0:
            return np.zeros(0, dtype=np.int64)
        start = np.repeat(lo - np.concatenate([[0], np.cumsum(n)[:-1]]), n)
        return np.arange(tot, dtype=np.int64) + start

    ridx = ranges(bro[idx], bro[idx + 1])
    jidx = ranges(bjo[idx], bjo[idx + 1])
    nseg = (rso[1:] - rso[:-1])[ridx]
    npair = (rpo[1:] - rpo[:-1])[ridx]
    sidx = ranges(rso[ridx], rso[ridx] + nseg)
    pidx = ranges(rpo[ridx], rpo[ridx] + npair)
    # introns of read n live at [r_seg_off[n] - n, r_seg_off[n+1] - (n+1))
    iidx = ranges(rso[ridx] - ridx, rso[ridx] - ridx + nseg - 1)
    out = {"b_start": a["b_start"][idx], "b_end": a["b_end"][idx]}
    for k, cnt in (("b_read_off", bro[idx + 1] - bro[idx]), ("b_junc_off", bjo[idx + 1] - bjo[idx]),
                   ("r_seg_off", nseg), ("r_pair_off", npair)):
        o = np.zeros(cnt.size + 1, dtype=np.int64)
        np.cumsum(cnt, out=o[1:])
        out[k] = o.astype(np.uint32)
    for k in ("r_start", "r_end", "r_len", "r_count", "r_nh", "r_strand", "r_flags"):
        out[k] = a[k][ridx]
    out["seg_start"], out["seg_end"] = a["seg_start"][sidx], a["seg_end"][sidx]
    out["rjunc_idx"] = a["rjunc_idx"][iidx]
    out["pair_idx"], out["pair_cnt"] = a["pair_idx"][pidx], a["pair_cnt"][pidx]
    for k in ("j_start", "j_end", "j_strand", "j_guide_match", "j_nreads", "j_nm", "j_mm"):
        out[k] = a[k][jidx]
    return input_view(out)


def shard(arrays: Dict[str, np.ndarray], world: int) -> Tuple[List[np.ndarray], np.ndarray]:
    """Bundle index lists per rank (genomic order preserved inside a rank) and the load per rank."""
    rank_of, cost = assign(arrays, world)
    return [np.nonzero(rank_of == r)[0] for r in range(world)], cost


def allreduce_totals(local: np.ndarray, dist=None) -> np.ndarray:
    """Sum of the per-rank totals {summed coverage, fragments, fragment length} over all ranks.  `dist` is
    torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(local, dtype=np.float64)
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.as_tensor(np.asarray(local, dtype=np.float64), device=dev)
    dist.all_reduce(t)
    return t.cpu().numpy()
