"""Host-side mirror of the reference boundary over libstgpu's C ABI (include/stgpu.h).

`Engine.infer_transcripts(bundle)` has the shape of the reference call `int infer_transcripts(BundleData*)`
(rlink.h:380); `Engine.upload/run/fetch_*` is the batched form a host uses to push thousands of bundles at once.
There is no CPU implementation behind this class: if libstgpu.so is missing or no B200 is visible it raises.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from typing import Dict, List, Optional, Tuple

import numpy as np

from .abi import (StgBundleResult, StgBundleView, StgPackedBatch, StgParams, bundle_view, c_f32p, c_f64p, c_i64p,
                  make_packed, ptr)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstgpu.so")

EXPORTS = [
    "stg_params_default", "stg_init", "stg_shutdown", "stg_last_error", "stg_abi_version", "stg_builder_create",
    "stg_builder_add", "stg_builder_packed", "stg_builder_reset", "stg_builder_destroy", "stg_upload", "stg_run",
    "stg_sync", "stg_free_batch", "stg_trim", "stg_cov_layout", "stg_fetch_bpcov", "stg_fetch_bpcov_all", "stg_device_bpcov",
    "stg_upload_wait", "stg_fetch_bpcov_all_begin", "stg_fetch_wait", "stg_fetch_bpcov_lane_begin", "stg_fetch_bpcov_bundle_lane",
    "stg_pack_bpcov", "stg_fetch_bpcov_packed_begin", "stg_unpack_bpcov", "stg_unpack_bpcov_bundle", "stg_unpack_bpcov_stream", "stg_release_device",
    "stg_fetch_junction_support", "stg_fetch_junction_pass", "stg_fetch_cov_totals", "stg_query_cov", "stg_trims_capacity", "stg_find_trims", "stg_infer_transcripts", "stg_infer_transcripts_batch",
    "stg_build_groups", "stg_groups_array_info", "stg_fetch_groups_array", "stg_neutral_predictions",
    "stg_assemble", "stg_asm_array_info", "stg_fetch_asm_array", "stg_rewind", "stg_reduce_totals",
    "stg_shard_assign", "stg_selftest_chain", "stg_pool_stats", "stg_set_profiling", "stg_kernel_count", "stg_kernel_time", "stg_launch_counter", "stg_batch_stats", "stg_stream",
]


# enum stg_groups_array (include/stgpu.h), in order
GROUPS_ARRAYS = [
    "R_STRAND", "PAIR_IDX", "SEG_START", "SEG_END", "RJUNC", "GROUP_OFF", "COLOR_OFF", "APP_OFF", "RG_OFF", "STARTGROUP",
    "NUM_FRAGMENTS", "FRAG_LEN", "G_START", "G_END", "G_COLOR", "G_NEXT", "G_MERGE", "G_COV", "G_MULTI", "G_ALIVE", "G_COLOR2",
    "G_NEG_PROP", "EQCOL", "EQCOL2", "EQNEG", "EQPOS", "G2B", "BUN_LEN", "BUN_START", "BUN_LAST", "BUN_COV", "BUN_MULTI",
    "BN_START", "BN_END", "BN_COV", "BN_NEXT", "N_BUN", "N_BN", "RG", "J_START", "J_END", "J_STRAND", "J_NREADS",
    "J_NREADS_GOOD", "J_LEFT", "J_RIGHT", "J_NM", "J_MM",
    "NP_OFF", "NP_START", "NP_END", "NP_COV", "NP_TLEN", "NP_GENENO",
]
NEUTRAL_ARRAYS = GROUPS_ARRAYS[-6:]
GROUPS_DTYPES = {k: np.int32 for k in GROUPS_ARRAYS}
GROUPS_DTYPES.update({k: np.uint32 for k in ("NP_OFF", "NP_START", "NP_END", "SEG_START", "SEG_END", "GROUP_OFF", "COLOR_OFF", "APP_OFF", "RG_OFF", "G_START", "G_END",
                                             "BN_START", "BN_END", "N_BUN", "N_BN", "J_START", "J_END")})
GROUPS_DTYPES.update({k: np.int8 for k in ("R_STRAND", "G_ALIVE", "J_STRAND")})
GROUPS_DTYPES.update({k: np.float32 for k in ("G_COV", "G_MULTI", "G_NEG_PROP", "BUN_COV", "BUN_MULTI", "BN_COV")})
GROUPS_DTYPES.update({k: np.float64 for k in ("NP_COV", "NUM_FRAGMENTS", "FRAG_LEN", "J_NREADS", "J_NREADS_GOOD", "J_LEFT", "J_RIGHT", "J_NM", "J_MM")})


# enum stg_asm_array (include/stgpu.h), in order
ASM_ARRAYS = ["PRED_OFF", "P_START", "P_END", "P_COV", "P_TLEN", "P_GENENO", "P_STRAND", "P_EXON_OFF", "E_START", "E_END", "E_COV",
              "G_BUNDLE", "G_STRAND", "G_K", "G_NODE_OFF", "N_START", "N_END", "N_COV", "G_NTF"]
ASM_DTYPES = {"PRED_OFF": np.uint32, "P_START": np.uint32, "P_END": np.uint32, "P_COV": np.float64, "P_TLEN": np.int32,
              "P_GENENO": np.int32, "P_STRAND": np.int8, "P_EXON_OFF": np.uint32, "E_START": np.uint32, "E_END": np.uint32,
              "E_COV": np.float32, "G_BUNDLE": np.int32, "G_STRAND": np.int8, "G_K": np.int32, "G_NODE_OFF": np.uint32,
              "N_START": np.uint32, "N_END": np.uint32, "N_COV": np.float32, "G_NTF": np.uint32}


class StgError(RuntimeError):
    pass


_lib = None


def load_library() -> C.CDLL:
    """Load libstgpu.so (built in-tree by __graft_entry__.build()).  Never falls back to anything else."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StgError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(the engine has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.stg_params_default.argtypes = [C.POINTER(StgParams)]
    L.stg_params_default.restype = None
    L.stg_init.argtypes = [C.POINTER(StgParams), C.c_int, C.POINTER(vp)]
    L.stg_shutdown.argtypes = [vp]
    L.stg_last_error.argtypes = [vp]
    L.stg_last_error.restype = C.c_char_p
    L.stg_builder_create.argtypes = [vp, C.POINTER(vp)]
    L.stg_builder_add.argtypes = [vp, C.POINTER(StgBundleView)]
    L.stg_builder_packed.argtypes = [vp, C.POINTER(StgPackedBatch)]
    L.stg_builder_reset.argtypes = [vp]
    L.stg_builder_destroy.argtypes = [vp]
    L.stg_upload.argtypes = [vp, C.POINTER(StgPackedBatch), C.POINTER(vp)]
    L.stg_run.argtypes = [vp, vp]
    L.stg_sync.argtypes = [vp]
    L.stg_free_batch.argtypes = [vp, vp]
    L.stg_trim.argtypes = [vp]
    L.stg_cov_layout.argtypes = [vp, c_i64p, c_i64p, c_i64p]
    L.stg_fetch_bpcov.argtypes = [vp, vp, C.c_int64, c_f32p]
    L.stg_fetch_bpcov_all.argtypes = [vp, vp, c_f32p]
    L.stg_fetch_bpcov_all_begin.argtypes = [vp, vp, c_f32p]
    L.stg_fetch_bpcov_lane_begin.argtypes = [vp, vp, C.c_int, c_f32p]
    L.stg_fetch_bpcov_bundle_lane.argtypes = [vp, vp, C.c_int64, C.c_int, c_f32p]
    L.stg_pack_bpcov.argtypes = [vp, vp, c_i64p, c_i64p]
    L.stg_release_device.argtypes = [vp, vp]
    L.stg_unpack_bpcov_bundle.argtypes = [vp, C.c_int64, C.POINTER(C.c_uint32), c_f32p, c_f32p, c_f32p]
    L.stg_unpack_bpcov_stream.argtypes = [vp, C.c_int64, C.c_int64, C.POINTER(C.c_uint32), c_f32p, c_f32p, c_f32p, C.c_int64, c_f64p]
    L.stg_fetch_bpcov_packed_begin.argtypes = [vp, vp, C.POINTER(C.c_uint32), c_f32p, c_f32p]
    L.stg_unpack_bpcov.argtypes = [vp, C.c_int64, C.c_int64, C.POINTER(C.c_uint32), c_f32p, c_f32p, c_f32p]
    L.stg_fetch_wait.argtypes = [vp, vp]
    L.stg_upload_wait.argtypes = [vp, vp]
    L.stg_device_bpcov.argtypes = [vp]
    L.stg_device_bpcov.restype = vp
    L.stg_fetch_junction_support.argtypes = [vp, vp, c_f64p, c_f64p, c_f64p]
    i8p, i32p = C.POINTER(C.c_int8), C.POINTER(C.c_int32)
    L.stg_fetch_junction_pass.argtypes = [vp, vp, i8p, c_f64p, c_f64p, c_f64p, c_f64p, c_f64p, i32p, i8p, i8p]
    L.stg_fetch_cov_totals.argtypes = [vp, vp, c_f64p]
    L.stg_query_cov.argtypes = [vp, vp, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int8), C.POINTER(C.c_uint32),
                                C.POINTER(C.c_uint32), C.c_int, c_f64p]
    u32p, u8p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
    L.stg_trims_capacity.argtypes = [C.c_uint32, C.c_uint32]
    L.stg_find_trims.argtypes = [vp, vp, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int8), u32p, u32p, u32p, u32p, u32p,
                                 c_f32p, u8p]
    L.stg_infer_transcripts.argtypes = [vp, C.POINTER(StgBundleView), C.POINTER(StgBundleResult)]
    L.stg_infer_transcripts_batch.argtypes = [vp, C.c_int64, C.POINTER(StgBundleView), C.POINTER(StgBundleResult)]
    L.stg_build_groups.argtypes = [vp, vp]
    L.stg_neutral_predictions.argtypes = [vp, vp]
    L.stg_groups_array_info.argtypes = [vp, C.c_int, c_i64p, C.POINTER(C.c_int32)]
    L.stg_fetch_groups_array.argtypes = [vp, vp, C.c_int, C.c_int64, C.c_int64, C.c_void_p]
    L.stg_assemble.argtypes = [vp, vp]
    L.stg_asm_array_info.argtypes = [vp, C.c_int, c_i64p, C.POINTER(C.c_int32)]
    L.stg_fetch_asm_array.argtypes = [vp, vp, C.c_int, C.c_int64, C.c_int64, C.c_void_p]
    L.stg_rewind.argtypes = [vp, vp]
    L.stg_pool_stats.argtypes = [vp, c_i64p]
    L.stg_selftest_chain.argtypes = [vp, C.c_int64, C.POINTER(C.c_uint32), c_f32p, c_f32p, c_f32p, C.c_int, c_f32p]
    L.stg_reduce_totals.argtypes = [vp, vp, vp]
    L.stg_shard_assign.argtypes = [C.c_int64, C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int,
                                   C.POINTER(C.c_int32), c_f64p]
    L.stg_set_profiling.argtypes = [vp, C.c_int]
    L.stg_kernel_count.argtypes = [vp]
    L.stg_kernel_time.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float)]
    L.stg_launch_counter.argtypes = [vp]
    L.stg_launch_counter.restype = C.c_int64
    L.stg_batch_stats.argtypes = [vp, vp, c_i64p, c_i64p, c_i64p]
    L.stg_stream.argtypes = [vp]
    L.stg_stream.restype = vp
    _lib = L
    return L


def default_params() -> StgParams:
    p = StgParams()
    load_library().stg_params_default(C.byref(p))
    return p


class DeviceBatch:
    """A packed batch resident in HBM together with its outputs (stg_dbatch)."""

    def __init__(self, engine: "Engine", handle: int, arrays_keepalive, n_bundles: int, n_juncs: int):
        self.engine = engine
        self.handle = C.c_void_p(handle)
        self._keep = arrays_keepalive
        self.n_bundles = n_bundles
        self.n_juncs = n_juncs
        self._layout: Optional[Tuple[np.ndarray, np.ndarray, int]] = None

    def layout(self):
        if self._layout is None:
            off = np.zeros(self.n_bundles, dtype=np.int64)
            stride = np.zeros(self.n_bundles, dtype=np.int64)
            tot = C.c_int64(0)
            self.engine._ck(self.engine.L.stg_cov_layout(self.handle, ptr(off), ptr(stride), C.byref(tot)))
            self._layout = (off, stride, tot.value)
        return self._layout

    def free(self):
        if self.handle:
            self.engine.L.stg_free_batch(self.engine.ctx, self.handle)
            self.handle = None


class Engine:
    def __init__(self, device: int = 0, params: Optional[StgParams] = None):
        self.L = load_library()
        self.ctx = C.c_void_p()
        p = params if params is not None else default_params()
        rc = self.L.stg_init(C.byref(p), device, C.byref(self.ctx))
        if rc != 0:
            raise StgError(f"stg_init failed ({rc}): {self.L.stg_last_error(None).decode()}")
        self.device = device

    def _ck(self, rc: int):
        if rc != 0:
            raise StgError(f"libstgpu error {rc}: {self.L.stg_last_error(self.ctx).decode()}")

    def trim(self):
        """Give the pooled device memory back to the driver (stg_trim)."""
        self._ck(self.L.stg_trim(self.ctx))

    def close(self):
        if self.ctx:
            self.L.stg_shutdown(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- batched path -----------------------------------------------------------------------------------
    def upload(self, arrays: Dict[str, np.ndarray]) -> DeviceBatch:
        pb, keep = make_packed(arrays)
        h = C.c_void_p()
        self._ck(self.L.stg_upload(self.ctx, C.byref(pb), C.byref(h)))
        return DeviceBatch(self, h.value, keep, int(pb.n_bundles), int(pb.n_juncs))

    def upload_packed(self, pb: StgPackedBatch, keep=None) -> DeviceBatch:
        h = C.c_void_p()
        self._ck(self.L.stg_upload(self.ctx, C.byref(pb), C.byref(h)))
        return DeviceBatch(self, h.value, keep, int(pb.n_bundles), int(pb.n_juncs))

    def run(self, batch: DeviceBatch, sync: bool = True):
        if os.environ.get("STGPU_RUN_DEBUG"):
            import time
            t0 = time.perf_counter()
            rc = self.L.stg_run(self.ctx, batch.handle)
            print("Engine.run: stg_run took %.3f ms of Python's clock" % (1e3 * (time.perf_counter() - t0)), file=sys.stderr)
            self._ck(rc)
        else:
            self._ck(self.L.stg_run(self.ctx, batch.handle))
        if sync:
            self.sync()

    def sync(self):
        self._ck(self.L.stg_sync(self.ctx))

    def fetch_bpcov(self, batch: DeviceBatch, b: int, length: int) -> np.ndarray:
        out = np.empty(3 * length, dtype=np.float32)
        self._ck(self.L.stg_fetch_bpcov(self.ctx, batch.handle, b, ptr(out)))
        return out

    def fetch_bpcov_all(self, batch: DeviceBatch) -> np.ndarray:
        _, _, tot = batch.layout()
        out = np.empty(tot, dtype=np.float32)
        self._ck(self.L.stg_fetch_bpcov_all(self.ctx, batch.handle, ptr(out)))
        return out

    def fetch_junction_support(self, batch: DeviceBatch):
        good = np.empty(batch.n_juncs, dtype=np.float64)
        left = np.empty(batch.n_juncs, dtype=np.float64)
        right = np.empty(batch.n_juncs, dtype=np.float64)
        self._ck(self.L.stg_fetch_junction_support(self.ctx, batch.handle, ptr(good), ptr(left), ptr(right)))
        return good, left, right

    def fetch_junction_pass(self, batch: DeviceBatch) -> Dict[str, np.ndarray]:
        """Row a7: the junction table after the junction pass of build_graphs + good_junc verdicts (stg_fetch_junction_pass)."""
        n = batch.n_juncs
        out = {"strand": np.zeros(n, np.int8), "nreads": np.zeros(n), "nreads_good": np.zeros(n),
               "leftsupport": np.zeros(n), "rightsupport": np.zeros(n), "mm": np.zeros(n),
               "ej": np.zeros(n, np.int32), "good": np.zeros(n, np.int8), "good_strand": np.zeros(n, np.int8)}
        i8p, i32p = C.POINTER(C.c_int8), C.POINTER(C.c_int32)
        self._ck(self.L.stg_fetch_junction_pass(
            self.ctx, batch.handle, out["strand"].ctypes.data_as(i8p), ptr(out["nreads"]), ptr(out["nreads_good"]),
            ptr(out["leftsupport"]), ptr(out["rightsupport"]), ptr(out["mm"]), out["ej"].ctypes.data_as(i32p),
            out["good"].ctypes.data_as(i8p), out["good_strand"].ctypes.data_as(i8p)))
        return out

    def find_trims(self, batch: DeviceBatch, bundle, sno, start, end):
        """find_all_trims (stg_find_trims) on regions; returns (off [n+1], pos, abundance, is_start), compacted."""
        bundle = np.ascontiguousarray(bundle, dtype=np.int32); sno = np.ascontiguousarray(sno, dtype=np.int8)
        start = np.ascontiguousarray(start, dtype=np.uint32); end = np.ascontiguousarray(end, dtype=np.uint32)
        n = bundle.size
        cap = np.array([self.L.stg_trims_capacity(int(a), int(z)) for a, z in zip(start, end)], dtype=np.int64)
        cap_off = np.zeros(n + 1, dtype=np.uint32)
        np.cumsum(cap, out=cap_off[1:])
        tot = int(cap_off[-1])
        cnt = np.zeros(n, np.uint32); pos = np.zeros(tot + 1, np.uint32); ab = np.zeros(tot + 1, np.float32)
        st = np.zeros(tot + 1, np.uint8)
        u32p, u8p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
        self._ck(self.L.stg_find_trims(self.ctx, batch.handle, n, bundle.ctypes.data_as(C.POINTER(C.c_int32)),
                                       sno.ctypes.data_as(C.POINTER(C.c_int8)), start.ctypes.data_as(u32p),
                                       end.ctypes.data_as(u32p), cap_off.ctypes.data_as(u32p), cnt.ctypes.data_as(u32p),
                                       pos.ctypes.data_as(u32p), ab.ctypes.data_as(c_f32p), st.ctypes.data_as(u8p)))
        off = np.zeros(n + 1, dtype=np.uint32)
        np.cumsum(cnt, out=off[1:])
        sel = np.concatenate([np.arange(int(cap_off[i]), int(cap_off[i]) + int(cnt[i])) for i in range(n)]) if n else np.zeros(0, int)
        sel = sel.astype(np.int64)
        return off, pos[sel], ab[sel], st[sel]

    def pool_stats(self) -> Dict[str, int]:
        out = np.zeros(5, dtype=np.int64)
        self._ck(self.L.stg_pool_stats(self.ctx, ptr(out)))
        return {"cudaMalloc_calls": int(out[0]), "cudaMalloc_bytes": int(out[1]), "cudaFree_calls": int(out[2]), "pool_flushes": int(out[3]),
                "pool_bytes": int(out[4])}

    def selftest_chain(self, off, x, s0, plain=False, want_ms=False):
        """Ordered float32 sums of the CSR lists (stg_selftest_chain): the kernels' integer form of the addition chain."""
        off = np.ascontiguousarray(off, dtype=np.uint32); x = np.ascontiguousarray(x, dtype=np.float32)
        s0 = np.ascontiguousarray(s0, dtype=np.float32)
        out = np.zeros(s0.size, np.float32)
        ms = C.c_float(0.0)
        self._ck(self.L.stg_selftest_chain(self.ctx, s0.size, off.ctypes.data_as(C.POINTER(C.c_uint32)), x.ctypes.data_as(c_f32p),
                                           s0.ctypes.data_as(c_f32p), out.ctypes.data_as(c_f32p), int(plain), C.byref(ms)))
        return (out, ms.value) if want_ms else out

    # ---- rows a8 + a9 -------------------------------------------------------------------------------------
    def build_groups(self, batch: DeviceBatch):
        """Read repair, groups, colours, bundles / bundlenodes of every bundle (stg_build_groups); results stay on the device."""
        self._ck(self.L.stg_build_groups(self.ctx, batch.handle))

    def fetch_groups_array(self, batch: DeviceBatch, which: str) -> np.ndarray:
        """One batch-wide result array of stg_build_groups by name (enum stg_groups_array without the STG_GA_ prefix)."""
        idx = GROUPS_ARRAYS.index(which)
        n, sz = C.c_int64(), C.c_int32()
        self._ck(self.L.stg_groups_array_info(batch.handle, idx, C.byref(n), C.byref(sz)))
        out = np.empty(n.value, dtype=GROUPS_DTYPES[which])
        assert out.itemsize == sz.value, which
        self._ck(self.L.stg_fetch_groups_array(self.ctx, batch.handle, idx, 0, n.value, C.c_void_p(out.ctypes.data)))
        return out

    def fetch_groups(self, batch: DeviceBatch) -> Dict[str, np.ndarray]:
        return {k: self.fetch_groups_array(batch, k) for k in GROUPS_ARRAYS if k not in NEUTRAL_ARRAYS}

    # ---- row a19 (de novo part) ----------------------------------------------------------------------------
    def neutral_predictions(self, batch: DeviceBatch) -> Dict[str, np.ndarray]:
        """Single-exon predictions of the neutral bundles (stg_neutral_predictions; needs build_groups): batch-wide
        arrays off [n_bundles+1], start, end, cov (f64), tlen, geneno."""
        self._ck(self.L.stg_neutral_predictions(self.ctx, batch.handle))
        return {k[3:].lower(): self.fetch_groups_array(batch, k) for k in NEUTRAL_ARRAYS}

    # ---- rows a10-a15 ---------------------------------------------------------------------------------------
    def assemble(self, batch: DeviceBatch) -> None:
        """Splice graphs, read -> transfrag mapping, flow decomposition, predictions (stg_assemble; needs build_groups
        and neutral_predictions)."""
        self._ck(self.L.stg_assemble(self.ctx, batch.handle))

    def rewind(self, batch: DeviceBatch) -> None:
        self._ck(self.L.stg_rewind(self.ctx, batch.handle))

    def reduce_totals(self, batch: DeviceBatch, dev_ptr: int) -> None:
        """num_fragments, frag_len and lane-1 coverage totals of the batch into 3 f64 at device address dev_ptr (async)."""
        self._ck(self.L.stg_reduce_totals(self.ctx, batch.handle, C.c_void_p(dev_ptr)))

    def fetch_asm_array(self, batch: DeviceBatch, which: str) -> np.ndarray:
        idx = ASM_ARRAYS.index(which)
        n, sz = C.c_int64(), C.c_int32()
        self._ck(self.L.stg_asm_array_info(batch.handle, idx, C.byref(n), C.byref(sz)))
        out = np.empty(n.value, dtype=ASM_DTYPES[which])
        assert out.itemsize == sz.value, which
        self._ck(self.L.stg_fetch_asm_array(self.ctx, batch.handle, idx, 0, n.value, C.c_void_p(out.ctypes.data)))
        return out

    def fetch_predictions(self, batch: DeviceBatch) -> Dict[str, np.ndarray]:
        """CPrediction list of the graphs of every bundle (after assemble): off [n_bundles+1], start, end, cov, tlen,
        geneno, strand, exon_off [n_pred+1], e_start, e_end, e_cov."""
        m = {"PRED_OFF": "off", "P_START": "start", "P_END": "end", "P_COV": "cov", "P_TLEN": "tlen", "P_GENENO": "geneno",
             "P_STRAND": "strand", "P_EXON_OFF": "exon_off", "E_START": "e_start", "E_END": "e_end", "E_COV": "e_cov"}
        return {v: self.fetch_asm_array(batch, k) for k, v in m.items()}

    def neutral_predictions_only(self, batch: DeviceBatch) -> None:
        self._ck(self.L.stg_neutral_predictions(self.ctx, batch.handle))

    def fetch_neutral(self, batch: DeviceBatch) -> Dict[str, np.ndarray]:
        return {k[3:].lower(): self.fetch_groups_array(batch, k) for k in NEUTRAL_ARRAYS}

    def cov_total_floats(self, batch: DeviceBatch) -> int:
        """Number of floats stg_fetch_bpcov_all writes (the padded 3-lane bpcov of every bundle)."""
        off = np.zeros(batch.n_bundles, dtype=np.int64)
        stride = np.zeros(batch.n_bundles, dtype=np.int64)
        tot = C.c_int64()
        self._ck(self.L.stg_cov_layout(batch.handle, ptr(off), ptr(stride), C.byref(tot)))
        return int(tot.value)

    def fetch_bpcov_all_into(self, batch: DeviceBatch, host_ptr: int) -> None:
        """bpcov of every bundle into caller memory (pinned memory makes it one DMA)."""
        self._ck(self.L.stg_fetch_bpcov_all(self.ctx, batch.handle, C.cast(C.c_void_p(host_ptr), c_f32p)))

    def fetch_bpcov_lane_begin(self, batch: DeviceBatch, lane: int, host_ptr: int) -> None:
        """One strand lane of every bundle (total_floats / 3 floats) on the copy-out stream (stg_fetch_bpcov_lane_begin)."""
        self._ck(self.L.stg_fetch_bpcov_lane_begin(self.ctx, batch.handle, lane, C.cast(C.c_void_p(host_ptr), c_f32p)))

    def fetch_bpcov_bundle_lane(self, batch: DeviceBatch, bundle: int, lane: int, host_ptr: int) -> None:
        self._ck(self.L.stg_fetch_bpcov_bundle_lane(self.ctx, batch.handle, bundle, lane, C.cast(C.c_void_p(host_ptr), c_f32p)))

    def release_device(self, batch: DeviceBatch) -> None:
        """Device memory of the batch back to the pool; the handle stays good for unpack_bpcov until batch.free()."""
        self._ck(self.L.stg_release_device(self.ctx, batch.handle))

    def pack_bpcov(self, batch: DeviceBatch):
        """bpcov packed on the device for the trip to the host (stg_pack_bpcov): returns (n_tiles, n_raw)."""
        nt, nr = C.c_int64(0), C.c_int64(0)
        self._ck(self.L.stg_pack_bpcov(self.ctx, batch.handle, C.byref(nt), C.byref(nr)))
        return int(nt.value), int(nr.value)

    def fetch_bpcov_packed_begin(self, batch: DeviceBatch, desc_ptr: int, val_ptr: int, raw_ptr: int) -> None:
        cast = lambda p, t: C.cast(C.c_void_p(p), t)
        self._ck(self.L.stg_fetch_bpcov_packed_begin(self.ctx, batch.handle, cast(desc_ptr, C.POINTER(C.c_uint32)), cast(val_ptr, c_f32p),
                                                     cast(raw_ptr, c_f32p)))

    def unpack_bpcov(self, batch: DeviceBatch, b0: int, b1: int, desc_ptr: int, val_ptr: int, raw_ptr: int, out_ptr: int) -> None:
        """Host side: bpcov of the bundles [b0, b1) from the packed form into a buffer laid out like fetch_bpcov_all's."""
        cast = lambda p, t: C.cast(C.c_void_p(p), t)
        self._ck(self.L.stg_unpack_bpcov(batch.handle, b0, b1, cast(desc_ptr, C.POINTER(C.c_uint32)), cast(val_ptr, c_f32p), cast(raw_ptr, c_f32p),
                                         cast(out_ptr, c_f32p)))

    def unpack_bpcov_bundle(self, batch: DeviceBatch, b: int, length: int, desc: np.ndarray, val: np.ndarray, raw: np.ndarray) -> np.ndarray:
        """BundleData::bpcov of bundle b ([3, length], length = end - start + 3) out of the packed form (stg_unpack_bpcov_bundle)."""
        out = np.zeros(3 * int(length), dtype=np.float32)
        self._ck(self.L.stg_unpack_bpcov_bundle(batch.handle, b, desc.ctypes.data_as(C.POINTER(C.c_uint32)), ptr(val), ptr(raw), ptr(out)))
        return out.reshape(3, -1)

    def unpack_bpcov_stream(self, batch: DeviceBatch, b0: int, b1: int, desc_ptr: int, val_ptr: int, raw_ptr: int, work: np.ndarray) -> float:
        """A worker thread's share [b0, b1) of a batch, bundle after bundle into its own working arrays (stg_unpack_bpcov_stream)."""
        cast = lambda p, t: C.cast(C.c_void_p(p), t)
        chk = C.c_double(0.0)
        self._ck(self.L.stg_unpack_bpcov_stream(batch.handle, b0, b1, cast(desc_ptr, C.POINTER(C.c_uint32)), cast(val_ptr, c_f32p), cast(raw_ptr, c_f32p),
                                                ptr(work), work.size, C.byref(chk)))
        return chk.value

    def fetch_bpcov_all_begin(self, batch: DeviceBatch, host_ptr: int) -> None:
        """Start the same copy on the copy-out stream (overlaps the group / assembly stages); fetch_wait completes it."""
        self._ck(self.L.stg_fetch_bpcov_all_begin(self.ctx, batch.handle, C.cast(C.c_void_p(host_ptr), c_f32p)))

    def fetch_wait(self, batch: DeviceBatch) -> None:
        self._ck(self.L.stg_fetch_wait(self.ctx, batch.handle))

    def upload_wait(self, batch: DeviceBatch) -> None:
        self._ck(self.L.stg_upload_wait(self.ctx, batch.handle))

    def fetch_cov_totals(self, batch: DeviceBatch) -> np.ndarray:
        out = np.empty(batch.n_bundles * 3, dtype=np.float64)
        self._ck(self.L.stg_fetch_cov_totals(self.ctx, batch.handle, ptr(out)))
        return out.reshape(batch.n_bundles, 3)

    def query_cov(self, batch: DeviceBatch, bundle, lane, start, end, sign: bool = False) -> np.ndarray:
        bundle = np.ascontiguousarray(bundle, dtype=np.int32)
        lane = np.ascontiguousarray(lane, dtype=np.int8)
        start = np.ascontiguousarray(start, dtype=np.uint32)
        end = np.ascontiguousarray(end, dtype=np.uint32)
        out = np.empty(bundle.size, dtype=np.float64)
        self._ck(self.L.stg_query_cov(self.ctx, batch.handle, bundle.size, bundle.ctypes.data_as(C.POINTER(C.c_int32)),
                                      lane.ctypes.data_as(C.POINTER(C.c_int8)),
                                      start.ctypes.data_as(C.POINTER(C.c_uint32)),
                                      end.ctypes.data_as(C.POINTER(C.c_uint32)), 1 if sign else 0, ptr(out)))
        return out

    def stats(self, batch: DeviceBatch) -> Tuple[int, int, int]:
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self._ck(self.L.stg_batch_stats(self.ctx, batch.handle, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    # ---- measurement hooks ------------------------------------------------------------------------------
    def set_profiling(self, on: bool):
        self._ck(self.L.stg_set_profiling(self.ctx, 1 if on else 0))

    def kernel_times(self) -> List[Tuple[str, float]]:
        out = []
        for i in range(self.L.stg_kernel_count(self.ctx)):
            name = C.c_char_p()
            ms = C.c_float()
            self._ck(self.L.stg_kernel_time(self.ctx, i, C.byref(name), C.byref(ms)))
            out.append((name.value.decode(), ms.value))
        return out

    def launch_counter(self) -> int:
        return int(self.L.stg_launch_counter(self.ctx))

    # ---- the reference-shaped call: one bundle, synchronous ------------------------------------------------
    def infer_transcripts(self, arrays: Dict[str, np.ndarray], b: int = 0):
        """Drop-in for `infer_transcripts(BundleData*)` on bundle b of a packed batch.  Returns what the stage
        writes back into BundleData: (bpcov[3][len], nreads_good, leftsupport, rightsupport)."""
        view, keep = bundle_view(arrays, b)
        length = view.end - view.start + 3
        cov = np.empty(3 * length, dtype=np.float32)
        good = np.zeros(view.n_juncs, dtype=np.float64)
        left = np.zeros(view.n_juncs, dtype=np.float64)
        right = np.zeros(view.n_juncs, dtype=np.float64)
        res = StgBundleResult(ptr(cov), ptr(good), ptr(left), ptr(right))
        self._ck(self.L.stg_infer_transcripts(self.ctx, C.byref(view), C.byref(res)))
        return cov.reshape(3, length), good, left, right

    def infer_transcripts_batch(self, arrays: Dict[str, np.ndarray], bundles):
        """stg_infer_transcripts_batch on the listed bundles: a list of (bpcov, predictions dict, num_fragments, frag_len), one
        per bundle, each as infer_transcripts_full returns it."""
        n = len(bundles)
        views = (StgBundleView * n)()
        results = (StgBundleResult * n)()
        keep, covs = [], []
        for i, b in enumerate(bundles):
            view, k = bundle_view(arrays, int(b))
            keep.append(k)
            views[i] = view
            length = view.end - view.start + 3
            cov = np.zeros(3 * length, dtype=np.float32)
            covs.append((cov, length))
            results[i] = StgBundleResult(ptr(cov), None, None, None)
            results[i].assemble = 1
        self._ck(self.L.stg_infer_transcripts_batch(self.ctx, n, views, results))
        cp = lambda p, k, dt: np.ctypeslib.as_array(p, shape=(k,)).astype(dt, copy=True) if k else np.empty(0, dtype=dt)
        out = []
        for i in range(n):
            res = results[i]
            m, ne = res.n_pred, res.n_exon
            pred = {"start": cp(res.p_start, m, np.int32), "end": cp(res.p_end, m, np.int32), "cov": cp(res.p_cov, m, np.float64),
                    "tlen": cp(res.p_tlen, m, np.int32), "geneno": cp(res.p_geneno, m, np.int32), "strand": cp(res.p_strand, m, np.int8),
                    "exon_off": cp(res.p_exon_off, m + 1, np.uint32) if m else np.zeros(1, dtype=np.uint32),
                    "e_start": cp(res.e_start, ne, np.uint32), "e_end": cp(res.e_end, ne, np.uint32), "e_cov": cp(res.e_cov, ne, np.float32)}
            out.append((covs[i][0].reshape(3, covs[i][1]), pred, res.num_fragments, res.frag_len))
        return out

    def infer_transcripts_full(self, arrays: Dict[str, np.ndarray], b: int = 0):
        """The whole of `infer_transcripts(BundleData*)` (rlink.h:380) for bundle b: returns (bpcov, predictions dict,
        num_fragments, frag_len); predictions = BundleData::pred in the reference's order."""
        view, keep = bundle_view(arrays, b)
        length = view.end - view.start + 3
        cov = np.zeros(3 * length, dtype=np.float32)
        good = np.zeros(view.n_juncs, dtype=np.float64)
        left = np.zeros(view.n_juncs, dtype=np.float64)
        right = np.zeros(view.n_juncs, dtype=np.float64)
        res = StgBundleResult(ptr(cov), ptr(good), ptr(left), ptr(right))
        res.assemble = 1
        self._ck(self.L.stg_infer_transcripts(self.ctx, C.byref(view), C.byref(res)))
        n, ne = res.n_pred, res.n_exon
        cp = lambda p, k, dt: np.ctypeslib.as_array(p, shape=(k,)).astype(dt, copy=True) if k else np.empty(0, dtype=dt)
        pred = {"start": cp(res.p_start, n, np.int32), "end": cp(res.p_end, n, np.int32), "cov": cp(res.p_cov, n, np.float64),
                "tlen": cp(res.p_tlen, n, np.int32), "geneno": cp(res.p_geneno, n, np.int32), "strand": cp(res.p_strand, n, np.int8),
                "exon_off": cp(res.p_exon_off, n + 1, np.uint32) if n else np.zeros(1, dtype=np.uint32),
                "e_start": cp(res.e_start, ne, np.uint32), "e_end": cp(res.e_end, ne, np.uint32), "e_cov": cp(res.e_cov, ne, np.float32)}
        return cov.reshape(3, length), pred, res.num_fragments, res.frag_len
