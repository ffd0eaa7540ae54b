"""The C-ABI library loads, exports every symbol include/stgpu.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from stringtie_b200 import abi, engine, synth
from tests.conftest import ROOT


def _header_functions():
    src = open(os.path.join(ROOT, "include", "stgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(stg_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = engine.load_library()
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"libstgpu.so does not export {n}"
    assert sorted(engine.EXPORTS) == names
    assert L.stg_abi_version() == 5


def test_params_default_match_reference_defaults():
    p = engine.default_params()
    # stringtie.cpp:138-163
    assert (p.junctionsupport, p.sserror, p.junctionthr, p.bundledist) == (10, 25, 1, 50)
    assert (p.mintranscriptlen, p.allowed_nodes) == (200, 1000)
    assert (p.readthr, p.singlethr, p.mcov) == (1.0, 4.75, 1.0)
    assert abs(p.isofrac - 0.01) < 1e-9
    assert (p.trim, p.isunitig, p.includesource) == (1, 1, 1)
    assert C.sizeof(abi.StgParams) == p.struct_size


def test_struct_sizes():
    assert C.sizeof(abi.StgPackedBatch) == 5 * 8 + 25 * 8
    assert C.sizeof(abi.StgBundleView) == 4 * 4 + 21 * 8


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(engine.StgError, match="no CUDA device"):
        engine.Engine(0)


def test_builder_packs_views_identically():
    """Bundle batcher host logic: BundleData-shaped views -> one packed batch == the original packed batch."""
    L = engine.load_library()
    a = synth.make_adversarial(seed=7, n_bundles=9, neg_links=True)
    bld = C.c_void_p()
    assert L.stg_builder_create(None, C.byref(bld)) == 0
    keep = []
    for b in range(a["b_start"].size):
        v, k = abi.bundle_view(a, b)
        keep.append(k)
        assert L.stg_builder_add(bld, C.byref(v)) == 0
    pb = abi.StgPackedBatch()
    assert L.stg_builder_packed(bld, C.byref(pb)) == 0
    assert pb.n_bundles == a["b_start"].size and pb.n_reads == a["r_start"].size
    assert pb.n_segs == a["seg_start"].size and pb.n_pairs == a["pair_idx"].size and pb.n_juncs == a["j_start"].size
    counts = {"b_start": pb.n_bundles, "b_end": pb.n_bundles, "b_read_off": pb.n_bundles + 1,
              "b_junc_off": pb.n_bundles + 1, "r_seg_off": pb.n_reads + 1, "r_pair_off": pb.n_reads + 1,
              "seg_start": pb.n_segs, "seg_end": pb.n_segs, "rjunc_idx": pb.n_segs - pb.n_reads,
              "pair_idx": pb.n_pairs, "pair_cnt": pb.n_pairs}
    for k in abi._PACKED_ORDER:
        n = counts.get(k, pb.n_reads if k.startswith("r_") else pb.n_juncs)
        got = np.ctypeslib.as_array(getattr(pb, k), shape=(n,)) if n else np.zeros(0, dtype=a[k].dtype)
        assert np.array_equal(got, a[k]), k
    assert L.stg_builder_destroy(bld) == 0


def test_builder_rejects_bad_view():
    L = engine.load_library()
    bld = C.c_void_p()
    assert L.stg_builder_create(None, C.byref(bld)) == 0
    v = abi.StgBundleView()
    v.start, v.end, v.n_reads, v.n_juncs = 100, 50, 0, 0
    assert L.stg_builder_add(bld, C.byref(v)) == -1
    assert b"bad view" in L.stg_last_error(None)
    L.stg_builder_destroy(bld)


def test_sass_has_no_contracted_multiply_add_and_uses_bulk_stores():
    """tools/sass_check.py on the built library: every FFMA / DFMA belongs to a division / square-root expansion (the build is
    -fmad=false: a contracted product-sum would change bits against the reference), and cov_fill stores its flat rows with
    cp.async.bulk (UBLKCP in SASS)."""
    import subprocess, sys, shutil
    if shutil.which("cuobjdump") is None:
        import pytest
        pytest.skip("cuobjdump not on PATH")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_check.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "cov_fill_kernel" in r.stdout.split("bulk copies (UBLKCP):")[1].splitlines()[0]
