"""Rows a8 + a9 without a GPU: the per-bundle code of stringtie_b200/csrc/groups_device.cuh is compiled for the HOST
(tools/groups_hostcheck.cu; development harness, not part of libstgpu) and compared with the oracle field by field —
the reference's loop restated (walk_read) and the kernel's schedule emulated lane by lane (classify_read against a
snapshot, commit of the leading run, steps of 32 and 128 reads).  Needs nvcc; no device."""
import ctypes as C
import os
import shutil
import sys

import numpy as np
import pytest

from stringtie_b200 import synth
from tests import orc
from tests.conftest import load_golden, GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = "/usr/local/cuda/bin/nvcc"
pytestmark = pytest.mark.skipif(not (os.path.exists(NVCC) or shutil.which("nvcc")), reason="nvcc not available")


@pytest.fixture(scope="module")
def hostcheck():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import groups_hostcheck as gh
    L = C.CDLL(gh.build())
    L.hostcheck_groups.argtypes = [C.POINTER(orc.StgPackedBatch), C.c_uint32, C.c_int32, C.c_uint32, C.c_int64, C.c_int64, C.c_int,
                                   C.POINTER(C.c_int64)]
    L.hostcheck_groups.restype = C.c_int
    return L


def _cases():
    from stringtie_b200.stgb import input_view
    for name in GOLDEN:
        yield name, input_view(load_golden(name))
    yield "synth", synth.make_batch(n_bundles=30, reads_per_bundle=1500, seed=5)
    yield "adversarial", synth.make_adversarial(seed=6, n_bundles=12)
    yield "sparse", synth.make_sparse(n_reads=600, seed=2, n_bundles=2)
    for seed in (38, 43):   # the repair appends junction rows and the table is re-sorted
        a = synth.make_adversarial(seed=seed, n_bundles=10, max_reads=250)
        rng = np.random.default_rng(seed)
        a["j_nm"] = np.where(rng.random(a["j_nm"].size) < 0.5, a["j_nreads"], a["j_nm"])
        yield f"appended{seed}", a


@pytest.mark.parametrize("steps", [0, 32, 128, -32, -96])   # < 0: reads that grow a group stay inside the runs
def test_device_code_on_host_equals_oracle(hostcheck, steps):
    for name, a in _cases():
        pb, keep = orc.make_packed(a)
        for bundledist in (50, 0):
            st = (C.c_int64 * 2)(0, 0)
            bad = hostcheck.hostcheck_groups(C.byref(pb), 10, 1, bundledist, 0, int(pb.n_bundles), steps, st)
            assert bad == 0, f"{name}, bundledist {bundledist}, steps {steps}: {bad} bundles differ"
        if steps and name == "synth":
            assert st[0] > 0.8 * st[1], "the parallel path should take most reads of an RNA-seq-like batch"
