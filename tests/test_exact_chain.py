"""Ordered float32 sums in integer form (stringtie_b200/csrc/exact_chain.cuh): on the CPU the shared increment function and
acceptance bounds against the plain loop; on the GPU warp_chain_add itself, through stg_selftest_chain, against numpy's
sequential float32 accumulation."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exact_chain_blocks_match_plain_loop():
    src = os.path.join(ROOT, "tests", "exact_chain_check.cc")
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "exact_chain_check")
        subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-msse2", "-mfpmath=sse", "-o", exe, src])
        out = subprocess.check_output([exe, "30000"], text=True)
    assert " 0 mismatches" in out, out


def _lists(rng, n_lists):
    wts = np.array([1.0, 0.5, 1.0 / 3.0, 0.25, 0.2, 1.0 / 7.0, 0.1, 2.0], dtype=np.float32)
    xs, s0 = [], []
    for i in range(n_lists):
        n = int(rng.choice([0, 1, 5, 31, 32, 33, 255, 256, 257, 700, 5000, 40000, 300000]))
        mode = i % 7
        if mode == 0:
            x = wts[rng.integers(0, wts.size, n)]                                            # coverage deposits
        elif mode == 1:
            x = wts[rng.integers(0, wts.size, n)] * rng.choice(np.array([1, -1], np.float32), n)   # +w / -w at one position
        elif mode == 2:
            x = (wts[rng.integers(0, 5, n)] * rng.integers(1, 300, n).astype(np.float32) * rng.random(n, dtype=np.float32)).astype(np.float32)
        elif mode == 3:
            x = np.ones(n, np.float32)                                                       # walks through every binade
        elif mode == 4:
            x = -wts[rng.integers(0, wts.size, n)]
        elif mode == 5:
            x = rng.standard_normal(n).astype(np.float32) * np.float32(10.0) ** rng.integers(-6, 6, n).astype(np.float32)
        else:
            x = np.where(rng.random(n) < 0.1, np.float32(2.0 ** -13) * rng.integers(1, 9, n).astype(np.float32), wts[rng.integers(0, 4, n)]).astype(np.float32)
        xs.append(x.astype(np.float32))
        s0.append(np.float32(rng.choice([0.0, 1.0, 1024.0, 16777000.0, 123456.7, -50000.3, 0.3333])))
    return xs, np.array(s0, np.float32)


@pytest.mark.gpu
def test_warp_chain_add_matches_sequential_float32():
    from stringtie_b200.engine import Engine
    rng = np.random.default_rng(99)
    xs, s0 = _lists(rng, 120)
    off = np.zeros(len(xs) + 1, np.uint32)
    np.cumsum([x.size for x in xs], out=off[1:])
    eng = Engine(0)
    allx = np.concatenate(xs) if off[-1] else np.zeros(0, np.float32)
    got = eng.selftest_chain(off, allx, s0)
    got_cta = eng.selftest_chain(off, allx, s0, plain=2)     # the form in which a whole CTA shares a list
    eng.close()
    for i, x in enumerate(xs):
        want = np.cumsum(np.concatenate([s0[i:i + 1], x]), dtype=np.float32)[-1]   # ufunc.accumulate: one add after the other
        assert got[i].view(np.uint32) == want.view(np.uint32), (i, x.size, float(got[i]), float(want))
        assert got_cta[i].view(np.uint32) == want.view(np.uint32), (i, x.size, float(got_cta[i]), float(want))
