#!/usr/bin/env python3
"""Fuzz of rows a8 + a9 + a19 on the GPU against the oracle: many seeded adversarial / RNA-seq-like / sparse batches, every
field of every bundle compared (tests/a9_util.assert_same).  usage: tools/groups_fuzz.py [first_seed] [count]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stringtie_b200 import synth
from stringtie_b200.engine import Engine
from tests import orc, a9_util

first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
count = int(sys.argv[2]) if len(sys.argv) > 2 else 40
eng = Engine(0)
t0 = time.time()
nb = nr = 0
for seed in range(first, first + count):
    kind = seed % 4
    if kind == 0:
        a = synth.make_adversarial(seed=seed, n_bundles=14)
    elif kind == 1:
        a = synth.make_adversarial(seed=seed, n_bundles=10, max_reads=250)
        rng = np.random.default_rng(seed)
        a["j_nm"] = np.where(rng.random(a["j_nm"].size) < 0.5, a["j_nreads"], a["j_nm"])   # demotions, appended junctions
    elif kind == 2:
        a = synth.make_batch(n_bundles=25, reads_per_bundle=int(400 + 300 * (seed % 7)), seed=seed)
    else:
        a = synth.make_batch(n_bundles=3, reads_per_bundle=25000, seed=seed)                 # deep: the 4-warp kernel
    cov, off, good, left, right = orc.run(a)
    want = orc.groups(a, cov, off, good, left, right)
    wantn = orc.neutral_preds(a, cov, off, good, left, right)
    b = eng.upload(a)
    try:
        eng.run(b); eng.build_groups(b)
        got = a9_util.split_engine_groups(eng.fetch_groups(b), a)
        N = eng.neutral_predictions(b)
    finally:
        b.free()
    a9_util.assert_same(got, want)
    noff = N["off"].astype(np.int64)
    for k, w in enumerate(wantn):
        for key in ("start", "end", "cov", "tlen", "geneno"):
            assert np.array_equal(N[key][noff[k]:noff[k + 1]], w[key].astype(N[key].dtype)), (seed, k, key)
    nb += len(want); nr += int(a["r_start"].size)
print(f"fuzz ok: seeds {first}..{first + count - 1}, {nb} bundles, {nr} reads, {time.time() - t0:.0f} s")
