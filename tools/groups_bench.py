#!/usr/bin/env python3
"""Time of the rows a8 + a9 stage (stg_build_groups) on the bench workload: one seeded C2 base (6000 loci, ~10^7 reads)
replicated `--replicas` times; kernel times from CUDA events (stg_set_profiling); the oracle's single-thread time on a
sample of the base beside it; a parity check of a sample of bundles against the oracle."""
import argparse, os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from stringtie_b200 import shard
from stringtie_b200.engine import Engine
from tests import orc, a9_util

ap = argparse.ArgumentParser()
ap.add_argument("--replicas", type=int, default=1)
ap.add_argument("--check", type=int, default=200, help="bundles of the base compared with the oracle")
ap.add_argument("--prewarm", type=int, default=0, help="passes of stg_run before the stage (the state bench.py measures it in)")
ap.add_argument("--torch", action="store_true", help="initialise torch.cuda first, as bench.py does")
args = ap.parse_args()
if args.torch:
    import torch
    torch.zeros(1, device="cuda")
base, work = bench.build_workload(args.replicas, bench.BASE_BUNDLES, 0)
eng = Engine(0)
eng.set_profiling(True)
res = {"replicas": args.replicas, "bundles": int(work["b_start"].size), "reads": int(work["r_start"].size)}
batch = eng.upload(work)
eng.run(batch)
for _ in range(args.prewarm):
    eng.run(batch)
eng.sync()
t = time.time()
eng.build_groups(batch)
eng.sync()
res["build_groups_wall_ms"] = (time.time() - t) * 1e3
res["kernel_ms"] = {n: ms for n, ms in eng.kernel_times() if n.startswith("groups")}
G = {k: eng.fetch_groups_array(batch, k) for k in ("GROUP_OFF", "COLOR_OFF", "APP_OFF", "RG_OFF", "N_BUN", "N_BN")}
res["groups"] = int(G["GROUP_OFF"][-1]); res["colors"] = int(G["COLOR_OFF"][-1]); res["appended_junctions"] = int(G["APP_OFF"][-1])
res["readgroup_entries"] = int(G["RG_OFF"][-1]); res["bundles_out"] = int(G["N_BUN"].sum()); res["bundlenodes"] = int(G["N_BN"].sum())
res["max_reads_in_bundle"] = int(np.diff(work["b_read_off"].astype(np.int64)).max())
batch.free()
if args.check:
    nb = int(base["b_start"].size)
    rng = np.random.default_rng(1)
    nreads = np.diff(base["b_read_off"].astype(np.int64))
    pick = np.unique(np.concatenate([rng.choice(nb, size=min(args.check, nb), replace=False), np.argsort(nreads)[-3:]]))
    sub = shard.take_bundles(base, pick)
    b2 = eng.upload(sub)
    eng.run(b2)
    eng.build_groups(b2)
    got = a9_util.split_engine_groups(eng.fetch_groups(b2), sub)
    b2.free()
    cov, off, good, left, right = orc.run(sub, threads=0)
    t = time.time()
    want = orc.groups(sub, cov, off, good, left, right)
    res["oracle_sample_s"] = time.time() - t
    res["oracle_sample_reads"] = int(sub["r_start"].size)
    a9_util.assert_same(got, want)
    res["sample_equal_to_oracle"] = True
print(json.dumps(res))
