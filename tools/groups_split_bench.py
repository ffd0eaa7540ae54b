#!/usr/bin/env python3
"""Where does the group stage's time go: the deepest bundle alone, the very deep class, the deep class, the shallow mass."""
import sys, os, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from stringtie_b200 import shard
from stringtie_b200.engine import Engine
base, a = bench.build_workload(bench.PARTS, bench.PART_BUNDLES, 0)
nr = np.diff(a["b_read_off"].astype(np.int64))
order = np.argsort(-nr)
eng = Engine(0)
eng.set_profiling(True)
def run(name, idx):
    sub = shard.take_bundles(a, np.sort(idx))
    for rep in range(2):
        b = eng.upload(sub); eng.run(b); eng.build_groups(b); eng.sync()
        k = {n: ms for n, ms in eng.kernel_times() if n.startswith("groups")}
        b.free()
    print(json.dumps({"case": name, "bundles": int(idx.size), "reads": int(nr[idx].sum()), "kernel_ms": k}))
run("deepest", order[:1])
run("second", order[1:2])
run("top63", order[:63])
run("deep_64_to_935", order[63:935])
run("shallow", order[935:])
run("all_but_deepest", order[1:])
