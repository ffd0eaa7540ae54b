#!/usr/bin/env python3
"""Small case for profiling the rows a8 + a9 walk under ncu: the deepest bundles of the seeded C2 base plus a sample."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from stringtie_b200 import shard
from stringtie_b200.engine import Engine

base, _ = bench.build_workload(1, bench.BASE_BUNDLES, 0)
nreads = np.diff(base["b_read_off"].astype(np.int64))
pick = np.unique(np.concatenate([np.argsort(nreads)[-2:], np.arange(0, base["b_start"].size, 40)]))
sub = shard.take_bundles(base, pick)
eng = Engine(0)
eng.set_profiling(True)
b = eng.upload(sub)
eng.run(b)
eng.build_groups(b)
eng.sync()
print({n: ms for n, ms in eng.kernel_times() if n.startswith("groups")}, int(sub["r_start"].size), "reads", int(nreads.max()), "deepest")
