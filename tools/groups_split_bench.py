#!/usr/bin/env python3
"""Where does the group stage's time go: single deep bundles alone, the classes of launch_groups_walk, the whole batch;
optionally for several cluster thresholds (STGPU_CLUSTER_READS is read when a context is created).
usage: groups_split_bench.py [threshold ...]"""
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
print(json.dumps({"top_reads": nr[order[:40]].tolist(), "ge_100k": int((nr >= 100000).sum()), "ge_200k": int((nr >= 200000).sum()),
                  "ge_50k": int((nr >= 50000).sum()), "ge_16k": int((nr >= 16384).sum())}))
def run(eng, name, idx, tag):
    sub = shard.take_bundles(a, np.sort(idx)) if idx is not None else a
    for rep in range(2):
        b = eng.upload(sub); eng.run(b); eng.build_groups(b); eng.sync()
        k = {n: round(ms, 2) for n, ms in eng.kernel_times() if n.startswith("groups")}
        b.free()
    print(json.dumps({"threshold": tag, "case": name, "bundles": int(nr.size if idx is None else idx.size), "kernel_ms": k}), flush=True)
ths = [int(x) for x in sys.argv[1:]] or [0]
for th in ths:
    if th:
        os.environ["STGPU_CLUSTER_READS"] = str(th)
    eng = Engine(0)
    eng.set_profiling(True)
    if th == ths[0]:
        for i in range(6):
            run(eng, f"bundle_rank{i}_reads{int(nr[order[i]])}", order[i:i + 1], th)
    run(eng, "all", None, th)
    eng.close()
