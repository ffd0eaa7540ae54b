#!/usr/bin/env python3
"""prepare: cut the deepest bundle (or the k-th deepest) out of the bench workload into /tmp/stg_deepest.npz;
run: push that one bundle through run + build_groups (+ neutral + assemble with --asm) — the command ncu wraps."""
import sys, os, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mode = sys.argv[1]
path = "/tmp/stg_deepest.npz"
if mode == "prepare":
    import bench
    from stringtie_b200 import shard
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    base, a = bench.build_workload(bench.PARTS, bench.PART_BUNDLES, 0)
    nr = np.diff(a["b_read_off"].astype(np.int64))
    order = np.argsort(-nr)
    sub = shard.take_bundles(a, order[k:k + 1])
    np.savez(path, **sub)
    print("bundle", int(order[k]), "reads", int(nr[order[k]]), "part", int(order[k]) // bench.PART_BUNDLES)
else:
    from stringtie_b200.engine import Engine
    with np.load(path) as z:
        sub = {k: z[k] for k in z.files}
    eng = Engine(0)
    eng.set_profiling(True)
    b = eng.upload(sub); eng.run(b); eng.build_groups(b)
    if "--asm" in sys.argv:
        eng.neutral_predictions_only(b); eng.assemble(b)
    eng.sync()
    print(json.dumps({n: ms for n, ms in eng.kernel_times()}))
