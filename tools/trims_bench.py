#!/usr/bin/env python3
"""Throughput of the find_trims kernel on the bundlenodes (and their halves) of one seeded C2 base (6000 loci):
regions come from the oracle's grouping stage; timing from CUDA events around the kernel (stg_set_profiling)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from stringtie_b200.engine import Engine
from tests import orc
from tests.test_gpu_parity import _regions_from_oracle

base, _ = bench.build_workload(1, bench.BASE_BUNDLES, 0)
t = time.time()
cov, off, good, left, right = orc.run(base, threads=0)
bun, sno, st, en = _regions_from_oracle(base, cov, off, good, left, right)
t_orc = time.time() - t
bases = int((en.astype(np.int64) - st + 1).sum())
eng = Engine(0)
batch = eng.upload(base)
eng.run(batch)
eng.set_profiling(True)
res = None
times = []
for _ in range(4):
    res = eng.find_trims(batch, bun, sno, st, en)
    times.append([ms for name, ms in eng.kernel_times() if name == "find_trims"][-1])
t = time.time()
want = orc.find_trims(base, cov, off, bun, sno, st, en)
t_cpu = time.time() - t
ok = all(np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a, b.view(np.uint32) if b.dtype == np.float32 else b)
         for a, b in zip(res, want))
print(json.dumps({"regions": int(bun.size), "bases": bases, "trimpoints": int(res[1].size), "kernel_ms": times,
                  "gbases_per_s": bases / (min(times) * 1e-3) / 1e9, "oracle_1thread_s": t_cpu, "equal_to_oracle": bool(ok)}))
