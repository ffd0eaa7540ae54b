#!/usr/bin/env python3
"""Stage timings of the whole pipeline (run -> build_groups -> neutral -> assemble) on the bench workload or a slice of it."""
import argparse, json, sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from stringtie_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--replicas", type=int, default=1)
ap.add_argument("--base-bundles", type=int, default=6000)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
base, arrays = bench.build_workload(a.replicas, a.base_bundles, 0)
eng = Engine(0)
eng.set_profiling(True)
for rep in range(a.reps):
    batch = eng.upload(arrays)
    t0 = time.perf_counter(); eng.run(batch); eng.sync(); t1 = time.perf_counter()
    k_run = dict(eng.kernel_times())
    eng.build_groups(batch); eng.sync(); t2 = time.perf_counter()
    k_grp = {k: v for k, v in eng.kernel_times() if k.startswith("groups")}
    eng.neutral_predictions(batch); eng.sync(); t3 = time.perf_counter()
    eng.assemble(batch); eng.sync(); t4 = time.perf_counter()
    k_asm = {k: v for k, v in eng.kernel_times() if k.startswith("asm")}
    pr = eng.fetch_predictions(batch)
    out = {"reads": int(arrays["r_start"].size), "bundles": int(arrays["b_start"].size), "wall_ms": {"run": 1e3 * (t1 - t0), "groups": 1e3 * (t2 - t1), "neutral": 1e3 * (t3 - t2), "assemble": 1e3 * (t4 - t3)},
           "run_ms": sum(k_run.values()), "groups_ms": sum(k_grp.values()), "asm_kernel_ms": k_asm, "asm_ms": sum(k_asm.values()),
           "graphs": int(eng.fetch_asm_array(batch, "G_BUNDLE").size), "nodes": int(eng.fetch_asm_array(batch, "N_START").size),
           "transfrags": int(eng.fetch_asm_array(batch, "G_NTF").sum()), "predictions": int(pr["start"].size), "exons": int(pr["e_start"].size)}
    print(json.dumps(out))
    batch.free()
