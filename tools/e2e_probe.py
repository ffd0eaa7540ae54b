#!/usr/bin/env python3
"""Where the end-to-end step of bench.py goes: upload alone, stages alone, stages beside an upload, stages beside the bpcov copy."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def main():
    base, arrays = bench.build_workload(bench.PARTS, bench.PART_BUNDLES, 0)
    import torch
    from stringtie_b200.abi import make_packed
    from stringtie_b200.engine import Engine
    from stringtie_b200.stgb import INPUT_FIELDS
    from concurrent.futures import ThreadPoolExecutor
    eng = Engine(0)
    pinned = {k: torch.from_numpy(np.ascontiguousarray(arrays[k])).pin_memory() for k in INPUT_FIELDS}
    parr = {k: v.numpy() for k, v in pinned.items()}
    pb, keep = make_packed(parr)
    b0 = eng.upload_packed(pb, keep); eng.upload_wait(b0)
    lane_floats = eng.cov_total_floats(b0) // 3
    cov_host = torch.empty(lane_floats, dtype=torch.float32).pin_memory()

    def stages(b, copy=False):
        eng.run(b, sync=False)
        if copy:
            eng.fetch_bpcov_lane_begin(b, 1, cov_host.data_ptr())
        eng.build_groups(b); eng.neutral_predictions_only(b); eng.assemble(b)
        eng.sync()
        if copy:
            eng.fetch_wait(b)

    stages(b0); b0.free()
    for rep in range(2):
        t = time.perf_counter(); b = eng.upload_packed(pb, keep); t1 = time.perf_counter(); eng.upload_wait(b); t2 = time.perf_counter()
        print(f"upload: call returns after {1e3 * (t1 - t):.1f} ms, data on the device after {1e3 * (t2 - t):.1f} ms")
        t = time.perf_counter(); stages(b); print(f"stages alone: {1e3 * (time.perf_counter() - t):.1f} ms")
        eng.rewind(b)
        t = time.perf_counter(); stages(b, copy=True); print(f"stages + lane-1 copy-out: {1e3 * (time.perf_counter() - t):.1f} ms")
        eng.rewind(b)
        ex = ThreadPoolExecutor(1)
        t = time.perf_counter(); fut = ex.submit(eng.upload_packed, pb, keep); stages(b); t1 = time.perf_counter(); b2 = fut.result(); eng.upload_wait(b2); t2 = time.perf_counter()
        print(f"stages beside an upload: {1e3 * (t1 - t):.1f} ms, the upload done after {1e3 * (t2 - t):.1f} ms")
        eng.rewind(b)
        t = time.perf_counter(); fut = ex.submit(eng.upload_packed, pb, keep); stages(b, copy=True); t1 = time.perf_counter(); b3 = fut.result(); eng.upload_wait(b3); t2 = time.perf_counter()
        print(f"stages + copy-out beside an upload: {1e3 * (t1 - t):.1f} ms, the upload done after {1e3 * (t2 - t):.1f} ms")
        b.free(); b2.free(); b3.free(); ex.shutdown()
    eng.close()


if __name__ == "__main__":
    main()
