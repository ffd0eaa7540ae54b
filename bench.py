#!/usr/bin/env python3
"""bench.py — throughput of the per-bundle assembly hot path (the stages built so far: the reference's
count_good_junctions = coverage + junction support, and the junction pass that opens build_graphs) on the C2 workload
of BASELINE.json:
synthetic 100 M 2x150 paired-end spliced alignments over ~60 k loci.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one rank per GPU under torchrun for N>1)
  python bench.py --impl reference [--gpus N] [--steps K] ...    the reference's own CPU code on the host cores

One "step" = one pass of the hot path over the whole batch.  `value` is measured with the batch resident in HBM,
`e2e` through the C ABI from pinned HOST buffers (H2D of all inputs + D2H of the stage's results inside the timed
region).  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = ("aligned reads/sec through assembly (stages built: count_good_junctions = coverage + junction support, "
          "and the junction pass of build_graphs)")
BASE_BUNDLES = 6000          # one seeded base of ~9.7 M alignments ...
BASE_READS_PER_BUNDLE = 1667
REPLICAS = 10                # ... x 10 coordinate-shifted replicas = ~97 M alignments over 60 000 loci
SEED = 20251019


def build_workload(replicas: int, base_bundles: int, rank: int):
    from stringtie_b200 import synth
    cache = os.path.join(tempfile.gettempdir(), f"stg_c2_base_{base_bundles}_{SEED + rank}.npz")
    if os.path.exists(cache):
        with np.load(cache) as z:
            base = {k: z[k] for k in z.files}
    else:
        base = synth.make_batch(n_bundles=base_bundles, reads_per_bundle=BASE_READS_PER_BUNDLE, seed=SEED + rank)
        try:
            np.savez(cache, **base)
        except Exception:
            pass
    return base, synth.replicate(base, replicas)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def reference_hot(arrays, nbundles: int, threads: int, reps: int):
    """Time the reference's own count_good_junctions on the first `nbundles` bundles via oracle/_ref/ref_hot
    (kind 'reference'); if that binary is absent fall back to the C port (kind 'port')."""
    from stringtie_b200.stgb import slice_bundles, write_stgb
    nb = min(nbundles, arrays["b_start"].size)
    sample = slice_bundles(arrays, 0, nb)
    nreads = int(sample["r_start"].size)
    tool = os.path.join(ROOT, "oracle", "_ref", "ref_hot")
    desc = f"first {nb} bundles ({nreads} alignments) of the workload x {reps} reps"
    if os.path.exists(tool):
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "sample.stgb")
            write_stgb(path, sample)
            out = subprocess.check_output([tool, path, "--threads", str(threads), "--reps", str(reps), "--a7"], text=True)
            out9 = subprocess.check_output([tool, path, "--threads", str(threads), "--reps", str(reps), "--a9b"], text=True)
            out19 = subprocess.check_output([tool, path, "--threads", str(threads), "--reps", str(reps), "--a19"], text=True)
        j = json.loads(out.strip().splitlines()[-1])
        j9 = json.loads(out9.strip().splitlines()[-1])
        j19 = json.loads(out19.strip().splitlines()[-1])
        return {"value": j["reads_per_s"], "unit": "reads/s", "cores": threads, "kind": "reference",
                "through_rows_a8_a9": {"reads_per_s": j9["reads_per_s"], "compute_s": j9["compute_s_max"],
                                       "note": "same sample, the reference run on to the end of its bundle creation (rlink.cpp:15433)"},
                "through_row_a19": {"reads_per_s": j19["reads_per_s"], "compute_s": j19["compute_s_max"],
                                    "note": "... and on to the end of its neutral-bundle predictions (rlink.cpp:15637)"},
                "sample": desc + "; reference count_good_junctions (rlink.cpp:16656) + build_graphs up to the end of its "
                          "junction pass (rlink.cpp:14165-14842) on rebuilt BundleData, "
                          "aggregate of per-thread reads / per-thread compute time",
                "bundles_per_s": j["reads_per_s"] * nb / max(nreads, 1), "compute_s": j["compute_s_max"]}
    from tests import orc
    t0 = time.perf_counter()
    for _ in range(reps):
        orc.run(sample, want_cov=True, threads=threads)
    dt = (time.perf_counter() - t0) / reps
    return {"value": nreads / dt, "unit": "reads/s", "cores": threads, "kind": "port", "sample": desc + "; oracle/liborc.so",
            "bundles_per_s": nb / dt, "compute_s": dt}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--replicas", type=int, default=REPLICAS)
    ap.add_argument("--base-bundles", type=int, default=BASE_BUNDLES)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-groups", action="store_true", help="do not run the rows a8 + a9 stage after the timed passes")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload_name = (f"C2 synthetic short-read: {args.replicas} x seeded base of {args.base_bundles} loci "
                     f"(~{args.replicas * args.base_bundles * BASE_READS_PER_BUNDLE / 1e6:.0f} M alignments, 2x150 PE, de novo)")
    config = {"workload": workload_name, "per_gpu": "one full copy of the workload per rank (weak scaling)",
              "l2": "inputs_larger_than_L2", "stages": ["cov_edge_add/add_read_to_cov", "junction support",
                                                         "blocked clamped scan", "get_cov totals",
                                                         "build_graphs junction pass + good_junc"],
              "scope": "value / e2e / roofline cover SURVEY rows a1-a7 (the reference arm times the same rows); rows a8 + a9 "
                       "(read loop, groups, bundles) and a19 (neutral predictions) are built too and reported beside "
                       "them under rows_a8_a9 (one pass outside the timed region; pipeline_a1_to_a9_reads_per_s)"}

    # ---------------- reference arm: the reference's own CPU code, all host threads, bounded sample ----------
    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        base, _ = build_workload(1, args.base_bundles, 0)
        # a step = one pass over a bounded sample sized for ~seconds of CPU work
        nb = min(2000, base["b_start"].size)
        for _ in range(args.warmup):
            reference_hot(base, nb, threads, 1)
        vals = [reference_hot(base, nb, threads, 1) for _ in range(max(1, args.steps))]
        v = statistics.median(x["value"] for x in vals)
        cb = vals[0]
        cb["value"] = v
        line = {"metric": METRIC, "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * statistics.median(x["compute_s"] for x in vals),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "impl": "reference", "cpu_baseline": cb,
                "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ---------------- our arm ----------------------------------------------------------------------------------
    import torch
    from stringtie_b200 import synth
    from stringtie_b200.abi import make_packed
    from stringtie_b200.engine import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    base, arrays = build_workload(args.replicas, args.base_bundles, rank)
    n_reads = int(arrays["r_start"].size)
    n_bundles = int(arrays["b_start"].size)
    in_bytes = synth.total_input_bytes(arrays)

    eng = Engine(local_rank)
    stream = torch.cuda.ExternalStream(eng.L.stg_stream(eng.ctx), device=torch.device("cuda", local_rank))
    totals_dev = torch.zeros(3, dtype=torch.float64, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-input measurement -------------------------------------------------------------------------
    batch = eng.upload(arrays)

    def step_resident():
        eng.run(batch, sync=False)
        if dist is not None:  # the path's only collective: the global totals needed for TPM normalisation
            eng.sync()
            t = eng.fetch_cov_totals(batch)
            totals_dev.copy_(torch.from_numpy(np.array([t[:, 1].sum(), float(n_reads), float(n_bundles)])))
            dist.all_reduce(totals_dev)

    # clocks / throttle reasons are sampled under this workload: the sampler (nvidia-smi -lms 100) needs a few hundred
    # ms to deliver its first line on a multi-GPU box, longer than the timed region, so it starts before the warm-up
    # and, if it still has too few lines afterwards, untimed passes of the same workload keep the load up for it
    sampler = ClockSampler(local_rank)
    for _ in range(args.warmup):
        step_resident()
    eng.set_profiling(True)
    barrier()
    l0 = eng.launch_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ktimes = {}
    ev0.record(stream)
    for _ in range(args.steps):
        step_resident()
        eng.sync()
        for name, ms in eng.kernel_times():
            ktimes.setdefault(name, []).append(ms)
    ev1.record(stream)
    barrier()
    launches = eng.launch_counter() - l0
    ms_total = ev0.elapsed_time(ev1)
    eng.set_profiling(False)
    t_extra, n_extra = time.perf_counter(), 0
    while len(sampler.rows) < 5 and time.perf_counter() - t_extra < 3.0:   # untimed: same kernels, for the clock sampler
        eng.run(batch, sync=True)                                           # (no collective: ranks loop independently)
        n_extra += 1
    clocks = sampler.stop()
    clocks["window"] = "warm-up + timed passes" + (f" + {n_extra} untimed passes of the same workload" if n_extra else "")
    if dist is not None:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    L, S, J = eng.stats(batch)
    kmean = {k: statistics.mean(v) for k, v in ktimes.items()}
    # ---- the next rows (SURVEY §8 a8 + a9: read repair, groups, colours, bundles / bundlenodes), measured once on
    # the resident batch, outside the timed region: reported beside the headline, not part of it
    groups_stage = None
    if not args.skip_groups:
        try:
            eng.set_profiling(True)
            eng.sync()
            t0 = time.perf_counter()
            eng.build_groups(batch)
            eng.sync()
            wall = time.perf_counter() - t0
            gk = {n: ms for n, ms in eng.kernel_times() if n.startswith("groups")}
            gms = sum(gk.values())
            groups_stage = {"rows": "a8 + a9 (stg_build_groups)", "kernel_ms": gk, "ms": gms, "wall_ms": 1e3 * wall,
                            "reads_per_s": n_reads / (gms * 1e-3) if gms else None,
                            "groups": int(eng.fetch_groups_array(batch, "GROUP_OFF")[-1]),
                            "bundlenodes": int(eng.fetch_groups_array(batch, "N_BN").sum()),
                            "note": "per GPU, one pass, outside the timed region; wall includes first-use cudaMalloc of the stage's buffers"}
            # row a19 (de novo part): single-exon predictions of the neutral bundles, from the device-resident bundlenodes
            eng.sync()
            t0 = time.perf_counter()
            npred = eng.neutral_predictions(batch)
            eng.sync()
            nk = {n: ms for n, ms in eng.kernel_times() if n.startswith("neutral") or n == "find_trims"}
            groups_stage["row_a19_neutral_predictions"] = {"kernel_ms": nk, "ms": sum(nk.values()),
                                                           "wall_ms": 1e3 * (time.perf_counter() - t0),
                                                           "predictions": int(npred["start"].size)}
        except Exception as ex:   # the headline must not depend on the newer stage
            groups_stage = {"error": str(ex)}
    eng.set_profiling(False)
    batch.free()

    # ---- end-to-end through the C ABI from pinned host buffers ------------------------------------------------
    e2e = None
    if not args.skip_e2e:
        from stringtie_b200.stgb import INPUT_FIELDS
        pinned = {}
        for k in INPUT_FIELDS:
            tns = torch.from_numpy(np.ascontiguousarray(arrays[k])).pin_memory()
            pinned[k] = tns
        parr = {k: v.numpy() for k, v in pinned.items()}
        pb, keep = make_packed(parr)
        d2h_bytes = (3 * 8 + 1 + 5 * 8 + 4 + 2) * int(pb.n_juncs) + 3 * 8 * n_bundles

        def step_e2e():
            b = eng.upload_packed(pb, keep)
            eng.run(b, sync=False)
            eng.fetch_junction_support(b)
            eng.fetch_junction_pass(b)
            eng.fetch_cov_totals(b)
            b.free()

        for _ in range(max(1, min(args.warmup, 2))):
            step_e2e()
        barrier()
        n_e2e = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(n_e2e):
            step_e2e()
        ev1.record(stream)
        barrier()
        wall = (time.perf_counter() - t0) / n_e2e
        if dist is not None:
            t = torch.tensor([wall], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall = float(t.item())
        e2e = {"value": world * n_reads / wall, "unit": "reads/s", "h2d_bytes_per_step": in_bytes,
               "d2h_bytes_per_step": d2h_bytes, "ms_per_step": 1e3 * wall, "steps": n_e2e,
               "timing": "host wall clock around stg_upload (H2D of every input array from pinned memory) + stg_run + D2H of "
                         "junction support, the junction-pass table and coverage totals + stg_free_batch (device blocks "
                         "recycled by the context pool), max over ranks; bpcov stays in HBM for the stages that follow. "
                         "(Streaming the workload as 5 batches over 2 contexts/threads measured 120 ms: PCIe is the "
                         "bound, splitting only adds fixed costs.)"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    t_tile = sum(kmean.get(k, 0.0) for k in ("deep_deposit", "cov_tile", "cov_fill")) or float("nan")
    alg_tile = 24.0 * L + 13.0 * S          # bytes the dominant kernel's work amounts to
    traffic = None
    try:  # DRAM bytes of the same kernels from the committed ncu --set full capture (only valid for the default workload)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if args.replicas == REPLICAS and args.base_bundles == BASE_BUNDLES:
            traffic = float(tj["traffic_bytes_per_pass"])
    except Exception:
        pass
    alg_stage = alg_tile + 8.0 * J
    t_stage = sum(kmean.values())
    roofline = {"bound": "hbm", "kernel": "deep_deposit + cov_tile + cov_fill (ordered deposit, recurrences and the bpcov write)", "achieved": alg_tile / (t_tile * 1e-3) / 1e9, "peak": peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 (B200_PROFILING.md)",
                "unit": "GB/s", "frac": alg_tile / (t_tile * 1e-3) / 1e9 / peak, "traffic": traffic,
                "algorithmic_bytes": alg_tile, "formula": "24*L + 13*S (SURVEY.md §8d; the 8*J term belongs to expand_count)",
                "L": L, "S": S, "J": J, "kernel_ms": t_tile,
                "stage": {"achieved": alg_stage / (t_stage * 1e-3) / 1e9, "frac": alg_stage / (t_stage * 1e-3) / 1e9 / peak,
                          "ms": t_stage, "formula": "24*L + 13*S + 8*J over all kernels of the stage"}}

    cpu_baseline = None
    if not args.skip_cpu_baseline:
        cpu_baseline = reference_hot(base, 2000, os.cpu_count() or 1, 3)

    line = {"metric": METRIC, "value": world * n_reads / (ms_step * 1e-3), "unit": "reads/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config, "bundles_per_sec": world * n_bundles / (ms_step * 1e-3), "reads_per_gpu": n_reads,
            "bundles_per_gpu": n_bundles, "gpu_launches": launches, "clocks": clocks, "e2e": e2e, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "kernel_ms": kmean, "rows_a8_a9": groups_stage}
    if groups_stage and groups_stage.get("ms"):
        # everything built so far in one figure: rows a1-a9 resident (the headline `value` stays the stages of round 1)
        groups_stage["pipeline_a1_to_a9_reads_per_s"] = world * n_reads / ((ms_step + groups_stage["ms"]) * 1e-3)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
