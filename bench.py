#!/usr/bin/env python3
"""bench.py — throughput of the per-bundle assembly hot path, infer_transcripts() (rlink.h:380): coverage + junction
support (count_good_junctions), then build_graphs to its end — junction pass, read repair, groups / bundles, neutral
predictions, splice graphs, read -> transfrag mapping, transfrag processing, flow decomposition, predictions — on the C2
workload of BASELINE.json: synthetic ~100 M 2x150 paired-end spliced alignments over 60 000 independently drawn loci.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--scaling weak|strong]   our arm (one rank per GPU under torchrun)
  python bench.py --impl reference [--gpus N] [--steps K] ...                   the reference's own CPU code, all host cores

One "step" = one pass of the whole path over the whole batch (every SURVEY §8 row built so far, a1-a15 + a19, is inside the
timed region).  `value` is measured with the batch resident in HBM; `e2e` through the C ABI from pinned HOST buffers:
H2D of all inputs + every stage + D2H of what the host's printResults needs (BundleData::pred, num_fragments, frag_len
and bpcov) inside the timed region.  Prints ONE JSON line (rank 0).
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

METRIC = "aligned reads/sec through assembly (infer_transcripts: count_good_junctions + build_graphs, short-read de novo)"
PART_BUNDLES = 6000          # loci per independently seeded part
READS_PER_BUNDLE = 1667
PARTS = 10                   # 10 parts x 6 000 loci = 60 000 loci, ~97 M alignments
SEED = 20251019
BASE_BUNDLES, REPLICAS = PART_BUNDLES, PARTS   # names the tests / tools of round 1 use


def _make_part(args):
    nb, seed, path = args
    from stringtie_b200 import synth
    a = synth.make_batch(n_bundles=nb, reads_per_bundle=READS_PER_BUNDLE, seed=seed)
    try:
        np.savez(path + ".tmp.npz", **a)
        os.replace(path + ".tmp.npz", path)
    except Exception:
        pass
    return path


def build_workload(parts: int, part_bundles: int, rank: int):
    """`parts` independently drawn sets of loci (seed SEED + 1000*rank + i), concatenated.  Parts are cached under the
    temp dir; missing ones are generated in parallel processes (call this before CUDA is initialised)."""
    from stringtie_b200 import synth
    paths = [os.path.join(tempfile.gettempdir(), f"stg_c2_part_{part_bundles}_{SEED + 1000 * rank + i}.npz") for i in range(parts)]
    todo = [(part_bundles, SEED + 1000 * rank + i, p) for i, p in enumerate(paths) if not os.path.exists(p)]
    if todo and rank == 0 and int(os.environ.get("LOCAL_RANK", "0")) != 0:
        # every rank of a multi-rank run assembles the SAME workload (seed of rank 0): local rank 0 generates the parts (atomic
        # rename), the others wait for the files instead of generating them again
        t_wait = time.time()
        while any(not os.path.exists(p) for p in paths) and time.time() - t_wait < 900:
            time.sleep(0.5)
        todo = [(part_bundles, SEED + i, p) for i, p in enumerate(paths) if not os.path.exists(p)]
    if todo:
        import multiprocessing as mp
        nproc = max(1, min(len(todo), os.cpu_count() or 2))   # (the other local ranks wait for these files)
        with mp.get_context("fork").Pool(nproc) as pool:
            pool.map(_make_part, todo)
    loaded = []
    for i, p in enumerate(paths):
        if os.path.exists(p):
            with np.load(p) as z:
                loaded.append({k: z[k] for k in z.files})
        else:   # the cache could not be written: generate in-process
            loaded.append(synth.make_batch(n_bundles=part_bundles, reads_per_bundle=READS_PER_BUNDLE, seed=SEED + 1000 * rank + i))
    return loaded[0], synth.concat(loaded)


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


def reference_hot(arrays, nbundles: int, threads: int, reps: int, stages=("--full",)):
    """Time the reference's own infer_transcripts body (count_good_junctions + build_graphs, rlink.cpp:17256-17258) on the
    first `nbundles` bundles via oracle/_ref/ref_hot (kind 'reference': the reference's sources compiled where they lie,
    BundleData rebuilt from the packed batch outside the timed part)."""
    from stringtie_b200.stgb import slice_bundles, write_stgb
    nb = min(nbundles, arrays["b_start"].size)
    sample = slice_bundles(arrays, 0, nb)
    nreads = int(sample["r_start"].size)
    tool = os.path.join(ROOT, "oracle", "_ref", "ref_hot")
    if not os.path.exists(tool):
        raise SystemExit("bench.py: oracle/_ref/ref_hot is missing (built by __graft_entry__.build() where /root/reference exists)")
    desc = f"first {nb} bundles ({nreads} alignments) of the workload x {reps} reps"
    res = {}
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "sample.stgb")
        write_stgb(path, sample)
        for st in stages:
            out = subprocess.check_output([tool, path, "--threads", str(threads), "--reps", str(reps), st], text=True)
            res[st] = json.loads(out.strip().splitlines()[-1])
    j = res["--full"]
    cb = {"value": j["reads_per_s"], "unit": "reads/s", "cores": threads, "kind": "reference",
          "sample": desc + "; the reference's count_good_junctions (rlink.cpp:16656) + build_graphs (rlink.cpp:14165) to its end "
                    "on rebuilt BundleData, aggregate of per-thread reads / per-thread compute time",
          "bundles_per_s": j["reads_per_s"] * nb / max(nreads, 1), "compute_s": j["compute_s_max"]}
    for st, key in (("--a7", "through_rows_a1_a7"), ("--a9b", "through_rows_a8_a9"), ("--a10", "through_row_a10"), ("--a11", "through_row_a11")):
        if st in res:
            cb[key] = {"reads_per_s": res[st]["reads_per_s"], "compute_s": res[st]["compute_s_max"]}
    return cb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank assembles its own C2-sized workload; strong: ONE workload is split across the ranks "
                         "by stg_shard_assign (greedy least-loaded by reads + span/64)")
    ap.add_argument("--parts", type=int, default=PARTS)
    ap.add_argument("--part-bundles", type=int, default=PART_BUNDLES)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--e2e-no-bpcov", action="store_true", help="leave bpcov in HBM in the end-to-end measurement")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload_name = (f"C2 synthetic short-read: {args.parts} x {args.part_bundles} independently drawn loci "
                     f"(~{args.parts * args.part_bundles * READS_PER_BUNDLE / 1e6:.0f} M alignments, 2x150 PE, de novo)")
    config = {"workload": workload_name,
              "per_gpu": ("every rank assembles the same full C2 workload: equal work per GPU (weak scaling)" if args.scaling == "weak" else
                          "ONE workload split across the ranks by stg_shard_assign (strong scaling)"),
              "l2": "inputs_larger_than_L2",
              "stages": ["a1-a5 coverage + junction support (count_good_junctions)", "a7 junction pass + good_junc",
                         "a8 + a9 read repair, groups, colours, bundles", "a19 neutral single-exon predictions",
                         "a10 create_graph", "a11 read -> transfrag", "a12 process_transfrags",
                         "a13-a15 find_transcripts / push_max_flow / store_transcript"],
              "scope": "value / e2e cover the whole of infer_transcripts (rows a1-a15 + a19, short-read de novo); the reference arm "
                       "times the reference's own count_good_junctions + build_graphs on the same bundles"}

    # ---------------- reference arm: the reference's own CPU code, all host threads, bounded sample ----------
    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        base, _ = build_workload(1, args.part_bundles, 0)
        nb = min(2000, base["b_start"].size)   # a step = one pass over a bounded sample sized for ~seconds of CPU work
        for _ in range(args.warmup):
            reference_hot(base, nb, threads, 1)
        vals = [reference_hot(base, nb, threads, 1) for _ in range(max(1, args.steps))]
        v = statistics.median(x["value"] for x in vals)
        cb = vals[0]
        cb["value"] = v
        line = {"metric": METRIC, "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * statistics.median(x["compute_s"] for x in vals),
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "impl": "reference", "cpu_baseline": cb,
                "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ---------------- our arm ----------------------------------------------------------------------------------
    # the workload first (worker processes are forked: no CUDA context may exist yet)
    # weak scaling: every rank assembles the same C2 workload (equal work per GPU: the stage's time is the latency of its deepest
    # bundle, and independently drawn workloads differ by 2x in that — measured 231 against 439 ms for two seeds —, which the
    # all-reduce of the totals would turn into the step time of every rank); strong scaling splits that one workload
    base, arrays = build_workload(args.parts, args.part_bundles, 0)
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
    shard_info = None
    if args.scaling == "strong" and world > 1:
        from stringtie_b200 import shard as shard_mod
        idx, cost = shard_mod.shard(arrays, world)
        total_reads = int(arrays["r_start"].size); total_bundles = int(arrays["b_start"].size)
        arrays = shard_mod.take_bundles(arrays, idx[rank])
        shard_info = {"rank_cost": [float(c) for c in cost], "bundles_this_rank": int(idx[rank].size)}
    n_reads = int(arrays["r_start"].size)
    n_bundles = int(arrays["b_start"].size)
    if args.scaling == "strong" and world > 1:
        job_reads, job_bundles = total_reads, total_bundles
    else:
        job_reads, job_bundles = world * n_reads, world * n_bundles
    in_bytes = synth.total_input_bytes(arrays)

    eng = Engine(local_rank)
    stream = torch.cuda.ExternalStream(eng.L.stg_stream(eng.ctx), device=torch.device("cuda", local_rank))
    totals_dev = torch.zeros(3, dtype=torch.float64, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    stage_wall = {}

    def all_stages(b, after_run=None, after_groups=None):
        t0 = time.perf_counter()
        eng.run(b, sync=False)
        t1 = time.perf_counter()
        if after_run is not None:
            after_run(b)
        t2 = time.perf_counter()
        eng.build_groups(b)
        t3 = time.perf_counter()
        if after_groups is not None:
            after_groups()
        eng.neutral_predictions_only(b)
        eng.assemble(b)
        t4 = time.perf_counter()
        for k, v in (("run_issue", t1 - t0), ("pack_bpcov", t2 - t1), ("build_groups", t3 - t2), ("neutral+assemble", t4 - t3)):
            stage_wall[k] = stage_wall.get(k, 0.0) + v
        if dist is not None:
            # the path's only collective (stringtie.cpp:1490-1497): the global totals of the FPKM / TPM pass, reduced on the
            # device and all-reduced from there — no host round trip
            eng.reduce_totals(b, totals_dev.data_ptr())
            with torch.cuda.stream(stream):
                dist.all_reduce(totals_dev)

    # ---- resident-input measurement -------------------------------------------------------------------------
    batch = eng.upload(arrays)

    def step_resident():
        eng.rewind(batch)
        all_stages(batch)

    sampler = ClockSampler(local_rank)
    for _ in range(args.warmup):
        step_resident()
    eng.sync()
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
        step_resident(); eng.sync()
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
    pr = eng.fetch_predictions(batch)
    results = {"graphs": int(eng.fetch_asm_array(batch, "G_BUNDLE").size), "graph_nodes": int(eng.fetch_asm_array(batch, "N_START").size),
               "transfrags": int(eng.fetch_asm_array(batch, "G_NTF").sum()), "predictions": int(pr["start"].size),
               "exons": int(pr["e_start"].size), "groups": int(eng.fetch_groups_array(batch, "GROUP_OFF")[-1])}
    batch.free()

    # ---- end-to-end through the C ABI from pinned host buffers ------------------------------------------------
    e2e = None
    if not args.skip_e2e:
        # this process drives the GPU from Python beside its uploader / unpacking threads: a thread that comes back from a C call
        # waits for the interpreter lock up to the switch interval (5 ms by default) — keep that out of the measurement
        sys.setswitchinterval(float(os.environ.get("STGPU_BENCH_SWITCH_INTERVAL", "1e-4")))
        import gc
        gc.collect(); gc.freeze(); gc.disable()   # no collector pauses inside the timed passes either
        from stringtie_b200.stgb import INPUT_FIELDS
        pinned = {k: torch.from_numpy(np.ascontiguousarray(arrays[k])).pin_memory() for k in INPUT_FIELDS}
        parr = {k: v.numpy() for k, v in pinned.items()}
        pb, keep = make_packed(parr)
        cov_host = None
        pk = None
        max_len = 0
        # the host's worker threads: this rank's share of the cores, less the one that drives the GPU (at most 6: more only get in its way)
        host_threads = max(1, min(6, (os.cpu_count() or 4) // max(1, world) - 1))
        if os.environ.get("STGPU_BENCH_UNPACK_THREADS"):
            host_threads = max(1, int(os.environ["STGPU_BENCH_UNPACK_THREADS"]))
        if not args.e2e_no_bpcov:
            try:   # BundleData::bpcov of every bundle comes back too: the reference's printResults reads it on the host
                b0 = eng.upload_packed(pb, keep)
                eng.run(b0)
                cov_floats = eng.cov_total_floats(b0)
                nt, nraw = eng.pack_bpcov(b0)
                b0.free()
                cov_host = True   # (bpcov comes back; the workers unpack it bundle by bundle into their own arrays)
                max_len = int((arrays["b_end"].astype(np.int64) - arrays["b_start"].astype(np.int64) + 3).max())
                cap = nraw + nraw // 8 + 64
                pk = [{"desc": torch.empty(3 * nt, dtype=torch.int32).pin_memory(), "val": torch.empty(3 * nt, dtype=torch.float32).pin_memory(),
                       "raw": torch.empty(cap * 512, dtype=torch.float32).pin_memory(), "cap": cap} for _ in range(2)]
            except Exception as ex:
                cov_host = None
                config["e2e_bpcov"] = f"not copied back (host allocation failed: {ex})"
        d2h = [0]
        pk_stats = {}

        from concurrent.futures import ThreadPoolExecutor
        uploader = ThreadPoolExecutor(1)   # stg_upload of batch i+1 (validation + H2D on the copy-in stream) beside batch i
        unpackers = ThreadPoolExecutor(host_threads)   # the host's worker threads: bpcov of their bundles out of the packed form
        pending = []                       # (batch, futures): unpacking of an earlier batch that may still run

        def finish_pending(keep_last=0):
            while len(pending) > keep_last:
                pb_, futs = pending.pop(0)
                for f in futs:
                    f.result()
                pb_.free()

        seq = [0]

        def start_cov_copy(b):
            if cov_host is None:
                return
            k = seq[0] % 2
            # the buffers of set k were last used two batches ago: that unpacking must be over
            finish_pending(keep_last=1)
            nt_, nraw_ = eng.pack_bpcov(b)      # on the device, right behind the coverage stage
            if nraw_ > pk[k]["cap"]:
                raise SystemExit("bench.py: packed bpcov outgrew its host buffer")
            pk_stats.update({"sub_tile_lanes": 3 * nt_, "lanes_sent_in_full": nraw_})
            if os.environ.get("STGPU_BENCH_D2H_EARLY") == "1":   # (experiment knob: the copy beside the group stage costs its walk kernels 40 ms)
                eng.fetch_bpcov_packed_begin(b, pk[k]["desc"].data_ptr(), pk[k]["val"].data_ptr(), pk[k]["raw"].data_ptr())

        def start_cov_copy_late(b):
            # the packed bpcov leaves on the copy-out stream beside the assembly stages, not beside the read loop of rows a8 + a9: its
            # latency-bound walk kernels lose a third of their speed to a concurrent device-to-host copy (measured: 173 against 121 ms)
            if cov_host is not None and os.environ.get("STGPU_BENCH_D2H_EARLY") != "1":
                k = seq[0] % 2
                eng.fetch_bpcov_packed_begin(b, pk[k]["desc"].data_ptr(), pk[k]["val"].data_ptr(), pk[k]["raw"].data_ptr())

        phases = {}

        def tick(name, t0):
            t = time.perf_counter()
            phases[name] = phases.get(name, 0.0) + (t - t0)
            return t

        def consume(b, after_groups=None):
            t = time.perf_counter()
            def behind_groups():
                start_cov_copy_late(b)
                if after_groups is not None:
                    after_groups()
            all_stages(b, after_run=start_cov_copy, after_groups=behind_groups)
            t = tick("stages", t)
            if e2e_prof:
                eng.sync()
                for name, ms in eng.kernel_times():
                    e2e_kernel.setdefault(name, []).append(ms)
            p = eng.fetch_predictions(b)
            n = eng.fetch_neutral(b)
            nf = eng.fetch_groups_array(b, "NUM_FRAGMENTS"); fl = eng.fetch_groups_array(b, "FRAG_LEN")
            nbytes = sum(x.nbytes for x in p.values()) + sum(x.nbytes for x in n.values()) + nf.nbytes + fl.nbytes
            t = tick("fetch_predictions", t)
            if cov_host is not None:
                k = seq[0] % 2
                eng.fetch_wait(b)
                nbytes += pk[k]["desc"].numel() * 4 + pk[k]["val"].numel() * 4 + pk_stats["lanes_sent_in_full"] * 512 * 4
                t = tick("wait_bpcov_copy", t)
                eng.release_device(b)      # device memory back to the pool; the handle keeps the layout tables for the unpacking
                nb_ = b.n_bundles
                step = (nb_ + 4 * host_threads - 1) // (4 * host_threads)
                def unpack_share(lo, hi, k=k, b=b):
                    # a worker thread: BundleData::bpcov of its bundles, one after the other, in its own (reused) arrays
                    return eng.unpack_bpcov_stream(b, lo, hi, pk[k]["desc"].data_ptr(), pk[k]["val"].data_ptr(), pk[k]["raw"].data_ptr(),
                                                   np.empty(3 * max_len, dtype=np.float32))
                futs = [unpackers.submit(unpack_share, lo, min(nb_, lo + step)) for lo in range(0, nb_, step)] if os.environ.get("STGPU_BENCH_SKIP_UNPACK") != "1" else []
                pending.append((b, futs))
                seq[0] += 1
            else:
                b.free()
            d2h[0] = nbytes
            tick("free", t)

        e2e_prof = os.environ.get("STGPU_BENCH_E2E_PROFILE") == "1"   # experiment: device time of every stage inside the e2e passes
        e2e_kernel = {}
        if e2e_prof:
            eng.set_profiling(True)
        late_upload = os.environ.get("STGPU_BENCH_LATE_UPLOAD") == "1"   # experiment: upload of batch i+1 only behind the group stage of batch i

        def e2e_pass(n):
            fut = uploader.submit(eng.upload_packed, pb, keep)
            for i in range(n):
                t = time.perf_counter()
                b = fut.result()
                tick("wait_upload_issue", t)
                nxt = {}
                def submit_next():
                    nxt["fut"] = uploader.submit(eng.upload_packed, pb, keep)
                if i + 1 < n and not late_upload:
                    submit_next()
                consume(b, after_groups=submit_next if (i + 1 < n and late_upload) else None)
                if "fut" in nxt:
                    fut = nxt["fut"]
            t = time.perf_counter()
            finish_pending()           # the last batch's bpcov is unpacked inside the timed region too
            tick("wait_last_unpack", t)

        e2e_pass(max(1, min(args.warmup, 2)))
        barrier()
        t0 = time.perf_counter()
        e2e_pass(1)
        barrier()
        single = time.perf_counter() - t0
        # (the first batch of a pass waits ~90 ms for its own upload and the last one's unpacking has nothing to hide behind: with
        # few batches this fill and drain is a fifth of the measurement, so the timed pass is 2 x steps batches, at most 10)
        n_e2e = max(1, min(2 * args.steps, 10))
        barrier()
        phases.clear(); stage_wall.clear()
        pool0 = eng.pool_stats()
        t0 = time.perf_counter()
        e2e_pass(n_e2e)
        barrier()
        wall = (time.perf_counter() - t0) / n_e2e
        pool1 = eng.pool_stats()
        uploader.shutdown(); unpackers.shutdown()
        if dist is not None:
            t = torch.tensor([wall, single], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall, single = float(t[0].item()), float(t[1].item())
        e2e = {"value": job_reads / wall, "unit": "reads/s", "h2d_bytes_per_step": in_bytes,
               "d2h_bytes_per_step": d2h[0], "ms_per_step": 1e3 * wall, "steps": n_e2e,
               "single_batch_ms": 1e3 * single,
               "host_phases_ms_per_step": {k: 1e3 * v / n_e2e for k, v in phases.items()},
               "stages_host_wall_ms_per_step": {k: 1e3 * v / n_e2e for k, v in stage_wall.items()},
               "device_ms_inside_e2e": {k: statistics.mean(v[-n_e2e:]) for k, v in e2e_kernel.items()} if e2e_prof else None,
               "bpcov_copied_back": cov_host is not None,
               "device_pool": {"cudaMalloc_calls_in_timed_region": pool1["cudaMalloc_calls"] - pool0["cudaMalloc_calls"],
                               "cudaFree_calls_in_timed_region": pool1["cudaFree_calls"] - pool0["cudaFree_calls"],
                               "pool_flushes_in_timed_region": pool1["pool_flushes"] - pool0["pool_flushes"], "pool_bytes": pool1["pool_bytes"]},
               "bpcov_packed": dict(pk_stats, host_threads_unpacking=host_threads) if cov_host is not None else None,
               "timing": "host wall clock around n batches, each: stg_upload (validation + H2D of every input array from pinned "
                         "memory, on the copy-in stream) + stg_run + stg_build_groups + stg_neutral_predictions + stg_assemble + D2H of "
                         "BundleData::pred (neutral and graph predictions with exons), num_fragments, frag_len and — unless "
                         "--e2e-no-bpcov — ALL of bpcov (the reference's printResults, which stays on the host, reads the three lanes): packed on "
                         "the device behind the coverage stage (a sub-tile lane that repeats one value travels as that value: "
                         "stg_pack_bpcov), copied on the copy-out stream into pinned memory and unpacked by the host's worker threads, "
                         "every bundle into the thread's own reused arrays as the reference's workers hold BundleData::bpcov "
                         "(stg_unpack_bpcov_stream), while the next batch computes + "
                         "stg_free_batch.  The upload of batch i+1 is issued while batch i computes, as the bundle queue of the "
                         "host does; every H2D and D2H byte of the n batches is inside the timed region.  single_batch_ms is one "
                         "batch alone (nothing to overlap its upload with).  Max over ranks"}

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
    cov_kernels = [k for k in kmean if not (k.startswith("groups") or k.startswith("asm") or k.startswith("neutral") or k == "find_trims")]
    t_cov = sum(kmean[k] for k in cov_kernels) or float("nan")
    t_tile = sum(kmean.get(k, 0.0) for k in ("deep_deposit", "cov_tile", "cov_fill")) or float("nan")
    alg_cov = 24.0 * L + 13.0 * S + 8.0 * J       # DESIGN.md §6 / SURVEY.md §8d: algorithmic bytes of the coverage stage
    traffic = None
    try:  # DRAM bytes of the bpcov kernels from the committed ncu --set full capture (valid for the default workload shape)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r4_traffic.json")))
        traffic = float(tj["traffic_bytes_per_pass"])
    except Exception:
        pass
    t_all = sum(kmean.values())
    dom = max(kmean, key=kmean.get) if kmean else None
    roofline = {"bound": "hbm", "kernel": "coverage stage: every kernel of count_good_junctions + the junction pass (rows a1-a7)",
                "achieved": alg_cov / (t_cov * 1e-3) / 1e9, "peak": peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 (B200_PROFILING.md)",
                "unit": "GB/s", "frac": alg_cov / (t_cov * 1e-3) / 1e9 / peak, "traffic": traffic,
                "algorithmic_bytes": alg_cov, "formula": "24*L + 13*S + 8*J (SURVEY.md §8d) over all kernels of the stage",
                "L": L, "S": S, "J": J, "kernel_ms": t_cov, "share_of_step": t_cov / t_all if t_all else None,
                "bpcov_kernels": {"kernels": "deep_deposit + cov_tile + cov_fill", "ms": t_tile,
                                  "achieved": (24.0 * L) / (t_tile * 1e-3) / 1e9, "formula": "24*L (difference array read + bpcov write)"},
                "dominant_kernel": {"name": dom, "ms": kmean.get(dom) if dom else None, "share_of_step": (kmean[dom] / t_all) if dom and t_all else None,
                                    "bound": "latency of the deepest bundle (one CTA / warp walks it in read order): not an HBM-bound kernel"}}

    cpu_baseline = None
    if not args.skip_cpu_baseline and world == 1:   # the CPU baseline is reported at N = 1 only (the reference arm covers every N)
        cpu_baseline = reference_hot(base, 2000, os.cpu_count() or 1, 2, stages=("--full", "--a7", "--a9b", "--a11"))

    line = {"metric": METRIC, "value": job_reads / (ms_step * 1e-3), "unit": "reads/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config, "bundles_per_sec": job_bundles / (ms_step * 1e-3), "reads_per_gpu": n_reads,
            "bundles_per_gpu": n_bundles, "gpu_launches": launches, "clocks": clocks, "e2e": e2e, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "kernel_ms": kmean, "kernel_ms_total": t_all, "results": results}
    if shard_info:
        line["shard"] = shard_info
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
