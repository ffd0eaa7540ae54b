#!/usr/bin/env python3
"""Wall time of the reference program on one SAM file: stock (oracle/_ref/stringtie, CPU) against the same program with
infer_transcripts() served by libstgpu through the coalescing adaptor (host/_build/stringtie_stgpu), for several -p.
The two GTFs are compared (up to gene numbering, which follows completion order with -p > 1).
usage: host_walltime.py [--loci N] [--pairs M] [--threads 1,4,16]"""
import argparse, json, os, re, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--loci", type=int, default=1000)
ap.add_argument("--pairs", type=int, default=400000, help="read pairs in total")
ap.add_argument("--threads", default="1,4,16")
ap.add_argument("--stgpu-threads", default="", help="extra -p values for the libstgpu build only: its workers wait for the GPU, so more of them than cores "
                "means larger device batches (the adaptor coalesces the concurrent calls)")
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
tmp = tempfile.mkdtemp()
sam = os.path.join(tmp, "in.sam")
subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_sam.py"), "--sam", sam, "--loci", str(a.loci), "--pairs", str(a.pairs),
                       "--seed", "21", "--long-intron-frac", "0.01", "--nh-frac", "0.1", "--jitter-frac", "0.03"], stderr=subprocess.DEVNULL)
ref = os.path.join(ROOT, "oracle", "_ref", "stringtie")
gpu = os.path.join(ROOT, "host", "_build", "stringtie_stgpu")
norm = lambda path: sorted(re.sub(r'(gene_id|transcript_id) "STRG\.[0-9.]+";', r'\1 "";', l) for l in open(path) if not l.startswith("#"))
out = {"sam_bytes": os.path.getsize(sam), "loci": a.loci, "pairs": a.pairs, "runs": []}
for p in [int(x) for x in a.threads.split(",")]:
    row = {"threads": p}
    for name, exe in (("reference", ref), ("stgpu", gpu)):
        gtf = os.path.join(tmp, f"{name}_{p}.gtf")
        best = None
        for rep in range(a.reps):
            t0 = time.perf_counter()
            r = subprocess.run([exe, sam, "-p", str(p), "-o", gtf], env=dict(os.environ, STGPU_VERBOSE="1"), capture_output=True, text=True)
            dt = time.perf_counter() - t0
            assert r.returncode == 0, r.stderr[-1500:]
            best = dt if best is None else min(best, dt)
        row[name + "_s"] = round(best, 3)
        if name == "stgpu":
            row["adaptor"] = [l for l in r.stderr.splitlines() if l.startswith("libstgpu:")][-1]
    row["same_gtf"] = norm(os.path.join(tmp, f"reference_{p}.gtf")) == norm(os.path.join(tmp, f"stgpu_{p}.gtf"))
    out["runs"].append(row)
ref_gtf = os.path.join(tmp, "reference_%d.gtf" % int(a.threads.split(",")[-1]))
for p in [int(x) for x in a.stgpu_threads.split(",") if x]:
    gtf = os.path.join(tmp, f"stgpu_x{p}.gtf")
    best = None
    for rep in range(a.reps):
        t0 = time.perf_counter()
        r = subprocess.run([gpu, sam, "-p", str(p), "-o", gtf], env=dict(os.environ, STGPU_VERBOSE="1"), capture_output=True, text=True)
        dt = time.perf_counter() - t0
        assert r.returncode == 0, r.stderr[-1500:]
        best = dt if best is None else min(best, dt)
    out["runs"].append({"threads": p, "stgpu_s": round(best, 3), "adaptor": [l for l in r.stderr.splitlines() if l.startswith("libstgpu:")][-1],
                        "same_gtf": norm(ref_gtf) == norm(gtf)})
print(json.dumps(out, indent=1))
