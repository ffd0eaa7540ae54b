"""Development tool: runs the per-bundle code of rows a8 + a9 (stringtie_b200/csrc/groups_device.cuh) on the HOST and
compares it with the oracle, bundle by bundle (see tools/groups_hostcheck.cu).  Not used by the library, tests or bench."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build():
    out = os.path.join(ROOT, "tools", "_build", "libgroups_hostcheck.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "orc"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O2", "-fmad=false",
                           "-Xcompiler", "-fPIC", "-shared", "-o", out, os.path.join(ROOT, "tools", "groups_hostcheck.cu"),
                           os.path.join(ROOT, "oracle", "liborc.so"), "-Xlinker", "-rpath," + os.path.join(ROOT, "oracle")])
    return out


def main():
    import orc
    from stringtie_b200 import synth
    L = C.CDLL(build())
    L.hostcheck_groups.argtypes = [C.POINTER(orc.StgPackedBatch), C.c_uint32, C.c_int32, C.c_uint32, C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_int64)]
    L.hostcheck_groups.restype = C.c_int
    cases = []
    gd = os.path.join(ROOT, "tests", "golden")
    for name in ("short_s11", "short_s12_deep", "short_s13_jitter"):
        z = np.load(os.path.join(gd, name + ".npz"))
        cases.append((name, {k: z[k] for k in z.files}))
    for seed in (1, 2, 3):
        cases.append((f"synth{seed}", synth.make_batch(n_bundles=40, seed=seed)))
    cases.append(("deep", synth.make_batch(n_bundles=6, reads_per_bundle=40000, seed=4)))
    cases.append(("adversarial", synth.make_adversarial(seed=5)))
    cases.append(("sparse", synth.make_sparse(n_reads=1500, seed=3, n_bundles=2)))
    for seed in (38, 43):   # the repair appends junction rows and the table is re-sorted
        a = synth.make_adversarial(seed=seed, n_bundles=10, max_reads=250)
        rng = np.random.default_rng(seed)
        a["j_nm"] = np.where(rng.random(a["j_nm"].size) < 0.5, a["j_nreads"], a["j_nm"])
        cases.append((f"appended{seed}", a))
    bad = 0
    for name, a in cases:
        pb, keep = orc.make_packed(a)
        for bd in (50, 0):
            for steps in (0, 32, 128, -32, -128):
                st = (C.c_int64 * 2)(0, 0)
                r = L.hostcheck_groups(C.byref(pb), 10, 1, bd, 0, int(pb.n_bundles), steps, st)
                print(f"{name}: bundledist {bd}, {('steps of %d%s' % (abs(steps), ', growth inside runs' if steps < 0 else '')) if steps else 'sequential'}: {int(pb.n_bundles)} bundles, {r} differ"
                      + (f", {st[0]}/{st[1]} reads simple" if steps else ""))
                bad += r
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
