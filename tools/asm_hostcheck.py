#!/usr/bin/env python3
"""Development aid: runs tools/_build/asm_hostcheck (the per-graph device code compiled for the host) on a reference dump
and compares its a10 / a11 / a12 / prediction arrays with the reference's.  Input: a golden fixture name or a .stgb dump
made by `oracle/_ref/ref_hot <batch.stgb> --full --dump <dump.stgb>`."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stringtie_b200.stgb import read_stgb, write_stgb  # noqa: E402


def load(path_or_name):
    if os.path.exists(path_or_name):
        return read_stgb(path_or_name)
    with np.load(os.path.join(ROOT, "tests", "golden", path_or_name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def compare(ref, got, prefix, keys, tol=None):
    bad = 0
    for k in keys:
        a, b = ref[prefix + k], got[prefix + k]
        if a.shape != b.shape:
            print(f"  {prefix}{k}: shape {a.shape} vs {b.shape}")
            bad += 1
            continue
        if a.dtype.kind == "f":
            same = np.array_equal(a.view(np.uint32 if a.dtype == np.float32 else np.uint64),
                                  b.view(np.uint32 if a.dtype == np.float32 else np.uint64))
        else:
            same = np.array_equal(a, b)
        if not same:
            idx = np.flatnonzero(a != b)
            print(f"  {prefix}{k}: {idx.size} of {a.size} differ, first at {idx[:5]}: ref {a[idx[:5]]} got {b[idx[:5]]}")
            bad += 1
    return bad


def main():
    src = sys.argv[1]
    stage = sys.argv[2] if len(sys.argv) > 2 else "15"
    ref = load(src)
    with tempfile.TemporaryDirectory() as td:
        inp = src if os.path.exists(src) else os.path.join(td, "in.stgb")
        if not os.path.exists(src):
            write_stgb(inp, ref)
        outp = os.path.join(td, "out.stgb")
        rc = subprocess.call([os.path.join(ROOT, "tools", "_build", "asm_hostcheck"), inp, outp, stage])
        got = read_stgb(outp)
    bad = compare(ref, got, "a10_", ["graph_off", "g_strand", "g_bundle", "g_graphno", "g_edgeno", "node_off", "n_start", "n_end",
                                      "n_child_off", "n_child", "tf_off", "tf_abund", "tf_node_off", "tf_nodes"])
    # parents: the sink's parent list is kept as a set here
    if not np.array_equal(ref["a10_n_par_off"], got["a10_n_par_off"]):
        print("  a10_n_par_off differs"); bad += 1
    else:
        off = ref["a10_n_par_off"]
        node_graph_last = set()
        no = ref["a10_node_off"]
        for g in range(no.size - 1):
            if no[g + 1] > no[g]:
                node_graph_last.add(int(no[g + 1]) - 1)
        for i in range(off.size - 1):
            a, b = ref["a10_n_par"][off[i]:off[i + 1]], got["a10_n_par"][off[i]:off[i + 1]]
            if i in node_graph_last:
                a, b = np.sort(a), np.sort(b)
            if not np.array_equal(a, b):
                print(f"  parents of node row {i}: ref {a} got {b}"); bad += 1
                if bad > 20: break
    for st, key in ((11, "a11_n_cov"), (12, "a12_tf_abund"), (15, "hp_start")):
        if int(stage) >= st and key not in got:
            print(f"  stage {st} arrays missing from the output"); bad += 1
    if int(stage) >= 11 and "a11_n_cov" in got:
        bad += compare(ref, got, "a11_", ["graph_off", "node_off", "n_cov", "tf_off", "tf_abund", "tf_node_off", "tf_nodes"])
        rc_cnt = np.diff(ref["a11_tf_bit_off"]).astype(np.int32)
        if not np.array_equal(rc_cnt, got["a11_tf_pcount"]):
            print("  a11 pattern counts differ"); bad += 1
    if int(stage) >= 12 and "a12_tf_abund" in got:
        bad += compare(ref, got, "a12_", ["graph_off", "g_strand", "g_bundle", "tf_off", "tf_abund", "tf_real", "tf_node_off", "tf_nodes",
                                          "tf_path_off", "tf_path_node", "tf_path_cont", "tf_path_ab", "node_off", "ntrf_off", "ntrf"])
    if int(stage) >= 15 and "hp_start" in got:
        n19 = ref["a19_pred_off"]
        # predictions of the graphs: the reference's list minus the neutral single-exon ones in front of it
        ro, go = ref["f_pred_off"], got["hp_pred_off"]
        for b in range(ro.size - 1):
            k = int(n19[b + 1] - n19[b])
            r0, r1 = int(ro[b]) + k, int(ro[b + 1])
            g0, g1 = int(go[b]), int(go[b + 1])
            if r1 - r0 != g1 - g0:
                print(f"  bundle {b}: {r1 - r0} graph predictions in the reference, {g1 - g0} here"); bad += 1
                continue
            for key, hk in (("p_start", "hp_start"), ("p_end", "hp_end"), ("p_tlen", "hp_tlen"), ("p_geneno", "hp_geneno"),
                            ("p_strand", "hp_strand")):
                a_, b_ = ref[key][r0:r1], got[hk][g0:g1]
                if key == "p_geneno" and a_.size:   # the neutral stage's gene counter is not replayed here: compare relative numbers
                    a_, b_ = a_ - a_[0], b_ - b_[0]
                if not np.array_equal(a_, b_):
                    print(f"  bundle {b}: {key} ref {ref[key][r0:r1]} got {got[hk][g0:g1]}"); bad += 1
            if not np.array_equal(ref["p_cov"][r0:r1].view(np.uint64), got["hp_cov"][g0:g1].view(np.uint64)):
                print(f"  bundle {b}: cov ref {ref['p_cov'][r0:r1]} got {got['hp_cov'][g0:g1]}"); bad += 1
            e0, e1 = int(ref["p_exon_off"][r0]), int(ref["p_exon_off"][r1])
            h0, h1 = int(got["hp_exon_off"][g0]), int(got["hp_exon_off"][g1])
            if not (np.array_equal(ref["pe_start"][e0:e1], got["hp_es"][h0:h1]) and np.array_equal(ref["pe_end"][e0:e1], got["hp_ee"][h0:h1])
                    and np.array_equal(ref["pe_cov"][e0:e1].view(np.uint32), got["hp_ecov"][h0:h1].view(np.uint32))):
                print(f"  bundle {b}: exons differ"); bad += 1
    print("hostcheck", src, "stage", stage, "rc", rc, "OK" if (bad == 0 and rc == 0) else f"{bad} DIFFERENCES")
    return 1 if bad or rc else 0


if __name__ == "__main__":
    sys.exit(main())
