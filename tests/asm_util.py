"""Helpers of the rows a10-a15 parity tests: the reference side is oracle/_ref/ref_hot --full (the reference's OWN
build_graphs from the packed batch to the CPrediction list, with the stage dumps a10 / a11 / a12 taken inside it), the
product side is Engine.assemble."""
import os
import subprocess
import tempfile

import numpy as np

from stringtie_b200.stgb import input_view, read_stgb, write_stgb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_HOT = os.path.join(ROOT, "oracle", "_ref", "ref_hot")


def have_ref_hot():
    return os.path.exists(REF_HOT)


def ref_full(arrays, threads=1):
    """Stage dumps + predictions of the reference for a packed batch (all of infer_transcripts, single thread)."""
    with tempfile.TemporaryDirectory() as td:
        inp, outp = os.path.join(td, "in.stgb"), os.path.join(td, "out.stgb")
        write_stgb(inp, input_view(arrays))
        subprocess.check_call([REF_HOT, inp, "--full", "--dump", outp], stdout=subprocess.DEVNULL)
        return read_stgb(outp)


def gpu_predictions(engine, arrays, keep=False):
    """BundleData::pred of every bundle as the engine leaves it: the neutral single-exon predictions, then the graphs'."""
    a = input_view(arrays)
    batch = engine.upload(a)
    try:
        engine.run(batch)
        engine.build_groups(batch)
        neu = engine.neutral_predictions(batch)
        engine.assemble(batch)
        pr = engine.fetch_predictions(batch)
        stage = {k: engine.fetch_asm_array(batch, k) for k in ("G_BUNDLE", "G_STRAND", "G_K", "G_NODE_OFF", "N_START", "N_END", "N_COV", "G_NTF")}
    finally:
        if not keep:
            batch.free()
    return neu, pr, stage


def assert_same_predictions(neu, pr, ref):
    nb = ref["f_pred_off"].size - 1
    assert pr["off"].size == nb + 1
    for b in range(nb):
        r0, r1 = int(ref["f_pred_off"][b]), int(ref["f_pred_off"][b + 1])
        n0, n1 = int(neu["off"][b]), int(neu["off"][b + 1])
        g0, g1 = int(pr["off"][b]), int(pr["off"][b + 1])
        assert (n1 - n0) + (g1 - g0) == r1 - r0, f"bundle {b}: {r1 - r0} predictions in the reference, {n1 - n0} + {g1 - g0} here"
        k = n1 - n0
        # the neutral ones first
        assert np.array_equal(ref["p_start"][r0:r0 + k], neu["start"][n0:n1].astype(np.int32)), f"bundle {b}: neutral starts"
        assert np.array_equal(ref["p_end"][r0:r0 + k], neu["end"][n0:n1].astype(np.int32)), f"bundle {b}: neutral ends"
        assert np.array_equal(ref["p_cov"][r0:r0 + k].view(np.uint64), neu["cov"][n0:n1].view(np.uint64)), f"bundle {b}: neutral cov"
        assert np.array_equal(ref["p_geneno"][r0:r0 + k], neu["geneno"][n0:n1]), f"bundle {b}: neutral gene numbers"
        q0 = r0 + k
        for rk, gk in (("p_start", "start"), ("p_end", "end"), ("p_tlen", "tlen"), ("p_geneno", "geneno"), ("p_strand", "strand")):
            assert np.array_equal(ref[rk][q0:r1].astype(np.int64), pr[gk][g0:g1].astype(np.int64)), \
                f"bundle {b}: {rk} ref {ref[rk][q0:r1]} got {pr[gk][g0:g1]}"
        assert np.array_equal(ref["p_cov"][q0:r1].view(np.uint64), pr["cov"][g0:g1].view(np.uint64)), \
            f"bundle {b}: cov ref {ref['p_cov'][q0:r1]} got {pr['cov'][g0:g1]}"
        e0, e1 = int(ref["p_exon_off"][q0]), int(ref["p_exon_off"][r1])
        h0, h1 = int(pr["exon_off"][g0]), int(pr["exon_off"][g1])
        assert np.array_equal(np.diff(ref["p_exon_off"][q0:r1 + 1]), np.diff(pr["exon_off"][g0:g1 + 1])), f"bundle {b}: exon counts"
        assert np.array_equal(ref["pe_start"][e0:e1], pr["e_start"][h0:h1]) and np.array_equal(ref["pe_end"][e0:e1], pr["e_end"][h0:h1]), \
            f"bundle {b}: exon coordinates"
        assert np.array_equal(ref["pe_cov"][e0:e1].view(np.uint32), pr["e_cov"][h0:h1].view(np.uint32)), f"bundle {b}: exon coverage"


def assert_same_graph_stage(stage, ref):
    """graphno of every candidate graph, node coordinates (a10) and node coverage after the read mapping (a11)."""
    gno = np.diff(stage["G_NODE_OFF"]).astype(np.int32)
    assert np.array_equal(gno, ref["a10_g_graphno"]), "graphno of the candidate graphs"
    assert np.array_equal(stage["G_STRAND"], ref["a10_g_strand"]) and np.array_equal(stage["G_K"], ref["a10_g_bundle"])
    assert np.array_equal(stage["N_START"], ref["a10_n_start"]) and np.array_equal(stage["N_END"], ref["a10_n_end"]), "node coordinates"
    assert np.array_equal(stage["N_COV"].view(np.uint32), ref["a11_n_cov"].view(np.uint32)), "node coverage"
    assert np.array_equal(stage["G_NTF"], np.diff(ref["a11_tf_off"])), "transfrags per graph after the read mapping"
