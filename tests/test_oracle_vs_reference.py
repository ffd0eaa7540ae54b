"""Pins the CPU restatement (oracle/stg_oracle.c) to the reference: golden vectors dumped from the reference binary
itself (tests/golden/make_golden.py) and, where oracle/_ref/ref_hot is present, live runs of the reference's own
count_good_junctions() on generated packed batches.  Bar: bit-exact float32 bpcov, exact f64 junction support."""
import os
import subprocess

import numpy as np
import pytest

from stringtie_b200 import synth
from stringtie_b200.stgb import read_stgb, write_stgb
from tests import orc
from tests.conftest import ROOT

REF_HOT = os.path.join(ROOT, "oracle", "_ref", "ref_hot")


def _check(arrays, ref):
    cov, off, good, left, right = orc.run(arrays)
    assert cov.size == ref["s1_cov"].size
    assert np.array_equal(cov.view(np.uint32), ref["s1_cov"].view(np.uint32)), "bpcov differs from the reference"
    assert np.array_equal(good, ref["s1_j_nreads_good"])
    assert np.array_equal(left, ref["s1_j_leftsupport"])
    assert np.array_equal(right, ref["s1_j_rightsupport"])


def test_oracle_matches_golden(golden):
    assert golden["b_cov_stage"].min() == 1  # dumped right after count_good_junctions
    assert golden["b_cov_final_same"].all()  # build_graphs never modifies bpcov
    _check(golden, golden)


A7 = (("strand", "a7_j_strand"), ("nreads", "a7_j_nreads"), ("nreads_good", "a7_j_nreads_good"),
      ("leftsupport", "a7_j_leftsupport"), ("rightsupport", "a7_j_rightsupport"), ("mm", "a7_j_mm"), ("ej", "a7_ej"),
      ("good", "a7_j_good"), ("good_strand", "a7_j_good_strand"))


def _check_a7(arrays, ref):
    """the junction table at the end of the junction pass of build_graphs (dumped from inside the
    reference, before rlink.cpp:14843) and the reference's own good_junc verdicts — exact equality."""
    cov, off, good, left, right = orc.run(arrays)
    jp = orc.junction_pass(arrays, cov, off, good, left, right)
    for k, rk in A7:
        bad = np.nonzero(jp[k] != ref[rk])[0]
        assert bad.size == 0, f"a7 {k}: {bad.size} rows differ, first {bad[:5]}: {jp[k][bad[:5]]} vs {ref[rk][bad[:5]]}"


def test_oracle_junction_pass_matches_golden(golden):
    _check_a7(golden, golden)


def test_jitter_golden_exercises_the_demotion_pass():
    from tests.conftest import load_golden
    g = load_golden("short_s13_jitter")
    assert (g["a7_j_nreads"] < 0).sum() >= 10 and (g["a7_j_nreads_good"] < 0).sum() >= 10   # demoted left / right
    assert (g["a7_j_mm"] < 0).sum() >= 10 and (g["a7_j_good"] == 0).sum() >= 10


def _check_a9(arrays, ref):
    """Rows a8 + a9 (first half): state after the read loop of build_graphs and the merging of neighbouring groups
    (dumped from inside the reference, before rlink.cpp:15194): reads, junction table, groups, colours, readgroup,
    fragment counters — exact equality."""
    from tests import a9_util
    cov, off, good, left, right = orc.run(arrays)
    a9_util.assert_same(orc.groups(arrays, cov, off, good, left, right), a9_util.split_reference_dump(ref, arrays))


def test_oracle_read_loop_and_groups_match_golden(golden):
    from stringtie_b200.stgb import input_view
    _check_a9(input_view(golden), golden)


def _check_trims(arrays, ref):
    """find_all_trims (part of row a10): the reference's own function, called from the dump hook on every bundlenode and
    on its halves, against the restatement — positions, abundances (float32 bits), source/sink flags, rejected points."""
    cov, off, *_ = orc.run(arrays)
    o, p, f, t = orc.find_trims(arrays, cov, off, ref["ft_bundle"], ref["ft_sno"], ref["ft_start"], ref["ft_end"])
    assert np.array_equal(o, ref["ft_off"]) and np.array_equal(p, ref["ft_pos"]) and np.array_equal(t, ref["ft_isstart"])
    assert np.array_equal(f.view(np.uint32), ref["ft_abund"].view(np.uint32))


def test_oracle_find_all_trims_matches_golden(golden):
    from stringtie_b200.stgb import input_view
    assert golden["ft_start"].size > 100
    _check_trims(input_view(golden), golden)


def _check_a19(arrays, ref):
    """Row a19 (de novo part): the single-exon predictions of the neutral bundles against the prediction list dumped
    inside the reference's build_graphs right after that section (hook before rlink.cpp:15637)."""
    cov, off, good, left, right = orc.run(arrays)
    got = orc.neutral_preds(arrays, cov, off, good, left, right)
    po = ref["a19_pred_off"].astype(np.int64)
    assert po.size == len(got) + 1
    for b, g in enumerate(got):
        w0, w1 = int(po[b]), int(po[b + 1])
        assert g["start"].size == w1 - w0, f"bundle {b}: {g['start'].size} predictions, reference {w1 - w0}"
        assert np.array_equal(g["start"], ref["a19_start"][w0:w1]) and np.array_equal(g["end"], ref["a19_end"][w0:w1])
        assert np.array_equal(g["cov"], ref["a19_cov"][w0:w1]), f"bundle {b}: coverage differs"   # f64, exactly
        assert np.array_equal(g["tlen"], ref["a19_tlen"][w0:w1]) and np.array_equal(g["geneno"], ref["a19_geneno"][w0:w1])
        assert np.array_equal(g["cov"].astype(np.float32), ref["a19_exoncov"][w0:w1])
    assert np.all(ref["a19_strand"] == ord("."))


def test_oracle_neutral_predictions_match_golden(golden):
    from stringtie_b200.stgb import input_view
    _check_a19(input_view(golden), golden)


def test_goldens_hold_neutral_predictions():
    from tests.conftest import load_golden, GOLDEN
    assert sum(load_golden(n)["a19_start"].size for n in GOLDEN) >= 15


def test_golden_has_inexact_weights(golden):
    # the fixtures must exercise order-dependent float accumulation (NH=3,5,6 -> 1/3, 1/5, 1/6)
    assert np.isin(golden["r_nh"], [3, 5, 6]).any()


@pytest.mark.skipif(not os.path.exists(REF_HOT), reason="reference harness not built (oracle/_ref is built from /root/reference)")
@pytest.mark.parametrize("case", ["synth", "wide", "adv1", "adv2"])
def test_oracle_matches_live_reference(tmp_path, case):
    if case == "synth":
        a = synth.make_batch(n_bundles=40, reads_per_bundle=800, seed=3)
    elif case == "wide":   # many loci: dozens of neutral bundles and single-exon predictions
        a = synth.make_batch(n_bundles=300, reads_per_bundle=500, seed=8)
    else:
        a = synth.make_adversarial(seed=int(case[-1]) + 10, n_bundles=14)
    src, dst = str(tmp_path / "in.stgb"), str(tmp_path / "out.stgb")
    write_stgb(src, a)
    subprocess.check_call([REF_HOT, src, "--dump", dst, "--a19"], stdout=subprocess.DEVNULL)
    ref = read_stgb(dst)
    _check(a, ref)
    _check_a7(a, ref)
    _check_a9(a, ref)
    _check_trims(a, ref)
    _check_a19(a, ref)


@pytest.mark.skipif(not os.path.exists(REF_HOT), reason="reference harness not built (oracle/_ref is built from /root/reference)")
@pytest.mark.parametrize("seed", [38, 43])
def test_oracle_read_loop_appended_junctions(tmp_path, seed):
    """Seeds on which the repair appends a junction row and the table is re-sorted with the reference's quicksort."""
    a = synth.make_adversarial(seed=seed, n_bundles=10, max_reads=250)
    rng = np.random.default_rng(seed)
    a["j_nm"] = np.where(rng.random(a["j_nm"].size) < 0.5, a["j_nreads"], a["j_nm"])
    src, dst = str(tmp_path / "in.stgb"), str(tmp_path / "out.stgb")
    write_stgb(src, a)
    subprocess.check_call([REF_HOT, src, "--dump", dst, "--a9b"], stdout=subprocess.DEVNULL)
    ref = read_stgb(dst)
    assert ref["a9_j_start"].size > a["j_start"].size
    _check_a9(a, ref)


def test_get_cov_against_bruteforce(golden):
    """get_cov on the blocked prefix layout == plain sum of the clamped coverage, across BSIZE block borders."""
    a = golden
    cov, off, *_ = orc.run(a)
    lens = a["b_end"].astype(np.int64) - a["b_start"] + 3
    b = int(np.argmax(lens))
    L = int(lens[b])
    assert L > 20000
    c = cov[off[b]:off[b + 1]].copy()
    lane1 = c[L:2 * L].astype(np.float64)
    # undo the blocked prefix sums to per-base coverage
    per = np.empty(L)
    for i in range(L):
        per[i] = lane1[i] - (lane1[i - 1] if i % 10000 else 0.0)
    rng = np.random.default_rng(0)
    for _ in range(200):
        s = int(rng.integers(0, L - 3))
        e = int(rng.integers(s, L - 2))
        got = orc.get_cov(c, L, 1, s, e)
        want = per[s + 1:e + 2].sum()
        # the stored prefix sums are float32: differences carry ~ulp(prefix) absolute error
        assert abs(got - want) <= 1e-3 * max(1.0, want) + 4 * np.finfo(np.float32).eps * lane1.max()


def test_stats_counts(golden):
    L, S, J = orc.stats(golden)
    lens = golden["b_end"].astype(np.int64) - golden["b_start"] + 3
    assert L == int(lens.sum())
    assert J == golden["seg_start"].size - golden["r_start"].size
    assert S > 0


def test_reference_graph_dump_is_consistent(golden):
    """Row a10 groundwork: the splice graphs dumped right after the reference's create_graph (hook before
    rlink.cpp:15784, arrays a10_*) are structurally sound and line up with the stranded bundles of the a9b dump — the
    next round pins its create_graph restatement on them."""
    nb = golden["b_start"].size
    go = golden["a10_graph_off"].astype(np.int64)
    bo = golden["a9b_bun_off"].astype(np.int64)
    assert go.size == nb + 1
    no, po, co = (golden[k].astype(np.int64) for k in ("a10_node_off", "a10_n_par_off", "a10_n_child_off"))
    seen_graph = False
    for b in range(nb):
        n_neg, n_pos = int(bo[3 * b + 1] - bo[3 * b]), int(bo[3 * b + 3] - bo[3 * b + 2])
        assert go[b + 1] - go[b] == n_neg + n_pos                       # one record per stranded bundle
        for g in range(int(go[b]), int(go[b + 1])):
            gn = int(golden["a10_g_graphno"][g])
            n0, n1 = int(no[g]), int(no[g + 1])
            if gn == 0:
                assert n1 == n0
                continue
            seen_graph = True
            assert n1 - n0 == gn                                        # source, the nodes, sink
            st, en = golden["a10_n_start"][n0:n1].astype(np.int64), golden["a10_n_end"][n0:n1].astype(np.int64)
            assert st[0] == 0 and en[0] == 0 and st[-1] == 0 and en[-1] == 0   # source and sink carry no interval
            ist, ien = st[1:-1], en[1:-1]
            assert np.all(ist >= golden["b_start"][b]) and np.all(ien <= golden["b_end"][b]) and np.all(ien >= ist)
            assert np.all(ist[1:] > ien[:-1])                           # nodes in genomic order, disjoint
            for i in range(n0, n1):                                     # parent / child lists mirror each other
                for c in golden["a10_n_child"][co[i]:co[i + 1]]:
                    if 0 <= c < n1 - n0:
                        assert (i - n0) in golden["a10_n_par"][po[n0 + c]:po[n0 + c + 1]]
    assert seen_graph
